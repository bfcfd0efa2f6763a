#!/bin/bash
# First GPU call of a round (run through gpurun from the repository root): everything that was written without hardware
# gets its first run, the headline numbers are refreshed, and the K1 CTA-shape experiment is measured A/B.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round_start.sh'
# Results land in gpurun_out/round_start/ (merged back by gpurun).
set -u
OUT=gpurun_out/round_start; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(quiet=True)" > $OUT/build.log 2>&1
# 1. the whole GPU suite; -rxX lists the non-gating tests (xfail / XPASS) of test_zz_gpu_*.py with their outcome
BDEPTH_FULLSIZE=1 timeout 2400 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
# 2. smoke + headline bench (N = 1)
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
# 3. K1 experiment: 4-warp CTAs for the streaming sub-launches (kernels.cuh k1_inflate_small); compare e2e.ms_per_step
BDEPTH_K1_STREAM_WARPS=4 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_n1_k1small.json 2> $OUT/bench_n1_k1small.err
BDEPTH_K1_LIT3=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_n1_k1lit3.json 2> $OUT/bench_n1_k1lit3.err             # compare stage_ms.k1_inflate
BDEPTH_K3_PREFETCH=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_n1_k3pre.json 2> $OUT/bench_n1_k3pre.err      # compare stage_ms.k3_coverage
for cb in 1664 3328 13312; do       # H2D chunk / sub-batch size sweep (default 6656 blocks), with and without the 4-warp CTAs
  BDEPTH_BENCH_CHUNK_BLOCKS=$cb timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $OUT/bench_n1_cb$cb.json 2> $OUT/bench_n1_cb$cb.err
  BDEPTH_BENCH_CHUNK_BLOCKS=$cb BDEPTH_K1_STREAM_WARPS=4 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $OUT/bench_n1_cb${cb}_k1small.json 2> $OUT/bench_n1_cb${cb}_k1small.err
done
# 4. -m on the bench workload with real pairs: cost of km_hash / km_link / km_fix (ms_mates) next to K3
timeout 900 python - > $OUT/mates_probe.log 2>&1 <<'PY'
import os, sys, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers, sambamba_b200 as sb
os.makedirs("/tmp/bdepth_bench", exist_ok=True)
p = "/tmp/bdepth_bench/pairs_chr20.bam"
if not os.path.exists(p):
    helpers.gen_bam(p, "--preset", "chr20", "--pairs", "24", "-t", "64")
with sb.BDepth(p) as b:
    b.set_fix_mates(True); b.stage()
    for i in range(3):
        b.run_resident(); st = b.stats()
        print(json.dumps({k: st[k] for k in ("ms_inflate", "ms_scan", "ms_coverage", "ms_mates", "mate_pairs", "mate_pair_columns", "mate_groups", "n_records_pass", "covered_positions")}))
PY
# 5. launch list of one staged pass (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
tail -3 $OUT/pytest_gpu.log; cat $OUT/bench_n1.json | head -c 600; echo; grep -o '"e2e": {"value": [0-9.]*' $OUT/bench_n1*.json; grep -o '"k3_coverage": [0-9.]*' $OUT/bench_n1.json $OUT/bench_n1_k3pre.json; grep -o '"k1_inflate": [0-9.]*' $OUT/bench_n1.json $OUT/bench_n1_k1lit3.json

# Second call, on two GPUs (the sharded path: sub-batches per rank and -m across shard boundaries have only run under the
# CPU emulation so far):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'python -m pytest tests/test_gpu_multi.py -m gpu -q -rxX -p no:cacheprovider > gpurun_out/round_start_multi.log 2>&1; \
#     python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err'
