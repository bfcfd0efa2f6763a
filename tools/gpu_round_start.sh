#!/bin/bash
# First GPU call of a round (run through gpurun from the repository root): the whole GPU suite (gating: no xfail markers
# are left), the headline bench with its bit-exact verification, the prepared kernel variants A/B, and ncu captures of
# the shipped K2 / K3 kernels.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round_start.sh'
# Results land in gpurun_out/round_start/ (merged back by gpurun).
set -u
OUT=gpurun_out/round_start; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(quiet=True)" > $OUT/build.log 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt; free -g >> $OUT/gpu.txt; df -h /tmp >> $OUT/gpu.txt
which ldc2 sambamba >> $OUT/gpu.txt 2>&1
# 1. the whole GPU suite, full-size parity included
timeout 2400 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
# 2. smoke + headline bench (N = 1), verified against the oracle after the timed region
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
# 3. prepared variants (A/B against bench_n1.json: e2e.ms_per_step, stage_ms)
for v in BDEPTH_K1_STREAM_WARPS=4 BDEPTH_K1_LIT3=1 BDEPTH_K3_PREFETCH=1; do
  env $v timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify > $OUT/bench_n1_${v%%=*}.json 2> $OUT/bench_n1_${v%%=*}.err
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
# 5. launch list of one staged pass (cold-cache, serialised: compare shares, not absolutes) and full captures of the shipped K2 / K3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify > $OUT/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k3_gather|k2_decode|k3_tile_index' -c 3 -f -o $OUT/r2_k2_k3 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify > $OUT/ncu_k2k3.log 2>&1
python tools/ncu_summary.py $OUT/r2_k2_k3.ncu-rep 30 > $OUT/r2_k2decode_k3gather_ncu_full_summary.txt 2>&1
tail -3 $OUT/pytest_gpu.log; head -c 900 $OUT/bench_n1.json; echo; grep -o '"verified": [a-z]*' $OUT/bench_n1.json; grep -o '"e2e": {"value": [0-9.]*' $OUT/bench_n1*.json; grep -o '"k3_coverage": [0-9.]*' $OUT/bench_n1*.json; grep -o '"k1_inflate": [0-9.]*' $OUT/bench_n1*.json
