#!/bin/bash
# Kernel A/B on the GPU box (one GPU): smoke under a short timeout first (a hung kernel must not eat the call), the parity / edge /
# CLI / sparse suites, the headline bench with the shipped kernels and with the round-1 ones (BDEPTH_K3=gather, BDEPTH_K1_ONEPHASE=1),
# the phase-1 instantiations, then ncu: launch list + full captures of the K1 and K3 kernels.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_ab.sh'
set -u
OUT=gpurun_out/ab; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(quiet=True)" > $OUT/build.log 2>&1
timeout 180 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
BDEPTH_SKIP_FULLSIZE=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_cli.py tests/test_gpu_sparse.py tests/test_gpu_multibam.py -m gpu -q -x -p no:cacheprovider --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $OUT/bench_default.json 2> $OUT/bench_default.err
BDEPTH_K3=gather timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify > $OUT/bench_k3gather.json 2> $OUT/bench_k3gather.err
BDEPTH_K1H_VARIANT=2 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify > $OUT/bench_v2.json 2> $OUT/bench_v2.err
BDEPTH_BENCH_CHUNK_BLOCKS=13312 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify > $OUT/bench_cb13312.json 2> $OUT/bench_cb13312.err
timeout 900 python bench.py --config wgs-shard --steps 3 --warmup 2 --no-cpu-baseline > $OUT/wgs_default.json 2> $OUT/wgs_default.err
BDEPTH_K1H_VARIANT=2 timeout 600 python bench.py --config wgs-shard --steps 3 --warmup 2 --no-cpu-baseline --no-verify > $OUT/wgs_v2.json 2> $OUT/wgs_v2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify > $OUT/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_huff|k1_lz|k3_tile$|k3_tile<' -c 3 -f -o $OUT/r2_k1_k3 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify > $OUT/ncu_k1k3.log 2>&1
python tools/ncu_summary.py $OUT/r2_k1_k3.ncu-rep 40 > $OUT/r2_k1_k3_ncu_full_summary.txt 2>&1
cat $OUT/smoke.log | tail -2; tail -3 $OUT/pytest.log; grep -o '"verified": [a-z]*' $OUT/*.json; grep -o '"k1_inflate": [0-9.]*' $OUT/*.json; grep -o '"k3_coverage": [0-9.]*' $OUT/*.json; grep -o '"e2e": {"value": [0-9.]*' $OUT/*.json; grep -o '"value": [0-9.]*, "unit": "GB/s", "n_gpus"' $OUT/*.json
