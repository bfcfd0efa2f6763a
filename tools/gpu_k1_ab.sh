#!/bin/bash
# K1 A/B on the GPU box: correctness of the two-phase inflater first (parity + edge cases), then the phase-1 instantiations
# against each other (BDEPTH_K1H_VARIANT; BDEPTH_K1_ONEPHASE=1 is the round-1 kernel), on the chr20 headline workload and on
# the GRCh38-shaped one (98 k blocks per batch: more warps than the SMs hold at once), then ncu: launch list + full captures.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_k1_ab.sh'
set -u
OUT=gpurun_out/k1_ab; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(quiet=True)" > $OUT/build.log 2>&1
BDEPTH_SKIP_FULLSIZE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_cli.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest_k1.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_k1.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $OUT/bench_v0.json 2> $OUT/bench_v0.err
for v in 2 3; do BDEPTH_K1H_VARIANT=$v timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify > $OUT/bench_v$v.json 2> $OUT/bench_v$v.err; done
timeout 900 python bench.py --config wgs-shard --steps 3 --warmup 2 --no-cpu-baseline > $OUT/wgs_v0.json 2> $OUT/wgs_v0.err
for v in 2 3; do BDEPTH_K1H_VARIANT=$v timeout 600 python bench.py --config wgs-shard --steps 3 --warmup 2 --no-cpu-baseline --no-verify > $OUT/wgs_v$v.json 2> $OUT/wgs_v$v.err; done
timeout 900 python bench.py --config window --steps 3 --warmup 2 > $OUT/window.json 2> $OUT/window.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify > $OUT/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_huff|k1_lz' -c 2 -f -o $OUT/r2_k1 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify > $OUT/ncu_k1.log 2>&1
python tools/ncu_summary.py $OUT/r2_k1.ncu-rep 45 > $OUT/r2_k1_two_phase_ncu_full_summary.txt 2>&1
tail -3 $OUT/pytest_k1.log; grep -o '"verified": [a-z]*' $OUT/*.json; grep -o '"k1_inflate": [0-9.]*' $OUT/*.json; grep -o '"e2e": {"value": [0-9.]*' $OUT/*.json; grep -o '"value": [0-9.]*, "unit": "GB/s", "n_gpus"' $OUT/*.json
