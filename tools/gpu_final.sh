#!/bin/bash
# Final single-GPU call of the round: the GPU tests that the per-change A/B calls did not repeat, the headline bench of the shipped build
# (verified, CPU baseline beside it), the phase-2 A/B, BASELINE configs[4] on one GPU, and the ncu evidence of exactly this build.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_final.sh'
set -u
OUT=gpurun_out/final; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(quiet=True)" > $OUT/build.log 2>&1
timeout 120 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
timeout 900 python -m pytest tests/test_zz_gpu_experiments.py tests/test_zz_gpu_mates.py tests/test_zz_gpu_filter.py tests/test_zz_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
BDEPTH_K1LZ=flat timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-verify > $OUT/bench_n1_k1lzflat.json 2> $OUT/bench_n1_k1lzflat.err
timeout 600 python bench.py --config exome --steps 3 --warmup 2 --no-cpu-baseline > $OUT/bench_exome_n1.json 2> $OUT/bench_exome_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify > $OUT/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_huff|k1_lz|k3_tile|k2_decode' -c 4 -f -o $OUT/r2_final python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify > $OUT/ncu_final.log 2>&1
python tools/ncu_summary.py $OUT/r2_final.ncu-rep 30 > $OUT/r2_final_ncu_full_summary.txt 2>&1
BDEPTH_K1LZ=flat timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k1_lz_flat' -c 1 -f -o $OUT/r2_k1lzflat python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify > $OUT/ncu_flat.log 2>&1
python tools/ncu_summary.py $OUT/r2_k1lzflat.ncu-rep 30 > $OUT/r2_k1lzflat_ncu_full_summary.txt 2>&1
tail -2 $OUT/smoke.log; tail -3 $OUT/pytest.log; grep -o '"verified": [a-z]*' $OUT/*.json; grep -o '"k1_inflate": [0-9.]*' $OUT/*.json; grep -o '"e2e": {"value": [0-9.]*' $OUT/*.json; grep -o '"value": [0-9.]*, "unit": "GB/s", "n_gpus"' $OUT/*.json; tail -2 $OUT/*.err | cut -c1-200
