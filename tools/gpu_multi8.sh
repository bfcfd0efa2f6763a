#!/bin/bash
# The 8-GPU call (charged 8x: every second counts): the sharded parity tests at 2 / 3 / 4 / 8 ranks, BASELINE configs[4]
# (`depth region -L exome.bed`, 8 GPUs) verified against the oracle's closed form, configs[3] (`depth base` sharded over 8 GPUs).
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 420 -- 'bash tools/gpu_multi8.sh'
set -u
N=8; OUT=gpurun_out/multi_n8; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(quiet=True)" > $OUT/build.log 2>&1
nvidia-smi -L > $OUT/gpus.txt 2>&1
python - > $OUT/gen.log 2>&1 <<'PY' &
import sys; sys.path.insert(0, ".")
import bench; print(bench.ensure_wgs(load=False))          # generate the GRCh38-shaped input while the tests run
PY
timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs -x -p no:cacheprovider --timeout 100 > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_multi.log
wait
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
BDEPTH_BENCH_WATCHDOG=170 timeout 180 $TR --master-port 29513 bench.py --config exome --gpus $N --steps 3 --warmup 2 --no-cpu-baseline > $OUT/bench_exome.json 2> $OUT/bench_exome.err
BDEPTH_BENCH_WATCHDOG=110 timeout 120 $TR --master-port 29512 bench.py --config wgs-shard --gpus $N --steps 3 --warmup 2 --no-cpu-baseline --no-verify > $OUT/bench_wgs_shard.json 2> $OUT/bench_wgs_shard.err
tail -4 $OUT/pytest_multi.log; for f in $OUT/bench_*.json; do echo $f; grep -o '"verified": [a-z]*' $f; grep -o '"value": [0-9.]*, "unit": "GB/s", "n_gpus": [0-9]*' $f; grep -o '"e2e": {"value": [0-9.]*' $f; done; tail -3 $OUT/*.err | cut -c1-300
