#!/bin/bash
# Multi-GPU run on one box (gpurun --gpus N): the sharded parity tests (tests/test_gpu_multi.py: every case whose rank count
# fits), then the benches as the driver launches them (torchrun, one rank per GPU): the headline config (weak scaling) and the
# GRCh38-shaped configs [3] (sharded depth base) and [4] (depth region -L) -- every line verified against the CPU oracle.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_multi.sh 2'
set -u
N=${1:-2}
OUT=gpurun_out/multi_n$N; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(quiet=True)" > $OUT/build.log 2>&1
nvidia-smi -L > $OUT/gpus.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs -p no:cacheprovider --timeout 600 > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_chr20.json 2> $OUT/bench_chr20.err
timeout 900 $TR --master-port 29512 bench.py --config wgs-shard --gpus $N --steps 3 --warmup 2 --no-cpu-baseline > $OUT/bench_wgs_shard.json 2> $OUT/bench_wgs_shard.err
timeout 900 $TR --master-port 29513 bench.py --config exome --gpus $N --steps 3 --warmup 2 --no-cpu-baseline > $OUT/bench_exome.json 2> $OUT/bench_exome.err
tail -4 $OUT/pytest_multi.log; for f in $OUT/bench_*.json; do echo $f; grep -o '"verified": [a-z]*' $f; grep -o '"value": [0-9.]*, "unit": "GB/s", "n_gpus": [0-9]*' $f; grep -o '"e2e": {"value": [0-9.]*' $f; grep -o '"ms_d2h": [0-9.]*' $f; done; tail -3 $OUT/*.err | cut -c1-300
