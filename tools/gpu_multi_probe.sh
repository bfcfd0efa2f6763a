#!/bin/bash
# Short, time-boxed probe of the multi-rank bench (small workload, watchdog that prints where a rank is stuck).
set -u
N=${1:-2}; OUT=gpurun_out/multi_probe; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(quiet=True)" > $OUT/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
BDEPTH_BENCH_WATCHDOG=150 timeout 200 $TR --master-port 29521 bench.py --gpus $N --steps 1 --warmup 1 --no-cpu-baseline --reads-per-unit 1500000 > $OUT/bench_small.json 2> $OUT/bench_small.err; echo "exit $?" >> $OUT/bench_small.err
BDEPTH_BENCH_WATCHDOG=240 timeout 300 $TR --master-port 29522 bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline --no-verify > $OUT/bench_full_noverify.json 2> $OUT/bench_full_noverify.err; echo "exit $?" >> $OUT/bench_full_noverify.err
tail -25 $OUT/bench_small.err | cut -c1-250; head -c 400 $OUT/bench_small.json; echo; tail -25 $OUT/bench_full_noverify.err | cut -c1-250; head -c 600 $OUT/bench_full_noverify.json
