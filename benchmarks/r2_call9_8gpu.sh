#!/bin/bash
# Round 2, GPU call 9 (EIGHT B200s): cp = 8 with per-head attention sets on two alternating streams -- bench + timeline.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29611 benchmarks/cp_layer_timeline.py > gpurun_out/r2c9_timeline_cp8.json 2> gpurun_out/r2c9_timeline_cp8.err
timeout 420 $TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2c9_bench_cp8.json 2> gpurun_out/r2c9_bench_cp8.err; echo "rc=$?" >> gpurun_out/r2c9_bench_cp8.err
head -c 300 gpurun_out/r2c9_bench_cp8.json; echo; head -c 300 gpurun_out/r2c9_timeline_cp8.json
exit 0
