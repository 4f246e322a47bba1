#!/bin/bash
# Round 2, GPU call 7 (EIGHT B200s): context parallel at cp = 8 / 4 on NCCL -- the bench with its parity gate, BASELINE
# configs[3] (720p, cp = 8) and configs[2] (50-step schedule), oracle parity checks, and the per-segment timeline.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2c7_gpus.txt
timeout 420 $TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2c7_bench_cp8.json 2> gpurun_out/r2c7_bench_cp8.err; echo "rc=$?" >> gpurun_out/r2c7_bench_cp8.err
timeout 420 $TR --nproc-per-node 8 --master-port 29602 bench.py --gpus 8 --res 720p --steps 2 --warmup 2 > gpurun_out/r2c7_bench_720p_cp8.json 2> gpurun_out/r2c7_bench_720p_cp8.err; echo "rc=$?" >> gpurun_out/r2c7_bench_720p_cp8.err
timeout 240 $TR --nproc-per-node 8 --master-port 29603 tests/cp_check.py > gpurun_out/r2c7_cp_check_world8.log 2>&1; echo "rc=$?" >> gpurun_out/r2c7_cp_check_world8.log
timeout 420 $TR --nproc-per-node 4 --master-port 29604 bench.py --gpus 4 --steps 3 --warmup 2 > gpurun_out/r2c7_bench_cp4.json 2> gpurun_out/r2c7_bench_cp4.err; echo "rc=$?" >> gpurun_out/r2c7_bench_cp4.err
timeout 300 $TR --nproc-per-node 8 --master-port 29605 benchmarks/cp_layer_timeline.py > gpurun_out/r2c7_timeline_cp8.json 2> gpurun_out/r2c7_timeline_cp8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29606 benchmarks/cp_layer_timeline.py --single-stream > gpurun_out/r2c7_timeline_cp8_single_stream.json 2> gpurun_out/r2c7_timeline_cp8_single_stream.err
timeout 420 $TR --nproc-per-node 8 --master-port 29607 bench.py --gpus 8 --schedule 50 > gpurun_out/r2c7_schedule50_cp8.json 2> gpurun_out/r2c7_schedule50_cp8.err; echo "rc=$?" >> gpurun_out/r2c7_schedule50_cp8.err
head -c 400 gpurun_out/r2c7_bench_cp8.json; echo; head -c 400 gpurun_out/r2c7_bench_720p_cp8.json; echo; tail -2 gpurun_out/r2c7_cp_check_world8.log | cut -c1-300
exit 0
