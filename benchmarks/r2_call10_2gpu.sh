#!/bin/bash
# Round 2, GPU call 10 (TWO B200s): first hardware run of the peer-memory Ulysses exchange (CUDA IPC windows, copy-engine
# pushes + flag words) -- parity against the oracle and against the NCCL exchange, then timeline + bench at cp = 2.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
timeout 300 $TR --master-port 29531 tests/cp_check.py > gpurun_out/r2c10_cp_check_world2.log 2>&1
echo "rc=$?" >> gpurun_out/r2c10_cp_check_world2.log
timeout 300 $TR --master-port 29532 tests/cp_step_check.py 2 > gpurun_out/r2c10_cp_step_check.log 2>&1
echo "rc=$?" >> gpurun_out/r2c10_cp_step_check.log
timeout 300 python -m pytest tests/test_gpu_cp.py tests/test_gpu_step.py -q -m gpu -k "cp or context or world" > gpurun_out/r2c10_pytest_cp.log 2>&1
echo "rc=$?" >> gpurun_out/r2c10_pytest_cp.log
timeout 300 $TR --master-port 29533 benchmarks/cp_layer_timeline.py --policies peer nccl > gpurun_out/r2c10_timeline_cp2.json 2> gpurun_out/r2c10_timeline_cp2.err
echo "rc=$?" >> gpurun_out/r2c10_timeline_cp2.err
timeout 400 $TR --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2c10_bench_cp2.json 2> gpurun_out/r2c10_bench_cp2.err
echo "rc=$?" >> gpurun_out/r2c10_bench_cp2.err
grep -E "cp_check|rc=|Error|error" gpurun_out/r2c10_cp_check_world2.log | cut -c1-220
grep -E "rank 0|rc=|Error" gpurun_out/r2c10_cp_step_check.log | tail -8
tail -3 gpurun_out/r2c10_pytest_cp.log
cut -c1-200 gpurun_out/r2c10_timeline_cp2.json; tail -3 gpurun_out/r2c10_timeline_cp2.err
head -c 400 gpurun_out/r2c10_bench_cp2.json; tail -3 gpurun_out/r2c10_bench_cp2.err
exit 0
