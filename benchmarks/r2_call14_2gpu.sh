#!/bin/bash
# Round 2, GPU call 14 (TWO B200s): peer exchange with local chunks split from the remote pushes (same-device copies run on SMs)
# parity, then the per-segment timeline against the kernel-flag version and NCCL in one process.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
timeout 300 $TR --master-port 29561 tests/cp_check.py peer > gpurun_out/r2c14_cp_check_world2.log 2>&1
echo "rc=$?" >> gpurun_out/r2c14_cp_check_world2.log
timeout 300 $TR --master-port 29562 tests/cp_step_check.py 2 > gpurun_out/r2c14_cp_step_check.log 2>&1
echo "rc=$?" >> gpurun_out/r2c14_cp_step_check.log
timeout 300 $TR --master-port 29563 benchmarks/cp_layer_timeline.py --policies peer nccl peer > gpurun_out/r2c14_timeline_cp2.json 2> gpurun_out/r2c14_timeline_cp2.err
echo "rc=$?" >> gpurun_out/r2c14_timeline_cp2.err
grep -E "cp_check|rc=|Error|error" gpurun_out/r2c14_cp_check_world2.log | cut -c1-220
grep -E "rank 0|rc=|Error" gpurun_out/r2c14_cp_step_check.log | tail -4
cut -c1-220 gpurun_out/r2c14_timeline_cp2.json; tail -5 gpurun_out/r2c14_timeline_cp2.err
exit 0
