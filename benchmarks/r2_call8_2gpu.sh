#!/bin/bash
# Round 2, GPU call 8 (TWO B200s): the per-head attention sets on alternating streams (odd heads per rank) vs the oracle.
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/cp_check.py > gpurun_out/r2c8_cp_check_world2.log 2>&1
echo "rc=$?" >> gpurun_out/r2c8_cp_check_world2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tests/cp_step_check.py 2 > gpurun_out/r2c8_cp_step_check.log 2>&1
echo "rc=$?" >> gpurun_out/r2c8_cp_step_check.log
grep -E "cp_check|rc=" gpurun_out/r2c8_cp_check_world2.log | cut -c1-200; grep -E "rank 0|rc=" gpurun_out/r2c8_cp_step_check.log
exit 0
