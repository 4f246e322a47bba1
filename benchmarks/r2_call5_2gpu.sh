#!/bin/bash
# Round 2, GPU call 5 (TWO B200s): bisect the world-2 step-level audio mismatch (with / without the audio side stream).
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/cp_step_check.py 4 > gpurun_out/r2c5_cp_step_check.log 2>&1
echo "rc=$?" >> gpurun_out/r2c5_cp_step_check.log
grep -E "rank 0|rc=|Error" gpurun_out/r2c5_cp_step_check.log | tail -20
exit 0
