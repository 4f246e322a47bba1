#!/bin/bash
# Round 2, GPU call 13 (TWO B200s): which copies of the peer exchange overlap a running attention kernel?
set -x
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29551 benchmarks/ce_overlap_probe.py > gpurun_out/r2c13_ce_overlap_probe.json 2> gpurun_out/r2c13_ce_overlap_probe.err
echo "rc=$?" >> gpurun_out/r2c13_ce_overlap_probe.err
tail -c 3000 gpurun_out/r2c13_ce_overlap_probe.json; tail -5 gpurun_out/r2c13_ce_overlap_probe.err
exit 0
