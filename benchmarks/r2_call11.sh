#!/bin/bash
# Round 2, GPU call 11 (one B200): the whole GPU suite on the current tree (first hardware run of csrc/peer.cu through the
# world-1 context-parallel tests), smoke(), then the default bench line and the reference arm as the driver runs them.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2c11_pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/r2c11_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c11_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/r2c11_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c11_bench.json 2> gpurun_out/r2c11_bench.err
echo "rc=$?" >> gpurun_out/r2c11_bench.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2c11_bench_reference.json 2> gpurun_out/r2c11_bench_reference.err
echo "rc=$?" >> gpurun_out/r2c11_bench_reference.err
tail -4 gpurun_out/r2c11_pytest_gpu.log; tail -2 gpurun_out/r2c11_smoke.log
head -c 600 gpurun_out/r2c11_bench.json; echo; tail -2 gpurun_out/r2c11_bench.err
head -c 400 gpurun_out/r2c11_bench_reference.json; echo; tail -2 gpurun_out/r2c11_bench_reference.err
exit 0
