#!/bin/bash
# Round 2, GPU call 4 (TWO B200s): context parallel on NCCL at world 2 -- parity against the CPU oracle (tests/cp_check.py,
# incl. an odd head count per rank), the world-2 GPU tests, and bench.py at N = 2 with its parity gate.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2c4_gpus.txt
( cd dualforce_b200/csrc && for shape in "2 0 300 520 136" "2 2 1000 1536 1536" "2 0 43120 15360 5120 5" "2 1 43120 13824 5120 5" "2 2 43120 5120 13824 5" "1 0 43120 15360 5120 5"; do echo "== gemm $shape"; timeout 60 ./selftest gemm $shape; done ) > gpurun_out/r2c4_gemm_l2hints.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/cp_check.py > gpurun_out/r2c4_cp_check_world2.log 2>&1
echo "cp_check rc=$?" >> gpurun_out/r2c4_cp_check_world2.log
timeout 600 python -m pytest tests -m gpu -q -rs -k "cp or world2 or parallel" 2>&1 | tail -15 > gpurun_out/r2c4_pytest_cp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2c4_bench_cp2.json 2> gpurun_out/r2c4_bench_cp2.err
echo "bench rc=$?" >> gpurun_out/r2c4_bench_cp2.err
tail -3 gpurun_out/r2c4_cp_check_world2.log; tail -5 gpurun_out/r2c4_pytest_cp.log; head -c 600 gpurun_out/r2c4_bench_cp2.json
exit 0
