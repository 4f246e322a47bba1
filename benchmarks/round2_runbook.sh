#!/bin/bash
# First GPU session of the next round: hardware verdicts on everything written after round 1's GPU budget was spent,
# cheapest first.  Each block is one `gpurun -- '<block>'` call; outputs land in gpurun_out/.
# (Nothing here is needed by the tests or the bench; it is the order of experiments, kept next to the code it drives.)
set -x
mkdir -p gpurun_out

# 1. the whole GPU suite with xfail / xpass reasons (step.cu kernels, inference_single_step, 720p sizes, variants)
python -m pytest tests -m gpu -q -rxX 2>&1 | tail -60 > gpurun_out/r2_pytest_gpu.log

# 2. attention schedules x bounded softmax at the 360p shape (each process latches its variant)
cd dualforce_b200/csrc
for v in v3 v7 v8; do
  for b in 0 1; do
    echo "== variant $v bounded $b"
    MOVA_ATTN_VARIANT=$v MOVA_ATTN_BOUNDED=$b timeout 60 ./selftest attn 1 4400 4400 40 5
    MOVA_ATTN_VARIANT=$v MOVA_ATTN_BOUNDED=$b timeout 60 ./selftest attn 1 43120 43120 40 3
  done
done > ../../gpurun_out/r2_attn_variants.log 2>&1
# polynomial-exp2 share on the best schedule (variants: MOVA_ATTN_EMU = 0 / 4 / 8; v3 also 2 / 6)
for e in 0 8; do
  echo "== v8 bounded EMU $e"
  MOVA_ATTN_VARIANT=v8 MOVA_ATTN_BOUNDED=1 MOVA_ATTN_EMU=$e timeout 60 ./selftest attn 1 43120 43120 40 3
done >> ../../gpurun_out/r2_attn_variants.log 2>&1
cd ../..

# 3. the bench with the winner (example: v8 + bounded), then the default for an A/B on the same box
MOVA_ATTN_VARIANT=v8 MOVA_ATTN_BOUNDED=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_v8b.json 2> gpurun_out/r2_bench_v8b.err
python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err

MOVA_V2A_SPLITS=5 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_v2a_split5.json 2> gpurun_out/r2_bench_v2a_split5.err

# 4. (gpurun --gpus 8) context parallel: default vs one-head groups with overlapped attention launches
#   torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 bench.py --gpus 8 --steps 3 --warmup 3
#   MOVA_CP_HEAD_GROUPS=5 MOVA_CP_ATTN_STREAMS=2 torchrun ... (same)
#   torchrun ... bench.py --gpus 8 --res 720p --steps 1 --warmup 3            # BASELINE configs[3]

# 5. BASELINE configs[4]
python benchmarks/bridge_microbench.py --quick > gpurun_out/r2_bridge_microbench.jsonl 2>&1

# 6. ncu evidence still missing from profiles/: the GEMM and the memory-bound kernels (tensor-pipe % / DRAM GB/s),
#    one capture each after the plain command has exited 0 (B200_PROFILING.md recipe)
cd dualforce_b200/csrc
./selftest gemm 1 0 43120 15360 5120 3 && ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 \
  -o ../../gpurun_out/r2_gemm_qkv ./selftest gemm 1 0 43120 15360 5120 3 > ../../gpurun_out/r2_ncu_gemm.log 2>&1
./selftest ln 43120 5120 0 1 && ncu --set full --clock-control none --import-source on -k regex:layernorm -c 1 \
  -o ../../gpurun_out/r2_layernorm ./selftest ln 43120 5120 0 1 > ../../gpurun_out/r2_ncu_ln.log 2>&1
./selftest rr 43120 5120 1 && ncu --set full --clock-control none --import-source on -k regex:rmsnorm_rope -c 1 \
  -o ../../gpurun_out/r2_rmsnorm_rope ./selftest rr 43120 5120 1 > ../../gpurun_out/r2_ncu_rr.log 2>&1
cd ../..
