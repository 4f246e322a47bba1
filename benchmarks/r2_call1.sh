#!/bin/bash
# Round 2, GPU call 1 (one B200): the same-box library bar, the hardware verdict on the attention variants written at
# the end of round 1, CG=1 vs CG=2 GEMM, the bridge microbench and ncu captures of the shipped kernels.
#   gpurun --timeout 1500 -- 'bash benchmarks/r2_call1.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
  --format=csv -lms 500 > gpurun_out/r2c1_clocks.csv &
SMI=$!

timeout 600 python benchmarks/kernels_vs_libs.py --iters 5 > gpurun_out/r2c1_kernels_vs_libs.jsonl 2> gpurun_out/r2c1_kernels_vs_libs.err

cd dualforce_b200/csrc
{
  for v in v3 v7 v8; do
    for b in 0 1; do
      echo "== variant $v bounded $b"
      MOVA_ATTN_VARIANT=$v MOVA_ATTN_BOUNDED=$b timeout 60 ./selftest attn 1 4400 4400 40 5
      MOVA_ATTN_VARIANT=$v MOVA_ATTN_BOUNDED=$b timeout 60 ./selftest attn 1 43120 43120 40 3
    done
  done
  for e in 0 2 6 8; do
    echo "== v3 EMU $e"
    MOVA_ATTN_EMU=$e timeout 60 ./selftest attn 1 43120 43120 40 3
  done
} > ../../gpurun_out/r2c1_attn_variants.log 2>&1
{
  for cg in 1 2; do
    for shape in "0 43120 15360 5120" "1 43120 13824 5120" "2 43120 5120 13824" "0 43120 5120 5120" "0 5390 15360 5120"; do
      echo "== gemm cg$cg $shape"
      timeout 60 ./selftest gemm $cg $shape 5
    done
  done
} > ../../gpurun_out/r2c1_gemm.log 2>&1
cd ../..

timeout 600 python benchmarks/bridge_microbench.py --quick --iters 10 > gpurun_out/r2c1_bridge_quick.jsonl 2> gpurun_out/r2c1_bridge_quick.err

# ncu: each capture directly after the same command exited 0 without ncu
cd dualforce_b200/csrc
NCU="ncu --set full --clock-control none --import-source on"
./selftest attn 1 43120 43120 40 2 > ../../gpurun_out/r2c1_plain_attn.log 2>&1 &&
  $NCU -k regex:attn_fwd -s 2 -c 1 -o ../../gpurun_out/r2c1_attn_v3 ./selftest attn 1 43120 43120 40 2 > ../../gpurun_out/r2c1_ncu_attn.log 2>&1
./selftest gemm 1 0 43120 15360 5120 3 > ../../gpurun_out/r2c1_plain_gemm1.log 2>&1 &&
  $NCU -k regex:gemm_bf16 -s 2 -c 1 -o ../../gpurun_out/r2c1_gemm_cg1 ./selftest gemm 1 0 43120 15360 5120 3 > ../../gpurun_out/r2c1_ncu_gemm1.log 2>&1
./selftest gemm 2 0 43120 15360 5120 3 > ../../gpurun_out/r2c1_plain_gemm2.log 2>&1 &&
  $NCU -k regex:gemm_bf16 -s 2 -c 1 -o ../../gpurun_out/r2c1_gemm_cg2 ./selftest gemm 2 0 43120 15360 5120 3 > ../../gpurun_out/r2c1_ncu_gemm2.log 2>&1
./selftest ln 43120 5120 0 1 > ../../gpurun_out/r2c1_plain_ln.log 2>&1 &&
  $NCU -k regex:layernorm -c 1 -o ../../gpurun_out/r2c1_layernorm ./selftest ln 43120 5120 0 1 > ../../gpurun_out/r2c1_ncu_ln.log 2>&1
./selftest rr 43120 5120 1 > ../../gpurun_out/r2c1_plain_rr.log 2>&1 &&
  $NCU -k regex:rmsnorm_rope -c 1 -o ../../gpurun_out/r2c1_rmsnorm_rope ./selftest rr 43120 5120 1 > ../../gpurun_out/r2c1_ncu_rr.log 2>&1
cd ../..
kill $SMI
ls -la gpurun_out
exit 0
