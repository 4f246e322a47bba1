#!/bin/bash
# Round 2, GPU call 3 (one B200): round-2 attention kernel after the softmax-phase tweaks (row sum inside the exp loop,
# token wait under the TMEM load), exp2-emulation share, short-KV crossover against the round-1 schedule, a trace and
# an ncu capture of the CTA-pair kernel at the full 360p shape, and a short bench with the CTA-pair GEMM as default.
set -x
mkdir -p gpurun_out
cd dualforce_b200/csrc
{
  for shape in "1 256 512 2" "2 300 403 3" "1 129 1300 1" "1 4400 4400 40"; do
    for v in 91 92; do
      for e in 4 6; do
        echo "== correctness variant $v emu $e shape $shape"
        timeout 60 ./selftest attn $shape 0 $v $e | grep -E "OK|FAIL|error"
      done
    done
  done
  for e in 0 4 6 8; do
    echo "== timing variant 92 emu $e 43120^2 H40"
    timeout 60 ./selftest attn 1 43120 43120 40 4 92 $e | grep timing
  done
  echo "== timing variant 91 emu 4 43120^2 H40"; timeout 60 ./selftest attn 1 43120 43120 40 4 91 4 | grep timing
  echo "== timing variant 92 emu 4 43120^2 H5";  timeout 60 ./selftest attn 1 43120 43120 5 8 92 4 | grep timing
  for skv in 512 1024 2048 4096 8192; do
    for v in 3 92; do
      echo "== timing variant $v 43120 x $skv H40"
      timeout 60 ./selftest attn 1 43120 $skv 40 20 $v 4 | grep timing
    done
  done
  echo "== default dispatch 43120x512 H40"; timeout 60 ./selftest attn 1 43120 512 40 20 | grep timing
  echo "== default dispatch 403x43120 H12"; timeout 60 ./selftest attn 1 403 43120 12 20 | grep timing
  timeout 60 ./selftest attn 1 43120 43120 40 1 92 4 ../../gpurun_out/r2c3_trace_v92_full.bin | grep -E "timing|trace"
} > ../../gpurun_out/r2c3_attn.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
./selftest attn 1 43120 43120 40 2 92 4 > ../../gpurun_out/r2c3_plain_attn.log 2>&1 &&
  $NCU -k regex:attn_pair -s 2 -c 1 -o ../../gpurun_out/r2c3_attn_v92 ./selftest attn 1 43120 43120 40 2 92 4 > ../../gpurun_out/r2c3_ncu_attn.log 2>&1
./selftest gemm 2 0 43120 15360 5120 3 > ../../gpurun_out/r2c3_plain_gemm2.log 2>&1 &&
  $NCU -k regex:gemm_bf16 -s 2 -c 1 -o ../../gpurun_out/r2c3_gemm_cg2 ./selftest gemm 2 0 43120 15360 5120 3 > ../../gpurun_out/r2c3_ncu_gemm2.log 2>&1
cd ../..
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
cat gpurun_out/r2c3_attn.log | tail -60
exit 0
