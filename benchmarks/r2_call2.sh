#!/bin/bash
# Round 2, GPU call 2 (one B200): first hardware run of the round-2 attention kernel (attn_pair.cu: variants 91 / 92)
# and of the repaired CTA-pair GEMM; if they are correct, the GPU test-suite and a short bench.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap \
  --format=csv -lms 500 > gpurun_out/r2c2_clocks.csv &
SMI=$!
cd dualforce_b200/csrc
ok=1
{
  for v in 91 92; do
    for shape in "1 128 128 1" "1 256 256 2" "1 256 512 2" "2 300 403 3" "1 403 403 12" "1 1000 512 4" "1 403 4400 12" "1 129 1300 1" "1 4400 4400 40"; do
      echo "== attn variant $v shape $shape"
      timeout 60 ./selftest attn $shape 0 $v 4 || { echo "   -> FAILED rc=$?"; ok=0; }
    done
    for e in 0 8; do
      echo "== attn variant $v emu $e shape 2 300 403 3"
      timeout 60 ./selftest attn 2 300 403 3 0 $v $e || { echo "   -> FAILED rc=$?"; ok=0; }
    done
  done
} > ../../gpurun_out/r2c2_attn_correct.log 2>&1
{
  for v in 3 91 92; do
    for e in 0 4 8; do
      echo "== timing variant $v emu $e  43120^2 H40"
      timeout 60 ./selftest attn 1 43120 43120 40 3 $v $e
    done
    echo "== timing variant $v  43120^2 H5"
    timeout 60 ./selftest attn 1 43120 43120 5 5 $v 4
    echo "== timing variant $v  43120x512 H40"
    timeout 60 ./selftest attn 1 43120 512 40 20 $v 4
    echo "== timing variant $v  403x43120 H12"
    timeout 60 ./selftest attn 1 403 43120 12 20 $v 4
  done
  timeout 60 ./selftest attn 1 4400 4400 2 1 91 4 ../../gpurun_out/r2c2_trace_v91.bin
  timeout 60 ./selftest attn 1 4400 4400 2 1 92 4 ../../gpurun_out/r2c2_trace_v92.bin
} > ../../gpurun_out/r2c2_attn_timing.log 2>&1
{
  for cg in 1 2; do
    for shape in "0 128 256 64" "0 300 520 136" "1 1000 1536 1536" "2 1000 1536 1536" "0 4400 5120 5120 10"; do
      echo "== gemm cg$cg $shape"
      timeout 60 ./selftest gemm $cg $shape || echo "   -> FAILED rc=$?"
    done
    for shape in "0 43120 15360 5120" "1 43120 13824 5120" "2 43120 5120 13824" "0 43120 5120 5120" "0 5390 15360 5120" "0 5390 5120 5120"; do
      echo "== gemm cg$cg $shape"
      timeout 60 ./selftest gemm $cg $shape 5
    done
  done
} > ../../gpurun_out/r2c2_gemm.log 2>&1
cd ../..
timeout 300 python benchmarks/kernels_vs_libs.py --iters 5 --only gemm > gpurun_out/r2c2_gemm_vs_cublas.jsonl 2> gpurun_out/r2c2_gemm_vs_cublas.err
if [ $ok = 1 ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -rxXs 2>&1 | tail -40 > gpurun_out/r2c2_pytest_gpu.log
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
else
  echo "attention selftests failed: pytest / bench skipped" > gpurun_out/r2c2_pytest_gpu.log
fi
kill $SMI
tail -5 gpurun_out/r2c2_attn_correct.log gpurun_out/r2c2_pytest_gpu.log
exit 0
