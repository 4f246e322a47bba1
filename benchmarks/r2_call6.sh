#!/bin/bash
# Round 2, GPU call 6 (one B200): 96-key blocks / 4 score buffers (variants 93 / 94) against 128-key / 3 buffers (92).
set -x
mkdir -p gpurun_out
cd dualforce_b200/csrc
{
  for v in 93 94; do
    for shape in "1 128 128 1" "1 256 96 2" "1 256 512 2" "2 300 403 3" "1 403 403 12" "1 1000 512 4" "1 403 4400 12" "1 129 1300 1" "1 4400 4400 40"; do
      echo "== correctness variant $v shape $shape"
      timeout 60 ./selftest attn $shape 0 $v 4 | grep -E "OK|FAIL|error|rror"
    done
  done
  echo "== correctness variant 94 emu 6"; timeout 60 ./selftest attn 2 300 403 3 0 94 6 | grep -E "OK|FAIL|rror"
  echo "== correctness variant 94 emu 0"; timeout 60 ./selftest attn 2 300 403 3 0 94 0 | grep -E "OK|FAIL|rror"
  for v in 92 94; do
    for e in 0 4 6; do
      echo "== timing variant $v emu $e 43120^2 H40"; timeout 60 ./selftest attn 1 43120 43120 40 4 $v $e | grep timing
    done
    echo "== timing variant $v emu 4 43120^2 H5"; timeout 60 ./selftest attn 1 43120 43120 5 8 $v 4 | grep timing
    echo "== timing variant $v emu 4 403x43120 H12"; timeout 60 ./selftest attn 1 403 43120 12 20 $v 4 | grep timing
  done
  echo "== timing variant 93 emu 4 43120^2 H40"; timeout 60 ./selftest attn 1 43120 43120 40 4 93 4 | grep timing
  timeout 60 ./selftest attn 1 43120 43120 40 1 94 4 ../../gpurun_out/r2c6_trace_v94_full.bin | grep -E "timing|trace"
} > ../../gpurun_out/r2c6_attn.log 2>&1
cd ../..
grep -E "^==|timing|OK|FAIL|rror" gpurun_out/r2c6_attn.log | grep -v "^+"
exit 0
