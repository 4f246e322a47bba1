#!/bin/bash
# Runs the device self-test matrix; each case in its own process under a timeout.
# usage: run_selftest.sh [gemm|attn|ew|all] ; log lines go to stdout
cd "$(dirname "$0")"
what=${1:-all}
fail=0
run() { echo "== $*"; timeout 120 ./selftest "$@"; rc=$?; if [ $rc -ne 0 ]; then echo "   -> exit $rc"; fail=1; fi; }
if [ "$what" = gemm ] || [ "$what" = all ]; then
  for cg in 1 2; do
    run gemm $cg 0 128 256 64
    run gemm $cg 0 256 256 128
    run gemm $cg 0 300 520 136
    run gemm $cg 1 1000 1536 1536
    run gemm $cg 2 1000 1536 1536
    run gemm $cg 0 4400 5120 5120 10
    run gemm $cg 1 43120 13824 5120 5
    run gemm $cg 2 43120 5120 13824 5
    run gemm $cg 0 43120 15360 5120 5
    run gemm $cg 0 403 4608 1536 10
  done
fi
if [ "$what" = ew ] || [ "$what" = all ]; then
  run ln 77 5120 0 1
  run ln 403 1536 1 0
  run ln 50 5120 1 1
  run ln 33 1536 0 0
  run rr 77 5120 1
  run rr 77 5120 2
  run rr 403 1536 1
  run rr 403 1536 2
  run rr 64 5120 0
  run merge 4 403 12
fi
if [ "$what" = attn ] || [ "$what" = all ]; then
  run attn 1 128 128 1
  run attn 1 256 256 2
  run attn 1 256 512 2
  run attn 2 300 403 3
  run attn 1 403 403 12
  run attn 1 1000 512 4
  run attn 1 403 4400 12
  run attn 1 4400 4400 40 5
  run attn 1 43120 43120 40 3
fi
exit $fail
