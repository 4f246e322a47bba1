#!/bin/bash
# Quick A/B of an attention schedule variant against the shipped one (correctness on ragged shapes + timing).
# usage: run_variant_check.sh v6
cd "$(dirname "$0")"
V=${1:-v6}
for a in "1 256 512 2" "2 300 403 3"; do
  MOVA_ATTN_VARIANT=$V timeout 30 ./selftest attn $a | grep -v "^  lse" || echo "   -> FAILED: $a"
done
echo "== $V EMU 4"
MOVA_ATTN_VARIANT=$V MOVA_ATTN_EMU=4 timeout 30 ./selftest attn 1 43120 43120 40 2
