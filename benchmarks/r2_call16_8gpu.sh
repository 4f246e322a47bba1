#!/bin/bash
# Round 2, GPU call 16 (EIGHT B200s): remote pushes of the peer exchange dealt over 1 / 2 / 4 / 7 streams (a copy-engine
# operation costs ~9 us on its stream); bench line only if more streams win by more than 0.5 %.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
timeout 300 $TR --master-port 29581 benchmarks/cp_layer_timeline.py --policies peer:1,3,1 peer@2:1,3,1 peer@4:1,3,1 peer@7:1,3,1 peer:1,3,1 > gpurun_out/r2c16_timeline_cp8.json 2> gpurun_out/r2c16_timeline_cp8.err
echo "rc=$?" >> gpurun_out/r2c16_timeline_cp8.err
BEST=$(python - <<'PY'
import json, sys
res = {}
for line in open("gpurun_out/r2c16_timeline_cp8.json"):
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print(d["policy"], d["exchange_used"], round(d["forward_ms_plain"], 2), file=sys.stderr)
    res.setdefault(d["policy"], []).append(d["forward_ms_plain"])
base = min(res.get("peer:1,3,1", [1e9]))
best, best_ms = 1, base
for k in (2, 4, 7):
    ms = min(res.get(f"peer@{k}:1,3,1", [1e9]))
    if ms < best_ms:
        best, best_ms = k, ms
print(best if best_ms < 0.995 * base else 1)
PY
)
echo "best push stream count: '$BEST'"
if [ "$BEST" != "1" ] && [ -n "$BEST" ]; then
  timeout 420 $TR --master-port 29582 bench.py --gpus 8 --steps 5 --warmup 3 --cp-push-streams $BEST > gpurun_out/r2c16_bench_cp8.json 2> gpurun_out/r2c16_bench_cp8.err
  echo "rc=$?" >> gpurun_out/r2c16_bench_cp8.err
  head -c 300 gpurun_out/r2c16_bench_cp8.json; echo; tail -3 gpurun_out/r2c16_bench_cp8.err
fi
tail -3 gpurun_out/r2c16_timeline_cp8.err
exit 0
