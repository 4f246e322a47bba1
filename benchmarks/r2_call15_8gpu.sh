#!/bin/bash
# Round 2, GPU call 15 (EIGHT B200s): peer-memory Ulysses exchange at cp = 8 -- parity record, attention-set policies and
# NCCL in one timeline process, then the bench line with the best set policy.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
timeout 200 $TR --master-port 29571 tests/cp_check.py peer > gpurun_out/r2c15_cp_check_world8.log 2>&1
echo "rc=$?" >> gpurun_out/r2c15_cp_check_world8.log
grep -E "cp_check|rc=" gpurun_out/r2c15_cp_check_world8.log | cut -c1-160
timeout 300 $TR --master-port 29572 benchmarks/cp_layer_timeline.py --policies peer:1,3,1 peer:1,4 peer:1,1,1,1,1 peer:2,3 peer:5 nccl:1,1,1,1,1 > gpurun_out/r2c15_timeline_cp8.json 2> gpurun_out/r2c15_timeline_cp8.err
echo "rc=$?" >> gpurun_out/r2c15_timeline_cp8.err
BEST=$(python - <<'PY'
import json
best = None
for line in open("gpurun_out/r2c15_timeline_cp8.json"):
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print(d["policy"], d["exchange_used"], round(d["forward_ms_plain"], 2), file=__import__("sys").stderr)
    if d["policy"].startswith("peer:") and "peer/memops" in d["exchange_used"]:
        if best is None or d["forward_ms_plain"] < best[0]:
            best = (d["forward_ms_plain"], d["policy"].split(":", 1)[1])
print(best[1] if best else "")
PY
)
echo "best peer set policy: '$BEST'"
if [ -n "$BEST" ]; then SETS="--cp-sets $BEST"; else SETS="--cp-exchange nccl"; fi
timeout 420 $TR --master-port 29573 bench.py --gpus 8 --steps 5 --warmup 3 $SETS > gpurun_out/r2c15_bench_cp8.json 2> gpurun_out/r2c15_bench_cp8.err
echo "rc=$?" >> gpurun_out/r2c15_bench_cp8.err
head -c 300 gpurun_out/r2c15_bench_cp8.json; echo; tail -3 gpurun_out/r2c15_bench_cp8.err; tail -3 gpurun_out/r2c15_timeline_cp8.err
exit 0
