#!/bin/bash
# Round 2, GPU call 17 (one B200): final state of the tree -- GPU suite, the default bench line, the CFG pair as one
# batched forward (A/B), and the ncu launch list of one bench step with the round-2 kernels.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2c17_pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/r2c17_pytest_gpu.log
tail -3 gpurun_out/r2c17_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err
echo "rc=$?" >> gpurun_out/r2c17_bench.err
head -c 300 gpurun_out/r2c17_bench.json; echo; tail -2 gpurun_out/r2c17_bench.err
timeout 400 python bench.py --steps 3 --warmup 3 --cfg-merge --no-cpu-baseline > gpurun_out/r2c17_bench_cfg_merge.json 2> gpurun_out/r2c17_bench_cfg_merge.err
echo "rc=$?" >> gpurun_out/r2c17_bench_cfg_merge.err
head -c 300 gpurun_out/r2c17_bench_cfg_merge.json; echo; tail -2 gpurun_out/r2c17_bench_cfg_merge.err
# launch list of ONE step (the timed one: 3034 launches of our kernels per step), after the same command ran without ncu
K='regex:attn_pair_kernel|attn_fwd_kernel|gemm_bf16_kernel|layernorm_kernel|rmsnorm_rope_kernel|add_to_f32_kernel|lse_merge_kernel|patchify_kernel|unpatchify_kernel|sinusoidal_kernel|gemv_f32_kernel|cfg_euler_kernel'
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2c17_plain_bench.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 3034 -c 3034 --csv \
  --log-file gpurun_out/r2c17_launches_bench_step.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2c17_ncu_bench.log 2>&1
echo "rc=$?" >> gpurun_out/r2c17_ncu_bench.log
tail -2 gpurun_out/r2c17_ncu_bench.log; wc -l gpurun_out/r2c17_launches_bench_step.csv
exit 0
