#!/bin/sh
# compute-sanitizer runs over the mbarrier / bulk-copy / tcgen05 heavy kernels (K1, chain2, grouped weight gradient, fused history encoder),
# driven by the golden-vector tests (small shapes: the tools slow kernels down 10-100 x).  Run on a GPU box:  sh tools/sanitize.sh
# Logs go to gpurun_out/sanitize_*.log; the summaries are copied to profiles/r2_sanitizer.md by hand.
mkdir -p gpurun_out
T="timeout -s KILL 600"
SAN=/usr/local/cuda/bin/compute-sanitizer
K_ENV='tests/test_gpu_env.py::test_env_step_matches_reference_golden'
K_PPO='tests/test_gpu_ppo.py::test_ppo_update_matches_reference_golden tests/test_gpu_ppo.py::test_policy_act_matches_reference_golden'
for tool in memcheck racecheck synccheck; do
  for what in env ppo; do
    if [ $what = env ]; then sel="$K_ENV"; else sel="$K_PPO"; fi
    $T $SAN --tool $tool --print-limit 20 python -m pytest $sel -q -m gpu -x -p no:cacheprovider > gpurun_out/sanitize_${tool}_${what}.log 2>&1
    echo "== $tool $what: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_${tool}_${what}.log | tail -4
  done
done
