#!/bin/bash
# One gpurun call: GPU parity tests (all of them, no -x, measured errors printed), the work-item plan sweep, one bench line.
# Every stage has its own timeout so that a hang cannot eat the box.  Outputs land in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/smi.txt 2>&1
timeout -k 10 700 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/gputest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputest.log
timeout -k 10 240 python tools/chain_split_sweep.py tf32x3 tf32 > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
echo "sweep rc=$?" >> gpurun_out/sweep.err
timeout -k 10 200 python tools/chain_profile.py update tf32x3 > gpurun_out/chain_stamps_x3.txt 2>&1
timeout -k 10 500 python bench.py > gpurun_out/bench_flat.json 2> gpurun_out/bench_flat.err
echo "bench rc=$?" >> gpurun_out/bench_flat.err
tail -5 gpurun_out/gputest.log; cat gpurun_out/sweep.jsonl; tail -c 1500 gpurun_out/bench_flat.json
