#!/bin/sh
# Installs the UNMODIFIED reference rsl_rl (PPO, ActorCritic, RolloutStorage: the update half of the hot path) into baseline/_ref
# (git-ignored; it travels to the GPU box with the snapshot).  Run in the authoring container, which has /root/reference.
# The env half (legged_gym's WidowGo1) imports the closed-source isaacgym package at module import and cannot be installed or run
# unmodified; tests/golden/ref_harness.py drives it through a fake isaacgym for the golden vectors only.
set -e
cd "$(dirname "$0")/.."
rm -rf /tmp/rsl_rl_src baseline/_ref
cp -r /root/reference/rsl_rl /tmp/rsl_rl_src            # /root/reference is read-only, the build writes an egg-info
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/rsl_rl_src
diff -rq -x __pycache__ /root/reference/rsl_rl/rsl_rl baseline/_ref/rsl_rl && echo "baseline/_ref/rsl_rl is identical to the reference tree"
