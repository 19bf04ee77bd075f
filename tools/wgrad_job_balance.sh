#!/bin/bash
# Per-job finish times of wgrad_tc_kernel and wait-time split of the dgrad epilogue warps on the C2 training step, from the
# experiments build (NERF_B200_EXPERIMENTS=1 python -c "import __graft_entry__ as g; g.build()" builds libnerf_b200_exp.so).
mkdir -p gpurun_out; rm -f gpurun_out/wgrad_prof.txt gpurun_out/dgrad_prof.txt
NERF_B200_EXPERIMENTS=1 NERF_B200_DBG_WGRAD_PROF=gpurun_out/wgrad_prof.txt NERF_B200_DBG_DGRAD_PROF=gpurun_out/dgrad_prof.txt timeout 300 python tools/train_step_time.py 4096 3 2>&1 | tail -1 | cut -c1-400
tail -26 gpurun_out/wgrad_prof.txt; tail -2 gpurun_out/dgrad_prof.txt
