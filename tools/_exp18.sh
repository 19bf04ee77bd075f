mkdir -p gpurun_out; rm -f gpurun_out/wgrad_prof*.txt
NERF_B200_DBG_WGRAD_PROF=gpurun_out/wgrad_prof.txt timeout 200 python tools/train_step_time.py 4096 3 2>&1 | tail -1 | cut -c1-300
tail -28 gpurun_out/wgrad_prof.txt
NERF_B200_DBG_NOAUX=7 NERF_B200_DBG_WGRAD_PROF=gpurun_out/wgrad_prof_noaux.txt timeout 200 python tools/train_step_time.py 4096 3 2>&1 | tail -1 | cut -c1-300
tail -14 gpurun_out/wgrad_prof_noaux.txt
