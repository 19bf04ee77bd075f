mkdir -p gpurun_out; rm -f gpurun_out/dgrad_prof*.txt
NERF_B200_DBG_DGRAD_PROF=gpurun_out/dgrad_prof.txt timeout 200 python tools/train_step_time.py 4096 3 2>&1 | tail -1 | cut -c1-100
tail -4 gpurun_out/dgrad_prof.txt
NERF_B200_DBG_EMIT=1 NERF_B200_DBG_DGRAD_PROF=gpurun_out/dgrad_prof_nostore.txt timeout 200 python tools/train_step_time.py 4096 3 2>&1 | tail -1 | cut -c1-100
tail -4 gpurun_out/dgrad_prof_nostore.txt
