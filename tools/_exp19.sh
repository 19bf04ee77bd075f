mkdir -p gpurun_out; rm -f gpurun_out/wgrad_prof*.txt
timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_backward.py -q -x --timeout 300 --timeout-method=thread 2>&1 | tail -2
NERF_B200_DBG_WGRAD_PROF=gpurun_out/wgrad_prof.txt timeout 200 python tools/train_step_time.py 4096 3 2>&1 | tail -1 | cut -c1-400
tail -26 gpurun_out/wgrad_prof.txt
timeout 200 python tools/train_step_time.py 4096 5 2>&1 | tail -1 | cut -c1-400
