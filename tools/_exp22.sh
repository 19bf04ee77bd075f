mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_trainer.py tests/test_gpu_backward.py tests/test_gpu_render.py -q -x --timeout 300 --timeout-method=thread 2>&1 | tail -3
run() { echo "== $1"; env $1 timeout 200 python tools/train_step_time.py 4096 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step_median'], d['kernel_ms_per_step'])"; }
run NERF_B200_DBG_X=0
run NERF_B200_DBG_EMIT=1
run NERF_B200_DBG_EMIT=5
timeout 300 python tools/bwd_overlap_sweep.py 0 2>&1 | tail -1
