mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -q -x --timeout 300 --timeout-method=thread 2>&1 | tail -3
run() { echo "== $1"; env $1 timeout 200 python tools/train_step_time.py 4096 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step_median'], d['kernel_ms_per_step'])"; }
run NERF_B200_DBG_X=0
run NERF_B200_DBG_WGRAD_ALPHA=0.75
run NERF_B200_DBG_WGRAD_ALPHA=0.5
run NERF_B200_DBG_WGRAD_ALPHA=0.25
run NERF_B200_DBG_WGRAD_ALPHA=0.0
run "NERF_B200_DBG_WGRAD=1 NERF_B200_DBG_NOAUX=7"
run "NERF_B200_DBG_WGRAD=1 NERF_B200_DBG_NOAUX=7 NERF_B200_DBG_WGRAD_ALPHA=0.5"
run "NERF_B200_DBG_WGRAD=1 NERF_B200_DBG_NOAUX=7 NERF_B200_DBG_WGRAD_ALPHA=0.0"
