mkdir -p gpurun_out
run() { echo "== $1"; env $1 timeout 200 python tools/train_step_time.py 4096 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step_median'], d['kernel_ms_per_step'])"; }
run NERF_B200_DBG_X=0
run NERF_B200_DBG_EMIT=1
run NERF_B200_DBG_EMIT=2
run "NERF_B200_DBG_EMIT=2 NERF_B200_DBG_WGRAD=2"
run NERF_B200_DBG_WGRAD=2
