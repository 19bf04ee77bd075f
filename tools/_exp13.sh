mkdir -p gpurun_out
python tools/hbm_probe.py | tee gpurun_out/hbm_probe.json
run() { echo "== $1"; env $1 timeout 200 python tools/train_step_time.py 4096 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step_median'], d['kernel_ms_per_step'])"; }
run NERF_B200_DBG_X=0
run NERF_B200_DBG_NOAUX=3
run NERF_B200_DBG_NOAUX=7
run NERF_B200_DBG_BWD_CTAS=148,128
run NERF_B200_DBG_BWD_CTAS=148,112
run NERF_B200_DBG_BWD_CTAS=148,96
run NERF_B200_DBG_BWD_CTAS=148,74
run NERF_B200_DBG_BWD_CTAS=128,148
run NERF_B200_DBG_BWD_CTAS=112,148
run NERF_B200_DBG_BWD_CTAS=96,148
run NERF_B200_DBG_BWD_CTAS=74,148
