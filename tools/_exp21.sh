mkdir -p gpurun_out
timeout 200 python tools/dram_stream_probe.py 2>&1 | tail -5 | tee gpurun_out/dram_scatter.jsonl
run() { echo "== $1"; env $1 timeout 200 python tools/train_step_time.py 4096 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step_median'], d['kernel_ms_per_step'])"; }
run NERF_B200_DBG_EMIT=1
run NERF_B200_DBG_EMIT=5
run NERF_B200_DBG_EMIT=13
run NERF_B200_DBG_EMIT=29
run NERF_B200_DBG_EMIT=9
