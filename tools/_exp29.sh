mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_trainer.py tests/test_gpu_dropin.py tests/test_gpu_parity.py -q -x --timeout 400 --timeout-method=thread 2>&1 | tail -3
timeout 200 python tools/train_step_time.py 4096 5 2>&1 | tail -1 | cut -c1-420
timeout 600 python tools/bwd_overlap_sweep.py default default 2>&1 | tail -2
