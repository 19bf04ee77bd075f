mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_trainer.py tests/test_gpu_dropin.py -q -x --timeout 400 --timeout-method=thread 2>&1 | tail -3
timeout 600 python tools/bwd_overlap_sweep.py default 0 default 0 default 0 2>&1 | tail -6
