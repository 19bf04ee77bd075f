mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -x -k "full_batch" --timeout 400 --timeout-method=thread 2>&1 | tail -12
cat gpurun_out/parity_train.json | python -c "import sys,json; d=json.load(sys.stdin); print(d.get('full_batch_4096'))"
timeout 300 python tools/bwd_overlap_sweep.py 0 0 0 2>&1 | tail -3
