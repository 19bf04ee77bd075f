#!/bin/bash
# GPU measurement pass (one B200): full -m gpu test suite, the bench line, ncu launch lists of the forward step and of the fused
# training step, one `ncu --set full` capture of the forward kernel and of the three training kernels, the HBM probes.
# Outputs -> gpurun_out/ (copied to profiles/r02_* by hand).   NCU=0 skips the profiler passes.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${STEPS:-30} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench.json") if l.startswith("{")][-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["e2e"]["eager_render"]["value"], d["roofline"]["frac"], d["gpu_launches"])
t=d["train"]; print("train", t["value"], t["ms_per_step"], t["e2e"]["value"], t["roofline"]["frac"], t["roofline"]["kernel_ms_per_step"], t["roofline"]["hbm"]["frac"], t["gpu_launches_per_step"])
print(d["torch_gpu"].get("speedup")); print(d["clocks"])
PY
timeout 120 python tools/hbm_probe.py > gpurun_out/hbm_probe.json 2>&1
timeout 200 python tools/dram_stream_probe.py > gpurun_out/dram_stream_probe.jsonl 2>&1
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_forward.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-train --no-torch-gpu > gpurun_out/ncu_launch_fwd.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fused_train_step.csv python tools/fused_train_steps.py 3 > gpurun_out/ncu_launch_train.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:march_tc -s 4 -c 2 -f -o gpurun_out/prof_fwd_r02 python bench.py --steps 1 --warmup 3 --no-cpu --no-train --no-torch-gpu > gpurun_out/ncu_full_fwd.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:march_tc2_kernel|dgrad_tc2_kernel|wgrad_tc_kernel" -s 18 -c 6 -f -o gpurun_out/prof_train_r02 python tools/fused_train_steps.py 4 > gpurun_out/ncu_full_train.log 2>&1
  ls -la gpurun_out | tail -20
fi
