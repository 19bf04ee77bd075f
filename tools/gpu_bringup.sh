#!/bin/bash
# Run the GPU test groups one by one under their own timeouts so that a hang in one group
# (e.g. a barrier deadlock) does not hide the results of the others.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-240} python -m pytest "$@" -q -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
run selftest tests/test_gpu_units.py -k selftest
run small tests/test_gpu_units.py -k "embed or raw2outputs or sample_pdf or coarse_z"
run net_fp32 tests/test_gpu_units.py -k "run_network and fp32"
run net_tc tests/test_gpu_units.py -k "run_network and not fp32"
run render_fp32 tests/test_gpu_render.py -k "fp32"
run render_tc tests/test_gpu_render.py -k "not fp32"
