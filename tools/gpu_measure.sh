#!/bin/bash
# GPU measurement pass: full -m gpu test suite, bench line, ncu launch list and one full capture
# of the dominant kernel.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps ${STEPS:-30} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:march_tc -s 4 -c 2 -f -o gpurun_out/prof_march python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
