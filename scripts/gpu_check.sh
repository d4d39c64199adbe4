#!/bin/bash
# Runs on the GPU box via gpurun: parity tests, smoke, a short bench, optional ncu passes.
# Usage: scripts/gpu_check.sh [tests] [smoke] [bench] [ncu_list] [ncu_full]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/host.txt
for stage in "$@"; do
case $stage in
tests)
  timeout -s KILL 300 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
  tail -25 gpurun_out/pytest_gpu.log ;;
tests_all)
  timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
  tail -60 gpurun_out/pytest_gpu.log ;;
smoke)
  timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log ;;
bench)
  timeout -s KILL 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
bench_ref)
  timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json ;;
ncu_list)
  PC_SYNC_DUMP=1 timeout -s KILL 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --ensemble 0 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; tail -3 gpurun_out/ncu_list.log ;;
ncu_full)
  # only the first launch (the device-resident run): a launch that hands dumps to the host cannot be replayed
  # by ncu (the replay passes run inside the intercepted launch call, the host never gets to acknowledge)
  timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:pc_run_kernel -c 1 -f -o gpurun_out/prof \
      python scripts/one_run.py > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log ;;
*) bash -c "$stage" ;;
esac
done
