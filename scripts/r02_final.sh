#!/bin/bash
# Round-2 measurement on the final tree: GPU tests, smoke, bench (+ reference arm), launch list, ncu captures.
set -u
mkdir -p gpurun_out
bash scripts/gpu_check.sh tests_all smoke bench bench_ref ncu_list
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pc_run_kernel -c 1 -f -o gpurun_out/single \
    python scripts/one_run.py > gpurun_out/ncu_single.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:pc_run_kernel -s 1 -c 1 -f -o gpurun_out/ens \
    python scripts/r02_ens_once.py 72 0 2 > gpurun_out/ncu_ens.log 2>&1
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:pc_run_kernel -c 1 -f -o gpurun_out/c4 \
    python scripts/r02_cfg_once.py C4 0 40000 > gpurun_out/ncu_c4.log 2>&1
PC_DEBUG=1 timeout -s KILL 200 python scripts/r02_cfg_once.py C2 > gpurun_out/diag_c2.log 2>&1
PC_DEBUG=1 timeout -s KILL 200 python scripts/r02_cfg_once.py C3 > gpurun_out/diag_c3.log 2>&1
PC_DEBUG=1 timeout -s KILL 200 python scripts/r02_cfg_once.py C4 > gpurun_out/diag_c4.log 2>&1
python scripts/r02_n8000_once.py > gpurun_out/n8000.log 2>&1
tail -n 3 gpurun_out/diag_c2.log gpurun_out/diag_c3.log gpurun_out/diag_c4.log gpurun_out/n8000.log gpurun_out/ncu_single.log
