#!/bin/bash
# Round-2 diagnostics on the GPU box: phase counters of every configuration, ncu captures (single, ensemble, C4).
set -u
mkdir -p gpurun_out
PC_DEBUG=1 timeout -s KILL 200 python scripts/r02_cfg_once.py C2 > gpurun_out/diag_c2.log 2>&1
PC_DEBUG=1 timeout -s KILL 200 python scripts/r02_cfg_once.py C3 > gpurun_out/diag_c3.log 2>&1
PC_DEBUG=1 timeout -s KILL 200 python scripts/r02_cfg_once.py C4 > gpurun_out/diag_c4.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pc_run_kernel -c 1 -f -o gpurun_out/single \
    python scripts/one_run.py > gpurun_out/ncu_single.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:pc_run_kernel -s 1 -c 1 -f -o gpurun_out/ens \
    python scripts/r02_ens_once.py 72 0 2 > gpurun_out/ncu_ens.log 2>&1
timeout -s KILL 500 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section LaunchStats --section Occupancy \
    --section MemoryWorkloadAnalysis --section InstructionStats --clock-control none --import-source on -k regex:pc_run_kernel -c 1 -f -o gpurun_out/c4 \
    python scripts/r02_cfg_once.py C4 0 40000 > gpurun_out/ncu_c4.log 2>&1
tail -3 gpurun_out/diag_c*.log gpurun_out/ncu_*.log
ls -la gpurun_out
