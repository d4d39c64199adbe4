#!/bin/bash
# ncu captures of the final tree, summarised ON the GPU box (the reports themselves exceed what gpurun copies back):
# raw metrics -> profiles-style JSON, source view -> shares per region of the engine source and per line.
set -u
mkdir -p gpurun_out/ncu
cap() {   # tag, workload, ncu args..., -- command
  local tag=$1 workload=$2; shift 2
  local args=(); while [ "$1" != "--" ]; do args+=("$1"); shift; done; shift
  timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:pc_run_kernel "${args[@]}" -f -o /tmp/$tag "$@" > gpurun_out/ncu/$tag.log 2>&1
  python scripts/ncu_summary.py /tmp/$tag.ncu-rep gpurun_out/ncu/r02_ncu_$tag.json "$workload" "ncu --set full --clock-control none --import-source on -k regex:pc_run_kernel ${args[*]} $*" "final tree of round 2" > /dev/null 2>&1
  ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source cuda,sass > /tmp/$tag.src.csv 2>/dev/null
  python scripts/ncu_regions.py /tmp/$tag.src.csv > gpurun_out/ncu/$tag.regions.txt 2>&1
  python scripts/ncu_lines.py /tmp/$tag.src.csv 40 > gpurun_out/ncu/$tag.lines.txt 2>&1
  rm -f /tmp/$tag.ncu-rep /tmp/$tag.src.csv
  tail -n 2 gpurun_out/ncu/$tag.log | cut -c1-200
}
cap gaussian20_nlive1000_R40_single "G20 nlive 1000, one run" -c 1 -- python scripts/one_run.py
cap gaussian20_nlive1000_R40_ensemble "G20 nlive 1000, 74-run ensemble" -s 1 -c 1 -- python scripts/r02_ens_once.py 74 0 2
cap corr_gaussian50_nlive4000_R250 "C50 nlive 4000 R 250, first 40000 deaths" -c 1 -- python scripts/r02_cfg_once.py C4 0 40000
cap rastrigin10_nlive2000_R50 "R10 nlive 2000 clustered, first launch" -c 1 -- python scripts/r02_cfg_once.py C3
python scripts/r02_ens_once.py 74 0 2 > gpurun_out/ens74.log 2>&1; tail -1 gpurun_out/ens74.log
python scripts/r02_ens_once.py 72 0 2 > gpurun_out/ens72.log 2>&1; tail -1 gpurun_out/ens72.log
ls -la gpurun_out/ncu
