set -u
mkdir -p gpurun_out/ncu
tag=gaussian20_nlive1000_R40_ensemble
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:pc_run_kernel -s 1 -c 1 -f -o /tmp/$tag python scripts/r02_ens_once.py 74 0 2 > gpurun_out/ncu/$tag.log 2>&1
python scripts/ncu_summary.py /tmp/$tag.ncu-rep gpurun_out/ncu/r02_ncu_$tag.json "G20 nlive 1000, 74-run ensemble" "ncu --set full --clock-control none --import-source on -k regex:pc_run_kernel -s 1 -c 1 python scripts/r02_ens_once.py 74 0 2" "final tree of round 2 (narrow phantom traffic)" > /dev/null 2>&1
ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source cuda,sass > /tmp/$tag.src.csv 2>/dev/null
python scripts/ncu_regions.py /tmp/$tag.src.csv > gpurun_out/ncu/$tag.regions.txt 2>&1
python scripts/ncu_lines.py /tmp/$tag.src.csv 40 > gpurun_out/ncu/$tag.lines.txt 2>&1
rm -f /tmp/$tag.ncu-rep /tmp/$tag.src.csv
tail -n 1 gpurun_out/ncu/$tag.log | cut -c1-200
