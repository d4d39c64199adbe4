"""Summarise an ncu report (one launch) into the JSON kept under profiles/:
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_ncu_<tag>.json "<workload>" "<command>" ["note"]
Reads the report here (no GPU needed) with `ncu -i ... --page raw --csv`; `dram_bytes_per_launch` is what bench.py's
roofline.traffic reads."""
import csv, json, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

rep, out, workload, command = sys.argv[1:5]
note = sys.argv[5] if len(sys.argv) > 5 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
m = {}
for k in WANT:
    if k in col:
        try:
            m[k] = {"value": float(vals[col[k]].replace(",", "")), "unit": units[col[k]]}
        except ValueError:
            pass
def nbytes(k):
    return m[k]["value"] * SCALE.get(m[k]["unit"], 1.0) if k in m else 0.0
res = {"kernel": vals[col["Kernel Name"]] if "Kernel Name" in col else None, "workload": workload, "command": command, "note": note,
       "dram_bytes_per_launch": nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"), "metrics": m}
json.dump(res, open(out, "w"), indent=1)
print(out, res["kernel"], res["dram_bytes_per_launch"])
