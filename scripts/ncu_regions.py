"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by region of the engine source (chain phase,
preparation, phase U, ...): share of the warp samples and of the executed instructions.  The regions are found from marker
lines in the source files of THIS tree (so the capture must be of this tree):
    python scripts/ncu_regions.py export.csv"""
import csv, sys
from pathlib import Path

CSRC = Path(__file__).resolve().parents[1] / "polychordlite_b200" / "csrc"
# (file, [(marker substring that starts the region, region name), ...]) in file order; a region runs to the next marker
MARKS = {
    "pc_chain.cuh": [("struct Model {", "Model (likelihood + prior of a point group)"), ("void gram_schmidt_block(", "prep: Gram-Schmidt (block, DMMA)"),
                     ("inline void prep_chain(", "prep: shuffle picks, uniforms"), ("auto fill_gauss = ", "prep: Gaussian deviates"),
                     ("const bool block = HAS_BLOCK", "prep: dispatch"), ("const int B = nb >= 4", "prep: Gram-Schmidt (vector at a time)"),
                     ("if (cs.uni && can_stage) draw_uniforms();", "prep: deck"), ("inline void whiten_chain(", "whiten"),
                     ("slow_uniform(unsigned seed", "slice_chain (speculative rounds)")],
    "pc_dense.cuh": [("inline void dense_table_fill(", "dense: table"), ("inline void dense_store(", "dense: store slice records"),
                     ("inline void slice_chains_dense(", "dense slice loop")],
    "pc_run_kernel.cuh": [("inline void init_phase(", "init"), ("------ phase S (CTA 0)", "phase S"), ("------ phase D (every CTA", "phase D"),
                          ("---- phase U", "U: boost / kept"), ("inline void phase_UA(", "phase UA"), ("inline void phase_UB(", "phase UB"),
                          ("inline bool finish_update(", "finish_update"), ("sharded run: last-baby exchange", "shard / dump"),
                          ("------ the persistent run kernel", "run kernel: generation head (barriers, publication, dead copies)"),
                          ("---- dense chain phase (pc_dense.cuh)", "run kernel: dense chain driver"),
                          ("} else if (p.paired) {", "run kernel: chain drivers, arrival, update calls"),
                          ("// ---------------------------------------------------------------- probes", "probes")],
    "pc_device.cuh": [("struct u4", "philox / uniform"), ("inv_normal_cdf_central(double p", "AS241"), ("logaddexp(double a", "device misc")],
    "pc_kernels.cuh": [("#pragma once", "barriers and waits (pc_kernels.cuh)")],
}
bounds = {}
for f, marks in MARKS.items():
    lines = (CSRC / f).read_text().splitlines()
    found = []
    for mk, name in marks:
        ln = next((i + 1 for i, l in enumerate(lines) if mk in l), None)
        if ln is not None:
            found.append((ln, name))
    bounds[f] = sorted(found)


def region(f, l):
    if f not in bounds:
        return f
    name = f + " (head)"
    for ln, nm in bounds[f]:
        if l >= ln:
            name = nm
    return name


rows = list(csv.reader(open(sys.argv[1])))
cur = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Line No"):
        continue
    if len(r) > 8 and r[2] == "-":
        try:
            line = int(r[0]); s = int(r[4]); i = int(r[7])
        except ValueError:
            continue
        a = agg.setdefault(region(cur, line), [0, 0])
        a[0] += s; a[1] += i
ts = sum(v[0] for v in agg.values()) or 1
ti = sum(v[1] for v in agg.values()) or 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:64s} samples {100 * v[0] / ts:5.1f}%  inst {100 * v[1] / ti:5.1f}%")
