"""Static SASS census of libchord.so per kernel:  cuobjdump -sass polychordlite_b200/lib/libchord.so | python scripts/sass_census.py"""
import re, sys, collections
MN = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DADD", "DMUL", "SHFL", "LDG", "STG", "LDS", "STS", "BAR", "ATOM", "RED", "UTMALDG", "UTCHMMA", "LDTM"]
cur = None
cnt = collections.defaultdict(collections.Counter)
for line in sys.stdin:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        cnt[cur]["_all"] += 1
        if op in MN: cnt[cur][op] += 1
import subprocess
def dem(n):
    try: return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception: return n
tot = collections.Counter()
for k, c in cnt.items(): tot.update(c)
print("total:", {k: tot[k] for k in ["_all"] + MN if tot[k]})
want = sys.argv[1:] or ["pc_run_kernelILi4ELi5ELi0ELi0", "pc_run_kernelILi4ELi5ELi0ELi1", "pc_run_kernelILi8ELi8ELi2ELi0", "pc_run_kernelILi4ELi4ELi1ELi0", "gram_schmidt_blockILi8", "moments"]
print("| kernel | instructions | " + " | ".join(MN[:15]) + " |")
print("|---|---|" + "---|" * 15)
for k, c in sorted(cnt.items()):
    if any(w in k for w in want):
        print(f"| `{dem(k)[:90]}` | {c['_all']} | " + " | ".join(str(c[m]) for m in MN[:15]) + " |")
