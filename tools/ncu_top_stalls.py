"""Top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view).
usage: ncu -i rep --page source --csv | python tools/ncu_top_stalls.py <kernel-substring> [instance] [top]"""
import csv, sys
pat, inst, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else -1, int(sys.argv[3]) if len(sys.argv) > 3 else 40
kernels, cur, hdr = [], None, None
for row in csv.reader(sys.stdin):
    if not row: continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}; kernels.append(cur); continue
    if row[0] == "Address": hdr = row; continue
    if cur is not None and hdr: cur["rows"].append(row)
sel = [k for k in kernels if pat in k["name"]]
k = sel[inst]
si, ai, ei = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ai] or 0) for r in k["rows"])
print(k["name"][:90], "instances", len(sel), "total samples", tot, "instructions", len(k["rows"]))
agg = {}
for r in k["rows"]:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print({a: round(100 * b / max(1, sum(agg.values())), 1) for a, b in sorted(agg.items(), key=lambda x: -x[1])[:8]})
rows = sorted(enumerate(k["rows"]), key=lambda ir: -int(ir[1][ai] or 0))[:top]
for idx, r in sorted(rows):
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{idx:5d} {100 * int(r[ai]) / max(1, tot):5.1f}%  exec {r[ei]:>8s}  {r[si].strip()[:70]:70s} {st}")
