#!/usr/bin/env python
"""Development aid: the SASS lines with the most warp-stall samples of one kernel in a .ncu-rep (ncu --set full --import-source on),
with the dominant stall reason and the shared-memory wavefront excess per line.   python tools/ncu_hot_lines.py rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
recs = [dict(zip(hdr, r)) for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r["# Samples"] or 0) for r in recs)
print("total samples", tot, "instructions", len(recs))
agg = {c: sum(int(r[c] or 0) for r in recs) for c in stall_cols}
print("stall totals:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01))
recs_s = sorted(enumerate(recs), key=lambda ir: -int(ir[1]["# Samples"] or 0))[:top]
for i, r in recs_s:
    s = int(r["# Samples"] or 0)
    dom = max(stall_cols, key=lambda c: int(r[c] or 0))
    exc = r.get("L1 Wavefronts Shared Excessive", "0")
    print(f"{i:5d} {100 * s / tot:5.2f}%  {dom[6:]:18s} exc_wf {exc:>9s}  {r['Source'][:110]}")
