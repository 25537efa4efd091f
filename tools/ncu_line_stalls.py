#!/usr/bin/env python
"""Development aid: warp-stall samples of one kernel per SOURCE LINE (innermost inlined frame).

`ncu --page source --csv` lists the samples per SASS instruction but has no line column; `nvdisasm --print-line-info-inline` has
the lines but no samples.  Both list the instructions of a function in the same order, so they are joined by instruction index.

  cuobjdump -xelf all gaudi_b200/csrc/tc_pred_edge.o          # -> tc_pred_edge.sm_100a.cubin (same build as the profiled library)
  python tools/ncu_line_stalls.py gpurun_out/prof_r2f_fwd.ncu-rep tc_pred_edge.sm_100a.cubin \\
         _ZN2gb23tc_pred_edge_fwd_kernelILi208ELb1EEEvNS_12PredEdgeArgsEPKfS3_i [top_n] [file:line]

With file:line (e.g. tc_common.cuh:568) the samples of that line are broken down by the CALL SITES it was inlined into.
This is how round 2f found that 14 % of the forward kernel's samples sat in the stage-availability spin of RingsH::put (the part
with the extra 13th chunk kept the others waiting) and 17 % of the backward's in svq_acquire.
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, cubin, fn = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 40
    target = next((a for a in sys.argv[4:] if ":" in a), None)
    out = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text." + fn + ":"))
    frames, mapping = [], []
    for l in lines[start + 1:]:
        if l.startswith("//-----") or l.startswith(".text."):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            frames.append((m.group(1).split("/")[-1], int(m.group(2))))      # innermost frame first
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            mapping.append(tuple(frames))
            frames = []
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    recs = [dict(zip(hdr, r)) for r in rows[hi + 1:] if len(r) == len(hdr)]
    print("SASS instructions", len(recs), "with line info", len(mapping))
    stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    agg = collections.defaultdict(collections.Counter)
    tot, last = 0, ()
    for i, r in enumerate(recs):
        s = int(r["# Samples"] or 0)
        tot += s
        fr = mapping[i] if i < len(mapping) and mapping[i] else last
        last = fr
        if not fr:
            continue
        if target:
            f, ln = target.split(":")
            if fr[0] == (f, int(ln)):
                agg[fr[1:3]]["samples"] += s
            continue
        agg[fr[0]]["samples"] += s
        agg[fr[0]]["inst"] += int(r["Instructions Executed"] or 0)
        for c in stall_cols:
            agg[fr[0]][c] += int(r[c] or 0)
    print("total samples", tot)
    for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        dom = sorted(((v, k[6:]) for k, v in c.items() if k.startswith("stall_")), reverse=True)[:2]
        print(f"{100 * c['samples'] / tot:5.2f}%  inst {c['inst']:10d}  {str(key):60s} {dom}")


if __name__ == "__main__":
    main()
