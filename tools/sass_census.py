#!/usr/bin/env python
"""Per-kernel SASS census of libgaudi_b200.so: counts of the mnemonics that prove the tcgen05 / TMEM / bulk-TMA path (and of the
shared / global memory instructions around it).  Regenerate with:   python tools/sass_census.py > profiles/r2_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gaudi_b200", "csrc", "libgaudi_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "BAR", "LDS", "STS", "LDG", "STG",
        "LD.E", "ST.E", "LDL", "STL", "MUFU", "FFMA2", "FMUL2", "FADD2", "F2FP", "HMMA", "ATOM", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, counts, order = None, {}, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            kern = re.sub(r"\(.*", "", kern)
            counts[kern] = collections.Counter(); order.append(kern)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and kern:
            op = m.group(1)
            counts[kern]["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k in ("LD.E", "ST.E") and op.startswith(k)):
                    counts[kern][k] += 1
    print(f"# SASS census of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); columns = instruction counts per kernel")
    print("# tensor core: UTCHMMA (tcgen05.mma kind::tf32 / kind::f16), UTCBAR (tcgen05.commit), LDTM / STTM (tcgen05.ld / st); TMA: UBLKCP (1-D bulk copy)")
    used = [k for k in KEYS if any(counts[n][k] for n in order)]
    print("kernel".ljust(58) + " total " + " ".join(k.rjust(7) for k in used))
    for n in order:
        if not any(counts[n][k] for k in ("UTCHMMA", "UBLKCP", "LDTM")) and "gb::" not in n:
            continue
        print(n[:57].ljust(58) + f"{counts[n]['_total']:6d} " + " ".join(str(counts[n][k]).rjust(7) for k in used))
    tot = collections.Counter()
    for n in order:
        tot.update(counts[n])
    print("ALL KERNELS".ljust(58) + f"{tot['_total']:6d} " + " ".join(str(tot[k]).rjust(7) for k in used))


if __name__ == "__main__":
    main()
