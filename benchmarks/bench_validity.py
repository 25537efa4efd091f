#!/usr/bin/env python
"""Geometric validity throughput: molecules/s of check_stability on the GPU for a sampler-sized batch,
with the CPU oracle (the reference's per-molecule Python algorithm) timed on a bounded sample beside it.  One JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from gaudi_b200 import analyze  # noqa: E402
import molgen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="cata")
    ap.add_argument("--batch", type=int, default=100000)
    ap.add_argument("--cpu-sample", type=int, default=200)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    ds = args.dataset
    x, rt, nm = molgen.batch(1, ds, 1000, 11 if ds == "cata" else 10)
    reps = (args.batch + 999) // 1000
    X, R, M = (torch.from_numpy(np.tile(a, (reps,) + (1,) * (a.ndim - 1))[:args.batch]).to(dev) for a in (x, rt, nm))
    analyze.check_stability_batch(X, R, M, 0.1, ds)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        flags = analyze.check_stability_batch(X, R, M, 0.1, ds)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    # end to end from host tensors (H2D of x / types / mask, D2H of the flags)
    Xh, Rh, Mh = X.cpu().pin_memory(), R.cpu().pin_memory(), M.cpu().pin_memory()
    t0 = time.perf_counter()
    f2 = analyze.check_stability_batch(Xh.to(dev, non_blocking=True), Rh.to(dev, non_blocking=True), Mh.to(dev, non_blocking=True), 0.1, ds).cpu()
    e2e = time.perf_counter() - t0
    import validity_oracle as VO
    n = args.cpu_sample
    t0 = time.perf_counter()
    for b in range(n):
        m = nm[b].astype(bool)
        VO.check_stability(torch.from_numpy(x[b][m]), torch.from_numpy(rt[b][m]), 0.1, ds)
    cpu = (time.perf_counter() - t0) / n
    bytes_per_mol = X.shape[1] * (12 + 4 + 4) + 8
    print(json.dumps({"workload": f"check_stability, {ds}, synthetic ring graphs", "batch": args.batch, "ms": ms,
                      "molecules_per_s": args.batch / (ms * 1e-3), "e2e_molecules_per_s": args.batch / e2e,
                      "algorithmic_GBps": args.batch * bytes_per_mol / (ms * 1e-3) / 1e9,
                      "stable_fraction": float(flags[:, 5].float().mean()),
                      "cpu_oracle": {"sample": n, "molecules_per_s": 1.0 / cpu, "cores": 1},
                      "speedup_vs_cpu_oracle": args.batch / (ms * 1e-3) * cpu}))


if __name__ == "__main__":
    main()
