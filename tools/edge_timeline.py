#!/usr/bin/env python
"""Development aid: print the clock64 timeline of CTA 0 of an edge kernel built with -DGB_TIMELINE.
  GAUDI_B200_LIB=/path/to/lib_tl.so python tools/edge_timeline.py den 0     (kernel family, `which` of gb_profile_kernel)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from gaudi_b200 import runtime, _lib


def main():
    fam, which = sys.argv[1], int(sys.argv[2])
    tiles = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    dev = torch.device("cuda", 0)
    gb, model, pred, nm, em = bench.build_product(dev, 10000)
    B, N, D = 10000, bench.N_RINGS, 4
    tf = gb.AffineTarget.max_gap(pred)
    sched, tvals, dec = model._tables(dev)
    w = (tf.weights * bench.SCALE).to(dev).contiguous()
    z = runtime.noise(nm.reshape(-1).contiguous(), B, N, D, 1.0, 7, 0)
    model.phi(z, tvals[500:501], nm, em, None)
    runtime.predictor_value_and_grad(pred, z, nm, em, tvals[500:501], w)
    L = _lib.lib()
    fn = getattr(L, f"gb_debug_timeline_{fam}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    h = runtime.denoiser_handle(model.dynamics) if which <= 1 else runtime.predictor_handle(pred)
    g = runtime.graph_for(nm, em, B, N)
    ws = runtime.workspace("den" if which <= 1 else "pred_grad", dev).buf
    torch.cuda.synchronize()
    fn(None, None, 1)
    _lib.check(L.gb_profile_kernel(h.handle, g.handle, which, 1, runtime._ptr(ws), ws.numel(), 1, runtime._stream()))
    torch.cuda.synchronize()
    out = np.zeros((5, 2048), dtype=np.uint64); n = np.zeros(5, dtype=np.uint32)
    fn(out.ctypes.data, n.ctypes.data, 0)
    t0 = min(int(out[r, 1]) for r in range(5) if n[r] > 0)
    for r in range(5):
        print(f"--- role {r}: {n[r]} records")
        prev = None; cnt = 0
        for i in range(int(n[r])):
            code, clk = int(out[r, 2 * i]), int(out[r, 2 * i + 1]) - t0
            d = 0 if prev is None else clk - prev
            print(f"  {code:4d} t={clk / 1.9:9.0f}ns  +{d / 1.9:7.0f}ns")
            prev = clk
            if code in ((60, 2, 70) if fam == 'den' else (70, 4)) or (r == 3 and code == 24):
                cnt += 1
                if cnt >= tiles + 1:
                    break


if __name__ == "__main__":
    main()
