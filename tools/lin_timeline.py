"""Per-tile timeline of the node-Linear kernel (CTA 0): build the library with `make -C gaudi_b200/csrc EXTRA=-DGB_TIMELINE` first.
BLD 200+q: builder warps published atom q; MMA 1000+q / 2000+q: atom q's A / W operands seen by the MMA warp;
EPI 300+t / 400+t: epilogue warps start / finish tile t.  Times in us at 1.9 GHz."""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
from gaudi_b200 import _lib, training
dev = torch.device('cuda:0')
M, K, N = 100000, 192, 192
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
for _ in range(3):
    y = training._linear(A, W, K, N, b, epi=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y = training._linear(A, W, K, N, b, epi=1)
e1.record(); torch.cuda.synchronize()
print("ms per call incl pack (10 calls)", e0.elapsed_time(e1) / 10)
buf = np.zeros((3, 1024), dtype=np.uint64)
lib = _lib.lib(); lib.gb_debug_timeline.argtypes = [C.c_void_p]; lib.gb_debug_timeline.restype = C.c_int
print("rc", lib.gb_debug_timeline(buf.ctypes.data_as(C.c_void_p)))
dur = np.zeros(512, dtype=np.uint64)
lib.gb_debug_timeline_dur.argtypes = [C.c_void_p]; lib.gb_debug_timeline_dur.restype = C.c_int
lib.gb_debug_timeline_dur(dur.ctypes.data_as(C.c_void_p))
d = np.sort(dur[dur > 0].astype(np.float64)) / 1.9e3
print(f"per-CTA entry-to-exit: n={len(d)} min {d[0]:.1f} median {d[len(d)//2]:.1f} max {d[-1]:.1f} us")
ev = []
for role in range(3):
    for i in range(512):
        code, clk = int(buf[role, 2 * i]), int(buf[role, 2 * i + 1])
        if clk: ev.append((clk, role, code))
ev.sort()
t0 = ev[0][0]
names = {0: "BLD ", 1: "MMA ", 2: "EPI "}
for clk, role, code in ev[:140]:
    print(f"{(clk - t0) / 1.9e3:9.2f} us  {names[role]} {code}")
