import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/benchmarks')
import gaudi_b200 as gb
from gaudi_b200 import _lib, runtime
from bench_shapes import models
dev = torch.device('cuda:0')
a, model, pred, nd, prop = models("cata", dev)
B = 10000
nm, em = gb.build_masks(torch.full((B,), 10), 10, False, device=dev)
z = runtime.noise(nm.reshape(-1).contiguous(), B, 10, 4, 1.0, 7, 0)
t = torch.full((1,), 0.5, device=dev)
w = torch.tensor([0., -1., 0., 0., 0.], device=dev)
for _ in range(2):
    runtime.predictor_value_and_grad(pred, z, nm, em, t, w)
torch.cuda.synchronize()
buf = np.zeros((3, 1024), dtype=np.uint64)
lib = _lib.lib(); lib.gb_debug_timeline.argtypes = [C.c_void_p]; lib.gb_debug_timeline.restype = C.c_int
print("rc", lib.gb_debug_timeline(buf.ctypes.data_as(C.c_void_p)))
ev = []
for role in range(3):
    for i in range(512):
        code, clk = int(buf[role, 2 * i]), int(buf[role, 2 * i + 1])
        if clk: ev.append((clk, role, code))
ev.sort()
t0 = ev[0][0]
names = {0: "TMA ", 1: "MMA ", 2: "WORK"}
for clk, role, code in ev[:int(sys.argv[1]) if len(sys.argv) > 1 else 100]:
    if role == 2: print(f"{(clk - t0) / 1.9e3:9.2f} us  {names[role]} {code}")
