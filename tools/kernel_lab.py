#!/usr/bin/env python
"""Development aid: quick parity + per-kernel timing of one build of the library (GAUDI_B200_LIB selects the .so).

  python tools/kernel_lab.py [--batch 10000] [--steps 6] [--tag name]
prints one JSON line: max-abs errors of one teacher-forced guided step vs the CPU oracle (cata ragged + hetro), the
CUDA-event time of each hot kernel (gb_profile_kernel) and the guided step time at the bench shape.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))


def parity(dev):
    import torch
    import gaudi_b200 as gb
    import gaudi_oracle as O
    from helpers import build_models, cpu_weights, oracle_cfgs, oracle_target
    out = {}
    for ds, nx, nmax in (("cata", [10, 9, 11, 7, 10, 3, 11, 11, 2, 1, 10, 10], 11), ("hetro", [10, 8, 3, 9, 10, 10], 10)):
        args, model, pred, prop = build_models(ds, dev)
        wd, wp = cpu_weights(model, pred)
        dcfg, pcfg = oracle_cfgs(ds)
        nxt = torch.tensor(nx)
        nm, em = gb.build_masks(nxt, nmax, ds != "cata", device=dev)
        B, N, D = nm.shape[0], nm.shape[1], 3 + dcfg.in_node_nf
        gen = torch.Generator().manual_seed(0)
        zt = O.draw_noise(B, N, D, nm.cpu(), generator=gen)
        noise = O.draw_noise(B, N, D, nm.cpu(), generator=gen)
        s = 499
        ref = O.guided_step(wd, dcfg, wp, pcfg, O.gamma_table(dcfg), s, zt, noise, nm.cpu(), em.cpu(), oracle_target(ds), 0.6)
        s_arr = torch.full((B, 1), s, device=dev) / model.T
        t_arr = torch.full((B, 1), s + 1, device=dev) / model.T
        tf = gb.AffineTarget.max_gap(pred) if ds == "cata" else gb.AffineTarget.opv(pred, prop)
        got = model.sample_p_zs_given_zt_guidance(s_arr, t_arr, zt.to(dev), nm, em, tf, 0.6, noise=noise.to(dev), return_parts=True)
        torch.cuda.synchronize()
        for k in ("eps", "zs_pre", "grad_raw", "zs"):
            out[f"{ds}_{k}"] = float((got[k].cpu() - ref[k]).abs().max())
    return out


def timing(dev, batch, steps):
    import torch
    import bench
    from gaudi_b200 import runtime
    gb, model, pred, nm, em = bench.build_product(dev, batch)
    B, N, D = batch, bench.N_RINGS, 4
    nmf = nm.reshape(-1).contiguous()
    tf = gb.AffineTarget.max_gap(pred)
    sched, tvals, dec = model._tables(dev)
    w = (tf.weights * bench.SCALE).to(dev).contiguous()
    z = runtime.noise(nmf, B, N, D, 1.0, 7, 0)
    T = bench.T_STEPS
    runtime.sample_loop(model.dynamics, pred, nm, em, z, T, T, T - 3, sched, tvals, w, None, 7, None, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    runtime.sample_loop(model.dynamics, pred, nm, em, z, T, T - 3, T - 3 - steps, sched, tvals, w, None, 7, None, False)
    e1.record()
    torch.cuda.synchronize()
    res = {"ms_per_step": e0.elapsed_time(e1) / steps, "finite": bool(torch.isfinite(z).all())}
    den_h, prd_h = runtime.denoiser_handle(model.dynamics), runtime.predictor_handle(pred)
    g = runtime.graph_for(nm, em, B, N)
    model.phi(z, tvals[500:501], nm, em, None)
    runtime.predictor_value_and_grad(pred, z, nm, em, tvals[500:501], w)
    L = runtime._lib.lib()
    for name, h, wsk, which in (("den_gcl", den_h, "den", 0), ("den_equiv", den_h, "den", 1), ("pred_fwd", prd_h, "pred_grad", 2),
                                ("pred_bwd", prd_h, "pred_grad", 3), ("node_lin", prd_h, "pred_grad", 4)):
        ws = runtime.workspace(wsk, dev).buf
        runtime._lib.check(L.gb_profile_kernel(h.handle, g.handle, which, 1, runtime._ptr(ws), ws.numel(), 1, runtime._stream()))
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        runtime._lib.check(L.gb_profile_kernel(h.handle, g.handle, which, 1, runtime._ptr(ws), ws.numel(), 5, runtime._stream()))
        k1.record()
        torch.cuda.synchronize()
        res[name] = round(k0.elapsed_time(k1) / 5, 4)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--tag", default=os.environ.get("GAUDI_B200_LIB", "default"))
    ap.add_argument("--no-parity", action="store_true")
    a = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    out = {"tag": a.tag}
    if not a.no_parity:
        out["parity"] = {k: float(f"{v:.2e}") for k, v in parity(dev).items()}
        out["parity_ok"] = all(v <= 1e-4 for v in out["parity"].values())
    out.update(timing(dev, a.batch, a.steps))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
