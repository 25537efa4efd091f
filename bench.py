#!/usr/bin/env python
"""bench.py -- guided molecules/sec (1000-step sampling) of the B200-native GaUDI hot path.

Contract (driver):  python bench.py --gpus N --steps K --warmup W   (N>1: launched under torchrun, one rank per GPU)
prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[1]): guided sampling, cc-PBH-shaped ring graphs (all molecules n = N = 10 rings,
90 directed edges), random-init ``args_edm`` denoiser (192 x 9) and ``prediction_args`` predictor (196 x 12, 5 outputs),
target -pred[:,1] (HOMO-LUMO gap), scale 0.6, batch 10 000 molecules PER GPU (weak scaling), synthetic data.

A bench "step" is ONE guided reverse-diffusion step (denoiser forward, z_s draw, predictor forward + input gradient,
guidance update) over the whole batch; every one of the T = 1000 steps of a sampling run launches the identical kernel
sequence, so   molecules/s = batch / (T * step_time + decode_time)   with decode_time (one more denoiser forward + the
decode kernel) measured in the same run.  ``--full`` runs one complete 1000-step sampling instead and reports it too.

value : steps timed with CUDA events with z resident in HBM (fused loop, gb_sample_loop).
e2e   : the same step through the reference-facing API ``EnVariationalDiffusion.sample_p_zs_given_zt_guidance`` with
        HOST (pinned) z_t / noise buffers: H2D copies, the step, and the D2H copy of z_s are all inside the timed region.
--impl reference : the CPU oracle port (oracle/gaudi_oracle.py, plain PyTorch ops in the reference's op order) on all
        host cores, same metric on a bounded batch.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_STEPS = 1000
N_RINGS = 10
SCALE = 0.6


def algorithmic_flops(Ev, N, F, H, L, kind, out=5):
    """2*MACs of the Linear layers as the REFERENCE modules define them, per molecule."""
    if kind == "denoiser":
        edge = Ev * ((2 * H + 2) * H + H * H + H)
        node = N * (2 * H * H + H * H)
        return 2 * (L * (edge + node + edge) + N * 2 * (F + 1) * H)
    edge = Ev * ((2 * H + 2) * H + H * H + H + H * H + H)
    node = N * (2 * H * H + H * H)
    return 2 * (L * (edge + node) + N * ((F + 1) * H + H * out))


def guided_step_flops(n=N_RINGS):
    Ev = n * (n - 1)
    d = algorithmic_flops(Ev, n, 1, 192, 9, "denoiser")
    p = algorithmic_flops(Ev, n, 1, 196, 12, "predictor")
    return d, p, d + 2 * p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return None
        sm, reasons, smax = [], set(), None
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def build_product(device, batch):
    import torch
    import gaudi_b200 as gb
    from argparse import Namespace
    a = gb.args_edm(dataset="cata", device="cpu", dp=False)
    p = gb.prediction_args(device="cpu", dp=False)
    ds = Namespace(num_node_features=1, num_targets=5, mean=torch.zeros(5), std=torch.ones(5))
    torch.manual_seed(0)
    model, _, prop = gb.get_model(a, Namespace(dataset=ds))
    torch.manual_seed(1)
    pred = gb.get_cond_predictor_model(p, ds)
    gb.switch_grad_off([model, pred])
    model, pred = model.to(device), pred.to(device)
    a.device = device
    nm, em = gb.build_masks(torch.full((batch,), N_RINGS), N_RINGS, False, device=device)
    return gb, model, pred, nm, em


def cpu_oracle_rate(batch, max_seconds, threads, steps=None, warmup=1):
    """molecules/s of the CPU oracle for 1000-step guided sampling, from consecutive teacher-forced guided steps."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gaudi_oracle as O
    import gaudi_b200 as gb
    from argparse import Namespace
    torch.set_num_threads(threads)
    a = gb.args_edm(dataset="cata", device="cpu", dp=False)
    p = gb.prediction_args(device="cpu", dp=False)
    ds = Namespace(num_node_features=1, num_targets=5, mean=torch.zeros(5), std=torch.ones(5))
    torch.manual_seed(0)
    model, _, _ = gb.get_model(a, Namespace(dataset=ds))
    torch.manual_seed(1)
    pred = gb.get_cond_predictor_model(p, ds)
    wd = {k: v.detach() for k, v in model.state_dict().items()}
    wp = {k: v.detach() for k, v in pred.state_dict().items()}
    dcfg, pcfg = O.DenoiserCfg(in_node_nf=1), O.PredictorCfg(in_node_nf=1)
    gamma = O.gamma_table(dcfg)
    nm, em = O.build_masks(torch.full((batch,), N_RINGS), N_RINGS, False)
    gen = torch.Generator().manual_seed(0)
    z = O.draw_noise(batch, N_RINGS, 4, nm, generator=gen)
    noise = O.draw_noise(batch, N_RINGS, 4, nm, generator=gen)
    times = []
    s = 500
    with torch.no_grad():
        for _ in range(warmup):
            O.guided_step(wd, dcfg, wp, pcfg, gamma, s, z, noise, nm, em, O.target_max_gap, SCALE)
        t_all = time.perf_counter()
        while True:
            t0 = time.perf_counter()
            O.guided_step(wd, dcfg, wp, pcfg, gamma, s, z, noise, nm, em, O.target_max_gap, SCALE)
            times.append(time.perf_counter() - t0)
            if steps is not None and len(times) >= steps:
                break
            if steps is None and (time.perf_counter() - t_all > max_seconds and len(times) >= 2):
                break
        t0 = time.perf_counter()
        O.decode(wd, dcfg, gamma, z, noise, nm, em)
        t_dec = time.perf_counter() - t0
    step = sum(times) / len(times)
    return batch / (T_STEPS * step + t_dec), step, t_dec, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = args.cpu_batch or 256
    rate, step, t_dec, n = cpu_oracle_rate(batch, 1e9, threads, steps=args.steps, warmup=max(1, args.warmup))
    line = {
        "impl": "reference", "metric": "guided molecules/sec (1000-step sampling)", "value": rate, "unit": "molecules/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "guided sampling cc-PBH-shaped (n=N=10 rings), args_edm denoiser + prediction_args "
                               "predictor, target -pred[:,1], scale 0.6; CPU oracle port of the reference's PyTorch path",
                   "batch": batch, "diffusion_steps": T_STEPS,
                   "step": "one guided reverse-diffusion step over the batch; molecules/s = batch/(1000*step+decode)"},
        "cpu_baseline": {"value": rate, "unit": "molecules/s", "cores": threads, "kind": "port",
                         "sample": f"{n} consecutive guided steps at batch {batch} (+1 decode), extrapolated to 1000 steps"},
        "e2e": {"value": rate, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=10000, help="molecules per GPU")
    ap.add_argument("--cpu-batch", type=int, default=0, help="0: try 64 and 256, report the better")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: min(steps, 5)")
    ap.add_argument("--full", action="store_true", help="also run one complete 1000-step guided sampling")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="few steps, no e2e / cpu legs (for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as tdist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        tdist.init_process_group("nccl", device_id=dev)
    W, K = max(args.warmup, 3), args.steps
    assert W + K + 8 < T_STEPS
    gb, model, pred, nm, em = build_product(dev, args.batch)
    from gaudi_b200 import runtime
    B, N, D = args.batch, N_RINGS, 4
    nmf = nm.reshape(-1).contiguous()
    tf = gb.AffineTarget.max_gap(pred)
    sched, tvals, dec = model._tables(dev)
    w = (tf.weights * SCALE).to(dev).contiguous()
    seed = 1234 + rank
    z = runtime.noise(nmf, B, N, D, 1.0, seed, 0)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ---------------------------------------------------------------------
    runtime.sample_loop(model.dynamics, pred, nm, em, z, T_STEPS, T_STEPS, T_STEPS - W, sched, tvals, w, None, seed, None, False)
    barrier()
    clocks = ClockSampler(local)
    runtime.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    runtime.sample_loop(model.dynamics, pred, nm, em, z, T_STEPS, T_STEPS - W, T_STEPS - W - K, sched, tvals, w, None, seed, None, False)
    ev1.record()
    barrier()
    launches = runtime.launch_count()
    ms_total = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    # decode (one denoiser forward at t=0 + decode kernel)
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    model.sample_p_xh_given_z0(z, nm, em, None, noise=z)
    torch.cuda.synchronize()
    d0.record()
    model.sample_p_xh_given_z0(z, nm, em, None, noise=z)
    d1.record()
    torch.cuda.synchronize()
    ms_dec = d0.elapsed_time(d1)
    t = torch.tensor([ms_total, ms_dec], device=dev, dtype=torch.float64)
    if world > 1:
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    ms_total, ms_dec = float(t[0]), float(t[1])
    ms_step = ms_total / K
    value = world * B / ((T_STEPS * ms_step + ms_dec) * 1e-3)

    # ---- end-to-end through the reference-facing API with host buffers ---------------------------------------
    e2e = None
    if not args.profile_only:
        Ke = args.e2e_steps or min(K, 5)
        zt_h = z.cpu().pin_memory()
        nz_h = runtime.noise(nmf, B, N, D, 1.0, seed, 999).cpu().pin_memory()
        zs_h = torch.empty_like(zt_h).pin_memory()
        s0 = T_STEPS - W - K - 1

        def e2e_step(s):
            s_arr = torch.full((1, 1), s, device=dev) / T_STEPS
            t_arr = torch.full((1, 1), s + 1, device=dev) / T_STEPS
            zt = zt_h.to(dev, non_blocking=True)
            nz = nz_h.to(dev, non_blocking=True)
            zs = model.sample_p_zs_given_zt_guidance(s_arr, t_arr, zt, nm, em, tf, SCALE, noise=nz)
            zs_h.copy_(zs, non_blocking=True)
        e2e_step(s0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(Ke):
            e2e_step(s0 - 1 - i)
            zt_h, zs_h = zs_h, zt_h
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1) / Ke], device=dev, dtype=torch.float64)
        if world > 1:
            tdist.all_reduce(te, op=tdist.ReduceOp.MAX)
        ms_e2e = float(te[0])
        nbytes = B * N * D * 4
        e2e = {"value": world * B / ((T_STEPS * ms_e2e + ms_dec) * 1e-3), "unit": "molecules/s",
               "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": nbytes, "ms_per_step": ms_e2e, "steps": Ke}

    # ---- roofline of the dominant kernels (live CUDA-event timing of single kernels) ---------------------------
    roof, kernels = None, {}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if peaks else "B200_PROFILING.md fallback (sustained)"
        den_h, prd_h = runtime.denoiser_handle(model.dynamics), runtime.predictor_handle(pred)
        g = runtime.graph_for(nm, em, B, N)
        Ev = N * (N - 1)
        # make sure both workspaces hold a finished forward / gradient
        model.phi(z, tvals[500:501], nm, em, None)
        runtime.predictor_value_and_grad(pred, z, nm, em, tvals[500:501], w)
        L = runtime._lib.lib()
        spec = [("den_edge_gcl", den_h, "den", 0, 2 * B * Ev * ((2 * 192 + 2) * 192 + 192 * 192 + 192), 9),
                ("den_edge_equiv", den_h, "den", 1, 2 * B * Ev * ((2 * 192 + 2) * 192 + 192 * 192 + 192), 9),
                ("pred_edge_fwd", prd_h, "pred_grad", 2, 2 * B * Ev * ((2 * 196 + 2) * 196 + 2 * 196 * 196 + 2 * 196), 12),
                ("pred_edge_bwd", prd_h, "pred_grad", 3, 2 * B * Ev * ((2 * 196 + 2) * 196 + 2 * 196 * 196 + 2 * 196), 12),
                ("node_linear", prd_h, "pred_grad", 4, 2 * B * N * (2 * 196 * 196), 0)]
        reps = 3
        for name, h, wsk, which, flops, per_step in spec:
            ws = runtime.workspace(wsk, dev).buf
            runtime._lib.check(L.gb_profile_kernel(h.handle, g.handle, which, 1, runtime._ptr(ws), ws.numel(), 1, runtime._stream()))
            torch.cuda.synchronize()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            runtime._lib.check(L.gb_profile_kernel(h.handle, g.handle, which, 1, runtime._ptr(ws), ws.numel(), reps, runtime._stream()))
            k1.record()
            torch.cuda.synchronize()
            ms = k0.elapsed_time(k1) / reps
            kernels[name] = {"ms": ms, "algorithmic_tflops": flops / ms * 1e-9, "launches_per_step": per_step,
                             "share_of_step": per_step * ms / ms_step}
        top = max((k for k in kernels if kernels[k]["launches_per_step"]), key=lambda k: kernels[k]["share_of_step"])
        ach = kernels[top]["algorithmic_tflops"]
        traffic = None
        try:      # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(top)
        except Exception:
            pass
        mode = os.environ.get("GAUDI_B200_GEMM", "tc")
        roof = {"bound": "tensor", "kernel": top, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": ach / peak_tf, "traffic": traffic, "peak_source": peak_src, "gemm_mode": mode,
                "note": "algorithmic FLOPs = the reference's un-factorised Linear layers (SURVEY 8d) per launch / live CUDA-event "
                        "duration. GEMMs run on tcgen05 as error-compensated 3xTF32 (3 MMAs per product, TF32 dense peak is "
                        "half the bf16 peak used as denominator); the kernel is bound by its CUDA-core build/epilogue phases "
                        "(see profiles/), not by the tensor pipe"}

    full = None
    if args.full:
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        x, h = model.sample_guidance(B, tf, nm, em, scale=SCALE)
        f1.record()
        barrier()
        full = {"seconds": f0.elapsed_time(f1) * 1e-3, "molecules_per_s": world * B / (f0.elapsed_time(f1) * 1e-3)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and not args.profile_only:
        threads = os.cpu_count() or 1
        best = None
        for cb in ([args.cpu_batch] if args.cpu_batch else [64, 256]):
            rate, step, t_dec, n = cpu_oracle_rate(cb, args.cpu_seconds / (1 if args.cpu_batch else 2), threads)
            if best is None or rate > best[0]:
                best = (rate, step, t_dec, n, cb)
        rate, step, t_dec, n, cb = best
        cpu = {"value": rate, "unit": "molecules/s", "cores": threads, "kind": "port",
               "sample": f"{n} consecutive guided steps at batch {cb} (+1 decode) of the CPU oracle port (best of the "
                         f"batches tried), extrapolated to 1000 steps", "ms_per_step": step * 1e3}

    if rank == 0:
        d_fl, p_fl, s_fl = guided_step_flops()
        line = {
            "metric": "guided molecules/sec (1000-step sampling)", "value": value, "unit": "molecules/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: guided sampling cc-PBH-shaped ring graphs (n=N=10 rings, 90 directed edges), "
                                   "random-init args_edm denoiser (192x9) + prediction_args predictor (196x12, 5 out), "
                                   "target -pred[:,1], scale 0.6",
                       "batch_per_gpu": B, "global_batch": world * B, "diffusion_steps": T_STEPS,
                       "step": "one guided reverse-diffusion step over the batch; molecules/s = batch/(1000*step+decode)",
                       "decode_ms": ms_dec, "noise": "in-kernel Philox", "l2": "per-step working set (saved activations "
                       f"{B * 90 * 196 * 4 * 3 * 12 / 1e9:.1f} GB) is far larger than the 126 MB L2",
                       "algorithmic_gflop_per_molecule_step": s_fl * 1e-9,
                       "algorithmic_tflops_achieved": world * B * s_fl / (ms_step * 1e-3) * 1e-12},
            "e2e": e2e, "gpu_launches": launches, "clocks": clk, "roofline": roof, "kernels": kernels,
            "cpu_baseline": cpu,
        }
        if full:
            line["full_run"] = full
        print(json.dumps(line), flush=True)
    if world > 1:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
