#!/usr/bin/env python
"""bench.py -- guided molecules/sec (1000-step sampling) of the B200-native GaUDI hot path.

Contract (driver):  python bench.py --gpus N --steps K --warmup W   (N>1: launched under torchrun, one rank per GPU)
prints ONE JSON line on rank 0.

--config 2 (default, BASELINE.json configs[1]): guided sampling, cc-PBH-shaped ring graphs (all molecules n = N = 10 rings, 90 directed
    edges), random-init ``args_edm`` denoiser (192 x 9) and ``prediction_args`` predictor (196 x 12, 5 outputs), target -pred[:,1]
    (HOMO-LUMO gap), scale 0.6, batch 10 000 molecules PER GPU (weak scaling), synthetic data.
--config 3 (configs[2]): PASs shape (10 rings + 10 orientation nodes, N 20, F 12, 110 directed edges), multi-objective target
    ip + ea + 3*gap on un-normalised predictions (``AffineTarget.opv``), 12 500 molecules per GPU (100 000 over 8 GPUs).
--config 4 (configs[3]): predictor forward + input-gradient microbenchmark (n = 10, batch 10 000, hidden 196; ``--sweep`` adds the
    nodes x batch x hidden grid of benchmarks/bench_shapes.py).
--config 5 (configs[4]): EDM training step (l2 denoising loss forward + backward + AdamW) on a synthetic cc-PBH batch of 512.

A bench "step" (configs 2 / 3) is ONE guided reverse-diffusion step (denoiser forward, z_s draw, predictor forward + input gradient,
guidance update) over the whole batch; every one of the T = 1000 steps of a sampling run launches the identical kernel sequence, so
    molecules/s = batch / (T * step_time + decode_time)
with decode_time (one more denoiser forward + the decode kernel) measured in the same run.

value    : steps timed with CUDA events with z resident in HBM (fused loop, gb_sample_loop).
e2e      : the same step through the reference-facing API ``EnVariationalDiffusion.sample_p_zs_given_zt_guidance`` with HOST (pinned)
           z_t / noise buffers: H2D copies, the step, and the D2H copy of z_s are all inside the timed region.
e2e_full : ONE complete ``gaudi_b200.sample_guidance(args, model, target, nodesxsample)`` call (masks, topology, 1000 steps as a replayed
           CUDA graph, decode, invariants, D2H of x / one-hot) at a batch sized for ~10 s; validates the per-step extrapolation.
sharded  : (N > 1) ``dist.sample_guidance_sharded`` on a short schedule: every rank samples its shard, then ONE NCCL all_gather of
           x / one-hot / node mask, timed separately (gather_ms, bytes).
--impl reference : the CPU oracle port (oracle/gaudi_oracle.py, plain PyTorch ops in the reference's op order; verified within 5 % of
           the unmodified reference's own step time by the round-1 review) on all host cores, same metric on a bounded batch.
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
SCALE = 0.6
PRECISION = ("fp32 activations and accumulation; every GEMM on tcgen05 as one TF32 MMA + one bf16 MMA carrying both correction terms per "
             "K step (error-compensated, fp32-class); saved SiLU derivatives of the input-gradient pass as 16-bit fixed-point codes "
             "(error 1e-5); max-abs error vs the fp32 reference 3e-6 per step (tolerance 1e-4)")

WORKLOADS = {
    2: dict(dataset="cata", rings=10, N=10, F=1, edges=90, batch=10000, target="max_gap",
            name="configs[1]: guided sampling cc-PBH-shaped ring graphs (n=N=10 rings, 90 directed edges), random-init args_edm denoiser "
                 "(192x9) + prediction_args predictor (196x12, 5 out), target -pred[:,1], scale 0.6"),
    3: dict(dataset="hetro", rings=10, N=20, F=12, edges=110, batch=12500, target="opv",
            name="configs[2]: guided sampling PASs-shaped graphs (10 rings + 10 orientation nodes, N=20, F=12, 110 directed edges), "
                 "random-init args_edm denoiser (192x9) + prediction_args predictor (196x12, 5 out), multi-objective target "
                 "ip+ea+3*gap on un-normalised predictions, scale 0.6; 100 000 molecules = 8 x 12 500"),
}


def algorithmic_flops(Ev, N, F, H, L, kind, out=5):
    """2*MACs of the Linear layers as the REFERENCE modules define them, per molecule."""
    if kind == "denoiser":
        edge = Ev * ((2 * H + 2) * H + H * H + H)
        node = N * (2 * H * H + H * H)
        return 2 * (L * (edge + node + edge) + N * 2 * (F + 1) * H)
    edge = Ev * ((2 * H + 2) * H + H * H + H + H * H + H)
    node = N * (2 * H * H + H * H)
    return 2 * (L * (edge + node) + N * ((F + 1) * H + H * out))


def guided_step_flops(wl):
    d = algorithmic_flops(wl["edges"], wl["N"], wl["F"], 192, 9, "denoiser")
    p = algorithmic_flops(wl["edges"], wl["N"], wl["F"], 196, 12, "predictor")
    return d, p, d + 2 * p


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


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


def build_nets(wl, device, timesteps=T_STEPS):
    """(gb, args, model, predictor, prop_dist, target) with random-init weights of the args_edm / prediction_args architectures."""
    import torch
    import gaudi_b200 as gb
    from argparse import Namespace
    a = gb.args_edm(dataset=wl["dataset"], device="cpu", dp=False, max_nodes=11 if wl["dataset"] == "cata" else 10,
                    diffusion_steps=timesteps)
    p = gb.prediction_args(device="cpu", dp=False)
    ds = Namespace(num_node_features=wl["F"], num_targets=5, mean=torch.zeros(5), std=torch.ones(5))
    torch.manual_seed(0)
    model, _, prop = gb.get_model(a, Namespace(dataset=ds))
    torch.manual_seed(1)
    pred = gb.get_cond_predictor_model(p, ds)
    gb.switch_grad_off([model, pred])
    model, pred = model.to(device), pred.to(device)
    a.device = device
    tf = gb.AffineTarget.max_gap(pred) if wl["target"] == "max_gap" else gb.AffineTarget.opv(pred, prop)
    return gb, a, model, pred, prop, tf


def build_product(device, batch, wl=None):
    """Kept for the development tools (tools/kernel_lab.py): config-2 networks and masks."""
    import torch
    wl = wl or WORKLOADS[2]
    gb, a, model, pred, prop, tf = build_nets(wl, device)
    nm, em = gb.build_masks(torch.full((batch,), wl["rings"]), wl["rings"], wl["dataset"] != "cata", device=device)
    return gb, model, pred, nm, em


N_RINGS = WORKLOADS[2]["rings"]


def cpu_oracle_rate(wl, batch, max_seconds, threads, steps=None, warmup=1):
    """molecules/s of the CPU oracle for 1000-step guided sampling, from consecutive teacher-forced guided steps."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gaudi_oracle as O
    torch.set_num_threads(threads)
    gb, a, model, pred, prop, tf = build_nets(wl, "cpu")
    wd = {k: v.detach() for k, v in model.state_dict().items()}
    wp = {k: v.detach() for k, v in pred.state_dict().items()}
    dcfg, pcfg = O.DenoiserCfg(in_node_nf=wl["F"]), O.PredictorCfg(in_node_nf=wl["F"])
    gamma = O.gamma_table(dcfg)
    nm, em = O.build_masks(torch.full((batch,), wl["rings"]), wl["rings"], wl["dataset"] != "cata")
    target = O.target_max_gap if wl["target"] == "max_gap" else O.make_target_opv(torch.zeros(5), torch.ones(5))
    gen = torch.Generator().manual_seed(0)
    D = 3 + wl["F"]
    z = O.draw_noise(batch, wl["N"], D, nm, generator=gen)
    noise = O.draw_noise(batch, wl["N"], D, nm, generator=gen)
    times = []
    s = 500
    with torch.no_grad():
        for _ in range(warmup):
            O.guided_step(wd, dcfg, wp, pcfg, gamma, s, z, noise, nm, em, target, SCALE)
        t_all = time.perf_counter()
        while True:
            t0 = time.perf_counter()
            O.guided_step(wd, dcfg, wp, pcfg, gamma, s, z, noise, nm, em, target, SCALE)
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
    wl = WORKLOADS[args.config if args.config in WORKLOADS else 2]
    threads = os.cpu_count() or 1
    batch = args.cpu_batch or 256
    rate, step, t_dec, n = cpu_oracle_rate(wl, batch, 1e9, threads, steps=args.steps, warmup=max(3, args.warmup))
    line = {
        "impl": "reference", "metric": "guided molecules/sec (1000-step sampling)", "value": rate, "unit": "molecules/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"] + "; CPU oracle port of the reference's PyTorch path", "batch": batch,
                   "diffusion_steps": T_STEPS,
                   "step": "one guided reverse-diffusion step over the batch; molecules/s = batch/(1000*step+decode)"},
        "cpu_baseline": {"value": rate, "unit": "molecules/s", "cores": threads, "cpu_model": cpu_model(), "kind": "port",
                         "sample": f"{n} consecutive guided steps at batch {batch} (+1 decode), extrapolated to 1000 steps"},
        "e2e": {"value": rate, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# configs 2 / 3: guided sampling
# ------------------------------------------------------------------------------------------------------------------
def run_guided(args, wl):
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
    B = args.batch or wl["batch"]
    gb, a_edm, model, pred, prop, tf = build_nets(wl, dev)
    from gaudi_b200 import runtime
    N, D = wl["N"], 3 + wl["F"]
    orient = wl["dataset"] != "cata"
    nm, em = gb.build_masks(torch.full((B,), wl["rings"]), wl["rings"], orient, device=dev)
    nmf = nm.reshape(-1).contiguous()
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
        for i in range(3):
            e2e_step(s0 + i)
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
        peak_tf = peaks.get("bf16_tflops", 1650.0)            # the kernel is timed alone: burst figure
        peak_src = "MEASURED_PEAKS.json bf16_tflops (measured, burst: the kernel is timed in isolation)" if peaks else "B200_PROFILING.md fallback (burst)"
        den_h, prd_h = runtime.denoiser_handle(model.dynamics), runtime.predictor_handle(pred)
        g = runtime.graph_for(nm, em, B, N)
        Ev = wl["edges"]
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
        roof = {"bound": "tensor", "kernel": top, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": ach / peak_tf, "traffic": traffic, "peak_source": peak_src, "gemm_mode": os.environ.get("GAUDI_B200_GEMM", "tc"),
                "note": "algorithmic FLOPs = the reference's un-factorised Linear layers (SURVEY 8d) per launch / live CUDA-event duration. "
                        "The kernels execute 2.3x fewer MACs (factorised first Linear) as one TF32 and one bf16 MMA per K step; their "
                        "limiter is the shared-memory / L1 data pipe (operand stores + tensor-core operand reads, 70-78 % busy in "
                        "profiles/r2*_ncu_*.txt), not the tensor pipe"}

    # ---- one complete sample_guidance() call -----------------------------------------------------------------------
    full = None
    if not args.profile_only and not args.no_full:
        Bf = args.full_batch or max(256, B // 5)
        nx = torch.full((Bf,), wl["rings"])
        model.set_seed(seed)
        gb.sample_guidance(a_edm, model, tf, nx[:64], scale=SCALE)          # warm-up: handles, graph capture path
        barrier()
        t0 = time.perf_counter()
        x, one_hot, nmk, _ = gb.sample_guidance(a_edm, model, tf, nx, scale=SCALE)
        x_h, oh_h = x.cpu(), one_hot.cpu()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        tt = torch.tensor([sec], device=dev, dtype=torch.float64)
        if world > 1:
            tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
        sec = float(tt[0])
        model.set_seed(None)
        full = {"batch_per_gpu": Bf, "seconds": sec, "value": world * Bf / sec, "unit": "molecules/s",
                "d2h_bytes": int(x_h.numel() * 4 + oh_h.numel() * 4), "finite": bool(torch.isfinite(x_h).all()),
                "what": "one gaudi_b200.sample_guidance() call: masks + topology, 1000 guided steps (replayed CUDA graph, Philox noise), "
                        "decode, invariant checks, D2H of x / one-hot; wall clock"}

    # ---- the product's sharded API: per-rank sampling + ONE all_gather --------------------------------------------------
    sharded = None
    if world > 1 and not args.profile_only:
        from gaudi_b200 import dist
        gb2, a2, model2, pred2, prop2, tf2 = build_nets(wl, dev, timesteps=8)
        Bs = min(B, 2000)
        nx_all = torch.full((world * Bs,), wl["rings"])
        timing = {}
        dist.sample_guidance_sharded(a2, model2, tf2, nx_all, scale=SCALE, seed=77)            # warm-up
        barrier()
        t0 = time.perf_counter()
        xs, ohs, nms = dist.sample_guidance_sharded(a2, model2, tf2, nx_all, scale=SCALE, seed=77, timing=timing)
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        ts = torch.tensor([sec, timing.get("gather_ms", 0.0)], device=dev, dtype=torch.float64)
        tdist.all_reduce(ts, op=tdist.ReduceOp.MAX)
        sharded = {"molecules": world * Bs, "diffusion_steps": 8, "seconds": float(ts[0]), "gather_ms": float(ts[1]),
                   "gather_bytes_per_rank": int(timing.get("gather_bytes", 0)), "gathered_shape": list(xs.shape),
                   "what": "dist.sample_guidance_sharded: contiguous shards, per-rank Philox seed, one NCCL all_gather of x / one-hot / mask"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and not args.profile_only:
        threads = os.cpu_count() or 1
        sweep = []
        batches = [args.cpu_batch] if args.cpu_batch else [64, 256, 1024]
        for cb in batches:
            rate, step, t_dec, n = cpu_oracle_rate(wl, cb, args.cpu_seconds / len(batches), threads, warmup=3 if cb <= 256 else 1)
            sweep.append({"batch": cb, "molecules_per_s": rate, "ms_per_step": step * 1e3, "steps": n})
        best = max(sweep, key=lambda r: r["molecules_per_s"])
        cpu = {"value": best["molecules_per_s"], "unit": "molecules/s", "cores": threads, "cpu_model": cpu_model(), "kind": "port",
               "sample": f"{best['steps']} consecutive guided steps at batch {best['batch']} (+1 decode) of the CPU oracle port (best of the "
                         f"batch sweep), extrapolated to 1000 steps", "ms_per_step": best["ms_per_step"], "sweep": sweep}

    if rank == 0:
        d_fl, p_fl, s_fl = guided_step_flops(wl)
        sv_gb = B * wl["edges"] * 196 * (2 + 4 + 2) * 12 / 1e9         # 16-bit SiLU-derivative codes x 2 + fp32 pre2, 12 layers
        line = {
            "metric": "guided molecules/sec (1000-step sampling)", "value": value, "unit": "molecules/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "bench_config": args.config, "precision": PRECISION,
                       "batch_per_gpu": B, "global_batch": world * B, "diffusion_steps": T_STEPS,
                       "step": "one guided reverse-diffusion step over the batch; molecules/s = batch/(1000*step+decode)",
                       "decode_ms": ms_dec, "noise": "in-kernel Philox",
                       "l2": f"per-step working set (saved activations {sv_gb:.1f} GB) is far larger than the 126 MB L2",
                       "algorithmic_gflop_per_molecule_step": s_fl * 1e-9,
                       "algorithmic_tflops_achieved": world * B * s_fl / (ms_step * 1e-3) * 1e-12},
            "e2e": e2e, "e2e_full": full, "sharded": sharded, "gpu_launches": launches, "clocks": clk, "roofline": roof,
            "kernels": kernels, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        tdist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
# config 4: predictor forward + input gradient
# ------------------------------------------------------------------------------------------------------------------
def run_config4(args):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    import bench_shapes
    from gaudi_b200 import runtime
    import gaudi_b200 as gb
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n, B, H = 10, args.batch or 10000, args.hidden
    a, model, pred, nodes_dist, prop = bench_shapes.models("cata", dev, nf_pred=H)
    nm, em = gb.build_masks(torch.full((B,), n), n, False, device=dev)
    z = runtime.noise(nm.reshape(-1).contiguous(), B, n, 4, 1.0, 3, 0)
    t = torch.full((1,), 0.5, device=dev)
    w = torch.tensor([0., -1., 0., 0., 0.], device=dev)
    for _ in range(max(args.warmup, 3)):
        runtime.predictor_value_and_grad(pred, z, nm, em, t, w)
    torch.cuda.synchronize()
    runtime.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        runtime.predictor_value_and_grad(pred, z, nm, em, t, w)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    ev = n * (n - 1)
    fl = 2 * algorithmic_flops(ev, n, 1, H, 12, "predictor") * B
    line = {"metric": "predictor forward + input-gradient molecules/sec", "value": B / (ms * 1e-3), "unit": "molecules/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[3]: EGNN predictor forward + input gradient, n = 10 rings, batch %d, hidden %d x 12 layers" % (B, H),
                       "bench_config": 4, "precision": PRECISION, "algorithmic_tflops_achieved": fl / ms * 1e-9},
            "gpu_launches": runtime.launch_count()}
    print(json.dumps(line), flush=True)
    if args.sweep:
        bench_shapes.config4(dev)


# ------------------------------------------------------------------------------------------------------------------
# config 5: EDM training step
# ------------------------------------------------------------------------------------------------------------------
def run_config5(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cmd = [sys.executable, os.path.join(ROOT, "benchmarks", "bench_train.py"), "--batch", str(args.batch or 512), "--steps", str(args.steps),
           "--warmup", str(max(args.warmup, 3)), "--graph"] + (["--no-cpu"] if args.no_cpu else [])
    out = subprocess.run(cmd, capture_output=True, text=True)
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    line = {"metric": "EDM training-step molecules/sec (l2 denoising loss fwd + bwd + fused AdamW)", "value": rec["molecules_per_s"],
            "unit": "molecules/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": rec["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[4]: EDM training step on a synthetic cc-PBH batch of 512 (ring-count histogram), whole step in one CUDA graph",
                       "bench_config": 5, "batch": rec["batch"], "edges": rec["edges"]},
            "gpu_launches": rec["gpu_launches_per_step"] * args.steps, "cpu_baseline": rec.get("cpu_oracle")}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json config to measure (2 = the headline line)")
    ap.add_argument("--batch", type=int, default=0, help="molecules per GPU (0: the config's own size)")
    ap.add_argument("--cpu-batch", type=int, default=0, help="0: sweep 64 / 256 / 1024 and report the best")
    ap.add_argument("--cpu-seconds", type=float, default=24.0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: min(steps, 5)")
    ap.add_argument("--full-batch", type=int, default=0, help="batch of the complete sample_guidance() call (0: batch / 5)")
    ap.add_argument("--no-full", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--hidden", type=int, default=196, help="config 4: predictor hidden width (196 = prediction_args, 192 / 256 = the sweep's widths)")
    ap.add_argument("--sweep", action="store_true", help="config 4: also print the nodes x batch x hidden sweep")
    ap.add_argument("--profile-only", action="store_true", help="few steps, no e2e / cpu legs (for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 4:
        return run_config4(args)
    if args.config == 5:
        return run_config5(args)
    return run_guided(args, WORKLOADS[args.config])


if __name__ == "__main__":
    main()
