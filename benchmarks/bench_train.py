#!/usr/bin/env python
"""BASELINE.json config 5: EDM training step (denoising loss forward + backward + AdamW) on a synthetic cc-PBH batch of 512.

One JSON line: CUDA-event time per step (zero_grad, loss, backward, adaptive clipping + optimizer step, as train_edm.py:71-82), molecules/s,
our kernel launches per step, and -- unless --no-cpu -- the CPU oracle's autograd step on a bounded sample beside it.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402

import gaudi_b200 as gb  # noqa: E402
from gaudi_b200 import _lib  # noqa: E402
from gaudi_b200 import train_utils as TU  # noqa: E402
from bench_shapes import models  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=32)
    ap.add_argument("--graph", action="store_true", help="capture the whole step (loss, backward, AdamW) in one CUDA graph and replay it")
    ap.add_argument("--torch-optim", action="store_true", help="torch.optim.AdamW (no clipping) instead of the fused clip + AdamW-amsgrad step")
    ap.add_argument("--model", choices=["denoiser", "predictor"], default="denoiser",
                    help="denoiser: EDM l2 loss (train_edm.py); predictor: l1 property loss on z_t (train_cond_predictor.py)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    a, model, pred, nodes_dist, prop = models("cata", dev)
    for n_, p_ in model.named_parameters():
        p_.requires_grad_(not n_.endswith("gamma.gamma"))
    model.train()
    torch.manual_seed(0)
    B, N = args.batch, 11
    nx = nodes_dist.sample(B)
    nm, em = gb.build_masks(nx, N, False, device=dev)
    x = torch.randn(B, N, 3, device=dev) * 2.4 * nm
    x = x - x.sum(1, keepdim=True) / nm.sum(1, keepdim=True) * nm
    h = {"categorical": torch.ones(B, N, 1, device=dev) * nm, "integer": torch.zeros(0, device=dev)}
    if args.model == "predictor":
        from gaudi_b200 import training
        for p_ in pred.parameters():
            p_.requires_grad_(True)
        pred.train()
        y = torch.randn(B, 5, device=dev)
        opt = (TU.FusedAdamWClip(pred.parameters(), lr=1e-3, weight_decay=1e-12, clip=True) if not args.torch_optim else
               torch.optim.AdamW(pred.parameters(), lr=1e-3, amsgrad=True, weight_decay=1e-12, capturable=args.graph))

        def step():
            opt.zero_grad()
            t = torch.randint(0, model.T + 1, size=(B, 1), device=dev).float() / model.T
            zt = training.sample_edm_t(x, h["categorical"], model, t, nm)
            loss = torch.nn.functional.l1_loss(pred(zt, nm, em, t), y)
            loss.backward()
            opt.step()
            return loss
    else:
        trainable = [p for p in model.parameters() if p.requires_grad]
        opt = (TU.FusedAdamWClip(trainable, lr=1e-4, weight_decay=1e-12, clip=True) if not args.torch_optim else
               torch.optim.AdamW(trainable, lr=1e-4, amsgrad=True, weight_decay=1e-12, capturable=args.graph))

        def step():
            opt.zero_grad()
            loss = model(x, h, nm, em).mean(0)
            loss.backward()
            opt.step()
            return loss

    eager_step = step
    if args.graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                eager_step()
        torch.cuda.current_stream().wait_stream(side)
        cg = torch.cuda.CUDAGraph()
        lc = _lib.lib().gb_launch_count(0)
        with torch.cuda.graph(cg):
            static_loss = eager_step()
        captured = _lib.lib().gb_launch_count(0) - lc          # our kernels inside one replay

        def step():
            cg.replay()
            return static_loss
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    l0 = _lib.lib().gb_launch_count(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    out = {"config": 5 if args.model == "denoiser" else "8f-2",
           "workload": "EDM training step (l2 denoising loss fwd+bwd+AdamW), cc-PBH shape, synthetic batch" if args.model == "denoiser"
           else "property-predictor training step (sample_edm_t + l1 loss fwd+bwd+AdamW), cc-PBH shape, synthetic batch",
           "batch": B, "edges": int(em.sum().item()), "ms_per_step": ms, "molecules_per_s": B / (ms * 1e-3),
           "gpu_launches_per_step": captured if args.graph else (_lib.lib().gb_launch_count(0) - l0) / args.steps, "loss": float(loss.detach()),
           "cuda_graph": bool(args.graph),
           "optimizer": "torch.optim.AdamW(amsgrad)" if args.torch_optim else "fused adaptive clip + AdamW(amsgrad) (gb_adamw_amsgrad_clip)"}
    if not args.no_cpu and args.model == "predictor":
        import gaudi_oracle as O
        cb = args.cpu_batch
        dcfg, pcfg = O.DenoiserCfg(in_node_nf=1), O.PredictorCfg(in_node_nf=1)
        w = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in pred.state_dict().items()}
        nmc, emc = nm[:cb].cpu(), em.view(B, N * N)[:cb].reshape(-1, 1).cpu()
        xc, hc, yc = x[:cb].cpu(), h["categorical"][:cb].cpu(), y[:cb].cpu()
        copt = torch.optim.AdamW(list(w.values()), lr=1e-3, amsgrad=True, weight_decay=1e-12)
        gamma = O.gamma_table(dcfg)

        def cpu_step():
            copt.zero_grad()
            t = torch.randint(0, 1001, (cb, 1)).float() / 1000
            zt = O.sample_edm_t(dcfg, gamma, xc, hc, nmc, t, O.draw_noise(cb, N, 4, nmc))
            torch.nn.functional.l1_loss(O.predictor_forward(w, pcfg, zt, nmc, emc, t), yc).backward()
            copt.step()

        cpu_step()
        t0 = time.perf_counter()
        for _ in range(3):
            cpu_step()
        dt = (time.perf_counter() - t0) / 3
        out["cpu_oracle"] = {"batch": cb, "s_per_step": dt, "molecules_per_s": cb / dt, "threads": torch.get_num_threads()}
        out["speedup_vs_cpu_oracle"] = out["molecules_per_s"] / (cb / dt)
    elif not args.no_cpu:
        import gaudi_oracle as O
        cb = args.cpu_batch
        dcfg = O.DenoiserCfg(in_node_nf=1)
        w = {k: v.detach().cpu().clone().requires_grad_(not k.endswith("gamma.gamma")) for k, v in model.state_dict().items()}
        nmc, emc = nm[:cb].cpu(), em.view(B, N * N)[:cb].reshape(-1, 1).cpu()
        xc, hc = x[:cb].cpu(), h["categorical"][:cb].cpu()
        copt = torch.optim.AdamW([v for v in w.values() if v.requires_grad], lr=1e-4, amsgrad=True, weight_decay=1e-12)
        gamma = O.gamma_table(dcfg)

        def cpu_step():
            copt.zero_grad()
            t_int = torch.randint(0, 1001, (cb, 1)).float()
            eps = O.draw_noise(cb, N, 4, nmc)
            l, _ = O.training_loss(w, dcfg, gamma, xc, hc, nmc, emc, t_int, eps)
            l.mean(0).backward()
            copt.step()

        cpu_step()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            cpu_step()
        dt = (time.perf_counter() - t0) / reps
        out["cpu_oracle"] = {"batch": cb, "s_per_step": dt, "molecules_per_s": cb / dt, "threads": torch.get_num_threads()}
        out["speedup_vs_cpu_oracle"] = out["molecules_per_s"] / (cb / dt)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
