#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs (not the driver's bench line; see bench.py for that).

  --config 1   unconditional sampling, cc-PBH shape, ring-count histogram, B 64 (full 1000-step run, graph replay)
  --config 3   guided, PASs shape (10 rings + 10 orientation nodes, N 20, F 12), multi-objective target, per-GPU share
               of the 100k batch (12 500), K guided steps timed like bench.py
  --config 4   predictor forward + input-gradient sweep (nodes x batch x hidden)
Every line is JSON; times are CUDA-event times on the launching stream after warm-up.
"""
import argparse
import json
import os
import sys
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import gaudi_b200 as gb  # noqa: E402
from gaudi_b200 import runtime  # noqa: E402


def models(dataset, dev, nf_den=192, nf_pred=196, out=5):
    F = 1 if dataset == "cata" else 12
    a = gb.args_edm(dataset=dataset, device="cpu", dp=False, nf=nf_den, max_nodes=11 if dataset == "cata" else 10)
    p = gb.prediction_args(device="cpu", dp=False, nf=nf_pred)
    ds = Namespace(num_node_features=F, num_targets=out, mean=torch.zeros(out), std=torch.ones(out))
    torch.manual_seed(0)
    model, nodes_dist, prop = gb.get_model(a, Namespace(dataset=ds))
    torch.manual_seed(1)
    pred = gb.get_cond_predictor_model(p, ds)
    gb.switch_grad_off([model, pred])
    a.device = dev
    return a, model.to(dev), pred.to(dev), nodes_dist, prop


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def config1(dev, batch):
    a, model, pred, nodes_dist, prop = models("cata", dev)
    torch.manual_seed(0)
    nx = nodes_dist.sample(batch)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gb.sample_pos_edm(a, model, nx)                       # warm-up (also builds handles / graph)
    torch.cuda.synchronize()
    e0.record()
    x, oh, nm, em = gb.sample_pos_edm(a, model, nx)
    e1.record()
    torch.cuda.synchronize()
    s = e0.elapsed_time(e1) * 1e-3
    print(json.dumps({"config": 1, "workload": "unconditional EDM sampling, cc-PBH shape, ring-count histogram, 1000 steps",
                      "batch": batch, "seconds": s, "molecules_per_s": batch / s, "cuda_graph": True}))


def config3(dev, batch, steps):
    a, model, pred, nodes_dist, prop = models("hetro", dev)
    nm, em = gb.build_masks(torch.full((batch,), 10), 10, True, device=dev)
    B, N, D = batch, 20, 15
    tf = gb.AffineTarget.opv(pred, prop)
    sched, tvals, dec = model._tables(dev)
    w = (tf.weights * 0.6).to(dev).contiguous()
    z = runtime.noise(nm.reshape(-1).contiguous(), B, N, D, 1.0, 7, 0)
    T = model.T
    runtime.sample_loop(model.dynamics, pred, nm, em, z, T, T, T - 3, sched, tvals, w, None, 7, None, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    runtime.sample_loop(model.dynamics, pred, nm, em, z, T, T - 3, T - 3 - steps, sched, tvals, w, None, 7, None, False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"config": 3, "workload": "guided, PASs shape (10 rings + 10 orientation nodes, N 20, F 12, 110 edges), "
                      "target ip+ea+3*gap on un-normalised predictions, scale 0.6", "batch_per_gpu": batch,
                      "ms_per_step": ms, "molecules_per_s_per_gpu": batch / (T * ms * 1e-3)}))


def config4(dev):
    """Batches whose saved activations do not fit one GPU are run as sequential chunks (what the sampler's _chunked does);
    the reported time is the sum over the chunks."""
    for hidden in (192, 256):
        for n in (4, 10, 20):
            for batch in (1000, 10000, 100000, 1000000):
                per_mol = n * (n - 1) * hidden * 12 * 3 * 4 * 1.2
                chunks = 1
                while batch / chunks * per_mol > 120e9:
                    chunks *= 2 if chunks < 8 else 1.25
                    chunks = int(chunks + 0.999)
                if chunks > 16 or (batch == 1000000 and (n > 4 or hidden > 192)):      # keep the sweep within minutes
                    continue
                b = (batch + chunks - 1) // chunks
                a, model, pred, nodes_dist, prop = models("cata", dev, nf_pred=hidden)
                nm, em = gb.build_masks(torch.full((b,), n), n, False, device=dev)
                z = runtime.noise(nm.reshape(-1).contiguous(), b, n, 4, 1.0, 3, 0)
                t = torch.full((1,), 0.5, device=dev)
                w = torch.tensor([0., -1., 0., 0., 0.], device=dev)
                fwd = timed(lambda: pred(z, nm, em, t)) * chunks
                both = timed(lambda: runtime.predictor_value_and_grad(pred, z, nm, em, t, w)) * chunks
                ev = n * (n - 1)
                fl = 2 * (12 * (ev * ((2 * hidden + 2) * hidden + 2 * hidden * hidden + 2 * hidden) + n * 3 * hidden * hidden)) * b * chunks
                print(json.dumps({"config": 4, "hidden": hidden, "nodes": n, "batch": b * chunks, "chunks": chunks, "fwd_ms": fwd,
                                  "fwd_grad_ms": both, "algorithmic_tflops_fwd": fl / fwd * 1e-9,
                                  "algorithmic_tflops_fwd_grad": 2 * fl / both * 1e-9}))
                runtime.release_workspaces()
                torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    if args.config == 1:
        config1(dev, args.batch or 64)
    elif args.config == 3:
        config3(dev, args.batch or 12500, args.steps)
    elif args.config == 4:
        config4(dev)
