"""Development aid: repeated full-size replica runs (the batch of tests/test_gpu_parity.py::test_full_size_batch_replicas...): every copy of the golden
molecules must get a bit-identical raw guidance gradient, in every trial.  This is the check that found the in-flight TMEM-load
register hazard of round 2f (GB_BWD_SPLIT_LD).   python tools/replica_trials.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, numpy as np
import gaudi_b200 as gb
from gaudi_b200 import runtime
from helpers import build_models, golden, product_target
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
for dataset, reps in (("cata", 2000), ("hetro", 4167)):
    g = golden(f"step_{dataset}.npz")
    args, model, pred, prop = build_models(dataset, dev)
    tf = product_target(dataset, pred, prop, True)
    nm0, em0 = torch.from_numpy(g["node_mask"]), torch.from_numpy(g["edge_mask"])
    b0, N = nm0.shape[0], nm0.shape[1]
    B = b0 * reps
    nm = nm0.repeat(reps, 1, 1).to(dev)
    em = em0.view(b0, N * N).repeat(reps, 1).reshape(-1, 1).to(dev)
    t = 500
    zt = torch.from_numpy(g[f"zt_{t}"]).repeat(reps, 1, 1).to(dev)
    noise = torch.from_numpy(g[f"noise_{t}"]).repeat(reps, 1, 1).to(dev)
    s_arr = torch.full((B, 1), t - 1, device=dev) / model.T
    t_arr = torch.full((B, 1), t, device=dev) / model.T
    for trial in range(3):
        out = model.sample_p_zs_given_zt_guidance(s_arr, t_arr, zt, nm, em, tf, float(g["scale"]), noise=noise, return_parts=True)
        v = out["grad_raw"].view(reps, b0, N, -1)
        d = (v - v[:1]).abs()
        bad = (d.amax(dim=(2, 3)) > 0)
        idx = bad.nonzero()
        print(dataset, "trial", trial, "bad (copy, mol) pairs:", idx.shape[0], "max diff", float(d.max()), "first:", idx[:6].tolist(), "nan:", bool(torch.isnan(v).any()))
    runtime.release_workspaces(); torch.cuda.empty_cache()
