"""Shared test helpers: seeded networks (bit-identical to the reference's, see tests/golden/meta.json), goldens."""
import hashlib
import json
import os

import numpy as np
import torch

import gaudi_b200 as gb
import gaudi_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
META = json.load(open(os.path.join(GOLD, "meta.json")))


def golden(name):
    return np.load(os.path.join(GOLD, name))


def n_feat(dataset):
    return 1 if dataset == "cata" else 12


def build_models(dataset, device="cpu", hidden=None, layers=None, timesteps=None):
    """(EnVariationalDiffusion, EGNN_predictor, prop_dist) seeded like tests/golden/make_golden.py."""
    F = n_feat(dataset)
    a = gb.args_edm(dataset=dataset, device="cpu", dp=False, max_nodes=11 if dataset == "cata" else 10)
    p = gb.prediction_args(device="cpu", dp=False)
    if hidden:
        a.nf, p.nf = hidden
    if layers:
        a.n_layers, p.n_layers = layers
    if timesteps:
        a.diffusion_steps = timesteps
    from argparse import Namespace
    ds = Namespace(num_node_features=F, num_targets=5, mean=torch.tensor(META["prop_mean"]), std=torch.tensor(META["prop_std"]))
    torch.manual_seed(META["seed_denoiser"])
    model, _, prop = gb.get_model(a, Namespace(dataset=ds), only_norm=True)
    torch.manual_seed(META["seed_predictor"])
    pred = gb.get_cond_predictor_model(p, ds)
    gb.switch_grad_off([model, pred])
    a.device = p.device = device
    return a, model.to(device), pred.to(device), prop


def digest(sd):
    return {k: hashlib.sha256(v.detach().cpu().contiguous().numpy().tobytes()).hexdigest() for k, v in sd.items()}


def oracle_cfgs(dataset, hidden=None, layers=None):
    F = n_feat(dataset)
    d = O.DenoiserCfg(in_node_nf=F)
    p = O.PredictorCfg(in_node_nf=F)
    if hidden:
        d.hidden_nf, p.hidden_nf = hidden
    if layers:
        d.n_layers, p.n_layers = layers
    return d, p


def oracle_target(dataset):
    if dataset == "cata":
        return O.target_max_gap
    return O.make_target_opv(torch.tensor(META["prop_mean"]), torch.tensor(META["prop_std"]))


def product_target(dataset, pred, prop, affine=True):
    if affine:
        return gb.AffineTarget.max_gap(pred) if dataset == "cata" else gb.AffineTarget.opv(pred, prop)
    if dataset == "cata":
        return lambda z, nm, em, t: -pred(z, nm, em, t)[:, 1]

    def opv(z, nm, em, t):
        q = prop.unnormalize(pred(z, nm, em, t))
        return q[:, 3] + q[:, 2] + 3 * q[:, 0]
    return opv


def cpu_weights(model, pred):
    wd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    wp = {k: v.detach().cpu() for k, v in pred.state_dict().items()}
    return wd, wp


def maxabs(a, b):
    return float((torch.as_tensor(a).detach().cpu().double() - torch.as_tensor(b).detach().cpu().double()).abs().max())
