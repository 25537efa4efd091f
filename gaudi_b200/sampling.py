"""Sampling helpers and model factories with GaUDI's signatures.

Mirrors ``sample_guidance`` / ``sample_pos_edm`` / ``node2edge_mask`` (sampling_edm.py:119-224), ``get_model`` /
``MyDataParallel`` / ``DistributionRings`` (models_edm.py:13-104), ``get_cond_predictor_model``
(cond_prediction/train_cond_predictor.py:183-203), ``switch_grad_off`` (utils/helpers.py:198-202) and the
``args_edm`` / ``prediction_args`` defaults (utils/args_edm.py:4-51, cond_prediction/prediction_args.py:4-51).
"""
from __future__ import annotations

from argparse import Namespace
from typing import Optional

import numpy as np
import torch
from torch import nn

from .diffusion import AffineTarget, EnVariationalDiffusion
from .egnn import EGNN_dynamics
from .graph import build_masks, node2edge_mask  # noqa: F401  (re-exported)
from .predictor import EGNN_predictor


def args_edm(**over) -> Namespace:
    """Defaults of utils/args_edm.py (the denoiser architecture the benchmarks use)."""
    a = Namespace(dataset="cata", rings_graph=True, max_nodes=11, name="cata-test", restore=None, lr=1e-3,
                  num_epochs=1000, normalize=True, num_workers=32, batch_size=256, sample_rate=1, dp=True,
                  clip_grad=True, n_layers=9, nf=192, tanh=True, attention=True, coords_range=4.0, norm_constant=1.0,
                  sin_embedding=False, inv_sublayers=1, normalization_factor=1.0, aggregation_method="sum",
                  diffusion_steps=1000, diffusion_noise_schedule="polynomial_2", diffusion_noise_precision=1e-5,
                  diffusion_loss_type="l2", normalize_factors=[3, 4, 10], save_dir="summary/", device="cuda",
                  exp_dir=None)
    a.__dict__.update(over)
    return a


def prediction_args(**over) -> Namespace:
    """Defaults of cond_prediction/prediction_args.py (the predictor architecture)."""
    a = Namespace(dataset="cata", rings_graph=True, max_nodes=11,
                  target_features="LUMO_eV,GAP_eV,Erel_eV,aIP_eV,aEA_eV", name="cata-test", restore=None, lr=6e-4,
                  num_epochs=1000, normalize=True, batch_size=256, sample_rate=1.0, num_workers=32, dp=True,
                  n_layers=12, nf=196, tanh=True, attention=True, coords_range=4.0, norm_constant=1.0,
                  normalization_factor=1.0, save_dir="prediction_summary/", device="cuda", exp_dir=None)
    a.__dict__.update(over)
    return a


# ring-count histograms of the two datasets (utils/helpers.py:64-95), used only to draw synthetic sizes
RING_COUNTS = {
    "cata": {11: 20559, 10: 5164, 9: 1349, 8: 363, 7: 108, 5: 11, 6: 32, 3: 2, 4: 3, 1: 1, 2: 1},
    "hetro": {10: 56617, 9: 111471, 8: 107610, 7: 66431, 5: 8622, 6: 28604, 4: 1829, 3: 329, 2: 51},
}


class DistributionRings:
    """Categorical over the number of rings (models_edm.py:21-58)."""

    def __init__(self, dataset="cata"):
        hist = RING_COUNTS[dataset]
        self.n_nodes = torch.tensor(list(hist.keys()))
        prob = np.array(list(hist.values()), dtype=np.float64)
        self.prob = torch.from_numpy(prob / prob.sum()).float()
        self.keys = {int(n): i for i, n in enumerate(hist.keys())}
        self.m = torch.distributions.Categorical(torch.tensor(prob / prob.sum()))

    def sample(self, n_samples=1):
        return self.n_nodes[self.m.sample((n_samples,))]

    def log_prob(self, batch_n_nodes):
        idx = torch.tensor([self.keys[int(i)] for i in batch_n_nodes], device=batch_n_nodes.device)
        return torch.log(self.prob + 1e-30).to(batch_n_nodes.device)[idx]


class DistributionProperty:
    """Normalisation constants of the predictor targets (models_edm.py:107-188, only_norm=True part)."""

    def __init__(self, mean, std):
        self.mean = torch.as_tensor(mean, dtype=torch.float32)
        self.std = torch.as_tensor(std, dtype=torch.float32)

    def unnormalize(self, val):
        return val * self.std.to(val.device) + self.mean.to(val.device)

    def normalize(self, val):
        return (val - self.mean.to(val.device)) / self.std.to(val.device)


class MyDataParallel(nn.DataParallel):
    """Attribute-forwarding DataParallel (models_edm.py:13-18): gives checkpoints their ``module.`` prefix.
    Sampling resolves through ``__getattr__`` to the bare module, i.e. one device per process; multi-GPU runs
    shard the batch across processes instead (``gaudi_b200.dist``)."""

    def __init__(self, module, device_ids=None, output_device=None, dim=0):
        if device_ids is None:
            # one device per process: nn.DataParallel would otherwise replicate over every visible GPU in forward(), and
            # the C library keeps per-process state (packed-weight handles, workspaces) that threads must not share
            p = next(module.parameters(), None)
            if p is not None and p.is_cuda:
                device_ids = [p.device.index if p.device.index is not None else torch.cuda.current_device()]
        super().__init__(module, device_ids=device_ids, output_device=output_device, dim=dim)

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)          # no scatter / replicate: multi-GPU runs shard across processes

    def __getattr__(self, name):
        if name == "module":
            return super().__getattr__("module")
        return getattr(self.module, name)


def load_state_dict_flexible(model: nn.Module, state_dict) -> None:
    """Load a checkpoint saved with or without the ``module.`` prefix into a wrapped or bare model."""
    want_prefix = isinstance(model, nn.DataParallel)
    has_prefix = all(k.startswith("module.") for k in state_dict)
    if want_prefix and not has_prefix:
        state_dict = {"module." + k: v for k, v in state_dict.items()}
    elif not want_prefix and has_prefix:
        state_dict = {k[len("module."):]: v for k, v in state_dict.items()}
    model.load_state_dict(state_dict)


def get_model(args, dataloader_train=None, only_norm=True, in_node_nf: Optional[int] = None):
    """(model, nodes_dist, prop_dist) as models_edm.get_model; random init unless args.restore."""
    ds = getattr(dataloader_train, "dataset", None)
    if in_node_nf is None:
        in_node_nf = ds.num_node_features
    prop_dist = DistributionProperty(getattr(ds, "mean", torch.zeros(1)), getattr(ds, "std", torch.ones(1)))
    nodes_dist = DistributionRings(getattr(args, "dataset", "cata"))
    dyn = EGNN_dynamics(in_node_nf=in_node_nf, n_dims=3, device=args.device, hidden_nf=args.nf,
                        act_fn=torch.nn.SiLU(), n_layers=args.n_layers, attention=args.attention, tanh=args.tanh,
                        norm_constant=args.norm_constant, inv_sublayers=args.inv_sublayers,
                        sin_embedding=args.sin_embedding, normalization_factor=args.normalization_factor,
                        aggregation_method=args.aggregation_method, coords_range=args.coords_range,
                        condition_time=True)
    model = EnVariationalDiffusion(dynamics=dyn, in_node_nf=in_node_nf, n_dims=3, timesteps=args.diffusion_steps,
                                   noise_schedule=args.diffusion_noise_schedule,
                                   noise_precision=args.diffusion_noise_precision,
                                   loss_type=args.diffusion_loss_type, norm_values=args.normalize_factors,
                                   include_charges=False, device=args.device)
    if getattr(args, "dp", False):
        model = MyDataParallel(model)
    if getattr(args, "restore", None):
        load_state_dict_flexible(model, torch.load(args.exp_dir + "/model.pt", map_location=args.device))
    return model, nodes_dist, prop_dist


def get_cond_predictor_model(args, dataset):
    pred = EGNN_predictor(in_nf=dataset.num_node_features, device=args.device, hidden_nf=args.nf,
                          out_nf=dataset.num_targets, act_fn=torch.nn.SiLU(), n_layers=args.n_layers, recurrent=True,
                          tanh=args.tanh, attention=args.attention, condition_time=True,
                          coords_range=args.coords_range)
    if getattr(args, "dp", False):
        pred = MyDataParallel(pred)
    if getattr(args, "restore", None):
        load_state_dict_flexible(pred, torch.load(args.exp_dir + "/model.pt", map_location=args.device))
    return pred


def switch_grad_off(models):
    for m in models:
        m.eval()
        for p in m.parameters():
            p.requires_grad = False


def _masks(args, nodesxsample, max_nodes):
    orientation = args.dataset != "cata"
    return build_masks(nodesxsample, max_nodes, orientation, device=args.device)


def _post_asserts(x, node_mask):
    assert float((x * (1 - node_mask)).abs().max().item()) < 1e-4, "Variables not masked properly."
    largest = float(x.abs().max().item())
    err = float(torch.sum(x, dim=1, keepdim=True).abs().max().item())
    assert err / (largest + 1e-10) < 1e-2, f"Mean is not zero, relative_error {err / (largest + 1e-10)}"


def sample_pos_edm(args, model, nodesxsample, std=0.7, noise=None):
    """Unconditional sampling helper (sampling_edm.py:128-169); pads to args.max_nodes."""
    assert int(torch.max(nodesxsample)) <= int(args.max_nodes)
    node_mask, edge_mask = _masks(args, nodesxsample, int(args.max_nodes))
    x, h = model.sample(len(nodesxsample), node_mask.size(1), node_mask, edge_mask, std=std, noise=noise)
    _post_asserts(x, node_mask)
    return x, h["categorical"], node_mask, edge_mask


def sample_guidance(args, model, target_function, nodesxsample, scale=1, std=1.0, noise=None, max_nodes=None):
    """Guided sampling helper (sampling_edm.py:172-224); pads to nodesxsample.max() as the reference does (:177).

    ``max_nodes`` overrides the padded ring count.  The predictor pools with ``mean`` over the PADDED node count
    (edm/egnn_predictor/models.py:456-457), so the guidance gradient depends on the padding: a shard of a larger batch
    must pad to the whole batch's maximum (``dist.sample_guidance_sharded`` passes it) to reproduce the single-process
    result."""
    nmax = int(torch.as_tensor(nodesxsample).max().item())
    if max_nodes is not None:
        if int(max_nodes) < nmax:
            raise ValueError(f"max_nodes={max_nodes} is smaller than the largest molecule ({nmax} rings)")
        nmax = int(max_nodes)
    node_mask, edge_mask = _masks(args, nodesxsample, nmax)
    x, h = model.sample_guidance(len(nodesxsample), target_function, node_mask, edge_mask, scale, fix_noise=False,
                                 std=std, noise=noise)
    _post_asserts(x, node_mask)
    return x, h["categorical"], node_mask, edge_mask


__all__ = ["args_edm", "prediction_args", "DistributionRings", "DistributionProperty", "MyDataParallel",
           "get_model", "get_cond_predictor_model", "switch_grad_off", "sample_pos_edm", "sample_guidance",
           "node2edge_mask", "build_masks", "AffineTarget", "load_state_dict_flexible"]
