"""Geometric validity of generated graphs of rings on the GPU.

Mirrors the reference's interface -- ``positions2adj`` (utils/helpers.py:172-196), ``check_stability`` and
``analyze_validity_for_molecules`` (analyze/analyze.py:50-100, 138-177) -- and adds the batched entry point the sampler's
output wants (``check_stability_batch``: padded ``x [B,N,3]``, one-hot ring types, node mask -> five flags per molecule).
All arithmetic runs in ``gb_check_stability`` / ``gb_positions2adj`` (csrc/validity.cu); the rdkit-based chemistry check
(``analyze_rdkit_validity_for_molecules``) is outside this package.

The ring statistics (distance ranges per ring pair, angle quantiles) are data constants of the reference, extracted by
``tools/extract_ring_tables.py`` into ``ring_tables.json``.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .runtime import _ptr, _stream, _need_cuda, _on_device

FLAG_NAMES = ("orientation_nodes", "dist_stable", "connected", "angels3", "angels4")
_MAXRANGE = 4
_F = C.c_float

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ring_tables.json")) as _f:
    RING_TABLES = json.load(_f)
RING_TABLES["peri"] = RING_TABLES["cata"]
RINGS_LIST = {k: v["rings"] for k, v in RING_TABLES.items()}
ring_distances = {k: {p: tuple(r) for p, r in v["distances"].items()} for k, v in RING_TABLES.items()}
angels3_dict = {k: v["angels3"] for k, v in RING_TABLES.items()}
angels4_dict = {k: v["angels4"] for k, v in RING_TABLES.items()}


class _Tables:
    """Thresholds of one (dataset, tol) pair in the layout of ``gb_check_stability``, on one device."""

    def __init__(self, dataset: str, tol: float, device):
        tab = RING_TABLES[dataset]
        rings = tab["rings"]
        T = len(rings)
        lo = np.full((T, T), np.inf, dtype=np.float64)
        hi = np.full((T, T), -np.inf, dtype=np.float64)
        for i, si in enumerate(rings):                       # key lookup order of helpers.py:183-188
            for j, sj in enumerate(rings):
                key = f"{si}-{sj}"
                if key not in tab["distances"]:
                    key = f"{sj}-{si}"
                if key in tab["distances"]:
                    lo[i, j] = tab["distances"][key][0] * (1 - tol)
                    hi[i, j] = tab["distances"][key][1] * (1 + tol)
        a3_lo = np.zeros((T, _MAXRANGE), dtype=np.float64)
        a3_hi = np.zeros((T, _MAXRANGE), dtype=np.float64)
        a3_cnt = np.full((T,), -1, dtype=np.int32)
        for i, s in enumerate(rings):
            if s in tab["angels3"]:
                ranges = tab["angels3"][s]
                assert len(ranges) <= _MAXRANGE
                a3_cnt[i] = len(ranges)
                for q, (ql, qh) in enumerate(ranges):
                    a3_lo[i, q], a3_hi[i, q] = ql * (1 - tol), qh * (1 + tol)
        self.n_types = T
        self.orientation_type = T - 1 if dataset == "hetro" else -1        # analyze.py:67
        # python doubles are compared against fp32 tensors in the reference, i.e. rounded to fp32 first
        f32 = lambda a: torch.from_numpy(a.astype(np.float32)).to(device).contiguous()
        self.pair_lo, self.pair_hi, self.a3_lo, self.a3_hi = f32(lo), f32(hi), f32(a3_lo), f32(a3_hi)
        self.a3_cnt = torch.from_numpy(a3_cnt).to(device)
        self.min_dist = float(np.float32(min(r[0] for r in tab["distances"].values()) * (1 - tol)))
        self.a4_hi = float(np.float32(tab["angels4"]["180"] * (1 - tol)))
        self.a4_lo = float(np.float32(tab["angels4"]["0"] * (1 + tol)))
        self.check_a4 = int(dataset != "hetro")                            # analyze.py:40-41
        self.rings = rings


_tables: Dict[Tuple, _Tables] = {}


def _get_tables(dataset: str, tol: float, device) -> _Tables:
    key = (dataset, float(tol), str(device))
    if key not in _tables:
        _tables[key] = _Tables(dataset, tol, device)
    return _tables[key]


def _ring_index(ring_type: torch.Tensor) -> torch.Tensor:
    if ring_type.dim() >= 2 and ring_type.dtype.is_floating_point:
        ring_type = ring_type.argmax(-1)
    return ring_type.to(torch.int32).contiguous()


@_on_device
def positions2adj(x: torch.Tensor, ring_type: torch.Tensor, tol: float = 0.1, dataset: str = "cata"):
    """(dist [B,N,N], adj [B,N,N]) -- utils/helpers.py:172-196.  ``ring_type`` [B,N] indices or [B,N,F] one-hot."""
    _need_cuda(x, "x")
    B, N, _ = x.shape
    rt = ring_type.argmax(2) if ring_type.dim() == 3 else ring_type
    rt = rt.to(device=x.device, dtype=torch.int32).contiguous()
    t = _get_tables(dataset, tol, x.device)
    xs = x.detach().to(torch.float32).contiguous()
    dist = torch.empty(B, N, N, dtype=torch.float32, device=x.device)
    adj = torch.empty_like(dist)
    _lib.check(_lib.lib().gb_positions2adj(_ptr(xs), _ptr(rt), B, N, t.n_types, _ptr(t.pair_lo), _ptr(t.pair_hi), _ptr(dist),
                                           _ptr(adj), _stream()))
    return dist, adj


@_on_device
def check_stability_batch(x: torch.Tensor, ring_type: torch.Tensor, node_mask: torch.Tensor, tol: float = 0.1,
                          dataset: str = "cata") -> torch.Tensor:
    """flags uint8 [B,8] for a padded batch: columns 0-4 = ``FLAG_NAMES``, 5 = molecule stable (all five), 6 = error bits,
    7 = number of rings.  Equivalent to ``check_stability(x[b][mask_b], ring_type[b][mask_b])`` per molecule."""
    _need_cuda(x, "x")
    B, N, _ = x.shape
    rt = _ring_index(ring_type).to(x.device)
    nm = node_mask.detach().reshape(B, N).to(device=x.device, dtype=torch.float32).contiguous()
    t = _get_tables(dataset, tol, x.device)
    xs = x.detach().to(torch.float32).contiguous()
    flags = torch.empty(B, 8, dtype=torch.uint8, device=x.device)
    _lib.check(_lib.lib().gb_check_stability(_ptr(xs), _ptr(rt), _ptr(nm), B, N, t.n_types, t.orientation_type, _ptr(t.pair_lo),
                                             _ptr(t.pair_hi), _F(t.min_dist), _ptr(t.a3_lo), _ptr(t.a3_hi), _ptr(t.a3_cnt),
                                             _F(t.a4_hi), _F(t.a4_lo), t.check_a4, _ptr(flags), _stream()))
    return flags


def _raise_on_errors(flags: torch.Tensor, dataset: str) -> None:
    err = flags[:, 6]
    if int(err.max()) == 0:
        return
    if int((err & 1).max()):
        raise ValueError("check_stability: a molecule has no ring nodes (networkx raises on the null graph)")
    if int((err & 2).max()):
        raise ValueError("check_stability: more than 16 rings per molecule are not supported")
    raise KeyError("check_stability: a ring symbol without angle statistics is the centre of a triplet "
                   f"(angels3_dict['{dataset}'])")


def validity_summary(flags: torch.Tensor) -> dict:
    """The fractions ``analyze_validity_for_molecules`` reports (analyze.py:164-172) from a flag tensor."""
    n = float(flags.shape[0])
    s = flags[:, :6].to(torch.float64).sum(0).cpu().tolist()
    return {"mol_stable": s[5] / n, "orientation_nodes": s[0] / n, "dist_stable": s[1] / n, "connected": s[2] / n,
            "angels3": s[3] / n, "angels4": s[4] / n, "molecule_stable_bool": flags[:, 5].bool().cpu().tolist()}


def eval_geometric_stability(x, one_hot, node_mask, tol: float = 0.1, dataset: str = "cata"):
    """Batched form for sampler output: (validity_dict, x_stable, one_hot_stable, node_mask_stable)."""
    flags = check_stability_batch(x, one_hot, node_mask, tol, dataset)
    _raise_on_errors(flags, dataset)
    keep = flags[:, 5].bool()
    return validity_summary(flags), x[keep], one_hot[keep], node_mask[keep]


def check_stability(positions, ring_type, tol: float = 0.1, dataset: str = "cata", device=None) -> dict:
    """One molecule, un-padded: ``positions`` [n,3], ``ring_type`` [n] or one-hot [n,F] -- analyze.py:50-100."""
    flags = _flags_for_list([(positions, ring_type)], tol, dataset, device)
    _raise_on_errors(flags, dataset)
    return {k: bool(flags[0, i]) for i, k in enumerate(FLAG_NAMES)}


def _flags_for_list(molecule_list: Sequence, tol: float, dataset: str, device=None) -> torch.Tensor:
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    B = len(molecule_list)
    N = max(1, max(int(np.asarray(p).shape[0]) if not torch.is_tensor(p) else int(p.shape[0]) for p, _ in molecule_list))
    xs = torch.zeros(B, N, 3, dtype=torch.float32)
    rt = torch.zeros(B, N, dtype=torch.int32)
    nm = torch.zeros(B, N, dtype=torch.float32)
    for b, (p, r) in enumerate(molecule_list):
        p = torch.as_tensor(np.asarray(p) if not torch.is_tensor(p) else p, dtype=torch.float32).cpu()
        r = torch.as_tensor(np.asarray(r) if not torch.is_tensor(r) else r).cpu()
        assert p.dim() == 2 and p.shape[1] == 3
        if r.dim() == 2:
            r = r.argmax(1)
        n = p.shape[0]
        xs[b, :n], rt[b, :n], nm[b, :n] = p, r.to(torch.int32), 1.0
    # un-padded molecules are laid out contiguously: valid nodes first (for hetro: rings then orientation nodes)
    return check_stability_batch(xs.to(device), rt.to(device), nm.to(device), tol, dataset)


def analyze_validity_for_molecules(molecule_list: List, tol: float = 0.1, dataset: str = "cata", device=None):
    """(validity_dict, molecule_stable_list) -- analyze.py:138-177, one kernel launch for the whole list."""
    flags = _flags_for_list(molecule_list, tol, dataset, device)
    _raise_on_errors(flags, dataset)
    out = validity_summary(flags)
    stable = [m for m, ok in zip(molecule_list, out["molecule_stable_bool"]) if ok]
    return out, stable
