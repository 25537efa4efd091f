"""``torch.library`` registration of the hot-path entry points (BASELINE.json north_star: "PyTorch custom ops calling a thin
C-ABI layer").

Each op is a one-to-one shim over a C-ABI function of ``include/gaudi_b200.h``: tensors in, tensors out, the opaque
``gb_net*`` / ``gb_graph*`` handles travel as 64-bit integers, the workspace is an explicit (mutated) tensor argument, the
launch goes to PyTorch's current CUDA stream.  The ops are what ``gaudi_b200.runtime`` calls, so module forwards
(``EGNN_dynamics._forward``, ``EGNN_predictor.forward``) dispatch through ``torch.ops.gaudi_b200.*`` and show up by name in
the PyTorch profiler and dispatcher; fake (meta) implementations give shape propagation without a device.  There is no CPU
kernel behind them: a CPU tensor raises.

    gaudi_b200::denoiser_forward     <- gb_denoiser_forward      (EGNN_dynamics._forward, edm/egnn/models.py:83-152)
    gaudi_b200::predictor_forward    <- gb_predictor_forward     (EGNN_predictor.forward, edm/egnn_predictor/models.py:433-457)
    gaudi_b200::predictor_input_grad <- gb_predictor_input_grad  (autograd.grad of the guidance term, en_diffusion.py:899-903)
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

_VP = C.c_void_p


def _ptr(t: Optional[torch.Tensor]) -> _VP:
    return _VP(0) if t is None else _VP(t.data_ptr())


def _stream() -> _VP:
    return _VP(torch.cuda.current_stream().cuda_stream)


def _cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"gaudi_b200: {what} must be a CUDA tensor (the sm_100a kernels have no CPU fallback)")


@torch.library.custom_op("gaudi_b200::denoiser_forward", mutates_args=("stats", "ws"))
def denoiser_forward(net: int, graph: int, z: torch.Tensor, t: torch.Tensor, t_per_mol: int, scrub_all: bool,
                     stats: Optional[torch.Tensor], ws: torch.Tensor) -> torch.Tensor:
    _cuda(z, "z")
    eps = torch.empty_like(z)
    _lib.check(_lib.lib().gb_denoiser_forward(_VP(net), _VP(graph), _ptr(z), _ptr(t), t_per_mol, _ptr(eps), int(scrub_all),
                                              _ptr(stats), _ptr(ws), ws.numel(), _stream()))
    return eps


@denoiser_forward.register_fake
def _(net, graph, z, t, t_per_mol, scrub_all, stats, ws):
    return torch.empty_like(z)


@torch.library.custom_op("gaudi_b200::predictor_forward", mutates_args=("ws",))
def predictor_forward(net: int, graph: int, z: torch.Tensor, t: torch.Tensor, t_per_mol: int, out_nf: int, save_for_grad: bool,
                      ws: torch.Tensor) -> torch.Tensor:
    _cuda(z, "z")
    out = torch.empty(z.shape[0], out_nf, dtype=torch.float32, device=z.device)
    _lib.check(_lib.lib().gb_predictor_forward(_VP(net), _VP(graph), _ptr(z), _ptr(t), t_per_mol, _ptr(out), int(save_for_grad),
                                               _ptr(ws), ws.numel(), _stream()))
    return out


@predictor_forward.register_fake
def _(net, graph, z, t, t_per_mol, out_nf, save_for_grad, ws):
    return z.new_empty((z.shape[0], out_nf), dtype=torch.float32)


@torch.library.custom_op("gaudi_b200::predictor_input_grad", mutates_args=("ws",))
def predictor_input_grad(net: int, graph: int, g_pred: torch.Tensor, broadcast: bool, B: int, N: int, D: int,
                         ws: torch.Tensor) -> torch.Tensor:
    """dL/dz from dL/dpred ([B, out] or, with ``broadcast``, one [out] row shared by every molecule) using the activations the
    last ``predictor_forward(save_for_grad=True)`` left in ``ws``."""
    _cuda(g_pred, "g_pred")
    gz = torch.empty(B, N, D, dtype=torch.float32, device=g_pred.device)
    _lib.check(_lib.lib().gb_predictor_input_grad(_VP(net), _VP(graph), _ptr(g_pred), int(broadcast), _ptr(gz), _ptr(ws),
                                                  ws.numel(), _stream()))
    return gz


@predictor_input_grad.register_fake
def _(net, graph, g_pred, broadcast, B, N, D, ws):
    return g_pred.new_empty((B, N, D), dtype=torch.float32)
