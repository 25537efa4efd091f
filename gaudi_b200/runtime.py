"""Host-side glue between the GaUDI-shaped PyTorch modules and the C ABI (``include/gaudi_b200.h``).

PyTorch is plumbing here: it owns device memory, the current CUDA stream and autograd bookkeeping; every
arithmetic op of the hot path is a hand-written sm_100a kernel reached through ``ctypes``.
There is deliberately no CPU / eager fallback: a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
import functools
import weakref
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib, ops
from .graph import Topology, build_topology

_VP = C.c_void_p


def _ptr(t: Optional[torch.Tensor]) -> _VP:
    return _VP(0) if t is None else _VP(t.data_ptr())


def _stream() -> _VP:
    return _VP(torch.cuda.current_stream().cuda_stream)


def _cuda_device_of(args, kwargs) -> Optional[torch.device]:
    """Device of the first CUDA tensor / module parameter among the arguments; all CUDA tensor arguments must share it."""
    dev = None
    for a in list(args) + list(kwargs.values()):
        d = None
        if isinstance(a, torch.Tensor):
            d = a.device if a.is_cuda else None
        elif isinstance(a, torch.nn.Module):
            p = next(a.parameters(), None)
            d = p.device if p is not None and p.is_cuda else None
        if d is None:
            continue
        if dev is None:
            dev = d
        elif d != dev:
            raise RuntimeError(f"gaudi_b200: arguments live on different CUDA devices ({dev} and {d})")
    return dev


def _on_device(fn):
    """Run a C-ABI entry point with the arguments' device current: the library launches on the current device's stream
    and allocates with cudaMalloc, so tensors on cuda:1 while cuda:0 is current (``args.device='cuda:1'`` without
    ``torch.cuda.set_device``, legal in the reference) would otherwise meet kernels of the wrong device."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = _cuda_device_of(args, kwargs)
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"gaudi_b200: {what} must be a CUDA tensor (the sm_100a kernels have no CPU fallback)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


# --------------------------------------------------------------------------------------------------
# packed-weights handles
# --------------------------------------------------------------------------------------------------
class NetHandle:
    def __init__(self, handle: int, tensors: List[torch.Tensor], versions: Tuple):
        self.handle = _VP(handle)
        self.versions = versions
        self._keep = tensors

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().gb_net_destroy(self.handle)
        except Exception:
            pass


def _param_list_denoiser(egnn) -> List[torch.Tensor]:
    ps = [egnn.embedding.weight, egnn.embedding.bias, egnn.embedding_out.weight, egnn.embedding_out.bias]
    for b in range(egnn.n_layers):
        blk = getattr(egnn, f"e_block_{b}")
        for s in range(blk.n_layers):
            g = getattr(blk, f"gcl_{s}")
            ps += [g.edge_mlp[0].weight, g.edge_mlp[0].bias, g.edge_mlp[2].weight, g.edge_mlp[2].bias,
                   g.node_mlp[0].weight, g.node_mlp[0].bias, g.node_mlp[2].weight, g.node_mlp[2].bias]
            if g.attention:
                ps += [g.att_mlp[0].weight, g.att_mlp[0].bias]
        e = blk.gcl_equiv
        ps += [e.coord_mlp[0].weight, e.coord_mlp[0].bias, e.coord_mlp[2].weight, e.coord_mlp[2].bias,
               e.coord_mlp[4].weight]
    return ps


def _param_list_predictor(egnn) -> List[torch.Tensor]:
    ps = [egnn.embedding.weight, egnn.embedding.bias, egnn.embedding_out.weight, egnn.embedding_out.bias]
    for l in range(egnn.n_layers):
        g = getattr(egnn, f"gcl_{l}")
        ps += [g.edge_mlp[0].weight, g.edge_mlp[0].bias, g.edge_mlp[2].weight, g.edge_mlp[2].bias,
               g.node_mlp[0].weight, g.node_mlp[0].bias, g.node_mlp[2].weight, g.node_mlp[2].bias,
               g.coord_mlp[0].weight, g.coord_mlp[0].bias, g.coord_mlp[2].weight]
        if g.attention:
            ps += [g.att_mlp[0].weight, g.att_mlp[0].bias]
    return ps


def _versions(ps: List[torch.Tensor]) -> Tuple:
    return tuple((p.data_ptr(), p._version) for p in ps)


@_on_device
def denoiser_handle(dyn) -> NetHandle:
    """Packed-weights handle of an ``EGNN_dynamics`` (re-packed when parameters were reloaded / moved)."""
    ps = _param_list_denoiser(dyn.egnn)
    _need_cuda(ps[0], "denoiser parameters")
    ver = _versions(ps)
    h = dyn.__dict__.get("_gb_handle")
    if h is None or h.versions != ver:
        tens = [_f32c(p) for p in ps]
        arr = (_VP * len(tens))(*[t.data_ptr() for t in tens])
        out = _VP(0)
        hy = dyn.hyper
        _lib.check(_lib.lib().gb_denoiser_create(
            C.byref(out), dyn.in_node_nf - 1, hy["hidden_nf"], hy["n_layers"], hy["inv_sublayers"],
            int(hy["attention"]), int(hy["tanh"]), hy["coords_range"], hy["norm_constant"],
            hy["normalization_factor"], arr, len(tens), _stream()))
        h = NetHandle(out.value, tens, ver)
        dyn.__dict__["_gb_handle"] = h
    return h


@_on_device
def predictor_handle(pred) -> NetHandle:
    ps = _param_list_predictor(pred.egnn)
    _need_cuda(ps[0], "predictor parameters")
    ver = _versions(ps)
    h = pred.__dict__.get("_gb_handle")
    if h is None or h.versions != ver:
        tens = [_f32c(p) for p in ps]
        arr = (_VP * len(tens))(*[t.data_ptr() for t in tens])
        out = _VP(0)
        hy = pred.hyper
        _lib.check(_lib.lib().gb_predictor_create(
            C.byref(out), hy["in_node_nf"] - 1, hy["out_nf"], hy["hidden_nf"], hy["n_layers"], int(hy["attention"]),
            int(hy["tanh"]), hy["coords_range"], arr, len(tens), _stream()))
        h = NetHandle(out.value, tens, ver)
        pred.__dict__["_gb_handle"] = h
    return h


# --------------------------------------------------------------------------------------------------
# graph handles (cached per mask pair)
# --------------------------------------------------------------------------------------------------
class GraphHandle:
    def __init__(self, topo: Topology):
        self.topo = topo
        out = _VP(0)
        # gb_graph_create launches its tile-table kernel on the legacy stream and synchronises the device; under a
        # non-blocking side stream the int32 conversions of build_topology must have finished before it reads them
        torch.cuda.current_stream(topo.rowptr.device).synchronize()
        _lib.check(_lib.lib().gb_graph_create(
            C.byref(out), topo.B, topo.N, topo.n_edges, topo.n_tiles, topo.n_tc, _ptr(topo.rowptr), _ptr(topo.erow),
            _ptr(topo.ecol), _ptr(topo.tile_ptr), _ptr(topo.tc_ptr), _ptr(topo.tc_node), _ptr(topo.tc_start),
            _ptr(topo.cperm), _ptr(topo.colptr), _ptr(topo.cedge), _ptr(topo.node_mask)))
        self.handle = out

    def __del__(self):
        try:
            _lib.lib().gb_graph_destroy(self.handle)
        except Exception:
            pass


_graph_cache: Dict[Tuple, GraphHandle] = {}
_GRAPH_CACHE_MAX = 8


@_on_device
def graph_for(node_mask: torch.Tensor, edge_mask: torch.Tensor, B: int, N: int) -> GraphHandle:
    _need_cuda(node_mask, "node_mask")
    _need_cuda(edge_mask, "edge_mask")
    key = (node_mask.data_ptr(), node_mask._version, edge_mask.data_ptr(), edge_mask._version, B, N,
           node_mask.device.index)
    g = _graph_cache.get(key)
    if g is None:
        if len(_graph_cache) >= _GRAPH_CACHE_MAX:
            _graph_cache.pop(next(iter(_graph_cache)))
        g = GraphHandle(build_topology(node_mask, edge_mask, B, N))
        g._masks = (node_mask, edge_mask)          # keep the keyed storage alive so data_ptr stays unique
        _graph_cache[key] = g
    return g


# --------------------------------------------------------------------------------------------------
# workspaces (grown on demand, one per purpose and device)
# --------------------------------------------------------------------------------------------------
class Workspace:
    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.generation = 0

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = None
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self.buf


_ws: Dict[Tuple[str, int], Workspace] = {}


def workspace(kind: str, device) -> Workspace:
    key = (kind, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _ws:
        _ws[key] = Workspace()
    return _ws[key]


def release_workspaces() -> None:
    _ws.clear()
    _graph_cache.clear()


# --------------------------------------------------------------------------------------------------
# top-level ops
# --------------------------------------------------------------------------------------------------
def _time_tensor(t, B: int, device) -> Tuple[torch.Tensor, int]:
    t = torch.as_tensor(t, dtype=torch.float32, device=device)
    if t.numel() == 1:
        return t.reshape(1).contiguous(), 0
    if t.numel() != B:
        raise ValueError(f"t must have 1 or B={B} elements, got {tuple(t.shape)}")
    return t.reshape(B).contiguous(), 1


@_on_device
def denoiser_forward(dyn, t, xh: torch.Tensor, node_mask: torch.Tensor, edge_mask: torch.Tensor,
                     scrub_all: bool = False, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """EGNN_dynamics._forward (edm/egnn/models.py:83-152)."""
    _need_cuda(xh, "xh")
    B, N, D = xh.shape
    net = denoiser_handle(dyn)
    g = graph_for(node_mask, edge_mask, B, N)
    z = _f32c(xh)
    tt, per_mol = _time_tensor(t, B, xh.device)
    nbytes = _lib.lib().gb_denoiser_workspace_bytes(net.handle, g.handle)
    ws = workspace("den", xh.device).get(nbytes, xh.device)
    return ops.denoiser_forward(net.handle.value, g.handle.value, z, tt, per_mol, bool(scrub_all), stats, ws)


class _PredictorFn(torch.autograd.Function):
    """pred = EGNN_predictor(xh); backward = hand-written input-gradient kernels (no weight gradients)."""

    @staticmethod
    def forward(ctx, xh, pred_module, node_mask, edge_mask, t):
        B, N, D = xh.shape
        net = predictor_handle(pred_module)
        g = graph_for(node_mask, edge_mask, B, N)
        z = _f32c(xh)
        tt, per_mol = _time_tensor(t, B, xh.device)
        need_grad = bool(ctx.needs_input_grad[0])
        nbytes = _lib.lib().gb_predictor_workspace_bytes(net.handle, g.handle, int(need_grad))
        wsp = workspace("pred_grad" if need_grad else "pred", xh.device)
        ws = wsp.get(nbytes, xh.device)
        out = ops.predictor_forward(net.handle.value, g.handle.value, z, tt, per_mol, pred_module.hyper["out_nf"], need_grad, ws)
        if need_grad:
            wsp.generation += 1
            ctx.gen = wsp.generation
            ctx.wsp, ctx.net, ctx.g, ctx.shape = wsp, net, g, (B, N, D)
        return out

    @staticmethod
    def backward(ctx, g_pred):
        wsp = ctx.wsp
        if wsp.generation != ctx.gen:
            raise RuntimeError("gaudi_b200: the predictor workspace was overwritten by a later forward; "
                               "call backward before running the predictor again")
        B, N, D = ctx.shape
        gp = _f32c(g_pred)
        with torch.cuda.device(gp.device):
            gz = ops.predictor_input_grad(ctx.net.handle.value, ctx.g.handle.value, gp, False, B, N, D, wsp.buf)
        return gz, None, None, None, None


@_on_device
def predictor_forward(pred_module, xh, node_mask, edge_mask, t) -> torch.Tensor:
    """EGNN_predictor.forward (edm/egnn_predictor/models.py:433-457), differentiable w.r.t. ``xh``."""
    _need_cuda(xh, "xh")
    return _PredictorFn.apply(xh, pred_module, node_mask, edge_mask, t)


@_on_device
def predictor_value_and_grad(pred_module, xh, node_mask, edge_mask, t, g_pred_row: torch.Tensor):
    """(pred, d<g_pred_row, sum_b pred_b>/dxh) without autograd: the affine-target fast path."""
    _need_cuda(xh, "xh")
    B, N, D = xh.shape
    net = predictor_handle(pred_module)
    g = graph_for(node_mask, edge_mask, B, N)
    z = _f32c(xh)
    tt, per_mol = _time_tensor(t, B, xh.device)
    nbytes = _lib.lib().gb_predictor_workspace_bytes(net.handle, g.handle, 1)
    wsp = workspace("pred_grad", xh.device)
    ws = wsp.get(nbytes, xh.device)
    wsp.generation += 1
    out = ops.predictor_forward(net.handle.value, g.handle.value, z, tt, per_mol, pred_module.hyper["out_nf"], True, ws)
    w = _f32c(g_pred_row).reshape(-1)
    gz = ops.predictor_input_grad(net.handle.value, g.handle.value, w, True, B, N, D, ws)
    return out, gz


# ---- step pieces -------------------------------------------------------------------------------------
@_on_device
def step_sample(zt, eps, noise, coef, node_mask_flat, project: bool, seed: int = 0, draw: int = 0) -> torch.Tensor:
    B, N, D = zt.shape
    zs = torch.empty_like(zt)
    _lib.check(_lib.lib().gb_step_sample(_ptr(zt), _ptr(eps), _ptr(noise), _ptr(coef), _ptr(node_mask_flat), B, N, D,
                                         seed, draw, int(project), _ptr(zs), _stream()))
    return zs


@_on_device
def step_guide(zs_pre, grad, coef, node_mask_flat, max_norm: float = 10.0) -> torch.Tensor:
    B, N, D = zs_pre.shape
    zs = torch.empty_like(zs_pre)
    _lib.check(_lib.lib().gb_step_guide(_ptr(zs_pre), _ptr(grad), _ptr(coef), _ptr(node_mask_flat), B, N, D,
                                        float(max_norm), _ptr(zs), _stream()))
    return zs


@_on_device
def decode(z0, eps, noise, coef, node_mask_flat, norm_x, norm_h, bias_h, seed: int = 0, draw: int = 0):
    B, N, D = z0.shape
    x = torch.empty(B, N, 3, dtype=torch.float32, device=z0.device)
    one_hot = torch.empty(B, N, D - 3, dtype=torch.float32, device=z0.device)
    cog = torch.zeros(1, dtype=torch.float32, device=z0.device)
    L = _lib.lib()
    _lib.check(L.gb_decode(_ptr(z0), _ptr(eps), _ptr(noise), _ptr(coef), _ptr(node_mask_flat), B, N, D, seed, draw,
                           float(norm_x), float(norm_h), float(bias_h), _ptr(x), _ptr(one_hot), _ptr(cog), _stream()))
    return x, one_hot, cog


@_on_device
def cog_fix(x, node_mask_flat, cog, thresh: float = 5e-2) -> None:
    B, N, _ = x.shape
    _lib.check(_lib.lib().gb_cog_fix(_ptr(x), _ptr(node_mask_flat), _ptr(cog), float(thresh), B, N, _stream()))


@_on_device
def noise(node_mask_flat, B: int, N: int, D: int, std: float, seed: int, draw: int) -> torch.Tensor:
    out = torch.empty(B, N, D, dtype=torch.float32, device=node_mask_flat.device)
    _lib.check(_lib.lib().gb_noise(_ptr(out), _ptr(node_mask_flat), B, N, D, float(std), seed, draw, _stream()))
    return out


@_on_device
def sample_loop(dyn, pred_module, node_mask, edge_mask, z, T: int, s_hi: int, s_lo: int, sched, tvals,
                target_w: Optional[torch.Tensor], noise_all: Optional[torch.Tensor], seed: int,
                stats: Optional[torch.Tensor], use_graph: bool) -> None:
    """In-place reverse diffusion of ``z`` over steps s_hi-1..s_lo (gb_sample_loop)."""
    B, N, D = z.shape
    den = denoiser_handle(dyn)
    prd = predictor_handle(pred_module) if pred_module is not None else None
    g = graph_for(node_mask, edge_mask, B, N)
    L = _lib.lib()
    nbytes = L.gb_sample_loop_workspace_bytes(den.handle, prd.handle if prd else _VP(0), g.handle)
    ws = workspace("loop", z.device).get(nbytes, z.device)
    _lib.check(L.gb_sample_loop(den.handle, prd.handle if prd else _VP(0), g.handle, _ptr(z), T, s_hi, s_lo,
                                _ptr(sched), _ptr(tvals), _ptr(target_w), _ptr(noise_all), seed, _ptr(stats),
                                _ptr(ws), ws.numel(), int(use_graph), _stream()))


def launch_count(reset: bool = False) -> int:
    return int(_lib.lib().gb_launch_count(int(reset)))


# ---- module-level forwards (sub-module API surface, inference only) ---------------------------------
def _dummy_linear(out_f, in_f, device, bias=True):
    w = torch.zeros(out_f, in_f, dtype=torch.float32, device=device)
    return [w, torch.zeros(out_f, dtype=torch.float32, device=device)] if bias else [w]


def _gcl_params(g):
    ps = [g.edge_mlp[0].weight, g.edge_mlp[0].bias, g.edge_mlp[2].weight, g.edge_mlp[2].bias,
          g.node_mlp[0].weight, g.node_mlp[0].bias, g.node_mlp[2].weight, g.node_mlp[2].bias]
    if g.attention:
        ps += [g.att_mlp[0].weight, g.att_mlp[0].bias]
    return ps


def _equiv_params(e):
    return [e.coord_mlp[0].weight, e.coord_mlp[0].bias, e.coord_mlp[2].weight, e.coord_mlp[2].bias, e.coord_mlp[4].weight]


def _sub_handle(mod, kind: str, ps: List[torch.Tensor], **cfg) -> NetHandle:
    """One-layer network handle built from a stand-alone sub-module's own parameters (cached on the module)."""
    _need_cuda(ps[0], "module parameters")
    ver = _versions(ps)
    h = mod.__dict__.get("_gb_handle")
    if h is None or h.versions != ver:
        tens = [_f32c(p) for p in ps]
        arr = (_VP * len(tens))(*[t.data_ptr() for t in tens])
        out = _VP(0)
        L = _lib.lib()
        if kind == "den":
            _lib.check(L.gb_denoiser_create(C.byref(out), 0, cfg["hidden"], cfg.get("n_layers", 1), cfg["n_sub"],
                                            int(cfg["attention"]), int(cfg["tanh"]), float(cfg["coords_range"]),
                                            float(cfg["norm_constant"]), float(cfg["normf"]), arr, len(tens), _stream()))
        else:
            _lib.check(L.gb_predictor_create(C.byref(out), 0, 1, cfg["hidden"], 1, int(cfg["attention"]),
                                             int(cfg["tanh"]), float(cfg["coords_range"]), arr, len(tens), _stream()))
        h = NetHandle(out.value, tens, ver)
        mod.__dict__["_gb_handle"] = h
    return h


def _flat_graph(h, edge_index, node_mask, edge_mask):
    """Graph handle for flattened [B*N, .] inputs; N is recovered from the dense edge list length."""
    _need_cuda(h, "h")
    n_nodes = h.shape[0]
    row = edge_index[0]
    n = row.numel() // n_nodes
    if n * n_nodes != row.numel():
        raise ValueError("edge_index is not the dense per-graph edge list of get_adj_matrix")
    B = n_nodes // n
    if node_mask is None:
        node_mask = torch.ones(n_nodes, 1, dtype=torch.float32, device=h.device)
    if edge_mask is None:
        edge_mask = torch.ones(row.numel(), 1, dtype=torch.float32, device=h.device)
    g = graph_for(node_mask, edge_mask, B, n)
    return g, B, n


def _den_ws(net, g, device):
    nbytes = _lib.lib().gb_denoiser_workspace_bytes(net.handle, g.handle)
    return workspace("den", device).get(nbytes, device)


def _edge_gather(t, g, width):
    return _f32c(t.reshape(-1, width)[g.topo.dense_idx])


@_on_device
def gcl_forward(mod, h, edge_index, edge_attr, node_mask, edge_mask):
    g, B, n = _flat_graph(h, edge_index, node_mask, edge_mask)
    H = mod.node_mlp[2].out_features
    ps = _dummy_linear(H, 1, h.device) + _dummy_linear(1, H, h.device) + _gcl_params(mod) + \
        _dummy_linear(H, 2 * H + 2, h.device) + _dummy_linear(H, H, h.device) + _dummy_linear(1, H, h.device, bias=False)
    net = _sub_handle(mod, "den", ps, hidden=H, n_sub=1, attention=mod.attention, tanh=False, coords_range=1.0,
                      norm_constant=1.0, normf=mod.normalization_factor)
    hin = _f32c(h)
    out = torch.empty_like(hin)
    ws = _den_ws(net, g, h.device)
    ea = _edge_gather(edge_attr, g, 2)            # keep alive across the call
    _lib.check(_lib.lib().gb_den_gcl_forward(net.handle, g.handle, 0, 0, _ptr(hin), _ptr(ea),
                                             _ptr(out), _ptr(ws), ws.numel(), _stream()))
    return out


@_on_device
def equiv_update_forward(mod, h, coord, edge_index, coord_diff, edge_attr, node_mask, edge_mask):
    g, B, n = _flat_graph(h, edge_index, node_mask, edge_mask)
    H = mod.coord_mlp[2].out_features
    ps = _dummy_linear(H, 1, h.device) + _dummy_linear(1, H, h.device) + _dummy_linear(H, 2 * H + 2, h.device) + \
        _dummy_linear(H, H, h.device) + _dummy_linear(H, 2 * H, h.device) + _dummy_linear(H, H, h.device) + _equiv_params(mod)
    net = _sub_handle(mod, "den", ps, hidden=H, n_sub=1, attention=False, tanh=mod.tanh, coords_range=mod.coords_range,
                      norm_constant=1.0, normf=mod.normalization_factor)
    hin, xin = _f32c(h), _f32c(coord)
    out = torch.empty_like(xin)
    ws = _den_ws(net, g, h.device)
    cd, ea = _edge_gather(coord_diff, g, 3), _edge_gather(edge_attr, g, 2)
    _lib.check(_lib.lib().gb_den_equiv_forward(net.handle, g.handle, 0, _ptr(hin), _ptr(xin), _ptr(cd), _ptr(ea),
                                               _ptr(out), _ptr(ws), ws.numel(), _stream()))
    return out


def _block_params(blk):
    ps = []
    for s in range(blk.n_layers):
        ps += _gcl_params(getattr(blk, f"gcl_{s}"))
    return ps + _equiv_params(blk.gcl_equiv)


@_on_device
def equiv_block_forward(mod, h, x, edge_index, node_mask, edge_mask, edge_attr):
    g, B, n = _flat_graph(h, edge_index, node_mask, edge_mask)
    H = mod.hidden_nf
    ps = _dummy_linear(H, 1, h.device) + _dummy_linear(1, H, h.device) + _block_params(mod)
    net = _sub_handle(mod, "den", ps, hidden=H, n_sub=mod.n_layers, attention=mod.gcl_0.attention, tanh=mod.gcl_equiv.tanh,
                      coords_range=mod.coords_range_layer, norm_constant=mod.norm_constant, normf=mod.normalization_factor)
    hin, xin = _f32c(h), _f32c(x)
    hout, xout = torch.empty_like(hin), torch.empty_like(xin)
    ws = _den_ws(net, g, h.device)
    d0 = _edge_gather(edge_attr, g, 1)
    _lib.check(_lib.lib().gb_den_block_forward(net.handle, g.handle, 0, _ptr(hin), _ptr(xin), _ptr(d0), _ptr(hout),
                                               _ptr(xout), _ptr(ws), ws.numel(), _stream()))
    return hout, xout


@_on_device
def egnn_forward(mod, h, x, edge_index, node_mask, edge_mask):
    g, B, n = _flat_graph(h, edge_index, node_mask, edge_mask)
    ps = _param_list_denoiser(mod)
    _need_cuda(ps[0], "EGNN parameters")
    ver = _versions(ps)
    net = mod.__dict__.get("_gb_handle")
    if net is None or net.versions != ver:
        tens = [_f32c(p) for p in ps]
        arr = (_VP * len(tens))(*[t.data_ptr() for t in tens])
        out = _VP(0)
        b0 = mod.e_block_0
        _lib.check(_lib.lib().gb_denoiser_create(
            C.byref(out), mod.embedding.in_features - 1, mod.hidden_nf, mod.n_layers, b0.n_layers, int(b0.gcl_0.attention),
            int(b0.gcl_equiv.tanh), float(b0.coords_range_layer), float(b0.norm_constant), float(mod.normalization_factor),
            arr, len(tens), _stream()))
        net = NetHandle(out.value, tens, ver)
        mod.__dict__["_gb_handle"] = net
    hin, xin = _f32c(h), _f32c(x)
    hout, xout = torch.empty(hin.shape[0], mod.embedding_out.out_features, dtype=torch.float32, device=h.device), torch.empty_like(xin)
    ws = _den_ws(net, g, h.device)
    _lib.check(_lib.lib().gb_den_egnn_forward(net.handle, g.handle, _ptr(hin), _ptr(xin), _ptr(hout), _ptr(xout), _ptr(ws),
                                              ws.numel(), _stream()))
    return hout, xout


def _egcl_params(g):
    ps = [g.edge_mlp[0].weight, g.edge_mlp[0].bias, g.edge_mlp[2].weight, g.edge_mlp[2].bias,
          g.node_mlp[0].weight, g.node_mlp[0].bias, g.node_mlp[2].weight, g.node_mlp[2].bias,
          g.coord_mlp[0].weight, g.coord_mlp[0].bias, g.coord_mlp[2].weight]
    if g.attention:
        ps += [g.att_mlp[0].weight, g.att_mlp[0].bias]
    return ps


def _pred_ws(net, g, device):
    nbytes = _lib.lib().gb_predictor_workspace_bytes(net.handle, g.handle, 0)
    return workspace("pred", device).get(nbytes, device)


@_on_device
def e_gcl_forward(mod, h, edge_index, coord, edge_attr, node_mask, edge_mask):
    g, B, n = _flat_graph(h, edge_index, node_mask, edge_mask)
    H = mod.node_mlp[2].out_features
    ps = _dummy_linear(H, 1, h.device) + _dummy_linear(1, H, h.device) + _egcl_params(mod)
    net = _sub_handle(mod, "pred", ps, hidden=H, attention=mod.attention, tanh=mod.tanh,
                      coords_range=getattr(mod, "coords_range", 1.0))
    hin, xin = _f32c(h), _f32c(coord)
    hout, xout = torch.empty_like(hin), torch.empty_like(xin)
    ws = _pred_ws(net, g, h.device)
    ea = _edge_gather(edge_attr, g, 1)
    _lib.check(_lib.lib().gb_pred_layer_forward(net.handle, g.handle, 0, _ptr(hin), _ptr(xin), _ptr(ea), _ptr(hout),
                                                _ptr(xout), _ptr(ws), ws.numel(), _stream()))
    return hout, xout


@_on_device
def pred_egnn_forward(mod, h, x, edges, edge_attr, node_mask, edge_mask):
    g, B, n = _flat_graph(h, edges, node_mask, edge_mask)
    ps = _param_list_predictor(mod)
    _need_cuda(ps[0], "EGNN parameters")
    ver = _versions(ps)
    net = mod.__dict__.get("_gb_handle")
    if net is None or net.versions != ver:
        tens = [_f32c(p) for p in ps]
        arr = (_VP * len(tens))(*[t.data_ptr() for t in tens])
        out = _VP(0)
        g0 = mod.gcl_0
        _lib.check(_lib.lib().gb_predictor_create(
            C.byref(out), mod.embedding.in_features - 1, mod.embedding_out.out_features, mod.hidden_nf, mod.n_layers,
            int(g0.attention), int(g0.tanh), float(mod.coords_range_layer * mod.n_layers), arr, len(tens), _stream()))
        net = NetHandle(out.value, tens, ver)
        mod.__dict__["_gb_handle"] = net
    hin, xin = _f32c(h), _f32c(x)
    hout = torch.empty(hin.shape[0], mod.embedding_out.out_features, dtype=torch.float32, device=h.device)
    xout = torch.empty_like(xin)
    ws = _pred_ws(net, g, h.device)
    ea = _edge_gather(edge_attr, g, 1)
    _lib.check(_lib.lib().gb_pred_egnn_forward(net.handle, g.handle, _ptr(hin), _ptr(xin), _ptr(ea),
                                               _ptr(hout), _ptr(xout), _ptr(ws), ws.numel(), _stream()))
    return hout, xout
