"""Training step of the EDM denoiser (BASELINE config 5).

Replaces torch autograd over ``EnVariationalDiffusion.forward`` (edm/equivariant_diffusion/en_diffusion.py:777-804,
644-775) -> ``EGNN_dynamics._forward`` (edm/egnn/models.py:76-152) -> ``EquivariantBlock`` (edm/egnn/egnn_new.py:42-235)
as driven by ``train_edm.compute_loss`` / ``train_epoch`` (train_edm.py:36-49, 71-75).

The batch is small (512 molecules), so unlike the sampler this path is unfused: each op below is one or a few
hand-written kernels behind the C ABI (``gb_gemm``, ``gb_edge_pre``, ...; csrc/train_ops.cu) with an explicit backward
that also produces the weight gradients.  ``torch.autograd.Function`` only chains them and accumulates ``.grad``;
``torch.optim`` stays the optimiser, as in the reference.  The message passing runs on the compacted edge list of
``graph.build_topology`` (masked edges contribute exact zeros in the reference).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from . import runtime
from .runtime import _ptr, _stream, _need_cuda, graph_for

_F = C.c_float


def _call(name: str, *args) -> None:
    _lib.check(getattr(_lib.lib(), name)(*args, _stream()))


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _gemm(mode: int, M: int, N: int, K: int, A, lda, B, ldb, Cm, ldc, bias=None, acc=False) -> None:
    _call("gb_gemm", mode, M, N, K, _ptr(A), lda, _ptr(B), ldb, _ptr(Cm), ldc, _ptr(bias), int(acc))


def _wgrad(G, X, out, M: int, N: int, K: int, colsum: bool = False):
    """out[M,N] (a view into a weight-gradient tensor, row stride out.stride(0)) = G[K,M]^T X[K,N]; with ``colsum`` also
    returns sum_k G[k,:] (the bias gradient of the same Linear; fused into the tensor-core kernel as a column of ones)."""
    if (_TC and M <= 256 and N < 256 and M % 4 == 0 and N % 4 == 0 and G.stride(0) % 4 == 0 and X.stride(0) % 4 == 0
            and G.data_ptr() % 16 == 0 and X.data_ptr() % 16 == 0):
        nbytes = _lib.lib().gb_wgrad_scratch_bytes(M, N)
        key = ("wgrad", G.device.index)
        if key not in _scratch or _scratch[key].numel() < nbytes:
            _scratch[key] = torch.empty(nbytes, dtype=torch.uint8, device=G.device)
        cs = torch.empty(M, dtype=torch.float32, device=G.device) if colsum else None
        _call("gb_wgrad", K, M, N, _ptr(G), G.stride(0), _ptr(X), X.stride(0), _ptr(out), out.stride(0), 0, _ptr(cs),
              _ptr(_scratch[key]), nbytes)
        return cs
    _gemm(2, M, N, K, G, G.stride(0), X, X.stride(0), out, out.stride(0))
    return _colsum(G, K, M) if colsum else None


def _colsum(X, M: int, N: int, w=None) -> torch.Tensor:
    out = torch.empty(N, dtype=torch.float32, device=X.device)
    _call("gb_colsum", _ptr(X), N, M, N, _ptr(w), _ptr(out), 0)
    return out


def _new(ref: torch.Tensor, *shape) -> torch.Tensor:
    return torch.empty(*shape, dtype=torch.float32, device=ref.device)


# GAUDI_B200_TRAIN_GEMM=fp32 keeps every Linear on the FP32 CUDA-core GEMM (gb_gemm); the default runs the hidden-width
# Linears (forward and dgrad) on the sampler's tcgen05 / 3xTF32 node-Linear kernel with its fused epilogues (gb_linear).
_TC = os.environ.get("GAUDI_B200_TRAIN_GEMM", "tc") != "fp32"
_scratch = {}
EPI_NONE, EPI_SILU, EPI_RES_MASK, EPI_MUL_DSILU, EPI_ADD = 0, 1, 2, 3, 4


def _linear(A1, W, K1, N, bias=None, A2=None, K2=0, transpose=False, epi=EPI_NONE, out2=None, aux=None, mask=None):
    """out = epi([A1 | A2] op(W) + bias); W may be a column/row view of a Linear weight (row stride W.stride(0))."""
    M, ldw = A1.shape[0], W.stride(0)
    out = _new(A1, M, N)
    if _TC and N <= 256 and N % 4 == 0 and K1 % 4 == 0 and K2 % 4 == 0:
        L = _lib.lib()
        nbytes = L.gb_linear_scratch_bytes(N, K1, K2)
        key = A1.device.index
        if key not in _scratch or _scratch[key].numel() < nbytes:
            _scratch[key] = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=A1.device)
        _call("gb_linear", M, N, K1, K2, _ptr(A1), A1.stride(0), _ptr(A2), 0 if A2 is None else A2.stride(0), _ptr(W), ldw,
              int(transpose), _ptr(bias), epi, _ptr(out), _ptr(out2), _ptr(aux), _ptr(mask), _ptr(_scratch[key]), nbytes)
        return out
    if epi == EPI_ADD:
        out.copy_(aux)
    _gemm(1 if transpose else 0, M, N, K1, A1, A1.stride(0), W, ldw, out, N, bias, acc=epi == EPI_ADD)
    if A2 is not None:
        _gemm(0, M, N, K2, A2, A2.stride(0), W[:, K1:], ldw, out, N, acc=True)
    if epi == EPI_SILU:
        if out2 is not None:
            out2.copy_(out)
        _call("gb_silu_fwd", _ptr(out), _ptr(out), M * N)
    elif epi == EPI_RES_MASK:
        _call("gb_resmask", _ptr(out), _ptr(aux), _ptr(mask), M, N, _ptr(out))
    elif epi == EPI_MUL_DSILU:
        _call("gb_silu_bwd", _ptr(aux), _ptr(out), _ptr(out), M * N)
    return out


# ------------------------------------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    """y = act(x W^T + b) [* mask]  -- nn.Linear, optionally followed by SiLU or the node mask of egnn_new.py:318."""

    @staticmethod
    def forward(ctx, x, W, b, mask, act=False):
        x, W = _c(x), _c(W)
        M, K = x.shape
        N = W.shape[0]
        pre = None
        if act:
            pre = _new(x, M, N)
            y = _linear(x, W, K, N, b, epi=EPI_SILU, out2=pre)
        else:
            y = _new(x, M, N)
            _gemm(0, M, N, K, x, K, W, K, y, N, b)
            if mask is not None:
                _call("gb_resmask", _ptr(y), None, _ptr(mask), M, N, _ptr(y))
        ctx.save_for_backward(x, W, mask, pre)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W, mask, pre = ctx.saved_tensors
        M, K = x.shape
        N = W.shape[0]
        gy = _c(gy)
        if mask is not None:
            gm = _new(gy, M, N)
            _call("gb_resmask", _ptr(gy), None, _ptr(mask), M, N, _ptr(gm))
            gy = gm
        if pre is not None:
            gp = _new(gy, M, N)
            _call("gb_silu_bwd", _ptr(pre), _ptr(gy), _ptr(gp), M * N)
            gy = gp
        gx = None
        if ctx.needs_input_grad[0]:
            gx = _linear(gy, W, N, K, transpose=True)
        gW = _new(gy, N, K)
        gb = _wgrad(gy, x, gW, N, K, M, colsum=ctx.has_bias)
        return gx, gW, gb, None, None


class _Geometry(torch.autograd.Function):
    """coord2diff (egnn_new.py:394-400) on the compacted edges: r [E], u [E,3]."""

    @staticmethod
    def forward(ctx, x, g, norm_constant):
        x = _c(x)
        E = g.topo.n_edges
        r, u = _new(x, E), _new(x, E, 3)
        _call("gb_geom_fwd", g.handle, _ptr(x), _F(norm_constant), _ptr(r), _ptr(u))
        ctx.save_for_backward(x)
        ctx.g, ctx.c = g, norm_constant
        return r, u

    @staticmethod
    def backward(ctx, g_r, g_u):
        (x,) = ctx.saved_tensors
        E = ctx.g.topo.n_edges
        g_r = None if g_r is None else _c(g_r)
        g_u = None if g_u is None else _c(g_u)
        scratch, gx = _new(x, E, 3), torch.empty_like(x)
        _call("gb_geom_bwd", ctx.g.handle, _ptr(x), _F(ctx.c), _ptr(g_r), _ptr(g_u), _ptr(scratch), _ptr(gx))
        return gx, None, None


class _EdgeMLP(torch.autograd.Function):
    """m = SiLU(W2 SiLU(W1 [h_i, h_j, r, d0] + b1) + b2) per edge (edge_mlp of GCL, egnn_new.py:43-47, and the first
    two layers of coord_mlp, :110-117).  The first Linear is factorised into per-node projections."""

    @staticmethod
    def forward(ctx, h, r, d0, W1, b1, W2, b2, g):
        h, W1, W2 = _c(h), _c(W1), _c(W2)
        n, H = h.shape
        E = g.topo.n_edges
        Pa = _linear(h, W1, H, H, b1)
        Pb = _linear(h, W1[:, H:], H, H)
        wr, wd = W1[:, 2 * H].contiguous(), W1[:, 2 * H + 1].contiguous()
        pre1, s1, pre2 = _new(h, E, H), _new(h, E, H), _new(h, E, H)
        _call("gb_edge_pre", g.handle, _ptr(Pa), _ptr(Pb), _ptr(r), _ptr(d0), _ptr(wr), _ptr(wd), H, _ptr(pre1), _ptr(s1))
        m = _linear(s1, W2, H, H, b2, epi=EPI_SILU, out2=pre2)
        ctx.save_for_backward(h, r, d0, W1, W2, pre1, pre2)     # SiLU(pre1) is recomputed in backward
        ctx.g = g
        return m

    @staticmethod
    def backward(ctx, g_m):
        h, r, d0, W1, W2, pre1, pre2 = ctx.saved_tensors
        g = ctx.g
        n, H = h.shape
        E, ld1 = g.topo.n_edges, W1.shape[1]
        g_m = _c(g_m)
        G2 = _new(h, E, H)
        _call("gb_silu_bwd", _ptr(pre2), _ptr(g_m), _ptr(G2), E * H)
        s1 = _new(h, E, H)
        _call("gb_silu_fwd", _ptr(pre1), _ptr(s1), E * H)
        gW2 = _new(h, H, H)
        gb2 = _wgrad(G2, s1, gW2, H, H, E, colsum=True)
        del s1
        G1 = _linear(G2, W2, H, H, transpose=True, epi=EPI_MUL_DSILU, aux=pre1)
        gPa, gPb = _new(h, n, H), _new(h, n, H)
        _call("gb_rowcol_reduce", g.handle, _ptr(G1), H, _F(1.0), _ptr(gPa), _ptr(gPb))
        gW1 = _new(h, H, ld1)
        gb1 = _wgrad(gPa, h, gW1, H, H, n, colsum=True)
        _wgrad(gPb, h, gW1[:, H:], H, H, n)
        gW1[:, 2 * H].copy_(_colsum(G1, E, H, r))
        gW1[:, 2 * H + 1].copy_(_colsum(G1, E, H, d0))
        gh = _linear(gPa, W1, H, H, transpose=True)
        gh = _linear(gPb, W1[:, H:], H, H, transpose=True, epi=EPI_ADD, aux=gh)
        g_r = _new(h, E)
        wr = W1[:, 2 * H].contiguous()
        _call("gb_rowdot", _ptr(G1), H, E, H, _ptr(wr), None, _ptr(g_r))
        return gh, g_r, None, gW1, gb1, gW2, gb2, None


class _AttAgg(torch.autograd.Function):
    """agg_i = sum_{j} m_ij * sigmoid(w_a . m_ij + b_a) / normalization_factor (egnn_new.py:49-56, 58-66, 403-419)."""

    @staticmethod
    def forward(ctx, m, wa, ba, g, normf):
        m = _c(m)
        E, H = m.shape
        n = g.topo.B * g.topo.N
        agg = _new(m, n, H)
        if wa is None:
            _call("gb_rowcol_reduce", g.handle, _ptr(m), H, _F(1.0 / normf), _ptr(agg), None)
            ctx.save_for_backward(m)
        else:
            wv = _c(wa.reshape(-1))
            logit, gate, ef = _new(m, E), _new(m, E), _new(m, E, H)
            _call("gb_rowdot", _ptr(m), H, E, H, _ptr(wv), _ptr(ba), _ptr(logit))
            _call("gb_gate_fwd", _ptr(m), _ptr(logit), E, H, _ptr(ef), _ptr(gate))
            _call("gb_rowcol_reduce", g.handle, _ptr(ef), H, _F(1.0 / normf), _ptr(agg), None)
            ctx.save_for_backward(m, wv, gate)
        ctx.g, ctx.normf, ctx.att = g, normf, wa is not None
        return agg

    @staticmethod
    def backward(ctx, g_agg):
        g = ctx.g
        g_agg = _c(g_agg)
        if not ctx.att:
            (m,) = ctx.saved_tensors
            E, H = m.shape
            gm = _new(m, E, H)
            _call("gb_gather_rows", g.handle, _ptr(g_agg), H, _F(1.0 / ctx.normf), _ptr(gm))
            return gm, None, None, None, None
        m, wv, gate = ctx.saved_tensors
        E, H = m.shape
        g_ef = _new(m, E, H)
        _call("gb_gather_rows", g.handle, _ptr(g_agg), H, _F(1.0 / ctx.normf), _ptr(g_ef))
        gm, coef = _new(m, E, H), _new(m, E)
        _call("gb_gate_bwd", _ptr(m), _ptr(gate), _ptr(wv), _ptr(g_ef), E, H, _ptr(gm), _ptr(coef))
        gwa = _colsum(m, E, H, coef).reshape(1, H)
        gba = _colsum(coef, E, 1)
        return gm, gwa, gba, None, None


class _Gate(torch.autograd.Function):
    """ef = m * sigmoid(w_a . m + b_a)  (gcl.py:232-237; the edge mask is implicit in the compacted edge list)."""

    @staticmethod
    def forward(ctx, m, wa, ba):
        m = _c(m)
        E, H = m.shape
        wv = _c(wa.reshape(-1))
        logit, gate, ef = _new(m, E), _new(m, E), _new(m, E, H)
        _call("gb_rowdot", _ptr(m), H, E, H, _ptr(wv), _ptr(ba), _ptr(logit))
        _call("gb_gate_fwd", _ptr(m), _ptr(logit), E, H, _ptr(ef), _ptr(gate))
        ctx.save_for_backward(m, wv, gate)
        return ef

    @staticmethod
    def backward(ctx, g_ef):
        m, wv, gate = ctx.saved_tensors
        E, H = m.shape
        gm, coef = _new(m, E, H), _new(m, E)
        _call("gb_gate_bwd", _ptr(m), _ptr(gate), _ptr(wv), _ptr(_c(g_ef)), E, H, _ptr(gm), _ptr(coef))
        return gm, _colsum(m, E, H, coef).reshape(1, H), _colsum(coef, E, 1)


class _SegSum(torch.autograd.Function):
    """agg_i = scale * sum_{e: row_e = i} ef_e  (unsorted_segment_sum, gcl.py:417-423)."""

    @staticmethod
    def forward(ctx, ef, g, scale):
        ef = _c(ef)
        E, H = ef.shape
        agg = _new(ef, g.topo.B * g.topo.N, H)
        _call("gb_rowcol_reduce", g.handle, _ptr(ef), H, _F(scale), _ptr(agg), None)
        ctx.g, ctx.scale, ctx.shape = g, scale, (E, H)
        return agg

    @staticmethod
    def backward(ctx, g_agg):
        E, H = ctx.shape
        g_ef = _new(g_agg, E, H)
        _call("gb_gather_rows", ctx.g.handle, _ptr(_c(g_agg)), H, _F(ctx.scale), _ptr(g_ef))
        return g_ef, None, None


class _PoolMean(torch.autograd.Function):
    """pred = h.view(B, N, C).mean(1)  (edm/egnn_predictor/models.py:456-457: mean over the padded N)."""

    @staticmethod
    def forward(ctx, h, B, N):
        h = _c(h)
        Cn = h.shape[1]
        pred = _new(h, B, Cn)
        _call("gb_pool_mean", _ptr(h), B, N, Cn, _ptr(pred))
        ctx.dims = (B, N, Cn)
        return pred

    @staticmethod
    def backward(ctx, g_pred):
        B, N, Cn = ctx.dims
        gh = _new(g_pred, B * N, Cn)
        _call("gb_pool_mean_bwd", _ptr(_c(g_pred)), B, N, Cn, _ptr(gh))
        return gh, None, None


class _NodeMLP(torch.autograd.Function):
    """h' = (h + W4 SiLU(W3 [h, agg] + b3) + b4) * mask (node_model, egnn_new.py:58-73, and the mask of :87-88)."""

    @staticmethod
    def forward(ctx, h, agg, W3, b3, W4, b4, mask):
        h, agg, W3, W4 = _c(h), _c(agg), _c(W3), _c(W4)
        n, H = h.shape
        pre = _new(h, n, H)
        sn = _linear(h, W3, H, H, b3, A2=agg, K2=H, epi=EPI_SILU, out2=pre)
        out = _linear(sn, W4, H, H, b4, epi=EPI_RES_MASK, aux=h, mask=mask)
        ctx.save_for_backward(h, agg, W3, W4, pre, sn, mask)
        return out

    @staticmethod
    def backward(ctx, g_out):
        h, agg, W3, W4, pre, sn, mask = ctx.saved_tensors
        n, H = h.shape
        gt = _new(h, n, H)
        _call("gb_resmask", _ptr(_c(g_out)), None, _ptr(mask), n, H, _ptr(gt))
        gW4 = _new(h, H, H)
        gb4 = _wgrad(gt, sn, gW4, H, H, n, colsum=True)
        gpre = _linear(gt, W4, H, H, transpose=True, epi=EPI_MUL_DSILU, aux=pre)
        gW3 = _new(h, H, 2 * H)
        gb3 = _wgrad(gpre, h, gW3, H, H, n, colsum=True)
        _wgrad(gpre, agg, gW3[:, H:], H, H, n)
        gh = _linear(gpre, W3, H, H, transpose=True, epi=EPI_ADD, aux=gt)      # residual branch + gpre W3[:, :H]
        gagg = _linear(gpre, W3[:, H:], H, H, transpose=True)
        return gh, gagg, gW3, gb3, gW4, gb4, None


class _CoordUpdate(torch.autograd.Function):
    """x' = (x + sum_j u_ij * tanh(w7 . c_ij) * range / normalization_factor) * mask (egnn_new.py:122-155)."""

    @staticmethod
    def forward(ctx, x, u, c2, w7, g, rng, use_tanh, normf):
        x, u, c2 = _c(x), _c(u), _c(c2)
        E, H = c2.shape
        wv = _c(w7.reshape(-1))
        phi, tau, xo = _new(x, E), _new(x, E), torch.empty_like(x)
        _call("gb_rowdot", _ptr(c2), H, E, H, _ptr(wv), None, _ptr(phi))
        _call("gb_coord_fwd", g.handle, _ptr(x), _ptr(u), _ptr(phi), _F(rng), int(use_tanh), _F(normf), _ptr(xo), _ptr(tau))
        ctx.save_for_backward(u, c2, wv, tau)
        ctx.g, ctx.cfg = g, (rng, use_tanh, normf)
        return xo

    @staticmethod
    def backward(ctx, g_xo):
        u, c2, wv, tau = ctx.saved_tensors
        g = ctx.g
        rng, use_tanh, normf = ctx.cfg
        E, H = c2.shape
        g_xo = _c(g_xo)
        g_phi, g_u, g_x = _new(u, E), _new(u, E, 3), torch.empty_like(g_xo)
        _call("gb_coord_bwd", g.handle, _ptr(u), _ptr(tau), _ptr(g_xo), _F(rng), int(use_tanh), _F(normf), _ptr(g_phi), _ptr(g_u),
              _ptr(g_x))
        g_c2 = _new(u, E, H)
        _call("gb_outer_dsilu", _ptr(g_phi), _ptr(wv), None, _ptr(g_c2), E, H)
        g_w7 = _colsum(c2, E, H, g_phi).reshape(1, H)
        return g_x, g_u, g_c2, g_w7, None, None, None, None


class _DenFinish(torch.autograd.Function):
    """eps = [remove_mean((x_fin - x_in) * mask), h3[:, :F]] (edm/egnn/models.py:116-152)."""

    @staticmethod
    def forward(ctx, x_fin, h3, x_in, mask, B, N, F):
        eps = _new(x_fin, B, N, 3 + F)
        x_fin, h3 = _c(x_fin), _c(h3)          # locals keep possible contiguous copies alive until the launch is enqueued
        _call("gb_den_finish_fwd", _ptr(x_fin), _ptr(x_in), _ptr(h3), _ptr(mask), B, N, F, _ptr(eps))
        ctx.save_for_backward(mask)
        ctx.dims = (B, N, F)
        return eps

    @staticmethod
    def backward(ctx, g_eps):
        (mask,) = ctx.saved_tensors
        B, N, F = ctx.dims
        gx, gh = _new(mask, B * N, 3), _new(mask, B * N, F + 1)
        _call("gb_den_finish_bwd", _ptr(_c(g_eps)), _ptr(mask), B, N, F, _ptr(gx), _ptr(gh))
        return gx, gh, None, None, None, None, None


class _TrainLoss(torch.autograd.Function):
    """loss [B] of compute_loss(t0_always=False) in train mode with loss_type 'l2' (en_diffusion.py:644-775)."""

    @staticmethod
    def forward(ctx, net, eps, zt, xh, mask, t_int, gamma_t, gamma_T, norm_h, bias_h):
        B, N, D = net.shape
        loss, g_net = _new(net, B), torch.empty_like(net)
        _call("gb_train_loss", _ptr(_c(net)), _ptr(eps), _ptr(zt), _ptr(xh), _ptr(mask), _ptr(t_int), _ptr(gamma_t), _F(gamma_T),
              _F(norm_h), _F(bias_h), B, N, D - 3, _ptr(loss), _ptr(g_net))
        ctx.save_for_backward(g_net)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        (g_net,) = ctx.saved_tensors
        B, N, D = g_net.shape
        out = torch.empty_like(g_net)
        _call("gb_resmask", _ptr(g_net), None, _ptr(_c(g_loss)), B, N * D, _ptr(out))
        return (out,) + (None,) * 9


# ------------------------------------------------------------------------------------------------------------------
@runtime._on_device
def denoiser_forward_train(dyn, t: torch.Tensor, xh: torch.Tensor, node_mask: torch.Tensor, edge_mask: torch.Tensor) -> torch.Tensor:
    """Differentiable (w.r.t. the parameters) ``EGNN_dynamics._forward`` (edm/egnn/models.py:76-152)."""
    _need_cuda(xh, "xh")
    B, N, D = xh.shape
    F = D - 3
    egnn, hy = dyn.egnn, dyn.hyper
    if getattr(egnn, "sin_embedding", None) is not None:
        raise NotImplementedError("sin_embedding is not implemented")
    g = graph_for(node_mask, edge_mask, B, N)
    mask = g.topo.node_mask                                            # [n] fp32
    n = B * N
    z = xh.detach().to(torch.float32).reshape(n, D)
    x_in = _new(z, n, 3)
    _call("gb_resmask", _ptr(_c(z[:, :3])), None, _ptr(mask), n, 3, _ptr(x_in))
    hf = _new(z, n, F)
    _call("gb_resmask", _ptr(_c(z[:, 3:])), None, _ptr(mask), n, F, _ptr(hf))
    tt = torch.as_tensor(t, dtype=torch.float32, device=z.device).reshape(-1)
    tcol = (tt.expand(B) if tt.numel() == 1 else tt).reshape(B, 1).expand(B, N).reshape(n, 1)
    h = _Linear.apply(torch.cat([hf, tcol], dim=1), egnn.embedding.weight, egnn.embedding.bias, None)
    d0 = _new(z, g.topo.n_edges)
    _call("gb_geom_fwd", g.handle, _ptr(x_in), _F(1.0), _ptr(d0), None)  # egnn_new.py:301: the EGNN-level radial uses norm_constant 1
    x = x_in
    normf = hy["normalization_factor"]
    for b in range(egnn.n_layers):
        blk = getattr(egnn, f"e_block_{b}")
        r, u = _Geometry.apply(x, g, float(blk.norm_constant))
        for s in range(blk.n_layers):
            gcl = getattr(blk, f"gcl_{s}")
            m = _EdgeMLP.apply(h, r, d0, gcl.edge_mlp[0].weight, gcl.edge_mlp[0].bias, gcl.edge_mlp[2].weight,
                               gcl.edge_mlp[2].bias, g)
            wa, ba = (gcl.att_mlp[0].weight, gcl.att_mlp[0].bias) if gcl.attention else (None, None)
            agg = _AttAgg.apply(m, wa, ba, g, float(normf))
            h = _NodeMLP.apply(h, agg, gcl.node_mlp[0].weight, gcl.node_mlp[0].bias, gcl.node_mlp[2].weight,
                               gcl.node_mlp[2].bias, mask)
        eq = blk.gcl_equiv
        c2 = _EdgeMLP.apply(h, r, d0, eq.coord_mlp[0].weight, eq.coord_mlp[0].bias, eq.coord_mlp[2].weight,
                            eq.coord_mlp[2].bias, g)
        x = _CoordUpdate.apply(x, u, c2, eq.coord_mlp[4].weight, g, float(eq.coords_range), bool(eq.tanh), float(normf))
    h3 = _Linear.apply(h, egnn.embedding_out.weight, egnn.embedding_out.bias, mask)
    return _DenFinish.apply(x, h3, x_in, mask, B, N, F)


@runtime._on_device
def training_loss(model, x, h, node_mask, edge_mask, t_int: Optional[torch.Tensor] = None,
                  eps: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``EnVariationalDiffusion.forward`` in train mode with loss_type 'l2' and include_charges False: loss [B].

    ``t_int`` [B,1] / ``eps`` [B,N,D] may be injected (tests); otherwise they are drawn like the reference does
    (torch.randint, en_diffusion.py:657-659; centre-of-gravity-free Gaussian noise, :677-679)."""
    _need_cuda(x, "x")
    if model.loss_type != "l2" or model.include_charges:
        raise NotImplementedError("the training step implements loss_type 'l2' without charges (the args_edm defaults)")
    B, N, _ = x.shape
    F = model.in_node_nf
    dev = x.device
    h_cat = h["categorical"] if isinstance(h, dict) else h
    if t_int is None:
        t_int = torch.randint(0, model.T + 1, size=(B, 1), device=dev)
    t_f = t_int.to(torch.float32).reshape(B).contiguous()
    g = graph_for(node_mask, edge_mask, B, N)
    mask = g.topo.node_mask
    if eps is None:
        eps = model.sample_combined_position_feature_noise(B, N, node_mask)
    eps = _c(eps.to(torch.float32))
    gamma = model.gamma.gamma.detach().to(torch.float32).contiguous()
    xh, zt, gamma_t = _new(eps, B, N, 3 + F), _new(eps, B, N, 3 + F), _new(eps, B)
    xs, hs = _c(x.to(torch.float32)), _c(h_cat.to(torch.float32))     # two temporaries in one call must not share a freed block
    _call("gb_make_zt", _ptr(xs), _ptr(hs), _ptr(mask), _ptr(eps), _ptr(gamma), _ptr(t_f),
          _F(model.norm_values[0]), _F(model.norm_values[1]), _F(model.norm_biases[1]), B, N, F, _ptr(xh), _ptr(zt), _ptr(gamma_t))
    net = denoiser_forward_train(model.dynamics, t_f / model.T, zt, node_mask, edge_mask)
    return _TrainLoss.apply(net, eps, zt, xh, mask, t_f, gamma_t, model._gamma_T(), float(model.norm_values[1]),
                            float(model.norm_biases[1]))


# ------------------------------------------------------------------------------------------------------------------
# property predictor: cond_prediction/train_cond_predictor.py:47-81
# ------------------------------------------------------------------------------------------------------------------
@runtime._on_device
def predictor_forward_train(pred_module, xh: torch.Tensor, node_mask: torch.Tensor, edge_mask: torch.Tensor, t) -> torch.Tensor:
    """``EGNN_predictor.forward`` (edm/egnn_predictor/models.py:433-457, 543-560; gcl.py:225-316), differentiable w.r.t. the
    PARAMETERS (the guidance path, differentiable w.r.t. the input, is runtime.predictor_forward)."""
    _need_cuda(xh, "xh")
    B, N, D = xh.shape
    F = D - 3
    egnn = pred_module.egnn
    g = graph_for(node_mask, edge_mask, B, N)
    mask = g.topo.node_mask
    n = B * N
    z = xh.detach().to(torch.float32).reshape(n, D)
    x = _new(z, n, 3)
    _call("gb_resmask", _ptr(_c(z[:, :3])), None, _ptr(mask), n, 3, _ptr(x))
    hf = _new(z, n, F)
    _call("gb_resmask", _ptr(_c(z[:, 3:])), None, _ptr(mask), n, F, _ptr(hf))
    if pred_module.condition_time:
        tt = torch.as_tensor(t, dtype=torch.float32, device=z.device).reshape(-1)
        tcol = (tt.expand(B) if tt.numel() == 1 else tt).reshape(B, 1).expand(B, N).reshape(n, 1)
        hf = torch.cat([hf, tcol], dim=1)
    a = _new(z, g.topo.n_edges)
    _call("gb_geom_fwd", g.handle, _ptr(x), _F(1.0), _ptr(a), None)       # models.py:452: edge_attr = |x_i - x_j|^2 of the input
    h = _Linear.apply(hf, egnn.embedding.weight, egnn.embedding.bias, None)
    for l in range(egnn.n_layers):
        lay = getattr(egnn, f"gcl_{l}")
        r, u = _Geometry.apply(x, g, 1.0)                                 # gcl.py:308-316 (norm_diff: / (sqrt(r + 1e-8) + 1))
        e = _EdgeMLP.apply(h, r, a, lay.edge_mlp[0].weight, lay.edge_mlp[0].bias, lay.edge_mlp[2].weight, lay.edge_mlp[2].bias, g)
        ef = _Gate.apply(e, lay.att_mlp[0].weight, lay.att_mlp[0].bias) if lay.attention else e
        c = _Linear.apply(ef, lay.coord_mlp[0].weight, lay.coord_mlp[0].bias, None, True)
        x_new = _CoordUpdate.apply(x, u, c, lay.coord_mlp[2].weight, g, float(getattr(lay, "coords_range", 1.0)), bool(lay.tanh), 1.0)
        agg = _SegSum.apply(ef, g, 1.0)
        h = _NodeMLP.apply(h, agg, lay.node_mlp[0].weight, lay.node_mlp[0].bias, lay.node_mlp[2].weight, lay.node_mlp[2].bias, mask)
        x = x_new
    hout = _Linear.apply(h, egnn.embedding_out.weight, egnn.embedding_out.bias, mask)
    return _PoolMean.apply(hout, B, N)


@runtime._on_device
def sample_edm_t(x, h, edm_model, t, node_mask, eps: Optional[torch.Tensor] = None) -> torch.Tensor:
    """z_t ~ q(z_t | x, h) at the per-sample times ``t`` [B,1] in [0,1] (cond_prediction/train_cond_predictor.py:47-62)."""
    _need_cuda(x, "x")
    B, N, _ = x.shape
    F = h.shape[2]
    mask = node_mask.detach().reshape(B * N).to(torch.float32).contiguous()
    if eps is None:
        eps = edm_model.sample_combined_position_feature_noise(B, N, node_mask)
    eps = _c(eps.to(torch.float32))
    t_idx = torch.round(t.reshape(B).to(torch.float32) * edm_model.T).contiguous()       # PredefinedNoiseSchedule.forward
    gamma = edm_model.gamma.gamma.detach().to(torch.float32).contiguous()
    xh, zt, gamma_t = _new(eps, B, N, 3 + F), _new(eps, B, N, 3 + F), _new(eps, B)
    xs, hs = _c(x.to(torch.float32)), _c(h.to(torch.float32))
    _call("gb_make_zt", _ptr(xs), _ptr(hs), _ptr(mask), _ptr(eps), _ptr(gamma), _ptr(t_idx),
          _F(edm_model.norm_values[0]), _F(edm_model.norm_values[1]), _F(edm_model.norm_biases[1]), B, N, F, _ptr(xh), _ptr(zt),
          _ptr(gamma_t))
    return zt


@runtime._on_device
def vlb_loss(model, x, h, node_mask, edge_mask, t_int: Optional[torch.Tensor] = None, eps: Optional[torch.Tensor] = None,
             eps0: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``EnVariationalDiffusion.forward`` in eval mode: the -log p(x,h) estimator compute_loss(t0_always=True) [B]
    (en_diffusion.py:644-804; used by train_edm.val_epoch / test).  Forward only: the two denoiser passes run on the
    sampler's fused inference kernels.  ``t_int`` [B,1] in 1..T, ``eps`` / ``eps0`` [B,N,D] optionally pin the draws."""
    from . import runtime
    _need_cuda(x, "x")
    if model.include_charges:
        raise NotImplementedError("include_charges=True is not used by GaUDI")
    B, N, _ = x.shape
    F = model.in_node_nf
    dev = x.device
    h_cat = h["categorical"] if isinstance(h, dict) else h
    with torch.no_grad():
        if t_int is None:
            t_int = torch.randint(1, model.T + 1, size=(B, 1), device=dev)
        t_f = t_int.to(torch.float32).reshape(B).contiguous()
        g = graph_for(node_mask, edge_mask, B, N)
        mask = g.topo.node_mask
        if eps is None:
            eps = model.sample_combined_position_feature_noise(B, N, node_mask)
        if eps0 is None:
            eps0 = model.sample_combined_position_feature_noise(B, N, node_mask)
        eps, eps0 = _c(eps.to(torch.float32)), _c(eps0.to(torch.float32))
        gamma = model.gamma.gamma.detach().to(torch.float32).contiguous()
        xs, hs = _c(x.to(torch.float32)), _c(h_cat.to(torch.float32))
        nv, nb = model.norm_values, model.norm_biases
        xh, zt, z0, gt = _new(eps, B, N, 3 + F), _new(eps, B, N, 3 + F), _new(eps, B, N, 3 + F), _new(eps, B)
        zeros = torch.zeros(B, dtype=torch.float32, device=dev)
        _call("gb_make_zt", _ptr(xs), _ptr(hs), _ptr(mask), _ptr(eps), _ptr(gamma), _ptr(t_f), _F(nv[0]), _F(nv[1]), _F(nb[1]), B, N, F,
              _ptr(xh), _ptr(zt), _ptr(gt))
        _call("gb_make_zt", _ptr(xs), _ptr(hs), _ptr(mask), _ptr(eps0), _ptr(gamma), _ptr(zeros), _F(nv[0]), _F(nv[1]), _F(nb[1]), B, N, F,
              _ptr(xh), _ptr(z0), _ptr(gt))
        net_t = runtime.denoiser_forward(model.dynamics, t_f / model.T, zt, node_mask, edge_mask)
        net_0 = runtime.denoiser_forward(model.dynamics, zeros, z0, node_mask, edge_mask)
        loss = _new(eps, B)
        _call("gb_vlb_loss", _ptr(net_t), _ptr(eps), _ptr(net_0), _ptr(eps0), _ptr(z0), _ptr(xh), _ptr(mask), _ptr(t_f), _ptr(gamma),
              int(model.T), _F(nv[0]), _F(nv[1]), _F(nb[1]), B, N, F, _ptr(loss), None)
    return loss
