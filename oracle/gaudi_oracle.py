"""CPU oracle for the GaUDI guided-sampling hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32 or fp64) functional restatement of the
reference algorithm.  It is NOT part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.
The product path (``gaudi_b200``) never imports anything from ``oracle/``.

Parity status: PINNED.  The reference has no golden vectors of its own
(SURVEY.md section 4), so the pins are outputs of the unmodified reference run
in the build container (``tests/golden/make_golden.py`` imports
``/root/reference`` and writes ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks every function here against them.

All weights come in as a flat ``state_dict``-style mapping with exactly the
reference's parameter names, so the same dict drives the
reference modules, this oracle and the CUDA path.

Citations are ``file:line`` relative to the reference checkout.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Weights = Dict[str, Tensor]


# --------------------------------------------------------------------------
# configuration presets (utils/args_edm.py:4-51, cond_prediction/prediction_args.py:4-51)
# --------------------------------------------------------------------------
@dataclass
class DenoiserCfg:
    in_node_nf: int = 1          # ring-type classes F (time column is added on top)
    hidden_nf: int = 192
    n_layers: int = 9
    coords_range: float = 4.0    # every block gets the full range (egnn_new.py:290)
    norm_constant: float = 1.0
    normalization_factor: float = 1.0
    inv_sublayers: int = 1
    timesteps: int = 1000
    noise_schedule: str = "polynomial_2"
    noise_precision: float = 1e-5
    norm_values: Tuple[float, float, float] = (3.0, 4.0, 10.0)
    norm_biases: Tuple[Optional[float], float, float] = (None, 0.0, 0.0)
    prefix: str = ""             # "module." when saved through MyDataParallel


@dataclass
class PredictorCfg:
    in_node_nf: int = 1
    out_nf: int = 5
    hidden_nf: int = 196
    n_layers: int = 12
    coords_range: float = 4.0    # divided by n_layers (egnn_predictor/models.py:515)
    prefix: str = ""


# --------------------------------------------------------------------------
# noise schedule (en_diffusion.py:32-61, 186-230, 365-373, 433-457)
# --------------------------------------------------------------------------
def polynomial_gamma(timesteps: int, precision: float, power: float) -> Tensor:
    """gamma table of PredefinedNoiseSchedule('polynomial_<power>').

    float64 numpy arithmetic, stored as float32 (en_diffusion.py:47-61, 207-218).
    """
    steps = timesteps + 1
    grid = np.linspace(0, steps, steps)
    a2 = (1.0 - np.power(grid / steps, power)) ** 2
    # clip_noise_schedule (en_diffusion.py:32-44)
    a2 = np.concatenate([np.ones(1), a2], axis=0)
    ratio = np.clip(a2[1:] / a2[:-1], a_min=0.001, a_max=1.0)
    a2 = np.cumprod(ratio, axis=0)
    a2 = (1.0 - 2.0 * precision) * a2 + precision
    s2 = 1.0 - a2
    gamma = -(np.log(a2) - np.log(s2))
    return torch.from_numpy(gamma).float()


def gamma_table(cfg: DenoiserCfg) -> Tensor:
    name = cfg.noise_schedule
    if "polynomial" not in name:
        raise ValueError(name)                     # en_diffusion.py:203
    parts = name.split("_")
    assert len(parts) == 2
    return polynomial_gamma(cfg.timesteps, cfg.noise_precision, float(parts[1]))


def time_value(step: int, T: int) -> Tensor:
    """fp32 value of ``torch.full((B,1), step) / T`` (en_diffusion.py:1035-1038)."""
    return (torch.full((1, 1), step) / T).reshape(())


def step_scalars(gamma: Tensor, s: int) -> Dict[str, Tensor]:
    """Scalars of one reverse step s <- t=s+1 in the reference's fp32 op order.

    en_diffusion.py:866-876 (lookup), :433-457 (sigma/alpha t|s), :365-373.
    """
    g_s, g_t = gamma[s], gamma[s + 1]
    sigma2_ts = -torch.expm1(F.softplus(g_s) - F.softplus(g_t))
    alpha_ts = torch.exp(0.5 * (F.logsigmoid(-g_t) - F.logsigmoid(-g_s)))
    sigma_ts = torch.sqrt(sigma2_ts)
    sigma_s = torch.sqrt(torch.sigmoid(g_s))
    sigma_t = torch.sqrt(torch.sigmoid(g_t))
    return {
        "alpha_ts": alpha_ts,
        "eps_coef": sigma2_ts / alpha_ts / sigma_t,   # en_diffusion.py:890
        "sigma": sigma_ts * sigma_s / sigma_t,         # en_diffusion.py:894
    }


# --------------------------------------------------------------------------
# masks (sampling_edm.py:119-125, 172-209)
# --------------------------------------------------------------------------
def build_masks(nodesxsample: Tensor, max_nodes: int, orientation: bool) -> Tuple[Tensor, Tensor]:
    """node_mask [B,N,1] and flattened edge_mask [B*N*N,1] (fp32 0/1).

    ``max_nodes`` is nodesxsample.max() for sample_guidance (sampling_edm.py:177)
    and args.max_nodes for sample_pos_edm (:141).  With ``orientation`` the PASs
    block structure [[ring-ring, I], [I, 0]] of :188-208 is produced.
    """
    B = len(nodesxsample)
    nm = torch.zeros(B, max_nodes)
    for b in range(B):
        nm[b, : int(nodesxsample[b])] = 1
    em = nm.unsqueeze(1) * nm.unsqueeze(2)
    em = em * (~torch.eye(max_nodes, dtype=torch.bool)).unsqueeze(0)
    nm = nm.unsqueeze(2)
    if orientation:
        eye = torch.eye(max_nodes).unsqueeze(0).repeat(B, 1, 1)
        left = torch.cat([em, eye], dim=1)
        right = torch.cat([torch.eye(max_nodes), torch.zeros(max_nodes, max_nodes)], dim=0)
        em = torch.cat([left, right.unsqueeze(0).repeat(B, 1, 1)], dim=2)
        nm = torch.cat([nm, nm], dim=1)
    return nm, em.reshape(-1, 1)


def dense_edges(B: int, N: int) -> Tuple[Tensor, Tensor]:
    """All (i,j) pairs incl. i==j, molecule-major, row-major (edm/egnn/models.py:154-175)."""
    base = torch.arange(B).repeat_interleave(N * N) * N
    i = torch.arange(N).repeat_interleave(N).repeat(B)
    j = torch.arange(N).repeat(N * B)
    return base + i, base + j


# --------------------------------------------------------------------------
# small shared pieces
# --------------------------------------------------------------------------
def remove_mean_with_mask(x: Tensor, node_mask: Tensor) -> Tensor:
    """edm/equivariant_diffusion/utils.py:33-44."""
    n = node_mask.sum(1, keepdims=True).clamp(min=1)
    mean = torch.sum(x, dim=1, keepdim=True) / n
    return x - mean * node_mask


def masked_max_abs(x: Tensor, node_mask: Tensor) -> float:
    """quantity checked by assert_correctly_masked (utils.py:62-65)."""
    return float((x * (1 - node_mask)).abs().max())


def cog_rel_error(x: Tensor, eps: float = 1e-10) -> float:
    """quantity checked by assert_mean_zero_with_mask (utils.py:52-59)."""
    return float(torch.sum(x, dim=1, keepdim=True).abs().max()) / (float(x.abs().max()) + eps)


def _lin(w: Weights, key: str, x: Tensor, bias: bool = True) -> Tensor:
    return F.linear(x, w[key + ".weight"], w[key + ".bias"] if bias else None)


def _segment_sum(data: Tensor, row: Tensor, n: int) -> Tensor:
    """scatter_add_ based unsorted_segment_sum (egnn_new.py:403-414, gcl.py:417-423)."""
    out = data.new_zeros((n, data.size(1)))
    out.scatter_add_(0, row.unsqueeze(-1).expand(-1, data.size(1)), data)
    return out


def _radial(x: Tensor, row: Tensor, col: Tensor, norm_constant: float) -> Tuple[Tensor, Tensor]:
    """coord2diff (egnn_new.py:394-400) / coord2radial (gcl.py:308-316)."""
    d = x[row] - x[col]
    r = torch.sum(d ** 2, 1).unsqueeze(1)
    return r, d / (torch.sqrt(r + 1e-8) + norm_constant)


def _append_time(h: Tensor, t: Tensor, B: int, N: int) -> Tensor:
    """time column (edm/egnn/models.py:97-105, egnn_predictor/models.py:442-450)."""
    if t.numel() == 1:
        col = torch.empty(h.shape[0], 1, dtype=h.dtype).fill_(t.item())
    else:
        col = t.view(B, 1).repeat(1, N).view(B * N, 1).to(h.dtype)
    return torch.cat([h, col], dim=1)


# --------------------------------------------------------------------------
# denoiser: EGNN_dynamics._forward -> EGNN -> EquivariantBlock -> GCL / EquivariantUpdate
# --------------------------------------------------------------------------
def denoiser_forward(w: Weights, cfg: DenoiserCfg, z: Tensor, t: Tensor,
                     node_mask: Tensor, edge_mask: Tensor) -> Tensor:
    """eps = phi(z_t, t)  (en_diffusion.py:352-355 -> edm/egnn/models.py:83-152)."""
    p = cfg.prefix + "dynamics.egnn."
    B, N, _ = z.shape
    row, col = dense_edges(B, N)
    nm = node_mask.reshape(B * N, 1).to(z.dtype)
    em = edge_mask.reshape(B * N * N, 1).to(z.dtype)
    xh = z.reshape(B * N, -1).clone() * nm                      # models.py:90
    x = xh[:, :3].clone()
    h = _append_time(xh[:, 3:].clone(), t, B, N)                # models.py:95-105
    x_in = x

    d0, _ = _radial(x, row, col, 1.0)                           # egnn_new.py:301 (initial radial)
    h = _lin(w, p + "embedding", h)                             # egnn_new.py:304
    for b in range(cfg.n_layers):
        q = f"{p}e_block_{b}."
        r, u = _radial(x, row, col, cfg.norm_constant)          # egnn_new.py:216
        e_attr = torch.cat([r, d0], dim=1)                      # egnn_new.py:219
        for s in range(cfg.inv_sublayers):                      # GCL, egnn_new.py:42-89
            g = f"{q}gcl_{s}."
            m = torch.cat([h[row], h[col], e_attr], dim=1)
            m = F.silu(_lin(w, g + "edge_mlp.0", m))
            m = F.silu(_lin(w, g + "edge_mlp.2", m))
            ef = m * torch.sigmoid(_lin(w, g + "att_mlp.0", m)) * em
            agg = _segment_sum(ef, row, B * N) / cfg.normalization_factor
            upd = torch.cat([h, agg], dim=1)
            upd = _lin(w, g + "node_mlp.2", F.silu(_lin(w, g + "node_mlp.0", upd)))
            h = (h + upd) * nm
        g = f"{q}gcl_equiv."                                    # EquivariantUpdate, egnn_new.py:119-155
        c = torch.cat([h[row], h[col], e_attr], dim=1)
        c = F.silu(_lin(w, g + "coord_mlp.0", c))
        c = F.silu(_lin(w, g + "coord_mlp.2", c))
        phi = _lin(w, g + "coord_mlp.4", c, bias=False)
        trans = u * torch.tanh(phi) * cfg.coords_range * em
        x = (x + _segment_sum(trans, row, B * N) / cfg.normalization_factor) * nm
        h = h * nm                                              # egnn_new.py:233-234
    h = _lin(w, p + "embedding_out", h) * nm                    # egnn_new.py:316-318

    vel = ((x - x_in) * nm).view(B, N, 3)                       # models.py:116-118,136
    vel = torch.where(torch.isnan(vel), torch.zeros_like(vel), vel)  # models.py:138-141 (no infs assumed)
    vel = remove_mean_with_mask(vel, nm.view(B, N, 1))          # models.py:146
    h_out = h[:, :-1].view(B, N, -1)                            # drop time column, models.py:132-134
    return torch.cat([vel, h_out], dim=2)


# --------------------------------------------------------------------------
# property predictor: EGNN_predictor.forward -> EGNN -> E_GCL
# --------------------------------------------------------------------------
def predictor_forward(w: Weights, cfg: PredictorCfg, z: Tensor, node_mask: Tensor,
                      edge_mask: Tensor, t: Tensor) -> Tensor:
    """pred [B,out]  (edm/egnn_predictor/models.py:433-457, 543-560; gcl.py:225-316)."""
    p = cfg.prefix + "egnn."
    B, N, _ = z.shape
    row, col = dense_edges(B, N)
    nm = node_mask.reshape(B * N, 1).to(z.dtype)
    em = edge_mask.reshape(B * N * N, 1).to(z.dtype)
    x = z[:, :, :3].reshape(B * N, -1).clone() * nm             # models.py:439
    h = z[:, :, 3:].reshape(B * N, -1).clone() * nm             # models.py:440
    h = _append_time(h, t, B, N)
    a = torch.sum((x[row] - x[col]) ** 2, dim=1, keepdim=True)  # models.py:452
    rng = float(cfg.coords_range) / cfg.n_layers                # models.py:515

    h = _lin(w, p + "embedding", h)
    for l in range(cfg.n_layers):
        g = f"{p}gcl_{l}."
        r, u = _radial(x, row, col, 1.0)                        # gcl.py:308-316
        e = torch.cat([h[row], h[col], r, a], dim=1)            # gcl.py:229
        e = F.silu(_lin(w, g + "edge_mlp.0", e))
        e = F.silu(_lin(w, g + "edge_mlp.2", e))
        ef = e * torch.sigmoid(_lin(w, g + "att_mlp.0", e)) * em        # gcl.py:232-237
        c = F.silu(_lin(w, g + "coord_mlp.0", ef))
        trans = u * torch.tanh(_lin(w, g + "coord_mlp.2", c, bias=False)) * rng * em  # gcl.py:256-262
        x_new = x + _segment_sum(trans, row, B * N)             # gcl.py:265,278
        agg = _segment_sum(ef, row, B * N)                      # gcl.py:242
        upd = torch.cat([h, agg], dim=1)
        upd = _lin(w, g + "node_mlp.2", F.silu(_lin(w, g + "node_mlp.0", upd)))
        h = (h + upd) * nm                                      # gcl.py:248-249,303-305
        x = x_new * nm
    h = _lin(w, p + "embedding_out", h) * nm                    # models.py:555-559
    return h.view(B, N, -1).mean(1)                             # models.py:456-457 (mean over padded N)


def predictor_input_grad(w: Weights, cfg: PredictorCfg, z: Tensor, node_mask: Tensor,
                         edge_mask: Tensor, t: Tensor,
                         target: Callable[[Tensor], Tensor], scale: float) -> Tuple[Tensor, Tensor]:
    """(pred, d(scale*sum_b target(pred)_b)/dz) -- en_diffusion.py:899-903."""
    with torch.enable_grad():
        zz = z.detach().clone().requires_grad_()
        pred = predictor_forward(w, cfg, zz, node_mask, edge_mask, t)
        energy = scale * target(pred).sum()
        grad = torch.autograd.grad(energy, zz)[0]
    return pred.detach(), grad


# --------------------------------------------------------------------------
# reverse-diffusion steps (en_diffusion.py:807-935) and decode (:533-560)
# --------------------------------------------------------------------------
def _project_x(z: Tensor, node_mask: Tensor) -> Tensor:
    return torch.cat([remove_mean_with_mask(z[:, :, :3], node_mask), z[:, :, 3:]], dim=2)


def unguided_step(w: Weights, cfg: DenoiserCfg, gamma: Tensor, s: int, zt: Tensor, noise: Tensor,
                  node_mask: Tensor, edge_mask: Tensor) -> Dict[str, Tensor]:
    """sample_p_zs_given_zt (en_diffusion.py:807-852)."""
    sc = step_scalars(gamma, s)
    t = time_value(s + 1, cfg.timesteps)
    eps = denoiser_forward(w, cfg, zt, t, node_mask, edge_mask)
    mu = zt / sc["alpha_ts"] - sc["eps_coef"] * eps
    zs = _project_x(mu + sc["sigma"] * noise, node_mask)
    return {"eps": eps, "zs": zs}


def clip_and_center_grad(grad: Tensor, node_mask: Tensor, max_norm: float = 10.0) -> Tensor:
    """en_diffusion.py:905-919."""
    coef = torch.clamp(max_norm / (grad.norm(dim=[1, 2]) + 1e-6), max=1.0)
    return _project_x(grad * coef[:, None, None], node_mask)


def guided_step(wd: Weights, dcfg: DenoiserCfg, wp: Weights, pcfg: PredictorCfg, gamma: Tensor, s: int,
                zt: Tensor, noise: Tensor, node_mask: Tensor, edge_mask: Tensor,
                target: Callable[[Tensor], Tensor], scale: float) -> Dict[str, Tensor]:
    """sample_p_zs_given_zt_guidance (en_diffusion.py:854-935).

    ``target`` maps predictor outputs [B,out] to the per-molecule objective [B]
    (the closures of generation_guidance.py:200-211 minus the predictor call).
    The gradient is taken at the freshly sampled z_s, conditioned on t (not s).
    """
    sc = step_scalars(gamma, s)
    t = time_value(s + 1, dcfg.timesteps)
    eps = torch.nan_to_num(denoiser_forward(wd, dcfg, zt, t, node_mask, edge_mask), 0.0)
    mu = zt / sc["alpha_ts"] - sc["eps_coef"] * eps
    zs_pre = mu + sc["sigma"] * noise
    pred, grad_raw = predictor_input_grad(wp, pcfg, zs_pre, node_mask, edge_mask, t, target, scale)
    grad = clip_and_center_grad(grad_raw, node_mask)
    zs = _project_x(zs_pre - sc["sigma"] * grad, node_mask)
    zs = torch.nan_to_num(zs, 0.0)
    return {"eps": eps, "zs_pre": zs_pre, "pred": pred, "grad_raw": grad_raw, "grad": grad, "zs": zs}


def decode(w: Weights, cfg: DenoiserCfg, gamma: Tensor, z0: Tensor, noise: Tensor,
           node_mask: Tensor, edge_mask: Tensor) -> Tuple[Tensor, Tensor]:
    """sample_p_xh_given_z0 with include_charges=False (en_diffusion.py:533-560, 493-505, 406-415)."""
    g0 = gamma[0]
    sigma_x = torch.exp(-(-0.5 * g0))                               # SNR(-0.5*gamma_0), :538,375-377
    eps = denoiser_forward(w, cfg, z0, torch.zeros(()), node_mask, edge_mask)
    sigma0, alpha0 = torch.sqrt(torch.sigmoid(g0)), torch.sqrt(torch.sigmoid(-g0))
    mu = 1.0 / alpha0 * (z0 - sigma0 * eps)                         # :501
    xh = mu + sigma_x * noise
    x = xh[:, :, :3] * cfg.norm_values[0]
    h_cat = (z0[:, :, 3:] * cfg.norm_values[1] + cfg.norm_biases[1]) * node_mask
    one_hot = F.one_hot(torch.argmax(h_cat, dim=2), cfg.in_node_nf) * node_mask  # float32 result
    return x, one_hot


def draw_noise(B: int, N: int, D: int, node_mask: Tensor, std: float = 1.0,
               generator: Optional[torch.Generator] = None) -> Tensor:
    """sample_combined_position_feature_noise (en_diffusion.py:937-956; utils.py:116-125,146-149)."""
    zx = torch.randn((B, N, 3), generator=generator) * std * node_mask
    zx = remove_mean_with_mask(zx, node_mask)
    zh = torch.randn((B, N, D - 3), generator=generator) * std * node_mask
    return torch.cat([zx, zh], dim=2)


def sample_chain(wd: Weights, dcfg: DenoiserCfg, noise: Tensor, node_mask: Tensor, edge_mask: Tensor,
                 wp: Optional[Weights] = None, pcfg: Optional[PredictorCfg] = None,
                 target: Optional[Callable[[Tensor], Tensor]] = None, scale: float = 1.0,
                 record_every: int = 0) -> Dict[str, Tensor]:
    """Full sampler with injected noise [T+2,B,N,D] (noise[0] is z_T, noise[k] the k-th draw).

    EnVariationalDiffusion.sample (en_diffusion.py:958-1008) when ``wp`` is None,
    else .sample_guidance (:1010-1067).
    """
    gamma = gamma_table(dcfg)
    T = dcfg.timesteps
    z = noise[0]
    rec = {}
    k = 1
    for s in reversed(range(T)):
        if wp is None:
            z = unguided_step(wd, dcfg, gamma, s, z, noise[k], node_mask, edge_mask)["zs"]
        else:
            z = guided_step(wd, dcfg, wp, pcfg, gamma, s, z, noise[k], node_mask, edge_mask, target, scale)["zs"]
        k += 1
        if record_every and s % record_every == 0:
            rec[f"z_{s}"] = z.clone()
    x, one_hot = decode(wd, dcfg, gamma, z, noise[k], node_mask, edge_mask)
    if float(torch.sum(x, dim=1, keepdim=True).abs().max()) > 5e-2:     # :1000-1006 / :1059-1065
        x = remove_mean_with_mask(x, node_mask)
    rec.update({"z0": z, "x": x, "one_hot": one_hot})
    return rec


# --------------------------------------------------------------------------
# cond_fn closures (generation_guidance.py:200-211)
# --------------------------------------------------------------------------
def target_max_gap(pred: Tensor) -> Tensor:
    return -pred[:, 1]


def make_target_opv(mean: Tensor, std: Tensor) -> Callable[[Tensor], Tensor]:
    def f(pred: Tensor) -> Tensor:
        p = pred * std + mean                                  # models_edm.py:186-188
        return p[:, 3] + p[:, 2] + 3 * p[:, 0]
    return f


# --------------------------------------------------------------------------
# sub-module restatements (used to test the stand-alone module forwards of the product)
# --------------------------------------------------------------------------
def gcl_layer(w: Weights, key: str, h: Tensor, row: Tensor, col: Tensor, edge_attr: Tensor, nm: Tensor, em: Tensor,
              normalization_factor: float = 1.0, attention: bool = True) -> Tensor:
    """GCL.forward (egnn_new.py:75-89) for parameters stored under ``key`` (e.g. '...gcl_0.')."""
    m = torch.cat([h[row], h[col], edge_attr], dim=1)
    m = F.silu(_lin(w, key + "edge_mlp.0", m))
    m = F.silu(_lin(w, key + "edge_mlp.2", m))
    ef = (m * torch.sigmoid(_lin(w, key + "att_mlp.0", m)) if attention else m) * em
    agg = _segment_sum(ef, row, h.size(0)) / normalization_factor
    upd = _lin(w, key + "node_mlp.2", F.silu(_lin(w, key + "node_mlp.0", torch.cat([h, agg], dim=1))))
    return (h + upd) * nm


def equiv_update_layer(w: Weights, key: str, h: Tensor, x: Tensor, row: Tensor, col: Tensor, coord_diff: Tensor,
                       edge_attr: Tensor, nm: Tensor, em: Tensor, coords_range: float, tanh: bool = True,
                       normalization_factor: float = 1.0) -> Tensor:
    """EquivariantUpdate.forward (egnn_new.py:119-155)."""
    c = torch.cat([h[row], h[col], edge_attr], dim=1)
    c = F.silu(_lin(w, key + "coord_mlp.0", c))
    c = F.silu(_lin(w, key + "coord_mlp.2", c))
    phi = _lin(w, key + "coord_mlp.4", c, bias=False)
    trans = (coord_diff * torch.tanh(phi) * coords_range if tanh else coord_diff * phi) * em
    return (x + _segment_sum(trans, row, x.size(0)) / normalization_factor) * nm


def equiv_block(w: Weights, key: str, h: Tensor, x: Tensor, row: Tensor, col: Tensor, d0: Tensor, nm: Tensor, em: Tensor,
                coords_range: float, norm_constant: float = 1.0, n_sub: int = 1) -> Tuple[Tensor, Tensor]:
    """EquivariantBlock.forward (egnn_new.py:214-235)."""
    r, u = _radial(x, row, col, norm_constant)
    e_attr = torch.cat([r, d0], dim=1)
    for s in range(n_sub):
        h = gcl_layer(w, f"{key}gcl_{s}.", h, row, col, e_attr, nm, em)
    x = equiv_update_layer(w, key + "gcl_equiv.", h, x, row, col, u, e_attr, nm, em, coords_range)
    return h * nm, x


def e_gcl_layer(w: Weights, key: str, h: Tensor, x: Tensor, row: Tensor, col: Tensor, edge_attr: Tensor, nm: Tensor,
                em: Tensor, coords_range: float) -> Tuple[Tensor, Tensor]:
    """E_GCL.forward (gcl.py:281-306) with attention and tanh."""
    r, u = _radial(x, row, col, 1.0)
    e = torch.cat([h[row], h[col], r, edge_attr], dim=1)
    e = F.silu(_lin(w, key + "edge_mlp.0", e))
    e = F.silu(_lin(w, key + "edge_mlp.2", e))
    ef = e * torch.sigmoid(_lin(w, key + "att_mlp.0", e)) * em
    c = F.silu(_lin(w, key + "coord_mlp.0", ef))
    trans = u * torch.tanh(_lin(w, key + "coord_mlp.2", c, bias=False)) * coords_range * em
    x_new = x + _segment_sum(trans, row, x.size(0))
    agg = _segment_sum(ef, row, h.size(0))
    upd = _lin(w, key + "node_mlp.2", F.silu(_lin(w, key + "node_mlp.0", torch.cat([h, agg], dim=1))))
    return (h + upd) * nm, x_new * nm


# --------------------------------------------------------------------------
# training loss: EnVariationalDiffusion.forward in train mode, loss_type 'l2',
# include_charges = False, as driven by train_edm.compute_loss (train_edm.py:36-49)
# --------------------------------------------------------------------------
def _sum_except_batch(x: Tensor) -> Tensor:
    return x.reshape(x.size(0), -1).sum(-1)


def _cdf_std_gaussian(x: Tensor) -> Tensor:
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2)))                 # en_diffusion.py:160-161


def training_loss(w: Weights, cfg: DenoiserCfg, gamma: Tensor, x: Tensor, h_cat: Tensor, node_mask: Tensor,
                  edge_mask: Tensor, t_int: Tensor, eps: Tensor) -> Tuple[Tensor, Tensor]:
    """(loss [B], z_t) of compute_loss(t0_always=False) with the two random draws (t_int [B,1] float, eps [B,N,D])
    injected.  en_diffusion.py:777-804 (forward), :384-404 (normalize), :644-775 (compute_loss), :459-491 (kl_prior),
    :568-642 (log_pxh_given_z0_without_constants), :507-520 (compute_error)."""
    B, N, _ = x.shape
    T = cfg.timesteps
    x = x / cfg.norm_values[0]                                        # normalize, :385
    hc = (h_cat.float() - cfg.norm_biases[1]) / cfg.norm_values[1] * node_mask
    t_is_zero = (t_int == 0).float()
    t = t_int / T                                                     # :666
    g_t = gamma[torch.round(t * T).long()].view(B, 1, 1)              # PredefinedNoiseSchedule.forward, :228-230
    alpha_t, sigma_t = torch.sqrt(torch.sigmoid(-g_t)), torch.sqrt(torch.sigmoid(g_t))
    xh = torch.cat([x, hc], dim=2)                                    # :683 (the integer part is empty)
    z_t = alpha_t * xh + sigma_t * eps                                # :685
    net_out = denoiser_forward(w, cfg, z_t, t, node_mask, edge_mask)   # per-sample time column (models.py:100-103)
    denom = (3 + cfg.in_node_nf) * N                                  # compute_error, :511-514 (training, l2)
    error = _sum_except_batch((eps - net_out) ** 2) / denom
    loss_t_larger_than_zero = 0.5 * error                             # SNR_weight = 1, :695-701
    # kl_prior, :459-491
    g_T = gamma[T].view(1, 1, 1).expand(B, 1, 1)
    alpha_T = torch.sqrt(torch.sigmoid(-g_T))
    mu_T = alpha_T * xh
    mu_T_x, mu_T_h = mu_T[:, :, :3], mu_T[:, :, 3:]
    sigma_T_x = torch.sqrt(torch.sigmoid(g_T)).squeeze()
    sigma_T_h = torch.sqrt(torch.sigmoid(g_T))
    kl_h = _sum_except_batch((torch.log(1.0 / sigma_T_h) + 0.5 * (sigma_T_h ** 2 + mu_T_h ** 2) - 0.5) * node_mask)  # :131-145
    d_sub = (node_mask.squeeze(2).sum(1) - 1) * 3                     # :379-382
    mu_norm2 = _sum_except_batch(mu_T_x ** 2)
    kl_x = d_sub * torch.log(1.0 / sigma_T_x) + 0.5 * (d_sub * sigma_T_x ** 2 + mu_norm2) - 0.5 * d_sub              # :148-157
    kl_prior = kl_x + kl_h
    # L0 term evaluated at (z_t, gamma_t) and selected by t == 0, :568-642
    eps_x, net_x = eps[:, :, :3], net_out[:, :, :3]
    log_p_x = -0.5 * _sum_except_batch((eps_x - net_x) ** 2) / denom  # compute_error on the x part (same denom: eps_t.shape[1] = N)
    sigma_0_cat = sigma_t * cfg.norm_values[1]
    onehot = hc * cfg.norm_values[1] + cfg.norm_biases[1]
    centered = z_t[:, :, 3:] * cfg.norm_values[1] + cfg.norm_biases[1] - 1
    log_prop = torch.log(_cdf_std_gaussian((centered + 0.5) / sigma_0_cat)
                         - _cdf_std_gaussian((centered - 0.5) / sigma_0_cat) + 1e-10)
    log_probs = log_prop - torch.logsumexp(log_prop, dim=2, keepdim=True)
    log_ph_cat = _sum_except_batch(log_probs * onehot * node_mask)
    loss_term_0 = -(log_p_x + log_ph_cat)                             # the integer part sums to 0 over an empty tensor
    tz = t_is_zero.squeeze(1)
    loss = kl_prior + loss_term_0 * tz + (1 - tz) * loss_t_larger_than_zero     # :744-760; constants and delta_log_px are zeroed
    return loss, z_t



# --------------------------------------------------------------------------
# predictor training inputs: cond_prediction/train_cond_predictor.py:47-81
# --------------------------------------------------------------------------
def sample_edm_t(cfg: DenoiserCfg, gamma: Tensor, x: Tensor, h_cat: Tensor, node_mask: Tensor, t: Tensor, eps: Tensor) -> Tensor:
    """z_t = alpha_t [x/3, h/4*mask] + sigma_t eps at per-sample t [B,1] (train_cond_predictor.py:47-62); the loss on top is
    l1_loss(predictor_forward(z_t, t), target) (:65-81)."""
    B = x.shape[0]
    xh = torch.cat([x / cfg.norm_values[0], (h_cat.float() - cfg.norm_biases[1]) / cfg.norm_values[1] * node_mask], dim=-1)
    g_t = gamma[torch.round(t * cfg.timesteps).long()].view(B, 1, 1)
    return torch.sqrt(torch.sigmoid(-g_t)) * xh + torch.sqrt(torch.sigmoid(g_t)) * eps


def validation_nll(w: Weights, cfg: DenoiserCfg, gamma: Tensor, x: Tensor, h_cat: Tensor, node_mask: Tensor, edge_mask: Tensor,
                   t_int: Tensor, eps: Tensor, eps0: Tensor) -> Tensor:
    """-log p(x,h) estimate [B] of forward() in eval mode = compute_loss(t0_always=True) with its draws injected
    (en_diffusion.py:644-775, 777-804; :517-531 log constants; :568-642 L0 term; :459-491 prior KL)."""
    B, N, _ = x.shape
    T = cfg.timesteps
    d_sub = (node_mask.squeeze(2).sum(1) - 1) * 3
    delta_log_px = -d_sub * np.log(cfg.norm_values[0])                 # normalize, :386-388
    xh = torch.cat([x / cfg.norm_values[0], (h_cat.float() - cfg.norm_biases[1]) / cfg.norm_values[1] * node_mask], dim=2)
    t, s = t_int / T, (t_int - 1) / T
    g_t = gamma[torch.round(t * T).long()].view(B, 1, 1)
    g_s = gamma[torch.round(s * T).long()].view(B, 1, 1)
    z_t = torch.sqrt(torch.sigmoid(-g_t)) * xh + torch.sqrt(torch.sigmoid(g_t)) * eps
    net_t = denoiser_forward(w, cfg, z_t, t, node_mask, edge_mask)
    error = _sum_except_batch((eps - net_t) ** 2)                      # compute_error outside training: denom = 1
    snr_w = (torch.exp(-(g_s - g_t)) - 1).squeeze(1).squeeze(1)
    loss_t = 0.5 * snr_w * error
    g_0 = gamma[0]
    neg_log_constants = -(d_sub * (-(0.5 * g_0) - 0.5 * np.log(2 * np.pi)))
    g_T = gamma[T].view(1, 1, 1).expand(B, 1, 1)
    mu_T = torch.sqrt(torch.sigmoid(-g_T)) * xh
    sig_T_x, sig_T_h = torch.sqrt(torch.sigmoid(g_T)).squeeze(), torch.sqrt(torch.sigmoid(g_T))
    kl_h = _sum_except_batch((torch.log(1.0 / sig_T_h) + 0.5 * (sig_T_h ** 2 + mu_T[:, :, 3:] ** 2) - 0.5) * node_mask)
    kl_x = d_sub * torch.log(1.0 / sig_T_x) + 0.5 * (d_sub * sig_T_x ** 2 + _sum_except_batch(mu_T[:, :, :3] ** 2)) - 0.5 * d_sub
    z_0 = torch.sqrt(torch.sigmoid(-g_0)) * xh + torch.sqrt(torch.sigmoid(g_0)) * eps0
    net_0 = denoiser_forward(w, cfg, z_0, torch.zeros(B, 1), node_mask, edge_mask)
    log_p_x = -0.5 * _sum_except_batch((eps0[:, :, :3] - net_0[:, :, :3]) ** 2)
    sigma_0_cat = torch.sqrt(torch.sigmoid(g_0)) * cfg.norm_values[1]
    onehot = xh[:, :, 3:] * cfg.norm_values[1] + cfg.norm_biases[1]
    centered = z_0[:, :, 3:] * cfg.norm_values[1] + cfg.norm_biases[1] - 1
    log_prop = torch.log(_cdf_std_gaussian((centered + 0.5) / sigma_0_cat) - _cdf_std_gaussian((centered - 0.5) / sigma_0_cat) + 1e-10)
    log_probs = log_prop - torch.logsumexp(log_prop, dim=2, keepdim=True)
    loss_term_0 = -(log_p_x + _sum_except_batch(log_probs * onehot * node_mask))
    return kl_x + kl_h + T * loss_t + neg_log_constants + loss_term_0 - delta_log_px
