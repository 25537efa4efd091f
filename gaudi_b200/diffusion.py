"""``EnVariationalDiffusion`` with GaUDI's sampling surface, running on the sm_100a kernels.

Mirrors edm/equivariant_diffusion/en_diffusion.py: ``PredefinedNoiseSchedule`` :186-230, ``polynomial_schedule`` /
``clip_noise_schedule`` :32-61, ``EnVariationalDiffusion`` :279 with ``phi`` :352, ``sigma``/``alpha`` :365-373,
``sigma_and_alpha_t_given_s`` :433-457, ``sample_p_zs_given_zt`` :807, ``sample_p_zs_given_zt_guidance`` :854,
``sample_combined_position_feature_noise`` :937, ``sample`` :958, ``sample_guidance`` :1010,
``sample_p_xh_given_z0`` :533, ``normalize``/``unnormalize`` :384-415.

What is different by design
  * the per-step scalars (alpha_t|s, sigma^2_t|s/alpha_t|s/sigma_t, sigma_t|s sigma_s/sigma_t) are a [T,3] table
    computed ONCE on the host with the reference's fp32 torch op order, so device results use identical constants;
  * the reference's per-step ``assert_mean_zero_with_mask`` / ``assert_correctly_masked`` (6 host syncs a step) are
    device-side running maxima checked once after the loop (same thresholds, same AssertionError);
  * a target function that is affine in the predictor outputs (``AffineTarget``; both closures of
    generation_guidance.py:200-211 are) runs the whole T-step loop inside ``gb_sample_loop`` (optionally as a
    replayed CUDA graph); any other Python callable goes through autograd with the hand-written backward.
  * with ``self.seed`` set (``dist`` sets ``seed + rank``) EVERY draw of a sampler call -- z_T, the per-step noise and the
    p(x|z0) noise -- comes from the in-kernel Philox stream of that seed (draw indices 0, 1..T, T+1), so a seed
    reproduces a run and ranks never share noise; with ``seed is None`` torch's generator is used as in the reference.
Out of scope (raise NotImplementedError): learned schedule, context conditioning, ``include_charges``.
"""
from __future__ import annotations

from typing import Callable, Optional

import os
import numpy as np
import torch
import torch.nn.functional as F

from . import runtime


def clip_noise_schedule(alphas2, clip_value=0.001):
    alphas2 = np.concatenate([np.ones(1), alphas2], axis=0)
    step = np.clip(alphas2[1:] / alphas2[:-1], a_min=clip_value, a_max=1.0)
    return np.cumprod(step, axis=0)


def polynomial_schedule(timesteps: int, s=1e-4, power=3.0):
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    alphas2 = (1 - np.power(x / steps, power)) ** 2
    alphas2 = clip_noise_schedule(alphas2, clip_value=0.001)
    return (1 - 2 * s) * alphas2 + s


def cosine_beta_schedule(timesteps, s=0.008, raise_to_power: float = 1):
    steps = timesteps + 2
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = np.clip(1 - (ac[1:] / ac[:-1]), a_min=0, a_max=0.999)
    ac = np.cumprod(1.0 - betas, axis=0)
    return np.power(ac, raise_to_power) if raise_to_power != 1 else ac


class PredefinedNoiseSchedule(torch.nn.Module):
    """Lookup table gamma[t_int] (float64 numpy -> float32 parameter, not trained)."""

    def __init__(self, noise_schedule, timesteps, precision):
        super().__init__()
        self.timesteps = timesteps
        if noise_schedule == "cosine":
            alphas2 = cosine_beta_schedule(timesteps)
        elif "polynomial" in noise_schedule:
            parts = noise_schedule.split("_")
            assert len(parts) == 2
            alphas2 = polynomial_schedule(timesteps, s=precision, power=float(parts[1]))
        else:
            raise ValueError(noise_schedule)
        sigmas2 = 1 - alphas2
        gamma = -(np.log(alphas2) - np.log(sigmas2))
        self.gamma = torch.nn.Parameter(torch.from_numpy(gamma).float(), requires_grad=False)

    def forward(self, t):
        t_int = torch.round(t * self.timesteps).long()
        return self.gamma[t_int]


class AffineTarget:
    """cond_fn of the form  f(z) = sum_o w[o] * pred(z)[o] (+ const)  over the predictor outputs.

    Callable with the reference's cond_fn signature ``(z, node_mask, edge_mask, t) -> [B]`` (so it also works on
    the generic autograd path and in the reference itself), and recognised by ``sample_guidance`` for the fused
    loop.  ``AffineTarget.max_gap(pred)`` and ``AffineTarget.opv(pred, prop_dist)`` are the two closures of
    generation_guidance.py:200-211.
    """

    def __init__(self, predictor, weights, bias: float = 0.0):
        self.predictor = predictor
        self.weights = torch.as_tensor(weights, dtype=torch.float32).reshape(-1)
        self.bias = float(bias)

    def __call__(self, z, node_mask, edge_mask, t):
        pred = self.predictor(z, node_mask, edge_mask, t)
        return (pred * self.weights.to(pred.device)).sum(-1) + self.bias

    @classmethod
    def max_gap(cls, predictor, index: int = 1, n_out: Optional[int] = None):
        n_out = n_out or predictor.hyper["out_nf"]
        w = torch.zeros(n_out)
        w[index] = -1.0
        return cls(predictor, w)

    @classmethod
    def opv(cls, predictor, prop_dist):
        mean, std = prop_dist.mean.float().cpu(), prop_dist.std.float().cpu()
        w = torch.zeros_like(std)
        w[3] += std[3]; w[2] += std[2]; w[0] += 3 * std[0]          # ip + ea + 3*gap on pred*std+mean
        return cls(predictor, w, float(mean[3] + mean[2] + 3 * mean[0]))


class EnVariationalDiffusion(torch.nn.Module):
    """The E(n) diffusion module (sampling surface)."""

    def __init__(self, dynamics, in_node_nf: int, n_dims: int, timesteps: int = 1000, parametrization="eps",
                 noise_schedule="learned", noise_precision=1e-4, loss_type="vlb", norm_values=(1.0, 1.0, 1.0),
                 norm_biases=(None, 0.0, 0.0), include_charges=True, device="cpu"):
        super().__init__()
        assert loss_type in {"vlb", "l2"}
        assert parametrization == "eps"
        if noise_schedule == "learned":
            raise NotImplementedError("learned noise schedule (GammaNetwork) is not on the sampling path of args_edm")
        if include_charges:
            raise NotImplementedError("include_charges=True is never used by GaUDI (models_edm.py:94)")
        self.loss_type = loss_type
        self.include_charges = include_charges
        self.gamma = PredefinedNoiseSchedule(noise_schedule, timesteps=timesteps, precision=noise_precision).to(device)
        self.dynamics = dynamics
        self.in_node_nf = in_node_nf
        self.n_dims = n_dims
        self.num_classes = self.in_node_nf - self.include_charges
        self.T = timesteps
        self.parametrization = parametrization
        self.norm_values = norm_values
        self.norm_biases = norm_biases
        self.register_buffer("buffer", torch.zeros(1).to(device))
        self.check_issues_norm_values()
        self.use_cuda_graph = True          # replay one captured step in the fused loop
        self.seed: Optional[int] = None     # Philox seed of the fused loop (None: drawn from torch's CPU generator)
        self.last_stats: Optional[torch.Tensor] = None

    # ---- schedule helpers (host, fp32, reference op order) ------------------------------------------
    def check_issues_norm_values(self, num_stdevs=8):
        zeros = torch.zeros((1, 1))
        gamma_0 = self.gamma.gamma.detach().cpu()[0].reshape(1, 1)
        sigma_0 = torch.sqrt(torch.sigmoid(gamma_0)).item()
        max_norm_value = max(self.norm_values[1], self.norm_values[2])
        if sigma_0 * num_stdevs > 1.0 / max_norm_value:
            raise ValueError(f"Value for normalization value {max_norm_value} probably too large with sigma_0 "
                             f"{sigma_0:.5f} and 1 / norm_value = {1. / max_norm_value}")

    def inflate_batch_array(self, array, target):
        return array.view((array.size(0),) + (1,) * (len(target.size()) - 1))

    def sigma(self, gamma, target_tensor):
        return self.inflate_batch_array(torch.sqrt(torch.sigmoid(gamma)), target_tensor)

    def alpha(self, gamma, target_tensor):
        return self.inflate_batch_array(torch.sqrt(torch.sigmoid(-gamma)), target_tensor)

    def SNR(self, gamma):
        return torch.exp(-gamma)

    def sigma_and_alpha_t_given_s(self, gamma_t, gamma_s, target_tensor):
        sigma2 = self.inflate_batch_array(-torch.expm1(F.softplus(gamma_s) - F.softplus(gamma_t)), target_tensor)
        log_a2 = F.logsigmoid(-gamma_t) - F.logsigmoid(-gamma_s)
        alpha = self.inflate_batch_array(torch.exp(0.5 * log_a2), target_tensor)
        return sigma2, torch.sqrt(sigma2), alpha

    def _tables(self, device):
        """([T,3] step coefficients, [T+1] time values, [3] decode coefficients) on ``device`` (cached)."""
        g = self.gamma.gamma.detach()
        key = (g.data_ptr(), g._version, str(device))
        cache = self.__dict__.get("_tab")
        if cache is None or cache[0] != key:
            gam = g.cpu().float()
            g_s, g_t = gam[:-1], gam[1:]
            sigma2 = -torch.expm1(F.softplus(g_s) - F.softplus(g_t))
            alpha_ts = torch.exp(0.5 * (F.logsigmoid(-g_t) - F.logsigmoid(-g_s)))
            sigma_ts = torch.sqrt(sigma2)
            sigma_s, sigma_t = torch.sqrt(torch.sigmoid(g_s)), torch.sqrt(torch.sigmoid(g_t))
            sched = torch.stack([alpha_ts, sigma2 / alpha_ts / sigma_t, sigma_ts * sigma_s / sigma_t], dim=1)
            tvals = torch.arange(self.T + 1) / self.T                       # == torch.full((B,1), k) / T
            g0 = gam[0]
            dec = torch.stack([torch.sqrt(torch.sigmoid(g0)), torch.sqrt(torch.sigmoid(-g0)), torch.exp(-(-0.5 * g0))])
            cache = (key, sched.contiguous().to(device), tvals.float().contiguous().to(device), dec.contiguous().to(device))
            self.__dict__["_tab"] = cache
        return cache[1], cache[2], cache[3]

    def _step_index(self, s: torch.Tensor) -> int:
        return int(torch.round(s.reshape(-1)[0] * self.T).item())

    # ---- network ---------------------------------------------------------------------------------------
    def phi(self, x, t, node_mask, edge_mask, context):
        return self.dynamics._forward(t, x, node_mask, edge_mask, context)

    # ---- (un)normalisation (en_diffusion.py:384-415) -------------------------------------------------------
    def normalize(self, x, h, node_mask):
        x = x / self.norm_values[0]
        n_nodes = torch.sum(node_mask.squeeze(2), dim=1)
        delta_log_px = -((n_nodes - 1) * self.n_dims) * np.log(self.norm_values[0])
        h_cat = (h["categorical"].float() - self.norm_biases[1]) / self.norm_values[1] * node_mask
        h_int = (h["integer"].float() - self.norm_biases[2]) / self.norm_values[2]
        if self.include_charges:
            h_int = h_int * node_mask
        return x, {"categorical": h_cat, "integer": h_int}, delta_log_px

    def unnormalize(self, x, h_cat, h_int, node_mask):
        x = x * self.norm_values[0]
        h_cat = (h_cat * self.norm_values[1] + self.norm_biases[1]) * node_mask
        h_int = h_int * self.norm_values[2] + self.norm_biases[2]
        if self.include_charges:
            h_int = h_int * node_mask
        return x, h_cat, h_int

    def _gamma_T(self) -> float:
        """gamma(1) as a host float (cached: the schedule table is frozen)."""
        g = self.gamma.gamma
        key = (g.data_ptr(), g._version)
        cache = self.__dict__.get("_gT")
        if cache is None or cache[0] != key:
            cache = (key, float(g.detach()[-1].cpu()))
            self.__dict__["_gT"] = cache
        return cache[1]

    def forward(self, x, h, node_mask=None, edge_mask=None, context=None, t_int=None, eps=None, eps0=None):
        """Loss [B] (en_diffusion.py:777-804).  Train mode with loss_type 'l2' (what train_edm.py optimises): the denoising
        loss with parameter gradients; eval mode: the variational bound compute_loss(t0_always=True) that val_epoch / test
        report (forward only).  ``t_int`` / ``eps`` (/ ``eps0``) optionally pin the random draws of compute_loss (tests)."""
        if context is not None:
            raise NotImplementedError("context conditioning is unused by GaUDI")
        from . import training
        if not self.training:
            return training.vlb_loss(self, x, h, node_mask, edge_mask, t_int=t_int, eps=eps, eps0=eps0)
        return training.training_loss(self, x, h, node_mask, edge_mask, t_int=t_int, eps=eps)

    # ---- noise -------------------------------------------------------------------------------------------------
    def sample_combined_position_feature_noise(self, n_samples, n_nodes, node_mask, std=1.0):
        """Centre-of-gravity-free x noise and masked h noise (en_diffusion.py:937-956), drawn with torch's
        generator on the mask's device (the fused loop uses the in-kernel Philox source instead)."""
        zx = torch.randn((n_samples, n_nodes, self.n_dims), device=node_mask.device) * std * node_mask
        n = node_mask.sum(1, keepdims=True).clamp(min=1)
        zx = zx - (torch.sum(zx, dim=1, keepdim=True) / n) * node_mask
        zh = torch.randn((n_samples, n_nodes, self.in_node_nf), device=node_mask.device) * std * node_mask
        return torch.cat([zx, zh], dim=2)

    def sample_normal(self, mu, sigma, node_mask, fix_noise=False):
        bs = 1 if fix_noise else mu.size(0)
        return mu + sigma * self.sample_combined_position_feature_noise(bs, mu.size(1), node_mask)

    # ---- single steps (reference signatures) ------------------------------------------------------------------------
    def _flat_mask(self, node_mask):
        return node_mask.reshape(-1).to(torch.float32).contiguous()

    def _step_noise(self, zt, node_mask, fix_noise, noise):
        if noise is not None:
            return noise.to(torch.float32).contiguous()
        bs = 1 if fix_noise else zt.size(0)
        eps = self.sample_combined_position_feature_noise(bs, zt.size(1), node_mask)
        return eps.expand_as(zt).contiguous()

    @torch.no_grad()
    def sample_p_zs_given_zt(self, s, t, zt, node_mask, edge_mask, context=None, fix_noise=False, noise=None,
                             stats=None):
        """zs ~ p(zs | zt), unguided (en_diffusion.py:807-852).  ``noise``: optional injected [B,N,D] draw."""
        sched, tvals, _ = self._tables(zt.device)
        si = self._step_index(s)
        zt = zt.to(torch.float32).contiguous()
        eps = runtime.denoiser_forward(self.dynamics, tvals[si + 1: si + 2], zt, node_mask, edge_mask, False, stats)
        return runtime.step_sample(zt, eps, self._step_noise(zt, node_mask, fix_noise, noise), sched[si],
                                   self._flat_mask(node_mask), project=True)

    def sample_p_zs_given_zt_guidance(self, s, t, zt, node_mask, edge_mask, target_function, scale, fix_noise=False,
                                      noise=None, stats=None, return_parts=False):
        """Guided step (en_diffusion.py:854-935): gradient of scale*target at the fresh z_s, conditioned on t."""
        sched, tvals, _ = self._tables(zt.device)
        si = self._step_index(s)
        nm = self._flat_mask(node_mask)
        t_dev = tvals[si + 1: si + 2]
        with torch.no_grad():
            zt = zt.to(torch.float32).contiguous()
            eps = runtime.denoiser_forward(self.dynamics, t_dev, zt, node_mask, edge_mask, True, stats)
            zs_pre = runtime.step_sample(zt, eps, self._step_noise(zt, node_mask, fix_noise, noise), sched[si], nm,
                                         project=False)
        if isinstance(target_function, AffineTarget):
            w = (target_function.weights * float(scale)).to(zt.device)
            pred, grad = runtime.predictor_value_and_grad(target_function.predictor, zs_pre, node_mask, edge_mask,
                                                          t_dev, w)
        else:
            with torch.enable_grad():
                zz = zs_pre.detach().requires_grad_()
                energy = scale * target_function(zz, node_mask, edge_mask, t).sum()
                grad = torch.autograd.grad(energy, zz)[0]
            pred = None
        with torch.no_grad():
            zs = runtime.step_guide(zs_pre, grad.to(torch.float32).contiguous(), sched[si], nm)
        if return_parts:
            return {"eps": eps, "zs_pre": zs_pre, "grad_raw": grad, "pred": pred, "zs": zs}
        return zs

    @torch.no_grad()
    def sample_p_xh_given_z0(self, z0, node_mask, edge_mask, context=None, fix_noise=False, noise=None):
        """x, h ~ p(x, h | z0) (en_diffusion.py:533-560)."""
        _, tvals, dec = self._tables(z0.device)
        z0 = z0.to(torch.float32).contiguous()
        eps = runtime.denoiser_forward(self.dynamics, tvals[0:1], z0, node_mask, edge_mask, False, None)
        nz = self._step_noise(z0, node_mask, fix_noise, noise)
        x, one_hot, cog = runtime.decode(z0, eps, nz, dec, self._flat_mask(node_mask), self.norm_values[0],
                                         self.norm_values[1], self.norm_biases[1])
        h_int = torch.zeros(z0.size(0), z0.size(1), 0, dtype=torch.float32, device=z0.device)
        self._last_cog = cog
        return x, {"integer": h_int, "categorical": one_hot}

    # ---- invariants: device-side maxima, checked once (utils.py:52-65) ---------------------------------------------
    @staticmethod
    def _check_stats(stats: torch.Tensor) -> None:
        st = stats.detach().cpu().view(-1, 8)
        assert not bool((st[:, 4] >= 1e-4).any()), "Variables not masked properly."
        for col_max, col_cog in ((0, 1), (2, 3)):
            rel = st[:, col_cog] / (st[:, col_max] + 1e-10)
            bad = ~(rel < 1e-2)
            assert not bool(bad.any()), f"Mean is not zero, relative_error {float(rel[bad].max()) if bad.any() else 0.0}"

    def _finish(self, z, node_mask, edge_mask, fix_noise, noise_last):
        x, h = self.sample_p_xh_given_z0(z, node_mask, edge_mask, None, fix_noise, noise_last)
        nm = self._flat_mask(node_mask)
        # assert_mean_zero_with_mask(x) + CoG drift projection (en_diffusion.py:998-1006 / 1057-1065)
        max_cog = float(self._last_cog.item())
        largest = float(x.abs().max().item())
        assert float((x * (1 - node_mask)).abs().max().item()) < 1e-4, "Variables not masked properly."
        rel = max_cog / (largest + 1e-10)
        assert rel < 1e-2, f"Mean is not zero, relative_error {rel}"
        if max_cog > 5e-2:
            print(f"Warning cog drift with error {max_cog:.3f}. Projecting the positions down.")
            runtime.cog_fix(x, nm, self._last_cog)
        return x, h

    def set_seed(self, seed: Optional[int]) -> None:
        """Fix the Philox base seed of the samplers and restart its call counter (same seed -> same molecules)."""
        self.seed = None if seed is None else int(seed)
        self.__dict__["_seed_base"], self.__dict__["_loop_calls"] = self.seed, 0

    def _call_seed(self) -> Optional[int]:
        """Philox seed of one sampler call: a function of the fixed base seed (e.g. seed + rank) and of how many calls /
        batch chunks were sampled under it so far, or None when ``self.seed`` is unset (torch's generator is used)."""
        if self.seed is None:
            return None
        if self.__dict__.get("_seed_base") != int(self.seed):          # a new seed restarts the stream
            self.__dict__["_seed_base"], self.__dict__["_loop_calls"] = int(self.seed), 0
        calls = self.__dict__.get("_loop_calls", 0)
        self.__dict__["_loop_calls"] = calls + 1
        return (int(self.seed) * 0x9E3779B1 + calls) & ((1 << 62) - 1)

    def _seeded_noise(self, n_samples, n_nodes, node_mask, std, seed: int, draw: int):
        """Same distribution as sample_combined_position_feature_noise, drawn by the Philox kernel (gb_noise)."""
        return runtime.noise(self._flat_mask(node_mask), n_samples, n_nodes, self.n_dims + self.in_node_nf, std, seed, draw)

    def _initial_z(self, n_samples, n_nodes, node_mask, fix_noise, std, noise, seed=None):
        if noise is not None:
            return noise[0].to(torch.float32).contiguous().clone()
        if seed is not None:
            return self._seeded_noise(n_samples, n_nodes, node_mask, std, seed, 0)
        bs = 1 if fix_noise else n_samples
        z = self.sample_combined_position_feature_noise(bs, n_nodes, node_mask, std)
        return z.expand(n_samples, -1, -1).contiguous()

    def _last_noise(self, n_samples, n_nodes, node_mask, fix_noise, noise, seed=None):
        if noise is not None:
            return noise[self.T + 1]
        if fix_noise:
            return None                                  # _step_noise draws the shared sample
        if seed is not None:
            return self._seeded_noise(n_samples, n_nodes, node_mask, 1.0, seed, self.T + 1)
        return self.sample_combined_position_feature_noise(n_samples, n_nodes, node_mask)

    def _fused_loop(self, z, node_mask, edge_mask, predictor, target_w, noise, stats, seed=None):
        sched, tvals, _ = self._tables(z.device)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        nz = None if noise is None else noise.to(torch.float32).contiguous()
        runtime.sample_loop(self.dynamics, predictor, node_mask, edge_mask, z, self.T, self.T, 0, sched, tvals,
                            target_w, nz, seed, stats, self.use_cuda_graph)
        return seed

    # ---- samplers ---------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, n_samples, n_nodes, node_mask, edge_mask, context=None, fix_noise=False, std=1.0, noise=None):
        """Unconditional sampling (en_diffusion.py:958-1008).  ``noise``: optional injected [T+2,B,N,D] draws."""
        if context is not None:
            raise NotImplementedError("context conditioning is unused by GaUDI's sampling path")
        if not fix_noise:
            out = self._chunked(lambda n, nm, em, nz: self.sample(n, n_nodes, nm, em, None, False, std, nz),
                                n_samples, node_mask, edge_mask, noise, None)
            if out is not None:
                return out
        # fix_noise (one shared draw for the whole batch, a visualisation mode) keeps torch's generator
        seed = None if (noise is not None or fix_noise) else self._call_seed()
        z = self._initial_z(n_samples, n_nodes, node_mask, fix_noise, std, noise, seed)
        stats = torch.zeros(self.T, 8, dtype=torch.float32, device=z.device)
        if fix_noise:
            for s in reversed(range(self.T)):
                s_arr = torch.full((n_samples, 1), fill_value=s, device=z.device) / self.T
                z = self.sample_p_zs_given_zt(s_arr, s_arr, z, node_mask, edge_mask, None, True, None, stats[s])
        else:
            self._fused_loop(z, node_mask, edge_mask, None, None, noise, stats, seed)
        last = self._last_noise(n_samples, n_nodes, node_mask, fix_noise, noise, seed)
        self.last_stats = stats
        self._check_stats(stats)
        return self._finish(z, node_mask, edge_mask, fix_noise, last)

    # ---- batch chunking: the input-gradient pass keeps SiLU'(pre1), pre2, SiLU'(pre3) per edge, column and layer -- 2 + 4 + 2
    #      bytes on the tensor-core engine (16-bit derivative codes, fp32 pre2: 16.9 GB per 10k cc-PBH molecules; 12 bytes
    #      = 25.4 GB on the FP32 engine and up to round 2c); batches whose workspace would not fit the free HBM are sampled
    #      in sequential chunks ----
    memory_fraction = 0.7

    def _max_chunk(self, node_mask, edge_mask, predictor, guided=False) -> int:
        B = node_mask.size(0)
        if not node_mask.is_cuda:
            return B
        edges_per_mol = float(edge_mask.sum().item()) / max(B, 1)
        n = node_mask.size(1)
        hd = self.dynamics.hyper["hidden_nf"]
        per_mol = 4.0 * n * (8 * hd + 64)                                   # denoiser node buffers
        if predictor is not None or guided:
            hyper = getattr(getattr(predictor, "module", predictor), "hyper", None) or {"hidden_nf": 256, "n_layers": 12}
            hp, lp = hyper["hidden_nf"], hyper["n_layers"]
            eng = os.environ.get("GAUDI_B200_GEMM", "tc")
            sv_bytes = 8.0 if (eng == "tc" or "pred" in eng) else 12.0
            per_mol += (sv_bytes * lp + 4.0) * hp * edges_per_mol * 1.1 + 4.0 * n * hp * (lp + 12)     # + g_pre1 [edges, hp]
        free, _ = torch.cuda.mem_get_info(node_mask.device)
        reusable = sum(w.buf.numel() for w in runtime._ws.values() if w.buf is not None)
        return max(1, int(self.memory_fraction * (free + reusable) / per_mol))

    def _chunked(self, fn, n_samples, node_mask, edge_mask, noise, predictor, guided=False):
        """Run ``fn(n, node_mask, edge_mask, noise)`` on batch slices that fit in memory and concatenate (x, h)."""
        chunk = self._max_chunk(node_mask, edge_mask, predictor, guided)
        if chunk >= n_samples:
            return None
        n = node_mask.size(1)
        em = edge_mask.reshape(n_samples, n * n)
        xs, cats = [], []
        for lo in range(0, n_samples, chunk):
            hi = min(n_samples, lo + chunk)
            nz = None if noise is None else noise[:, lo:hi].contiguous()
            x, h = fn(hi - lo, node_mask[lo:hi].contiguous(), em[lo:hi].reshape(-1, 1).contiguous(), nz)
            xs.append(x); cats.append(h["categorical"])
        x = torch.cat(xs, dim=0)
        return x, {"integer": torch.zeros(n_samples, n, 0, dtype=torch.float32, device=x.device),
                   "categorical": torch.cat(cats, dim=0)}

    @torch.no_grad()
    def sample_guidance(self, n_samples, target_function: Callable, node_mask, edge_mask, scale=1, fix_noise=False,
                        std=1.0, noise=None):
        """Guided sampling (en_diffusion.py:1010-1067)."""
        predictor = getattr(target_function, "predictor", None)
        if not fix_noise:
            out = self._chunked(lambda n, nm, em, nz: self.sample_guidance(n, target_function, nm, em, scale, False, std, nz),
                                n_samples, node_mask, edge_mask, noise, predictor, guided=True)
            if out is not None:
                return out
        n_nodes = node_mask.size(1)
        seed = None if (noise is not None or fix_noise) else self._call_seed()
        z = self._initial_z(n_samples, n_nodes, node_mask, fix_noise, std, noise, seed)
        stats = torch.zeros(self.T, 8, dtype=torch.float32, device=z.device)
        if isinstance(target_function, AffineTarget) and not fix_noise:
            w = (target_function.weights * float(scale)).to(z.device).contiguous()
            self._fused_loop(z, node_mask, edge_mask, target_function.predictor, w, noise, stats, seed)
        else:
            for s in reversed(range(self.T)):
                s_arr = torch.full((n_samples, 1), fill_value=s, device=z.device) / self.T
                t_arr = torch.full((n_samples, 1), fill_value=s + 1, device=z.device) / self.T
                nz = None if noise is None else noise[self.T - s]
                if nz is None and seed is not None:
                    nz = self._seeded_noise(n_samples, n_nodes, node_mask, 1.0, seed, self.T - s)
                z = self.sample_p_zs_given_zt_guidance(s_arr, t_arr, z, node_mask, edge_mask, target_function, scale,
                                                       fix_noise, nz, stats[s])
        self.last_stats = stats
        self._check_stats(stats)
        last = self._last_noise(n_samples, n_nodes, node_mask, fix_noise, noise, seed)
        return self._finish(z, node_mask, edge_mask, fix_noise, last)

    @torch.no_grad()
    def sample_chain(self, n_samples, n_nodes, node_mask, edge_mask, context=None, keep_frames=None, std=1.0, noise=None):
        """Unguided sampling that keeps intermediate states (en_diffusion.py:1118-1174): returns the un-normalised frames
        flattened to [n_samples * keep_frames, N, D]; frame 0 is the final (x, h).  Runs the eager per-step path."""
        seed = None if noise is not None else self._call_seed()
        z = self._initial_z(n_samples, n_nodes, node_mask, False, std, noise, seed)
        keep_frames = self.T if keep_frames is None else keep_frames
        assert keep_frames <= self.T
        chain = torch.zeros((keep_frames,) + z.size(), device=z.device)
        stats = torch.zeros(self.T, 8, dtype=torch.float32, device=z.device)
        for s in reversed(range(self.T)):
            s_arr = torch.full((n_samples, 1), fill_value=s, device=z.device) / self.T
            nz = None if noise is None else noise[self.T - s]
            if nz is None and seed is not None:
                nz = self._seeded_noise(n_samples, n_nodes, node_mask, 1.0, seed, self.T - s)
            z = self.sample_p_zs_given_zt(s_arr, s_arr + 1.0 / self.T, z, node_mask, edge_mask, context, noise=nz,
                                          stats=stats[s])
            x, h_cat, _ = self.unnormalize(z[:, :, :self.n_dims], z[:, :, self.n_dims:], z[:, :, :0], node_mask)
            chain[(s * keep_frames) // self.T] = torch.cat([x, h_cat], dim=2)
        self._check_stats(stats)
        last = self._last_noise(n_samples, n_nodes, node_mask, False, noise, seed)
        x, h = self._finish(z, node_mask, edge_mask, False, last)
        chain[0] = torch.cat([x, h["categorical"], h["integer"]], dim=2)
        return chain.view(n_samples * keep_frames, *z.size()[1:])
