"""Host-side pieces of the training loops: adaptive gradient clipping and experiment loading.

  Queue, gradient_clipping      edm/utils.py:31-70 (used by train_edm.py:77-79)
  save_model, load_model        edm/utils.py:20-27
  get_edm_args, get_cond_predictor_args   utils/helpers.py:204-224 (args.txt of an experiment directory)
  FusedAdamWClip                the two together as ONE fused device-side step (gb_adamw_amsgrad_clip): global norm, adaptive clip
                                window, AdamW(amsgrad) over a flat parameter bucket, no host synchronisation; `fit` uses it on CUDA.
``Queue`` / ``gradient_clipping`` + ``torch.optim.AdamW`` (the reference's own combination) remain available.
"""
from __future__ import annotations

import json
from argparse import Namespace

import numpy as np
import torch

from .sampling import args_edm, prediction_args


class Queue:
    """Sliding window (newest first) of recent gradient norms."""

    def __init__(self, max_len: int = 50):
        self.items = []
        self.max_len = max_len

    def __len__(self):
        return len(self.items)

    def add(self, item) -> None:
        self.items.insert(0, item)
        del self.items[self.max_len:]

    def mean(self):
        return np.mean(self.items)

    def std(self):
        return np.std(self.items)


def gradient_clipping(flow, gradnorm_queue: Queue):
    """Clip to 1.5 x mean + 2 x std of the recent norms; the queue records the norm actually applied."""
    max_grad_norm = 1.5 * gradnorm_queue.mean() + 2 * gradnorm_queue.std()
    grad_norm = torch.nn.utils.clip_grad_norm_(flow.parameters(), max_norm=max_grad_norm, norm_type=2.0)
    clipped = float(grad_norm) > max_grad_norm
    gradnorm_queue.add(float(max_grad_norm) if clipped else float(grad_norm))
    if clipped:
        print(f"Clipped gradient with value {grad_norm:.1f} while allowed {max_grad_norm:.1f}")
    return grad_norm


class FusedAdamWClip:
    """``gradient_clipping`` (edm/utils.py:51-70) + ``torch.optim.AdamW(amsgrad=True).step()`` (train_edm.py:22-24, 71-82) as one
    device-side step over a flat bucket: the parameters (and their gradients) are re-pointed to views of two contiguous fp32
    buffers, so the global norm, the clip window (1.5 mean + 2 std of the last ``window`` applied norms, seeded with
    ``first_norm`` = 3000 as in train_edm.py) and the update are three kernels and no ``float(grad_norm)`` host sync.
    ``clip=False`` is plain AdamW-amsgrad.  ``last_grad_norm`` / ``last_max_norm`` are device scalars (read them lazily)."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-12, clip=True, window=50, first_norm=3000.0):
        from . import _lib
        self.params = [p for p in params if p.requires_grad]
        if not self.params or not all(p.is_cuda and p.dtype == torch.float32 for p in self.params):
            raise RuntimeError("FusedAdamWClip needs fp32 CUDA parameters (the sm_100a kernels have no CPU fallback)")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            p.grad = self.flat_g[off:off + k].view_as(p)
            off += k
        self.m, self.v, self.vmax = (torch.zeros_like(self.flat_p) for _ in range(3))
        L = _lib.lib()
        self.window = int(window)
        self.state = torch.zeros(int(L.gb_adamw_state_doubles(self.window)), dtype=torch.float64, device=dev)
        if clip:
            self.state[4] = 1.0
            self.state[5] = 1.0 % self.window
            self.state[8] = float(first_norm)
        self.scratch = torch.zeros(int(L.gb_adamw_scratch_doubles()), dtype=torch.float64, device=dev)
        self.lr, self.betas, self.eps, self.weight_decay, self.clip = float(lr), betas, float(eps), float(weight_decay), bool(clip)

    def zero_grad(self):
        self.flat_g.zero_()                              # gradients stay views of the flat bucket (set_to_none would detach them)

    @torch.no_grad()
    def step(self):
        from . import _lib
        from .runtime import _ptr, _stream
        for p in self.params:                            # autograd may have replaced a .grad view (e.g. after set_to_none)
            if p.grad is None or p.grad.data_ptr() < self.flat_g.data_ptr() or p.grad.data_ptr() >= self.flat_g.data_ptr() + 4 * self.flat_g.numel():
                raise RuntimeError("FusedAdamWClip: a parameter's .grad no longer lives in the flat bucket; use this optimizer's zero_grad()")
        with torch.cuda.device(self.flat_p.device):
            _lib.check(_lib.lib().gb_adamw_amsgrad_clip(_ptr(self.flat_p), _ptr(self.flat_g), _ptr(self.m), _ptr(self.v), _ptr(self.vmax),
                                                        self.flat_p.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                                        _ptr(self.state), self.window, int(self.clip), _ptr(self.scratch), _stream()))

    @property
    def last_grad_norm(self) -> torch.Tensor:
        return self.state[1]

    @property
    def last_max_norm(self) -> torch.Tensor:
        return self.state[2]


def save_model(model, path) -> None:
    torch.save(model.state_dict(), path)


def load_model(model, path):
    model.load_state_dict(torch.load(path))
    model.eval()
    return model


def _load_args(preset: Namespace, exp_dir_path: str) -> Namespace:
    with open(exp_dir_path + "/args.txt", "r") as f:
        preset.__dict__ = json.load(f)
    preset.restore = True
    preset.exp_dir = exp_dir_path
    preset.device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    return preset


def get_edm_args(exp_dir_path: str) -> Namespace:
    return _load_args(args_edm(), exp_dir_path)


def get_cond_predictor_args(exp_dir_path: str) -> Namespace:
    return _load_args(prediction_args(), exp_dir_path)


# --------------------------------------------------------------------------------------------------------------------
# epoch loops of train_edm.py (:36-49 compute_loss, :52-92 train_epoch, :95-141 val_epoch, :144-191 main), dataset-agnostic:
# a loader is any iterable of (x [B,N,3], node_mask [B,N], edge_mask [B,N,N] or flat, node_features [B,N,F], y)
# --------------------------------------------------------------------------------------------------------------------
def remove_mean_with_mask(x: torch.Tensor, node_mask: torch.Tensor) -> torch.Tensor:
    n = node_mask.sum(1, keepdim=True).clamp(min=1)
    return x - x.sum(1, keepdim=True) / n * node_mask


def compute_loss(model, x, h, node_mask, edge_mask) -> torch.Tensor:
    bs, n_nodes, _ = x.size()
    assert float((x * (1 - node_mask)).abs().max()) < 1e-4, "Variables not masked properly."
    h = {"categorical": h, "integer": torch.zeros(0, device=x.device)}
    return model(x, h, node_mask, edge_mask.view(bs, n_nodes * n_nodes)).mean(0)


def _batch_to(batch, device):
    x, node_mask, edge_mask, h = batch[0].to(device), batch[1].to(device).unsqueeze(2), batch[2].to(device), batch[3].to(device)
    return remove_mean_with_mask(x, node_mask), h, node_mask, edge_mask


def train_epoch(model, loader, optimizer, device, gradnorm_queue: Queue = None, clip_grad: bool = True):
    """One pass of train_edm.train_epoch: loss -> backward -> adaptive clipping -> optimizer step.  Returns (mean loss, mean grad norm).
    With a ``FusedAdamWClip`` optimizer clipping is part of its step and losses / norms are read back once per epoch."""
    model.train()
    fused = isinstance(optimizer, FusedAdamWClip)
    losses, norms = [], []
    for batch in loader:
        x, h, node_mask, edge_mask = _batch_to(batch, device)
        loss = compute_loss(model, x, h, node_mask, edge_mask)
        optimizer.zero_grad()
        loss.backward()
        if fused:
            optimizer.step()
            losses.append(loss.detach())
            if optimizer.clip:
                norms.append(optimizer.last_grad_norm.clone())
            continue
        if clip_grad and gradnorm_queue is not None:
            norms.append(float(gradient_clipping(model, gradnorm_queue)))
        optimizer.step()
        losses.append(float(loss.detach()))
    if fused:
        losses = [float(v) for v in torch.stack(losses).cpu()]
        norms = [float(v) for v in torch.stack(norms).cpu()] if norms else []
    return float(np.mean(losses)), (float(np.mean(norms)) if norms else float("nan"))


@torch.no_grad()
def val_epoch(model, loader, device) -> float:
    """train_edm.val_epoch without the sampling side effects: mean of the eval-mode bound over the loader."""
    model.eval()
    losses = []
    for batch in loader:
        x, h, node_mask, edge_mask = _batch_to(batch, device)
        losses.append(float(compute_loss(model, x, h, node_mask, edge_mask)))
    return float(np.mean(losses))


def fit(model, train_loader, val_loader, device, num_epochs: int, lr: float = 1e-4, exp_dir: str = None, clip_grad: bool = True,
        fused: bool = None):
    """train_edm.main's loop: AdamW(amsgrad, weight_decay 1e-12), gradient-norm queue seeded with 3000, checkpoint
    (`exp_dir/model.pt`) whenever the validation bound improves; the best weights are restored at the end.
    ``fused`` (default: on CUDA) runs clipping + AdamW as the fused device-side step (``FusedAdamWClip``)."""
    trainable = [p for p in model.parameters() if p.requires_grad]
    if fused is None:
        fused = bool(trainable) and trainable[0].is_cuda
    if fused:
        optimizer = FusedAdamWClip(trainable, lr=lr, weight_decay=1e-12, clip=clip_grad, window=50, first_norm=3000.0)
    else:
        optimizer = torch.optim.AdamW(trainable, lr=lr, weight_decay=1e-12, amsgrad=True)
    queue = Queue(max_len=50)
    queue.add(3000)
    best_val, best_epoch, best_state, history = 1e9, 0, None, []
    for epoch in range(num_epochs):
        tr, gn = train_epoch(model, train_loader, optimizer, device, queue, clip_grad)
        va = val_epoch(model, val_loader, device)
        history.append({"epoch": epoch, "train_loss": tr, "grad_norm": gn, "val_loss": va})
        if va < best_val:
            best_val, best_epoch = va, epoch
            if exp_dir is not None:
                save_model(model, exp_dir + "/model.pt")
            else:
                best_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    if exp_dir is not None:
        model.load_state_dict(torch.load(exp_dir + "/model.pt"))
    elif best_state is not None:
        model.load_state_dict(best_state)
    return {"best_epoch": best_epoch, "best_val_loss": best_val, "history": history}
