"""Host-side pieces of the training loops: adaptive gradient clipping and experiment loading.

  Queue, gradient_clipping      edm/utils.py:31-70 (used by train_edm.py:77-79)
  save_model, load_model        edm/utils.py:20-27
  get_edm_args, get_cond_predictor_args   utils/helpers.py:204-224 (args.txt of an experiment directory)
Gradient-norm clipping and AdamW(amsgrad) stay torch utilities, exactly as in the reference.
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
    """One pass of train_edm.train_epoch: loss -> backward -> adaptive clipping -> optimizer step.  Returns (mean loss, mean grad norm)."""
    model.train()
    losses, norms = [], []
    for batch in loader:
        x, h, node_mask, edge_mask = _batch_to(batch, device)
        loss = compute_loss(model, x, h, node_mask, edge_mask)
        optimizer.zero_grad()
        loss.backward()
        if clip_grad and gradnorm_queue is not None:
            norms.append(float(gradient_clipping(model, gradnorm_queue)))
        optimizer.step()
        losses.append(float(loss.detach()))
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


def fit(model, train_loader, val_loader, device, num_epochs: int, lr: float = 1e-4, exp_dir: str = None, clip_grad: bool = True):
    """train_edm.main's loop: AdamW(amsgrad, weight_decay 1e-12), gradient-norm queue seeded with 3000, checkpoint
    (`exp_dir/model.pt`) whenever the validation bound improves; the best weights are restored at the end."""
    optimizer = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=lr, weight_decay=1e-12, amsgrad=True)
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
