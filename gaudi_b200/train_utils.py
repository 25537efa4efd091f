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
