"""Multi-GPU: the molecule batch shards trivially (no op couples molecules).

One process per GPU (``torch.distributed``, NCCL over NVLink on the GPU box, gloo in the CPU tests).  Each rank
samples its contiguous chunk of ``nodesxsample`` with its own seed (``seed + rank``) and a single all-gather of the
generated coordinates, one-hot ring types and node masks follows the loop; there is no per-step communication
(the reference has no distributed sampling at all: ``MyDataParallel`` forwards ``sample_guidance`` to the bare
module, models_edm.py:13-18).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as tdist


def shard_bounds(total: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, near-equal chunks [lo, hi) per rank; the first ``total % world`` ranks get one extra item."""
    base, extra = divmod(total, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def world() -> Tuple[int, int]:
    if tdist.is_available() and tdist.is_initialized():
        return tdist.get_rank(), tdist.get_world_size()
    return 0, 1


def gather_padded(t: torch.Tensor, counts: Sequence[int], pad_nodes: Optional[int] = None) -> torch.Tensor:
    """all_gather of per-rank tensors [b_r, N_r, C] into [sum b_r, N_max, C] (zero padded) on every rank."""
    rank, ws = world()
    if ws == 1:
        return t
    bmax = max(counts)
    nmax = pad_nodes if pad_nodes is not None else t.shape[1]
    buf = t.new_zeros((bmax, nmax) + tuple(t.shape[2:]))
    buf[: t.shape[0], : t.shape[1]] = t
    outs = [torch.empty_like(buf) for _ in range(ws)]
    tdist.all_gather(outs, buf.contiguous())
    return torch.cat([o[:c] for o, c in zip(outs, counts)], dim=0)


def sample_guidance_sharded(args, model, target_function, nodesxsample: torch.Tensor, scale=1.0, std=1.0,
                            seed: int = 0, noise=None, sampler=None, timing: Optional[dict] = None):
    """``sampling_edm.sample_guidance`` over all ranks: returns the FULL (x, one_hot, node_mask) on every rank.

    ``sampler`` defaults to ``gaudi_b200.sampling.sample_guidance``; ``noise`` (parity mode) is the injected
    [T+2, B_total, N, D] tensor, sliced per rank.  ``timing`` (optional dict) receives ``gather_ms`` (CUDA-event time of the
    three all_gathers on this rank) and ``gather_bytes`` (payload this rank contributes).
    """
    from . import sampling
    rank, ws = world()
    bounds = shard_bounds(len(nodesxsample), ws)
    lo, hi = bounds[rank]
    counts = [b - a for a, b in bounds]
    nmax_global = int(torch.as_tensor(nodesxsample).max().item())
    local = nodesxsample[lo:hi]
    fn = sampler or sampling.sample_guidance
    if hi > lo:
        inner = getattr(model, "module", model)
        if noise is None:                              # per-rank Philox stream for z_T, every step and the final decode
            if hasattr(inner, "set_seed"):
                inner.set_seed(int(seed) + rank)
            else:
                inner.seed = int(seed) + rank
        # every shard pads to the ring count of the WHOLE batch (sampling_edm.py:177): the predictor's mean pooling runs over
        # the padded nodes (egnn_predictor/models.py:456-457), so local padding would rescale the guidance gradient by
        # nmax_global / nmax_local (1.10x for a shard whose largest molecule has 10 of 11 rings)
        local_noise = None if noise is None else noise[:, lo:hi].contiguous()
        x, one_hot, node_mask, _ = fn(args, model, target_function, local, scale=scale, std=std, noise=local_noise,
                                      max_nodes=nmax_global)
    else:
        dev = args.device
        x = torch.zeros(0, 1, 3, device=dev); one_hot = torch.zeros(0, 1, 1, device=dev); node_mask = torch.zeros(0, 1, 1, device=dev)
    orient = 2 if args.dataset != "cata" else 1
    pad = nmax_global * orient
    if orient == 2 and x.shape[1] != pad and x.shape[0] > 0:
        # ring nodes then orientation nodes: re-pad each half separately
        h = x.shape[1] // 2
        def repad(t):
            out = t.new_zeros((t.shape[0], pad) + tuple(t.shape[2:]))
            out[:, :h] = t[:, :h]; out[:, nmax_global:nmax_global + h] = t[:, h:]
            return out
        x, one_hot, node_mask = repad(x), repad(one_hot), repad(node_mask)
    F = one_hot.shape[2]
    if ws > 1:                                        # empty ranks must still agree on the feature width
        ft = torch.tensor([F], device=x.device)
        tdist.all_reduce(ft, op=tdist.ReduceOp.MAX)
        F = int(ft.item())
        if one_hot.shape[0] == 0:
            one_hot = one_hot.new_zeros(0, 1, F)
    if timing is not None and x.is_cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    out = (gather_padded(x, counts, pad), gather_padded(one_hot, counts, pad), gather_padded(node_mask, counts, pad))
    if timing is not None:
        timing["gather_bytes"] = 4 * max(counts) * pad * (3 + F + 1)
        if x.is_cuda:
            e1.record()
            torch.cuda.synchronize()
            timing["gather_ms"] = e0.elapsed_time(e1)
    return out


def average_gradients(parameters) -> int:
    """Data-parallel training step (each rank runs forward/backward on its shard of the batch): ONE all-reduce of all
    gradients as a single flat bucket (3 M parameters = 12 MB: latency-, not bandwidth-sized), then the mean is copied
    back.  The reference trains with single-process ``DataParallel`` (models_edm.py:13-18); this is its one-process-per-GPU
    equivalent.  Returns the number of parameters reduced; a no-op outside ``torch.distributed``."""
    rank, ws = world()
    grads = [p.grad for p in parameters if p.grad is not None]
    if ws == 1 or not grads:
        return sum(g.numel() for g in grads)
    flat = torch.cat([g.reshape(-1) for g in grads])
    tdist.all_reduce(flat, op=tdist.ReduceOp.SUM)
    flat /= ws
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return off
