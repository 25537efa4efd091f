"""Multi-GPU test (needs >= 2 CUDA devices; skipped otherwise): batch sharding over NCCL ranks reproduces the
single-process result when every rank is fed its slice of the same injected noise."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, torch, torch.distributed as tdist
sys.path.insert(0, os.environ["GB_ROOT"]); sys.path.insert(0, os.path.join(os.environ["GB_ROOT"], "tests")); sys.path.insert(0, os.path.join(os.environ["GB_ROOT"], "oracle"))
import gaudi_b200 as gb
import gaudi_oracle as O
from gaudi_b200 import dist
from helpers import build_models
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dev = torch.device("cuda", local)
tdist.init_process_group("nccl", device_id=dev)
rank, ws = tdist.get_rank(), tdist.get_world_size()
args, model, pred, prop = build_models("cata", dev, hidden=(64, 64), layers=(2, 2), timesteps=30)
nx = torch.tensor([10, 9, 11, 4, 7, 10])
nm, em = O.build_masks(nx, 11, False)
gen = torch.Generator().manual_seed(11)
noise = torch.stack([O.draw_noise(len(nx), 11, 4, nm, generator=gen) for _ in range(model.T + 2)]).to(dev)
tf = gb.AffineTarget.max_gap(pred)
x, oh, m = dist.sample_guidance_sharded(args, model, tf, nx, scale=0.6, noise=noise)
assert x.shape == (6, 11, 3) and oh.shape == (6, 11, 1) and torch.equal(m.cpu(), nm)
# single-process reference on this rank
xr, ohr, _, _ = gb.sample_guidance(args, model, tf, nx, scale=0.6, noise=noise)
err = float((x - xr).abs().max()) / max(1.0, float(xr.abs().max()))
assert err < 1e-3, err
assert torch.equal(oh, ohr)
# Philox mode: different ranks draw different noise, results are finite and masked
x2, oh2, m2 = dist.sample_guidance_sharded(args, model, tf, nx, scale=0.6, seed=5)
assert torch.isfinite(x2).all() and float((x2 * (1 - m2)).abs().max()) == 0.0
tdist.destroy_process_group()
print("multi ok", rank, err)
"""


def test_sharded_guided_sampling_matches_single_process(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, GB_ROOT=ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("multi ok") == 2
