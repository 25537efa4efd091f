"""Geometric validity: oracle vs the reference's golden flags on CPU; CUDA kernel vs golden and
oracle on the GPU.  Flags, adjacency and masks are integer results: the bar is bit-exact."""
import numpy as np
import pytest
import torch

import molgen
import validity_oracle as VO
from helpers import golden


def _oracle_flags(x, rt, nm, ds):
    out = np.zeros((x.shape[0], 5), np.uint8)
    for b in range(x.shape[0]):
        m = nm[b].astype(bool)
        out[b] = VO.flags_of(VO.check_stability(torch.from_numpy(x[b][m]), torch.from_numpy(rt[b][m]), 0.1, ds))[:5]
    return out


@pytest.mark.parametrize("ds", ["cata", "hetro"])
def test_oracle_matches_reference_golden(ds):
    g = golden(f"validity_{ds}.npz")
    n = 150                                                   # the python oracle is slow; the GPU test covers the rest
    assert np.array_equal(_oracle_flags(g["x"][:n], g["ring_type"][:n], g["node_mask"][:n], ds), g["flags"][:n])
    dist, adj = VO.positions2adj(torch.from_numpy(g["x"][:16]), torch.from_numpy(g["ring_type"][:16]), 0.1, ds)
    assert np.array_equal(dist.numpy(), g["dist16"]) and np.array_equal(adj.numpy(), g["adj16"])
    first = 0 if ds == "hetro" else 1                           # cata has no orientation nodes: that flag is always True
    assert g["flags"][:, first:].min(0).max() == 0 and g["flags"].max(0).min() == 1     # every flag takes both values


@pytest.mark.gpu
@pytest.mark.parametrize("ds", ["cata", "hetro"])
def test_kernel_matches_reference_golden(ds):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    from gaudi_b200 import analyze
    dev = torch.device("cuda:0")
    g = golden(f"validity_{ds}.npz")
    x, rt, nm = (torch.from_numpy(g[k]).to(dev) for k in ("x", "ring_type", "node_mask"))
    flags = analyze.check_stability_batch(x, rt, nm, 0.1, ds).cpu().numpy()
    assert int(flags[:, 6].max()) == 0
    assert np.array_equal(flags[:, :5], g["flags"])
    assert np.array_equal(flags[:, 5], g["flags"].all(1).astype(np.uint8))
    one_hot = torch.nn.functional.one_hot(rt, len(analyze.RINGS_LIST[ds])).float()        # one-hot input, like the sampler returns
    assert torch.equal(analyze.check_stability_batch(x, one_hot, nm.unsqueeze(2), 0.1, ds)[:, :6].cpu(), torch.from_numpy(flags[:, :6]))
    dist, adj = analyze.positions2adj(x[:16], rt[:16], 0.1, ds)
    assert np.array_equal(adj.cpu().numpy(), g["adj16"])
    assert np.abs(dist.cpu().numpy() - g["dist16"]).max() <= 1e-6
    s = analyze.validity_summary(torch.from_numpy(flags))
    assert abs(s["mol_stable"] - g["flags"].all(1).mean()) < 1e-12 and abs(s["connected"] - g["flags"][:, 2].mean()) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("ds,seed", [("cata", 7), ("hetro", 8)])
def test_kernel_matches_oracle_and_list_api(ds, seed):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    from gaudi_b200 import analyze
    dev = torch.device("cuda:0")
    x, rt, nm = molgen.batch(seed, ds, 120, 11 if ds == "cata" else 10)
    ref = _oracle_flags(x, rt, nm, ds)
    flags = analyze.check_stability_batch(torch.from_numpy(x).to(dev), torch.from_numpy(rt).to(dev), torch.from_numpy(nm).to(dev), 0.1, ds)
    assert np.array_equal(flags[:, :5].cpu().numpy(), ref)
    mols = [(x[b][nm[b].astype(bool)], rt[b][nm[b].astype(bool)]) for b in range(40)]
    vd, stable = analyze.analyze_validity_for_molecules(mols, tol=0.1, dataset=ds)
    assert vd["molecule_stable_bool"] == [bool(r.all()) for r in ref[:40]] and len(stable) == int(ref[:40].all(1).sum())
    for b in (0, 1, 2):
        res = analyze.check_stability(*mols[b], tol=0.1, dataset=ds)
        assert [int(res[k]) for k in analyze.FLAG_NAMES] == ref[b].tolist()
    empty = (np.zeros((0, 3), np.float32), np.zeros((0,), np.int64))
    if ds == "cata":                      # networkx raises on the null graph; hetro fails the orientation test first
        with pytest.raises(ValueError):
            analyze.check_stability(*empty, dataset=ds)
    else:
        assert analyze.check_stability(*empty, dataset=ds) == VO.check_stability(*empty, dataset=ds)


@pytest.mark.gpu
def test_large_batch_properties():
    """Size-independent properties at the sampler's batch size: permutation of the batch permutes the flags; a rigid
    rotation + translation of every molecule leaves the distance-derived flags unchanged."""
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    from gaudi_b200 import analyze
    dev = torch.device("cuda:0")
    x, rt, nm = molgen.batch(3, "cata", 500, 11)
    reps = 20
    X, R, M = (torch.from_numpy(np.tile(a, (reps,) + (1,) * (a.ndim - 1))).to(dev) for a in (x, rt, nm))
    f = analyze.check_stability_batch(X, R, M, 0.1, "cata")
    assert torch.equal(f[:500], f[-500:])
    perm = torch.randperm(X.shape[0], device=dev)
    assert torch.equal(analyze.check_stability_batch(X[perm], R[perm], M[perm], 0.1, "cata"), f[perm])
    shift = torch.tensor([0.5, -0.25, 0.125], device=dev)
    f2 = analyze.check_stability_batch((X + shift) * M.unsqueeze(2), R, M, 0.1, "cata")
    assert (f2[:, :3] != f[:, :3]).float().mean() < 1e-3
