"""GPU tests of the stand-alone sub-module forwards (GCL, EquivariantUpdate, EquivariantBlock, EGNN, E_GCL, predictor
EGNN) against a plain PyTorch fp32 restatement of the same op (oracle/gaudi_oracle.py), tolerance 1e-4 max-abs."""
import pytest
import torch

import gaudi_b200 as gb
import gaudi_oracle as O
from helpers import build_models, maxabs

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _setup(dataset="cata"):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    dev = torch.device("cuda:0")
    args, model, pred, prop = build_models(dataset, dev)
    nx = torch.tensor([10, 7, 11, 3, 9])
    nm, em = O.build_masks(nx, 11, dataset != "cata")
    B, N = nm.shape[0], nm.shape[1]
    row, col = O.dense_edges(B, N)
    gen = torch.Generator().manual_seed(21)
    nmf, emf = nm.reshape(B * N, 1), em.reshape(-1, 1)
    x = torch.randn(B * N, 3, generator=gen) * nmf
    return dev, model, pred, B, N, row, col, nmf, emf, x, gen


def test_denoiser_submodules_match_torch_reference():
    dev, model, pred, B, N, row, col, nm, em, x, gen = _setup()
    egnn = model.dynamics.egnn
    w = {k: v.detach().cpu() for k, v in egnn.state_dict().items()}
    H = 192
    h = torch.randn(B * N, H, generator=gen) * 0.5 * nm
    d0, _ = O._radial(x * 1.3, row, col, 1.0)
    r, u = O._radial(x, row, col, 1.0)
    eattr = torch.cat([r, d0], dim=1)
    edges = [row.to(dev), col.to(dev)]
    to = lambda t: t.to(dev)
    blk = egnn.e_block_3
    # GCL
    got, mij = blk.gcl_0(to(h), edges, edge_attr=to(eattr), node_mask=to(nm), edge_mask=to(em))
    ref = O.gcl_layer(w, "e_block_3.gcl_0.", h, row, col, eattr, nm, em)
    assert mij is None and maxabs(got, ref) <= TOL
    # EquivariantUpdate
    got = blk.gcl_equiv(to(h), to(x), edges, to(u), to(eattr), to(nm), to(em))
    ref = O.equiv_update_layer(w, "e_block_3.gcl_equiv.", h, x, row, col, u, eattr, nm, em, 4.0)
    assert maxabs(got, ref) <= TOL
    # EquivariantBlock
    gh, gx = blk(to(h), to(x), edges, node_mask=to(nm), edge_mask=to(em), edge_attr=to(d0))
    rh, rx = O.equiv_block(w, "e_block_3.", h, x, row, col, d0, nm, em, 4.0)
    assert maxabs(gh, rh) <= TOL and maxabs(gx, rx) <= TOL
    # EGNN (whole stack on flattened inputs)
    hin = torch.cat([torch.randn(B * N, 1, generator=gen) * nm, torch.full((B * N, 1), 0.37)], dim=1)
    gh, gx = egnn(to(hin), to(x), edges, node_mask=to(nm), edge_mask=to(em))
    hh = O._lin(w, "embedding", hin)
    xx = x
    dd0, _ = O._radial(x, row, col, 1.0)
    for b in range(9):
        hh, xx = O.equiv_block(w, f"e_block_{b}.", hh, xx, row, col, dd0, nm, em, 4.0)
    hh = O._lin(w, "embedding_out", hh) * nm
    assert maxabs(gh, hh) <= TOL and maxabs(gx, xx) <= TOL


def test_predictor_submodules_match_torch_reference():
    dev, model, pred, B, N, row, col, nm, em, x, gen = _setup()
    egnn = pred.egnn
    w = {k: v.detach().cpu() for k, v in egnn.state_dict().items()}
    H = 196
    h = torch.randn(B * N, H, generator=gen) * 0.5 * nm
    a = torch.sum((x[row] - x[col]) ** 2, dim=1, keepdim=True) * 0.8
    edges = [row.to(dev), col.to(dev)]
    to = lambda t: t.to(dev)
    gh, gx, ga = egnn.gcl_5(to(h), edges, to(x), edge_attr=to(a), node_mask=to(nm), edge_mask=to(em))
    rh, rx = O.e_gcl_layer(w, "gcl_5.", h, x, row, col, a, nm, em, 4.0 / 12)
    assert maxabs(gh, rh) <= TOL and maxabs(gx, rx) <= TOL and torch.equal(ga.cpu(), a)
    hin = torch.cat([torch.randn(B * N, 1, generator=gen) * nm, torch.full((B * N, 1), 0.61)], dim=1)
    gh, gx = egnn(to(hin), to(x), edges, edge_attr=to(a), node_mask=to(nm), edge_mask=to(em))
    hh, xx = O._lin(w, "embedding", hin), x
    for l in range(12):
        hh, xx = O.e_gcl_layer(w, f"gcl_{l}.", hh, xx, row, col, a, nm, em, 4.0 / 12)
    hh = O._lin(w, "embedding_out", hh) * nm
    assert maxabs(gh, hh) <= TOL and maxabs(gx, xx) <= TOL
