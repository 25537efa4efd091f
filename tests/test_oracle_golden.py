"""CPU tests pinning the oracle (oracle/gaudi_oracle.py) to outputs of the UNMODIFIED reference.

The fixtures in tests/golden/ were produced by tests/golden/make_golden.py, which imports /root/reference in the build
container.  Weights are not stored: the product's modules are constructed under the recorded seeds and their sha256
digests must equal the reference's, which proves identical initialisation order and state_dict layout.
"""
import numpy as np
import pytest
import torch

import gaudi_oracle as O
from helpers import META, build_models, cpu_weights, digest, golden, maxabs, oracle_cfgs, oracle_target


@pytest.fixture(scope="module", params=["cata", "hetro"])
def nets(request):
    ds = request.param
    args, model, pred, prop = build_models(ds, "cpu")
    return ds, model, pred


def test_seeded_weights_and_state_dict_layout_match_reference(nets):
    ds, model, pred = nets
    ref_d, ref_p = META[f"digest_denoiser_{ds}"], META[f"digest_predictor_{ds}"]
    got_d = {"module." + k: v for k, v in digest(model.state_dict()).items()}
    got_p = {"module." + k: v for k, v in digest(pred.state_dict()).items()}
    assert set(got_d) == set(ref_d) and len(got_d) == 141
    assert set(got_p) == set(ref_p) and len(got_p) == 160
    assert got_d == ref_d and got_p == ref_p


def test_masks_bit_exact():
    for ds in ("cata", "hetro"):
        g = golden(f"masks_{ds}.npz")
        nx = torch.from_numpy(g["nodesxsample"])
        nm, em = O.build_masks(nx, int(nx.max()), ds != "cata")
        assert torch.equal(nm, torch.from_numpy(g["node_mask"]))
        assert torch.equal(em, torch.from_numpy(g["edge_mask"]))


def test_gamma_table_matches_reference_parameter(nets):
    ds, model, pred = nets
    dcfg, _ = oracle_cfgs(ds)
    assert torch.equal(O.gamma_table(dcfg), model.gamma.gamma.detach())


@pytest.mark.parametrize("t", [1000, 500, 1])
def test_guided_and_unguided_step_match_reference(nets, t):
    ds, model, pred = nets
    g = golden(f"step_{ds}.npz")
    wd, wp = cpu_weights(model, pred)
    dcfg, pcfg = oracle_cfgs(ds)
    gamma = O.gamma_table(dcfg)
    nm, em = torch.from_numpy(g["node_mask"]), torch.from_numpy(g["edge_mask"])
    zt, noise = torch.from_numpy(g[f"zt_{t}"]), torch.from_numpy(g[f"noise_{t}"])
    out = O.guided_step(wd, dcfg, wp, pcfg, gamma, t - 1, zt, noise, nm, em, oracle_target(ds), float(g["scale"]))
    for k in ("eps", "zs_pre", "pred"):
        assert maxabs(out[k], g[f"{k}_{t}"]) == 0.0, k            # same ops, same order: bit-exact
    assert maxabs(out["grad_raw"], g[f"grad_raw_{t}"]) <= 1e-7     # autograd accumulation order may differ
    assert maxabs(out["zs"], g[f"zs_{t}"]) <= 1e-6
    un = O.unguided_step(wd, dcfg, gamma, t - 1, zt, noise, nm, em)
    assert maxabs(un["zs"], g[f"zs_unguided_{t}"]) == 0.0


def test_decode_matches_reference(nets):
    ds, model, pred = nets
    g = golden(f"step_{ds}.npz")
    wd, _ = cpu_weights(model, pred)
    dcfg, _ = oracle_cfgs(ds)
    nm, em = torch.from_numpy(g["node_mask"]), torch.from_numpy(g["edge_mask"])
    x, oh = O.decode(wd, dcfg, O.gamma_table(dcfg), torch.from_numpy(g["dec_z0"]), torch.from_numpy(g["dec_noise"]), nm, em)
    assert maxabs(x, g["dec_x"]) == 0.0
    assert torch.equal(oh, torch.from_numpy(g["dec_one_hot"])) and oh.dtype == torch.float32


def test_free_running_chain_prefix_matches_reference():
    """First 100 steps (s = 999..900) of the reference's own unguided run with the same injected noise."""
    g = golden("chain_cata_unguided.npz")
    args, model, pred, prop = build_models("cata", "cpu")
    wd, _ = cpu_weights(model, pred)
    dcfg, _ = oracle_cfgs("cata")
    gamma = O.gamma_table(dcfg)
    nm, em = torch.from_numpy(g["node_mask"]), torch.from_numpy(g["edge_mask"])
    noise = torch.from_numpy(g["noise"])
    z = noise[0]
    for k, s in enumerate(range(999, 899, -1)):
        z = O.unguided_step(wd, dcfg, gamma, s, z, noise[k + 1], nm, em)["zs"]
    ref = torch.from_numpy(g["z_900"])
    assert maxabs(z, ref) <= 1e-5 * max(1.0, float(ref.abs().max()))
    # the reference's sample_chain (en_diffusion.py:1118-1174) on the same masks / noise: frame k of keep_frames = 8 holds the
    # un-normalised z after step s = 125 k, so frame 7 (s = 875) continues this very prefix
    f = golden("chain_frames_cata.npz")
    import hashlib
    assert bytes(f["noise_digest"]) == hashlib.sha256(g["noise"].tobytes()).digest(), "fixtures must share their injected noise"
    for k, s in enumerate(range(899, 874, -1)):
        z = O.unguided_step(wd, dcfg, gamma, s, z, noise[101 + k], nm, em)["zs"]
    frame = torch.cat([z[:, :, :3] * 3.0, z[:, :, 3:] * 4.0 * nm], dim=2)           # unnormalize_z with norm_values (3, 4, 10)
    ref7 = torch.from_numpy(f["frames"][7])
    assert maxabs(frame, ref7) <= 1e-5 * max(1.0, float(ref7.abs().max()))
    assert maxabs(torch.from_numpy(f["frames"][0][:, :, :3]), torch.from_numpy(g["x"])) == 0.0    # frame 0 = the final molecules


def test_fp64_oracle_agrees_with_fp32():
    """The restatement is dtype-generic; fp64 gives the noise floor quoted in DESIGN.md."""
    g = golden("step_cata.npz")
    args, model, pred, prop = build_models("cata", "cpu")
    wd, wp = cpu_weights(model, pred)
    dcfg, pcfg = oracle_cfgs("cata")
    nm, em = torch.from_numpy(g["node_mask"]), torch.from_numpy(g["edge_mask"])
    z = torch.from_numpy(g["zt_500"])
    t = O.time_value(500, 1000)
    e32 = O.denoiser_forward(wd, dcfg, z, t, nm, em)
    e64 = O.denoiser_forward({k: v.double() for k, v in wd.items()}, dcfg, z.double(), t.double(), nm.double(), em.double())
    assert maxabs(e32, e64) <= 1e-5


@pytest.mark.parametrize("ds", ["cata", "hetro"])
def test_training_loss_and_gradients_match_reference(ds):
    """Row a19: loss [B] of the reference's train-mode forward (t_int / eps pinned, incl. t = 0 and t = T samples) and
    its parameter gradients (all 139 norms, 7 full tensors)."""
    g = golden(f"train_{ds}.npz")
    args, model, pred, prop = build_models(ds, "cpu")
    dcfg, _ = oracle_cfgs(ds)
    w = {k: v.detach().clone().requires_grad_(not k.endswith("gamma.gamma")) for k, v in model.state_dict().items()}
    nm, em = O.build_masks(torch.from_numpy(g["nodesxsample"]), 11 if ds == "cata" else 10, ds == "hetro")
    loss, _ = O.training_loss(w, dcfg, O.gamma_table(dcfg), torch.from_numpy(g["x"]), torch.from_numpy(g["h"]), nm, em,
                              torch.from_numpy(g["t_int"]).float(), torch.from_numpy(g["eps"]))
    assert maxabs(loss, g["loss_b"]) <= 1e-6
    loss.mean(0).backward()
    assert abs(float(loss.mean(0)) - float(g["loss"])) <= 1e-6
    for name, norm in zip(g["grad_names"], g["grad_norms"]):
        assert abs(float(w[str(name)].grad.double().norm()) - norm) <= 1e-5 * max(norm, 1e-6), name
    for k in g.files:
        if k.startswith("grad:"):
            ref = torch.from_numpy(g[k])
            assert maxabs(w[k[5:]].grad, ref) <= 1e-6 * max(1.0, float(ref.abs().max())), k


@pytest.mark.parametrize("ds", ["cata", "hetro"])
def test_predictor_training_loss_and_gradients_match_reference(ds):
    """8f rank 2: sample_edm_t + l1 loss of the reference's predictor training step and its parameter gradients."""
    g = golden(f"pred_train_{ds}.npz")
    args, model, pred, prop = build_models(ds, "cpu")
    dcfg, pcfg = oracle_cfgs(ds)
    w = {k: v.detach().clone().requires_grad_(True) for k, v in pred.state_dict().items()}
    nm, em = O.build_masks(torch.from_numpy(g["nodesxsample"]), 11 if ds == "cata" else 10, ds == "hetro")
    t = torch.from_numpy(g["t_int"]).float() / dcfg.timesteps
    zt = O.sample_edm_t(dcfg, O.gamma_table(dcfg), torch.from_numpy(g["x"]), torch.from_numpy(g["h"]), nm, t, torch.from_numpy(g["eps"]))
    assert maxabs(zt, g["z_t"]) == 0.0
    p = O.predictor_forward(w, pcfg, zt, nm, em, t)
    loss = torch.nn.functional.l1_loss(p, torch.from_numpy(g["y"]))
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-6
    assert maxabs((p.detach() - torch.from_numpy(g["y"])).abs(), g["abs_err"]) <= 1e-6
    loss.backward()
    for name, norm in zip(g["grad_names"], g["grad_norms"]):
        assert abs(float(w[str(name)].grad.double().norm()) - norm) <= 1e-5 * max(norm, 1e-6), name
    for k in g.files:
        if k.startswith("grad:"):
            ref = torch.from_numpy(g[k])
            assert maxabs(w[k[5:]].grad, ref) <= 1e-6 * max(1.0, float(ref.abs().max())), k


@pytest.mark.parametrize("ds", ["cata", "hetro"])
def test_eval_mode_nll_matches_reference(ds):
    """forward() in eval mode (compute_loss(t0_always=True), the bound val_epoch reports) with pinned draws."""
    g = golden(f"nll_{ds}.npz")
    args, model, pred, prop = build_models(ds, "cpu")
    dcfg, _ = oracle_cfgs(ds)
    wd, _ = cpu_weights(model, pred)
    nm, em = O.build_masks(torch.from_numpy(g["nodesxsample"]), 11 if ds == "cata" else 10, ds == "hetro")
    with torch.no_grad():
        nll = O.validation_nll(wd, dcfg, O.gamma_table(dcfg), torch.from_numpy(g["x"]), torch.from_numpy(g["h"]), nm, em,
                               torch.from_numpy(g["t_int"]).float(), torch.from_numpy(g["eps"]), torch.from_numpy(g["eps0"]))
    assert maxabs(nll, g["nll"]) <= 1e-6 * float(np.abs(g["nll"]).max())
