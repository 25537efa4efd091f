"""Generate golden fixtures by running the UNMODIFIED reference (tomer196/GaUDI).

Run in the build container only (needs /root/reference; the GPU box has no copy):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Writes small .npz / .json fixtures next to this file.  Weights are never stored:
both networks are PyTorch-default-initialised under a fixed seed, and the
fixtures carry a sha256 per parameter so the tests can prove that the product's
own modules (constructed under the same seed) hold bit-identical weights.

Recipe follows SURVEY.md appendix C: rdkit/matplotlib/imageio are stubbed with
MagicMock (they are only touched by plotting / chemistry post-processing).
"""
import hashlib
import json
import os
import sys
from argparse import Namespace
from unittest.mock import MagicMock

import numpy as np
import torch

REF = os.environ.get("GAUDI_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
for _m in ["rdkit", "rdkit.Chem", "rdkit.Chem.Draw", "matplotlib", "matplotlib.pyplot", "imageio", "torch.utils.tensorboard"]:
    sys.modules[_m] = MagicMock()

from utils.args_edm import Args_EDM  # noqa: E402
from cond_prediction.prediction_args import PredictionArgs  # noqa: E402
from models_edm import get_model  # noqa: E402
from cond_prediction.train_cond_predictor import get_cond_predictor_model  # noqa: E402
from utils.helpers import switch_grad_off  # noqa: E402
import sampling_edm  # noqa: E402
from edm.equivariant_diffusion.utils import remove_mean_with_mask  # noqa: E402

torch.set_num_threads(8)
SEED_DEN, SEED_PRED = 0, 1
STD5 = torch.tensor([0.7, 1.3, 0.9, 1.1, 0.8])
MEAN5 = torch.tensor([0.2, -0.4, 0.1, 0.3, -0.1])


def build(dataset):
    F_in = 1 if dataset == "cata" else 12
    args = Args_EDM().parse_args([])
    args.device, args.exp_dir, args.dataset = "cpu", None, dataset
    args.max_nodes = 11 if dataset == "cata" else 10
    pargs = PredictionArgs().parse_args([])
    pargs.device = "cpu"
    ds = Namespace(num_node_features=F_in, num_targets=5, mean=MEAN5.clone(), std=STD5.clone())
    torch.manual_seed(SEED_DEN)
    model, _, prop = get_model(args, Namespace(dataset=ds), only_norm=True)
    torch.manual_seed(SEED_PRED)
    pred = get_cond_predictor_model(pargs, ds)
    switch_grad_off([model, pred])
    return args, model, pred, prop


def digest(sd):
    return {k: hashlib.sha256(v.detach().cpu().contiguous().numpy().tobytes()).hexdigest() for k, v in sd.items()}


def targets(pred, prop):
    def max_gap(_input, _node_mask, _edge_mask, _t):          # generation_guidance.py:200-203
        return -pred(_input, _node_mask, _edge_mask, _t)[:, 1]

    def opv(_input, _node_mask, _edge_mask, _t):              # generation_guidance.py:205-211
        p = prop.unnormalize(pred(_input, _node_mask, _edge_mask, _t))
        return p[:, 3] + p[:, 2] + 3 * p[:, 0]
    return {"max_gap": max_gap, "opv": opv}


def ref_masks(args, nodesxsample):
    """Masks exactly as sampling_edm.sample_guidance builds them (captured from its model call)."""
    got = {}

    class Fake:
        def sample_guidance(self, bs, tf, node_mask, edge_mask, scale, fix_noise=False, std=1.0):
            got["nm"], got["em"] = node_mask.clone(), edge_mask.clone()
            n = node_mask.shape[1]
            return torch.zeros(bs, n, 3), {"categorical": torch.zeros(bs, n, 1), "integer": torch.zeros(0)}

    sampling_edm.sample_guidance(args, Fake(), None, nodesxsample)
    return got["nm"], got["em"]


def processed_noise(gen, shape_bnd, node_mask, std=1.0):
    B, N, D = shape_bnd
    zx = torch.randn((B, N, 3), generator=gen) * std * node_mask
    zx = remove_mean_with_mask(zx, node_mask)
    zh = torch.randn((B, N, D - 3), generator=gen) * std * node_mask
    return torch.cat([zx, zh], dim=2)


def data_like_z(gen, model, node_mask, F_in, t_int):
    """z_t = alpha_t*xh + sigma_t*eps for a synthetic, normalised ring-graph-like xh."""
    B, N, _ = node_mask.shape
    x = remove_mean_with_mask(torch.randn((B, N, 3), generator=gen) * node_mask, node_mask)
    cls = torch.randint(0, F_in, (B, N), generator=gen)
    h = torch.nn.functional.one_hot(cls, F_in).float() / 4.0 * node_mask
    xh = torch.cat([x, h], dim=2)
    g = model.gamma.gamma[t_int]
    a, s = torch.sqrt(torch.sigmoid(-g)), torch.sqrt(torch.sigmoid(g))
    return a * xh + s * processed_noise(gen, (B, N, 3 + F_in), node_mask)


def teacher_forced(dataset, nodesxsample, target_name, scale, steps, seed):
    args, model, pred, prop = build(dataset)
    F_in = 1 if dataset == "cata" else 12
    inner = model.module
    nm, em = ref_masks(args, nodesxsample)
    B, N, _ = nm.shape
    D = 3 + F_in
    tf = targets(pred, prop)[target_name]
    gen = torch.Generator().manual_seed(seed)
    out = {"nodesxsample": nodesxsample.numpy(), "node_mask": nm.numpy(), "edge_mask": em.numpy(),
           "steps_t": np.array(steps), "scale": np.float32(scale)}
    for t_int in steps:
        s_int = t_int - 1
        zt = data_like_z(gen, inner, nm, F_in, t_int)
        noise = processed_noise(gen, (B, N, D), nm)
        s_arr = torch.full((B, 1), s_int) / inner.T
        t_arr = torch.full((B, 1), s_int + 1) / inner.T
        cap = {}

        def tf_rec(_i, _n, _e, _t):
            cap["zs_pre"] = _i.detach().clone()
            return tf(_i, _n, _e, _t)

        inner.sample_combined_position_feature_noise = lambda n_s, n_n, node_mask, std=1.0: noise
        zs = inner.sample_p_zs_given_zt_guidance(s_arr, t_arr, zt, nm, em, tf_rec, scale)
        zs_un = inner.sample_p_zs_given_zt(s_arr, t_arr, zt, nm, em, None)
        with torch.no_grad():
            eps = inner.phi(zt, t_arr, nm, em, None)
            pr = pred(cap["zs_pre"], nm, em, t_arr)
        with torch.enable_grad():
            zz = cap["zs_pre"].clone().requires_grad_()
            energy = scale * tf(zz, nm, em, t_arr).sum()
            graw = torch.autograd.grad(energy, zz)[0]
        for k, v in dict(zt=zt, noise=noise, eps=eps, zs_pre=cap["zs_pre"], pred=pr, grad_raw=graw,
                         zs=zs, zs_unguided=zs_un).items():
            out[f"{k}_{t_int}"] = v.detach().numpy()
    # final decode from a data-like z_0
    z0 = data_like_z(gen, inner, nm, F_in, 0)
    noise = processed_noise(gen, (B, N, D), nm)
    inner.sample_combined_position_feature_noise = lambda n_s, n_n, node_mask, std=1.0: noise
    with torch.no_grad():
        x, h = inner.sample_p_xh_given_z0(z0, nm, em, None)
    out.update(dec_z0=z0.numpy(), dec_noise=noise.numpy(), dec_x=x.numpy(), dec_one_hot=h["categorical"].numpy())
    del inner.sample_combined_position_feature_noise
    return out, digest(model.state_dict()), digest(pred.state_dict())


def chain(dataset, nodesxsample, guided, target_name, scale, std, seed, record_every=100):
    args, model, pred, prop = build(dataset)
    F_in = 1 if dataset == "cata" else 12
    inner = model.module
    T = inner.T
    if guided:
        nm, em = ref_masks(args, nodesxsample)
    else:                                                     # sample_pos_edm pads to args.max_nodes
        got = {}

        class Fake:
            def sample(self, bs, n_nodes, node_mask, edge_mask, std=1.0):
                got["nm"], got["em"] = node_mask.clone(), edge_mask.clone()
                return torch.zeros(bs, n_nodes, 3), {"categorical": torch.zeros(bs, n_nodes, 1)}
        sampling_edm.sample_pos_edm(args, Fake(), nodesxsample)
        nm, em = got["nm"], got["em"]
    B, N, _ = nm.shape
    D = 3 + F_in
    gen = torch.Generator().manual_seed(seed)
    noise = torch.stack([processed_noise(gen, (B, N, D), nm, std if k == 0 else 1.0) for k in range(T + 2)])
    it = iter(noise)
    inner.sample_combined_position_feature_noise = lambda n_s, n_n, node_mask, std=1.0: next(it)
    rec = {}
    if guided:
        tf = targets(pred, prop)[target_name]
        orig = inner.sample_p_zs_given_zt_guidance

        def hook(s, t, zt, *a, **k):
            z = orig(s, t, zt, *a, **k)
            si = int(round(float(s[0, 0]) * T))
            if si % record_every == 0:
                rec[f"z_{si}"] = z.detach().numpy().copy()
            return z
        inner.sample_p_zs_given_zt_guidance = hook
        x, one_hot, nm2, em2 = sampling_edm.sample_guidance(args, model, tf, nodesxsample, scale=scale, std=std)
    else:
        orig = inner.sample_p_zs_given_zt

        def hook(s, t, zt, *a, **k):
            z = orig(s, t, zt, *a, **k)
            si = int(round(float(s[0, 0]) * T))
            if si % record_every == 0:
                rec[f"z_{si}"] = z.detach().numpy().copy()
            return z
        inner.sample_p_zs_given_zt = hook
        x, one_hot, nm2, em2 = sampling_edm.sample_pos_edm(args, model, nodesxsample, std=std)
    assert torch.equal(nm2, nm) and torch.equal(em2, em)
    rec.update(nodesxsample=nodesxsample.numpy(), node_mask=nm.numpy(), edge_mask=em.numpy(),
               noise=noise.numpy(), x=x.numpy(), one_hot=one_hot.numpy(), scale=np.float32(scale),
               std=np.float32(std))
    return rec


def chain_frames(dataset, nodesxsample, std, seed, keep_frames):
    """EnVariationalDiffusion.sample_chain (en_diffusion.py:1118-1174) with the SAME masks and injected noise as the unguided
    `chain` fixture (same seed -> same generator stream), so the frames fixture does not need to store the noise again."""
    args, model, pred, prop = build(dataset)
    F_in = 1 if dataset == "cata" else 12
    inner = model.module
    T = inner.T
    got = {}

    class Fake:
        def sample(self, bs, n_nodes, node_mask, edge_mask, std=1.0):
            got["nm"], got["em"] = node_mask.clone(), edge_mask.clone()
            return torch.zeros(bs, n_nodes, 3), {"categorical": torch.zeros(bs, n_nodes, 1)}
    sampling_edm.sample_pos_edm(args, Fake(), nodesxsample)
    nm, em = got["nm"], got["em"]
    B, N, _ = nm.shape
    D = 3 + F_in
    gen = torch.Generator().manual_seed(seed)
    noise = torch.stack([processed_noise(gen, (B, N, D), nm, std if k == 0 else 1.0) for k in range(T + 2)])
    it = iter(noise)
    inner.sample_combined_position_feature_noise = lambda n_s, n_n, node_mask, std=1.0: next(it)
    frames = inner.sample_chain(B, N, nm, em, None, keep_frames=keep_frames, std=std)
    return {"frames": frames.view(keep_frames, B, N, D).numpy(), "keep_frames": np.int64(keep_frames), "nodesxsample": nodesxsample.numpy(),
            "noise_digest": np.frombuffer(hashlib.sha256(noise.numpy().tobytes()).digest(), dtype=np.uint8)}


def train_step(dataset, nodesxsample, t_list, seed):
    """One training-loss evaluation + backward of the reference (train_edm.py:36-49, 71-75) with the two random draws of
    compute_loss (en_diffusion.py:657-659 t_int, :677-679 eps) pinned so that the oracle / product can be fed the same."""
    args, model, pred, prop = build(dataset)
    for n_, p_ in model.named_parameters():
        p_.requires_grad_(not n_.endswith("gamma.gamma"))          # the noise table is a frozen Parameter (en_diffusion.py:214-216)
    model.train()
    F_in = 1 if dataset == "cata" else 12
    gen = torch.Generator().manual_seed(seed)
    nm, em = ref_masks(args, nodesxsample)
    B, N = nm.shape[0], nm.shape[1]
    x = remove_mean_with_mask(torch.randn((B, N, 3), generator=gen) * 2.5 * nm, nm)
    cls = torch.randint(0, F_in, (B, N), generator=gen)
    h = torch.nn.functional.one_hot(cls, F_in).float() * nm
    t_int = torch.tensor(t_list, dtype=torch.int64).view(B, 1)
    eps = processed_noise(gen, (B, N, 3 + F_in), nm)
    real_randint = torch.randint
    torch.randint = lambda *a, **k: t_int.clone()
    inner = model.module if hasattr(model, "module") else model     # get_model may wrap the model (MyDataParallel)
    inner.sample_combined_position_feature_noise = lambda n_samples, n_nodes, node_mask, std=1.0: eps.clone()
    try:
        hd = {"categorical": h, "integer": torch.zeros(0)}          # train_edm.py:42
        loss_b = model(x, hd, nm, em.view(B, N * N))
        loss = loss_b.mean(0)                                       # train_edm.py:47
        loss.backward()
    finally:
        torch.randint = real_randint
    full = ["dynamics.egnn.embedding.weight", "dynamics.egnn.embedding.bias", "dynamics.egnn.embedding_out.weight",
            "dynamics.egnn.e_block_0.gcl_0.att_mlp.0.weight", "dynamics.egnn.e_block_4.gcl_0.edge_mlp.0.bias",
            "dynamics.egnn.e_block_8.gcl_equiv.coord_mlp.4.weight", "dynamics.egnn.e_block_0.gcl_equiv.coord_mlp.0.weight"]
    names, norms = [], []
    out = dict(nodesxsample=nodesxsample.numpy(), x=x.numpy(), h=h.numpy(), t_int=t_int.numpy(), eps=eps.numpy(),
               loss_b=loss_b.detach().numpy(), loss=loss.detach().numpy())
    for n_, p_ in model.named_parameters():
        if p_.grad is None:
            continue
        n_ = n_[len("module."):] if n_.startswith("module.") else n_
        names.append(n_)
        norms.append(float(p_.grad.double().norm()))
        if n_ in full:
            out["grad:" + n_] = p_.grad.numpy()
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms, dtype=np.float64)
    return out


def predictor_train_step(dataset, nodesxsample, t_list, seed):
    """One loss + backward of the reference's predictor training (cond_prediction/train_cond_predictor.py:65-81, 104-111)
    with its two random draws (t_int, eps of sample_edm_t) pinned."""
    from cond_prediction import train_cond_predictor as TCP
    args, model, pred, prop = build(dataset)
    for p_ in pred.parameters():
        p_.requires_grad_(True)
    pred.train()
    inner = model.module if hasattr(model, "module") else model
    F_in = 1 if dataset == "cata" else 12
    gen = torch.Generator().manual_seed(seed)
    nm, em = ref_masks(args, nodesxsample)
    B, N = nm.shape[0], nm.shape[1]
    x = remove_mean_with_mask(torch.randn((B, N, 3), generator=gen) * 2.5 * nm, nm)
    h = torch.nn.functional.one_hot(torch.randint(0, F_in, (B, N), generator=gen), F_in).float() * nm
    y = torch.randn((B, 5), generator=gen)
    t_int = torch.tensor(t_list, dtype=torch.int64).view(B, 1)
    eps = processed_noise(gen, (B, N, 3 + F_in), nm)
    real_randint = torch.randint
    torch.randint = lambda *a, **k: t_int.clone()
    inner.sample_combined_position_feature_noise = lambda n_samples, n_nodes, node_mask, std=1.0: eps.clone()
    try:
        loss, err = TCP.compute_loss(pred, x, h, nm, em, y, inner, Namespace(diffusion_steps=inner.T))
        loss.backward()
        with torch.no_grad():
            z_t = TCP.sample_edm_t(x, h, inner, t_int.float() / inner.T, nm)
    finally:
        torch.randint = real_randint
        del inner.sample_combined_position_feature_noise
    out = dict(nodesxsample=nodesxsample.numpy(), x=x.numpy(), h=h.numpy(), y=y.numpy(), t_int=t_int.numpy(), eps=eps.numpy(),
               loss=loss.detach().numpy(), abs_err=err.numpy(), z_t=z_t.numpy())
    full = ["egnn.embedding.weight", "egnn.embedding_out.weight", "egnn.embedding_out.bias", "egnn.gcl_0.att_mlp.0.weight",
            "egnn.gcl_5.coord_mlp.0.bias", "egnn.gcl_10.coord_mlp.2.weight", "egnn.gcl_3.edge_mlp.0.weight"]
    names, norms = [], []
    for n_, p_ in pred.named_parameters():
        n_ = n_[len("module."):] if n_.startswith("module.") else n_
        if p_.grad is None:                      # the last layer's coordinate branch never reaches the output
            continue
        names.append(n_)
        norms.append(float(p_.grad.double().norm()))
        if n_ in full:
            out["grad:" + n_] = p_.grad.numpy()
    out["grad_names"], out["grad_norms"] = np.array(names), np.array(norms, dtype=np.float64)
    return out


def eval_nll(dataset, nodesxsample, t_list, seed):
    """forward() of the reference in eval mode (the bound val_epoch reports, en_diffusion.py:777-804 with t0_always=True),
    draws pinned: t_int (randint(1, T+1)), then eps for z_t, then eps_0 for z_0."""
    args, model, pred, prop = build(dataset)
    model.eval()
    inner = model.module if hasattr(model, "module") else model
    F_in = 1 if dataset == "cata" else 12
    gen = torch.Generator().manual_seed(seed)
    nm, em = ref_masks(args, nodesxsample)
    B, N = nm.shape[0], nm.shape[1]
    x = remove_mean_with_mask(torch.randn((B, N, 3), generator=gen) * 2.5 * nm, nm)
    h = torch.nn.functional.one_hot(torch.randint(0, F_in, (B, N), generator=gen), F_in).float() * nm
    t_int = torch.tensor(t_list, dtype=torch.int64).view(B, 1)
    draws = [processed_noise(gen, (B, N, 3 + F_in), nm), processed_noise(gen, (B, N, 3 + F_in), nm)]
    it = iter(draws)
    real_randint = torch.randint
    torch.randint = lambda *a, **k: t_int.clone()
    inner.sample_combined_position_feature_noise = lambda n_samples, n_nodes, node_mask, std=1.0: next(it).clone()
    try:
        with torch.no_grad():
            nll = model(x, {"categorical": h, "integer": torch.zeros(0)}, nm, em.view(B, N * N))
    finally:
        torch.randint = real_randint
        del inner.sample_combined_position_feature_noise
    return dict(nodesxsample=nodesxsample.numpy(), x=x.numpy(), h=h.numpy(), t_int=t_int.numpy(), eps=draws[0].numpy(),
                eps0=draws[1].numpy(), nll=nll.numpy())


def main():
    meta = {"torch": torch.__version__, "seed_denoiser": SEED_DEN, "seed_predictor": SEED_PRED,
            "prop_mean": MEAN5.tolist(), "prop_std": STD5.tolist()}
    steps = [1000, 999, 750, 500, 250, 2, 1]

    if "--only-nll" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "nll_cata.npz"), **eval_nll("cata", torch.tensor([11, 10, 9, 11, 7, 2]), [1000, 613, 1, 2, 250, 77], seed=888))
        np.savez_compressed(os.path.join(HERE, "nll_hetro.npz"), **eval_nll("hetro", torch.tensor([10, 8, 3]), [777, 1, 12], seed=889))
        return
    if "--only-pred-train" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "pred_train_cata.npz"),
                            **predictor_train_step("cata", torch.tensor([11, 10, 9, 11, 7, 2]), [1000, 613, 0, 1, 250, 0], seed=666))
        np.savez_compressed(os.path.join(HERE, "pred_train_hetro.npz"),
                            **predictor_train_step("hetro", torch.tensor([10, 8, 3]), [777, 0, 12], seed=667))
        return
    if "--only-chain-frames" in sys.argv:   # sample_chain frames on the masks / noise of chain_cata_unguided.npz
        np.savez_compressed(os.path.join(HERE, "chain_frames_cata.npz"), **chain_frames("cata", torch.tensor([11, 10, 8]), 0.7, 78, 8))
        return
    if "--only-train" in sys.argv:          # add the training fixtures without regenerating the (slow) chains
        np.savez_compressed(os.path.join(HERE, "train_cata.npz"),
                            **train_step("cata", torch.tensor([11, 10, 9, 11, 7, 2]), [1000, 613, 0, 1, 250, 0], seed=555))
        np.savez_compressed(os.path.join(HERE, "train_hetro.npz"),
                            **train_step("hetro", torch.tensor([10, 8, 3]), [777, 0, 12], seed=556))
        return

    # masks, both datasets, ragged sizes incl. minimum
    for ds, nx in [("cata", [11, 10, 9, 11, 7, 2]), ("hetro", [10, 8, 10, 3, 1])]:
        args = Args_EDM().parse_args([])
        args.device, args.dataset = "cpu", ds
        nm, em = ref_masks(args, torch.tensor(nx))
        np.savez_compressed(os.path.join(HERE, f"masks_{ds}.npz"), nodesxsample=np.array(nx),
                            node_mask=nm.numpy(), edge_mask=em.numpy())

    out, dd, dp = teacher_forced("cata", torch.tensor([11, 10, 9, 11, 7]), "max_gap", 0.6, steps, seed=1234)
    np.savez_compressed(os.path.join(HERE, "step_cata.npz"), **out)
    meta["digest_denoiser_cata"], meta["digest_predictor_cata"] = dd, dp
    print("step_cata done", flush=True)

    out, dd, dp = teacher_forced("hetro", torch.tensor([10, 8, 3]), "opv", 0.6, steps, seed=4321)
    np.savez_compressed(os.path.join(HERE, "step_hetro.npz"), **out)
    meta["digest_denoiser_hetro"], meta["digest_predictor_hetro"] = dd, dp
    print("step_hetro done", flush=True)

    rec = chain("cata", torch.tensor([10, 9]), True, "max_gap", 0.6, 1.0, seed=77)
    np.savez_compressed(os.path.join(HERE, "chain_cata_guided.npz"), **rec)
    print("chain guided done", flush=True)

    rec = chain("cata", torch.tensor([11, 10, 8]), False, None, 0.0, 0.7, seed=78)
    np.savez_compressed(os.path.join(HERE, "chain_cata_unguided.npz"), **rec)
    print("chain unguided done", flush=True)
    np.savez_compressed(os.path.join(HERE, "chain_frames_cata.npz"), **chain_frames("cata", torch.tensor([11, 10, 8]), 0.7, 78, 8))

    np.savez_compressed(os.path.join(HERE, "train_cata.npz"),
                        **train_step("cata", torch.tensor([11, 10, 9, 11, 7, 2]), [1000, 613, 0, 1, 250, 0], seed=555))
    np.savez_compressed(os.path.join(HERE, "train_hetro.npz"),
                        **train_step("hetro", torch.tensor([10, 8, 3]), [777, 0, 12], seed=556))
    np.savez_compressed(os.path.join(HERE, "pred_train_cata.npz"),
                        **predictor_train_step("cata", torch.tensor([11, 10, 9, 11, 7, 2]), [1000, 613, 0, 1, 250, 0], seed=666))
    np.savez_compressed(os.path.join(HERE, "pred_train_hetro.npz"),
                        **predictor_train_step("hetro", torch.tensor([10, 8, 3]), [777, 0, 12], seed=667))
    np.savez_compressed(os.path.join(HERE, "nll_cata.npz"), **eval_nll("cata", torch.tensor([11, 10, 9, 11, 7, 2]), [1000, 613, 1, 2, 250, 77], seed=888))
    np.savez_compressed(os.path.join(HERE, "nll_hetro.npz"), **eval_nll("hetro", torch.tensor([10, 8, 3]), [777, 1, 12], seed=889))
    print("train done", flush=True)

    with open(os.path.join(HERE, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
