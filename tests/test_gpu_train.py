"""GPU parity of the training step: loss and parameter gradients of the CUDA path (C ABI ops chained
by gaudi_b200/training.py) against the reference goldens (tests/golden/train_*.npz) and the CPU oracle's autograd.

Tolerances: loss max-abs <= 1e-4; every parameter gradient within 1e-4 * max(1, max|ref|) absolute AND 2e-4 relative
in norm (fp32, different summation order: the reductions over ~60 edges/node and ~10^3 rows are re-associated).
"""
import numpy as np
import pytest
import torch

import gaudi_b200 as gb
import gaudi_oracle as O
from gaudi_b200 import _lib, training
from helpers import build_models, golden, maxabs, oracle_cfgs

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch.device("cuda:0")


def _train_model(ds, dev, **kw):
    args, model, pred, prop = build_models(ds, dev, **kw)
    for n_, p_ in model.named_parameters():
        p_.requires_grad_(not n_.endswith("gamma.gamma"))
    model.train()
    return args, model


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gemm_modes_match_fp64(mode):
    dev = _dev()
    torch.manual_seed(mode)
    for (M, N, K) in [(70, 65, 33), (192, 386, 5000), (1, 7, 130), (300, 192, 192)]:
        a = torch.randn((K, M) if mode == 2 else (M, K), device=dev)
        b = torch.randn((N, K) if mode == 0 else (K, N), device=dev)
        bias = torch.randn(N, device=dev) if mode == 0 else None
        c = torch.empty(M, N, device=dev)
        training._gemm(mode, M, N, K, a, a.shape[1], b, b.shape[1], c, N, bias)
        ad, bd = a.double(), b.double()
        ref = (ad @ bd.T + bias.double()) if mode == 0 else (ad @ bd if mode == 1 else ad.T @ bd)
        assert maxabs(c, ref) <= 2e-5 * max(1.0, float(ref.abs().max()))
        c2 = c.clone()
        training._gemm(mode, M, N, K, a, a.shape[1], b, b.shape[1], c2, N, None, acc=True)
        ref2 = c.double() + (ref - (bias.double() if mode == 0 else 0))
        assert maxabs(c2, ref2) <= 4e-5 * max(1.0, float(ref2.abs().max()))


@pytest.mark.parametrize("epi", [0, 1, 2, 3, 4])
def test_tensor_core_linear_matches_fp64(epi):
    """gb_linear (tcgen05, 3xTF32) with every epilogue the training step uses, forward and dgrad orientation."""
    dev = _dev()
    torch.manual_seed(epi)
    for (M, N, K1, K2, tr) in [(1000, 192, 192, 0, False), (777, 192, 192, 192, False), (130, 64, 64, 0, True), (5000, 192, 192, 0, True)]:
        a1, a2 = torch.randn(M, K1, device=dev), (torch.randn(M, K2, device=dev) if K2 else None)
        Wfull = torch.randn(N + 5, K1 + K2 + 3, device=dev) if not tr else torch.randn(K1 + 2, N + 7, device=dev)
        W = Wfull[:N, :K1 + K2] if not tr else Wfull[:K1, 3:3 + N]        # strided views, like W1[:, H:2H]
        bias = torch.randn(N, device=dev)
        aux, mask = torch.randn(M, N, device=dev), (torch.rand(M, device=dev) > 0.3).float()
        out2 = torch.empty(M, N, device=dev) if epi == 1 else None
        y = training._linear(a1, W, K1, N, bias, A2=a2, K2=K2, transpose=tr, epi=epi, out2=out2,
                             aux=aux if epi in (2, 3, 4) else None, mask=mask if epi == 2 else None)
        A = torch.cat([a1, a2], 1).double() if K2 else a1.double()
        pre = A @ (W.double() if tr else W.double().T) + bias.double()
        if epi == 1:
            ref = torch.nn.functional.silu(pre)
            assert maxabs(out2, pre) <= 2e-5 * float(pre.abs().max())
        elif epi == 2:
            ref = (aux.double() + pre) * mask.double()[:, None]
        elif epi == 3:
            sg = torch.sigmoid(aux.double())
            ref = pre * (sg * (1 + aux.double() * (1 - sg)))
        elif epi == 4:
            ref = pre + aux.double()
        else:
            ref = pre
        assert maxabs(y, ref) <= 2e-5 * max(1.0, float(ref.abs().max())), (epi, M, N, K1, K2, tr)


@pytest.mark.parametrize("gemm", ["tc", "fp32"])
@pytest.mark.parametrize("ds", ["cata", "hetro"])
def test_training_loss_and_gradients_match_reference_golden(ds, gemm, monkeypatch):
    dev = _dev()
    monkeypatch.setattr(training, "_TC", gemm == "tc")
    g = golden(f"train_{ds}.npz")
    args, model = _train_model(ds, dev)
    nm, em = gb.build_masks(torch.from_numpy(g["nodesxsample"]), 11 if ds == "cata" else 10, ds == "hetro", device=dev)
    x, h = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["h"]).to(dev)
    before = _lib.lib().gb_launch_count(0)
    loss_b = model(x, {"categorical": h, "integer": torch.zeros(0, device=dev)}, nm, em,
                   t_int=torch.from_numpy(g["t_int"]).to(dev), eps=torch.from_numpy(g["eps"]).to(dev))
    loss = loss_b.mean(0)                                    # train_edm.py:47
    loss.backward()
    assert _lib.lib().gb_launch_count(0) - before > 500      # the CUDA ops ran (no silent fallback)
    assert maxabs(loss_b, g["loss_b"]) <= 1e-4
    assert abs(float(loss) - float(g["loss"])) <= 1e-4
    params = dict(model.named_parameters())
    worst_rel = 0.0
    for name, norm in zip(g["grad_names"], g["grad_norms"]):
        gn = float(params[str(name)].grad.double().norm())
        rel = abs(gn - norm) / max(norm, 1e-6)
        worst_rel = max(worst_rel, rel)
        assert rel <= 2e-4, f"{name}: |grad| {gn:.6e} vs {norm:.6e}"       # measured 4e-5 (tcgen05 3xTF32) / 2e-6 (fp32 GEMM)
    worst_abs = 0.0
    for k in g.files:
        if k.startswith("grad:"):
            ref = torch.from_numpy(g[k])
            e = maxabs(params[k[5:]].grad, ref) / max(1.0, float(ref.abs().max()))
            worst_abs = max(worst_abs, e)
            assert e <= 1e-4, f"{k}: {e:.3e}"
    print(f"[train parity {ds} {gemm}] loss max-abs {maxabs(loss_b, g['loss_b']):.2e}, worst grad-norm rel {worst_rel:.2e}, "
          f"worst full-tensor abs {worst_abs:.2e}")


def test_all_parameter_gradients_match_oracle_autograd():
    """Every gradient tensor (not only norms) against the CPU oracle's autograd on a ragged seeded batch."""
    dev = _dev()
    ds = "cata"
    args, model = _train_model(ds, dev)
    dcfg, _ = oracle_cfgs(ds)
    torch.manual_seed(11)
    nx = torch.tensor([11, 3, 8, 10, 2, 11, 5, 9])
    nm_c, em_c = O.build_masks(nx, 11, False)
    B, N = nm_c.shape[:2]
    x = O.remove_mean_with_mask(torch.randn(B, N, 3) * 2.0 * nm_c, nm_c)
    h = torch.ones(B, N, 1) * nm_c
    t_int = torch.tensor([0, 1000, 1, 500, 37, 0, 999, 250]).view(B, 1)
    eps = O.draw_noise(B, N, 4, nm_c)
    w = {k: v.detach().cpu().clone().requires_grad_(not k.endswith("gamma.gamma")) for k, v in model.state_dict().items()}
    lo, _ = O.training_loss(w, dcfg, O.gamma_table(dcfg), x, h, nm_c, em_c, t_int.float(), eps)
    lo.mean(0).backward()
    nm, em = nm_c.to(dev), em_c.to(dev)
    lp = model(x.to(dev), {"categorical": h.to(dev), "integer": torch.zeros(0, device=dev)}, nm, em, t_int=t_int.to(dev),
               eps=eps.to(dev))
    lp.mean(0).backward()
    assert maxabs(lp, lo) <= 1e-4
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        ref = w[name].grad
        assert p.grad is not None, name
        assert maxabs(p.grad, ref) <= 1e-4 * max(1.0, float(ref.abs().max())), name
        rn = float(ref.double().norm())
        assert abs(float(p.grad.double().norm()) - rn) <= 2e-4 * rn + 2e-9, name          # + fp32 noise floor of ~1e-6-sized gradients


def test_optimizer_steps_reduce_the_loss():
    """train_epoch's inner loop (train_edm.py:71-82) on one fixed batch with pinned draws: AdamW steps drive the loss down."""
    dev = _dev()
    args, model = _train_model("cata", dev, hidden=(64, 64), layers=(3, 3))
    torch.manual_seed(5)
    nx = torch.tensor([11, 9, 10, 7] * 4)
    nm, em = gb.build_masks(nx, 11, False, device=dev)
    B, N = nm.shape[:2]
    x = torch.randn(B, N, 3, device=dev) * nm
    x = x - x.sum(1, keepdim=True) / nm.sum(1, keepdim=True) * nm
    h = torch.ones(B, N, 1, device=dev) * nm
    t_int = torch.randint(1, 1001, (B, 1), device=dev)
    eps = model.sample_combined_position_feature_noise(B, N, nm)
    hd = {"categorical": h, "integer": torch.zeros(0, device=dev)}
    l32 = model(x, hd, nm, em, t_int=t_int, eps=eps).detach()
    l64 = model(x.double(), {"categorical": h.double(), "integer": torch.zeros(0, device=dev)}, nm, em, t_int=t_int, eps=eps.double()).detach()
    assert torch.equal(l32, l64)                              # non-fp32 inputs are converted once, result unchanged
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-3, amsgrad=True, weight_decay=1e-12)
    losses = []
    for _ in range(30):
        opt.zero_grad()
        loss = model(x, {"categorical": h, "integer": torch.zeros(0, device=dev)}, nm, em, t_int=t_int, eps=eps).mean(0)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert np.isfinite(losses).all() and losses[-1] < 0.7 * losses[0], losses


@pytest.mark.parametrize("gemm", ["tc", "fp32"])
@pytest.mark.parametrize("ds", ["cata", "hetro"])
def test_predictor_training_step_matches_reference_golden(ds, gemm, monkeypatch):
    """8f rank 2: sample_edm_t, predictor forward in training mode, l1 loss, parameter gradients (train_cond_predictor.py:47-81)."""
    dev = _dev()
    monkeypatch.setattr(training, "_TC", gemm == "tc")
    g = golden(f"pred_train_{ds}.npz")
    args, model, pred, prop = build_models(ds, dev)
    for p_ in pred.parameters():
        p_.requires_grad_(True)
    pred.train()
    nm, em = gb.build_masks(torch.from_numpy(g["nodesxsample"]), 11 if ds == "cata" else 10, ds == "hetro", device=dev)
    x, h, y = (torch.from_numpy(g[k]).to(dev) for k in ("x", "h", "y"))
    t = torch.from_numpy(g["t_int"]).to(dev).float() / model.T
    zt = training.sample_edm_t(x, h, model, t, nm, eps=torch.from_numpy(g["eps"]).to(dev))
    assert maxabs(zt, g["z_t"]) <= 1e-6
    p = pred(zt, nm, em, t)
    loss = torch.nn.functional.l1_loss(p, y)                 # train_cond_predictor.py:80
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-4
    assert maxabs((p.detach() - y).abs(), g["abs_err"]) <= 1e-4
    params = dict(pred.named_parameters())
    worst = 0.0
    for name, norm in zip(g["grad_names"], g["grad_norms"]):
        gn = float(params[str(name)].grad.double().norm())
        worst = max(worst, abs(gn - norm) / max(norm, 1e-6))
        assert abs(gn - norm) <= 2e-4 * norm + 2e-9, f"{name}: {gn:.6e} vs {norm:.6e}"          # measured 9e-5 (tc) / 4e-6 (fp32)
    for k in g.files:
        if k.startswith("grad:"):
            ref = torch.from_numpy(g[k])
            assert maxabs(params[k[5:]].grad, ref) <= 1e-4 * max(1.0, float(ref.abs().max())), k
    # the guidance path (input gradient) is untouched by parameters that require grad
    z2 = zt.clone().requires_grad_()
    (gz,) = torch.autograd.grad(pred(z2, nm, em, t)[:, 1].sum(), z2)
    assert gz.shape == zt.shape and torch.isfinite(gz).all()
    print(f"[predictor train parity {ds} {gemm}] loss diff {abs(float(loss.detach()) - float(g['loss'])):.2e}, worst grad-norm rel {worst:.2e}")


@pytest.mark.parametrize("ds", ["cata", "hetro"])
def test_eval_mode_nll_matches_reference_golden(ds):
    """forward() in eval mode: variational bound with two fused denoiser passes (en_diffusion.py:644-804, t0_always=True).
    Tolerance: 1e-4 relative (the t-term multiplies the squared error by T/2 (SNR(gamma_s-gamma_t)-1), up to ~1e4)."""
    dev = _dev()
    g = golden(f"nll_{ds}.npz")
    args, model, pred, prop = build_models(ds, dev)
    model.eval()
    nm, em = gb.build_masks(torch.from_numpy(g["nodesxsample"]), 11 if ds == "cata" else 10, ds == "hetro", device=dev)
    x, h = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["h"]).to(dev)
    nll = model(x, {"categorical": h, "integer": torch.zeros(0, device=dev)}, nm, em, t_int=torch.from_numpy(g["t_int"]).to(dev),
                eps=torch.from_numpy(g["eps"]).to(dev), eps0=torch.from_numpy(g["eps0"]).to(dev))
    ref = torch.from_numpy(g["nll"])
    rel = ((nll.cpu() - ref).abs() / ref.abs().clamp(min=1.0)).max()
    print(f"[nll parity {ds}] worst relative error {float(rel):.2e}")
    assert float(rel) <= 1e-4
    assert torch.isfinite(model(x, {"categorical": h, "integer": torch.zeros(0, device=dev)}, nm, em)).all()   # own draws


def test_sample_chain_frames():
    """sample_chain (en_diffusion.py:1118-1174): frame k holds the un-normalised z of the last step written to it,
    frame 0 the final (x, h); same injected noise as sample() gives the same molecules."""
    dev = _dev()
    args, model, pred, prop = build_models("cata", dev, timesteps=20)
    T, keep = model.T, 5
    nm, em = gb.build_masks(torch.tensor([11, 9, 4]), 11, False, device=dev)
    B, N = nm.shape[:2]
    torch.manual_seed(3)
    noise = torch.stack([model.sample_combined_position_feature_noise(B, N, nm) for _ in range(T + 2)])
    chain = model.sample_chain(B, N, nm, em, None, keep_frames=keep, noise=noise).view(keep, B, N, 4)
    x, h = model.sample(B, N, nm, em, noise=noise)
    assert maxabs(chain[0, :, :, :3], x) <= 1e-6 and torch.equal(chain[0, :, :, 3:], h["categorical"])
    z = noise[0].clone()
    frames = {}
    for s in reversed(range(T)):
        s_arr = torch.full((B, 1), s, device=dev) / T
        z = model.sample_p_zs_given_zt(s_arr, s_arr + 1.0 / T, z, nm, em, None, noise=noise[T - s])
        frames[(s * keep) // T] = torch.cat([z[:, :, :3] * model.norm_values[0], z[:, :, 3:] * model.norm_values[1] * nm], 2)
    for k in range(1, keep):
        assert maxabs(chain[k], frames[k]) <= 1e-6


def test_sample_chain_matches_reference_golden():
    """sample_chain against the frames the UNMODIFIED reference produced (tests/golden/make_golden.py --only-chain-frames) on the
    masks and injected noise of the unguided chain fixture: 8 frames of a 1000-step run, frame 0 = the final (x, h)."""
    dev = _dev()
    f, g = golden("chain_frames_cata.npz"), golden("chain_cata_unguided.npz")
    args, model, pred, prop = build_models("cata", dev)
    nm, em = torch.from_numpy(g["node_mask"]).to(dev), torch.from_numpy(g["edge_mask"]).to(dev)
    noise = torch.from_numpy(g["noise"]).to(dev)
    B, N = nm.shape[:2]
    keep = int(f["keep_frames"])
    chain = model.sample_chain(B, N, nm, em, None, keep_frames=keep, std=float(g["std"]), noise=noise).view(keep, B, N, -1)
    ref = torch.from_numpy(f["frames"])
    worst = 0.0
    for k in range(keep):
        rel = maxabs(chain[k], ref[k]) / max(1.0, float(ref[k].abs().max()))
        worst = max(worst, rel)
        assert rel <= 1e-4, (k, rel)
    print(f"[sample_chain vs reference] worst relative frame error {worst:.2e}")
    assert torch.equal(chain[0, :, :, 3:].cpu(), ref[0, :, :, 3:])                  # one-hot ring types of the final molecules


def test_fused_adamw_amsgrad_clip_matches_torch():
    """gb_adamw_amsgrad_clip vs the reference's combination: edm/utils.gradient_clipping (Queue seeded with 3000, numpy mean / std,
    torch.nn.utils.clip_grad_norm_) followed by torch.optim.AdamW(amsgrad=True, weight_decay=1e-12) (train_edm.py:22-24, 71-82),
    over steps whose gradient norms first fill the window and then get clipped."""
    dev = _dev()
    from gaudi_b200 import train_utils as TU
    torch.manual_seed(0)
    shapes = [(192, 386), (192,), (1, 192), (3, 5, 7)]
    ref_p = [torch.nn.Parameter(torch.randn(*s_, device=dev) * 0.1) for s_ in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref_opt = torch.optim.AdamW(ref_p, lr=1e-3, weight_decay=1e-12, amsgrad=True)
    queue = TU.Queue(max_len=5)
    queue.add(3000)
    holder = torch.nn.ParameterList(ref_p)
    ours = TU.FusedAdamWClip(our_p, lr=1e-3, weight_decay=1e-12, clip=True, window=5, first_norm=3000.0)
    gen = torch.Generator(device=dev).manual_seed(1)
    scales = [1.0, 1.2, 0.8, 1.1, 0.9, 1.0, 40.0, 1.0, 0.05, 25.0, 1.0, 1.0]      # steps 6 and 9 exceed 1.5 mean + 2 std -> clipped
    for step, sc in enumerate(scales):
        grads = [torch.randn(p.shape, device=dev, generator=gen) * sc for p in ref_p]
        ref_opt.zero_grad(); ours.zero_grad()
        for p, q, g_ in zip(ref_p, our_p, grads):
            p.grad = g_.clone(); q.grad.copy_(g_)
        max_norm = 1.5 * queue.mean() + 2 * queue.std()
        ref_norm = float(TU.gradient_clipping(holder, queue))
        ref_opt.step()
        ours.step()
        assert abs(float(ours.last_grad_norm) - ref_norm) <= 1e-5 * ref_norm, step
        assert abs(float(ours.last_max_norm) - max_norm) <= 1e-6 * max_norm, step
        for p, q in zip(ref_p, our_p):
            assert maxabs(q, p) <= 2e-6 * max(1.0, float(p.abs().max())), (step, float(maxabs(q, p)))
    assert sorted(float(v) for v in ours.state[8:8 + 5].cpu()) == pytest.approx(sorted(queue.items), rel=1e-5)
    # plain AdamW-amsgrad (clip off) over a few more steps
    ref2 = [torch.nn.Parameter(torch.randn(64, 64, device=dev))]
    our2 = [torch.nn.Parameter(ref2[0].detach().clone())]
    o_ref, o_ours = torch.optim.AdamW(ref2, lr=3e-3, weight_decay=1e-2, amsgrad=True), TU.FusedAdamWClip(our2, lr=3e-3, weight_decay=1e-2, clip=False)
    for _ in range(5):
        g_ = torch.randn(64, 64, device=dev, generator=gen)
        ref2[0].grad = g_.clone(); our2[0].grad.copy_(g_)
        o_ref.step(); o_ours.step()
    assert maxabs(our2[0], ref2[0]) <= 2e-6


def test_tensor_core_wgrad_matches_fp64():
    """gb_wgrad (tcgen05, MN-major operands, 3xTF32, split-K): C = G^T X on strided views, with accumulation, through the
    deterministic two-phase reduction (scratch) and through the atomic path (no scratch)."""
    dev = _dev()
    torch.manual_seed(0)
    L = _lib.lib()
    for (K, M, N) in [(16, 192, 192), (1000, 192, 192), (52820, 192, 192), (777, 64, 64), (5000, 196, 196), (300, 256, 128), (9, 4, 8)]:
        Gf, Xf = torch.randn(K, M + 8, device=dev), torch.randn(K, N + 4, device=dev)
        G, X = Gf[:, 4:4 + M], Xf[:, :N]
        ref = G.double().T @ X.double()
        tol = 3e-6 * max(1.0, float(ref.abs().max())) * max(1.0, (K / 1000) ** 0.5)
        nbytes = L.gb_wgrad_scratch_bytes(M, N)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        results = []
        for sc, sb in ((scratch, nbytes), (None, 0)):
            Cf = torch.full((M, N + 6), 7.0, device=dev)
            C = Cf[:, 2:2 + N]
            args = (K, M, N, training._ptr(G), Gf.stride(0), training._ptr(X), Xf.stride(0), training._ptr(C), Cf.stride(0))
            cs = torch.full((M,), 3.0, device=dev) if (sc is not None and N < 256) else None
            _lib.check(L.gb_wgrad(*args, 0, training._ptr(cs), training._ptr(sc), sb, training._stream()))
            if cs is not None:                      # fused bias gradient: column sums of G through a column of ones
                assert maxabs(cs, G.double().sum(0)) <= 3e-6 * max(1.0, float(G.double().sum(0).abs().max())) * max(1.0, (K / 1000) ** 0.5)
            assert maxabs(C, ref) <= tol, (K, M, N, sc is None, maxabs(C, ref))
            assert float((Cf[:, :2] - 7).abs().max()) == 0 and float((Cf[:, 2 + N:] - 7).abs().max()) == 0     # nothing outside the view
            results.append(C.clone())
            _lib.check(L.gb_wgrad(*args, 1, None, training._ptr(sc), sb, training._stream()))
            assert maxabs(C, 2 * ref) <= 2 * tol
        Cf = torch.empty(M, N, device=dev)
        _lib.check(L.gb_wgrad(K, M, N, training._ptr(G), Gf.stride(0), training._ptr(X), Xf.stride(0), training._ptr(Cf), N, 0,
                              None, training._ptr(scratch), nbytes, training._stream()))
        assert torch.equal(Cf, results[0])                 # the two-phase path is bit-reproducible


def test_fit_loop_checkpoints_the_best_validation_epoch(tmp_path):
    """train_edm.main's loop (train_epoch / val_epoch / best-val checkpoint) on a tiny synthetic loader."""
    from gaudi_b200 import train_utils as TU
    dev = _dev()
    args, model = _train_model("cata", dev, hidden=(64, 64), layers=(2, 2))
    torch.manual_seed(9)

    def make_batch(sizes):
        nm, em = gb.build_masks(torch.tensor(sizes), 11, False)
        B, N = nm.shape[:2]
        x = torch.randn(B, N, 3) * 2.0 * nm
        return x, nm.squeeze(2), em.view(B, N, N), torch.ones(B, N, 1) * nm, torch.zeros(B)
    train = [make_batch([11, 10, 9, 8]), make_batch([11, 11, 7, 5])]
    val = [make_batch([10, 9, 11, 6])]
    torch.manual_seed(10)
    out = TU.fit(model, train, val, dev, num_epochs=3, lr=1e-3, exp_dir=str(tmp_path))
    assert len(out["history"]) == 3 and all(np.isfinite(h["train_loss"]) and np.isfinite(h["val_loss"]) for h in out["history"])
    assert (tmp_path / "model.pt").exists()
    best = min(out["history"], key=lambda h: h["val_loss"])
    assert out["best_epoch"] == best["epoch"] and abs(out["best_val_loss"] - best["val_loss"]) < 1e-9
    sd = torch.load(str(tmp_path / "model.pt"))
    assert all(torch.equal(v.cpu(), model.state_dict()[k].cpu()) for k, v in sd.items())      # best weights restored
