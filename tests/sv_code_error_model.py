"""CPU error model (test infrastructure, not collected by pytest): effect of storing the predictor's saved activations
SiLU'(pre1), pre2, SiLU'(pre3) at reduced precision on the raw guidance gradient, evaluated with the oracle's autograd on the
reference goldens (tests/golden/step_*.npz).  This is the experiment behind the 16-bit fixed-point / 24-bit codes of
gaudi_b200/csrc/tc_common.cuh:   python tests/sv_code_error_model.py
"""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gaudi_oracle as O
import torch.nn.functional as F
from helpers import build_models, cpu_weights, oracle_cfgs, oracle_target
torch.set_num_threads(8)
MODE = {"d": None, "p2": None}
class QSiLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind):
        s = torch.sigmoid(x)
        if kind == "d":   # save derivative quantised
            d = s * (1 + x * (1 - s))
            if MODE["d"] is not None: d = d.to(MODE["d"]).float()
            ctx.save_for_backward(d)
        else:             # save pre2 quantised; derivative from it
            xq = x if MODE["p2"] is None else x.to(MODE["p2"]).float()
            sq = torch.sigmoid(xq)
            ctx.save_for_backward(sq * (1 + xq * (1 - sq)))
        return x * s
    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return g * d, None
class Gate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e, eq, W, b):
        gate = torch.sigmoid(e @ W.t() + b)
        ctx.save_for_backward(eq, gate, W)
        return e * gate
    @staticmethod
    def backward(ctx, g):
        eq, gate, W = ctx.saved_tensors
        pdot = (g * eq).sum(1, keepdim=True)
        return g * gate + pdot * gate * (1 - gate) * W, None, None, None
def predictor_forward(w, cfg, z, node_mask, edge_mask, t):
    p = cfg.prefix + "egnn."
    B, N, _ = z.shape
    row, col = O.dense_edges(B, N)
    nm = node_mask.reshape(B * N, 1).to(z.dtype)
    em = edge_mask.reshape(B * N * N, 1).to(z.dtype)
    x = z[:, :, :3].reshape(B * N, -1).clone() * nm
    h = z[:, :, 3:].reshape(B * N, -1).clone() * nm
    h = O._append_time(h, t, B, N)
    a = torch.sum((x[row] - x[col]) ** 2, dim=1, keepdim=True)
    rng = float(cfg.coords_range) / cfg.n_layers
    h = O._lin(w, p + "embedding", h)
    for l in range(cfg.n_layers):
        g = f"{p}gcl_{l}."
        r, u = O._radial(x, row, col, 1.0)
        e = torch.cat([h[row], h[col], r, a], dim=1)
        e = QSiLU.apply(O._lin(w, g + "edge_mlp.0", e), "d")
        pre2 = O._lin(w, g + "edge_mlp.2", e)
        # pre2: attention uses q = silu(pre2q) in backward too; model: forward exact, backward from quantised pre2
        e = QSiLU.apply(pre2, "p2")
        pq = pre2 if MODE["p2"] is None else pre2.to(MODE["p2"]).float()
        ef = Gate.apply(e, F.silu(pq).detach(), w[g + "att_mlp.0.weight"], w[g + "att_mlp.0.bias"]) * em
        c = QSiLU.apply(O._lin(w, g + "coord_mlp.0", ef), "d")
        trans = u * torch.tanh(O._lin(w, g + "coord_mlp.2", c, bias=False)) * rng * em
        x_new = x + O._segment_sum(trans, row, B * N)
        agg = O._segment_sum(ef, row, B * N)
        upd = torch.cat([h, agg], dim=1)
        upd = O._lin(w, g + "node_mlp.2", F.silu(O._lin(w, g + "node_mlp.0", upd)))
        h = (h + upd) * nm
        x = x_new * nm
    h = O._lin(w, p + "embedding_out", h) * nm
    return h.view(B, N, -1).mean(1)


def run():
    for ds in ("cata", "hetro"):
        G = np.load(os.path.join(ROOT, "tests", "golden", f"step_{ds}.npz"))
        args, model, pred, prop = build_models(ds, "cpu")
        wd, wp = cpu_weights(model, pred)
        dcfg, pcfg = oracle_cfgs(ds)
        nm = torch.tensor(G["node_mask"]); em = torch.tensor(G["edge_mask"])
        tgt = oracle_target(ds); scale = float(G["scale"])
        for t in (1000, 500, 1):
            z = torch.tensor(G[f"zs_pre_{t}"]); gref = torch.tensor(G[f"grad_raw_{t}"])
            s = t - 1
            tv = O.time_value(s, dcfg.timesteps) if hasattr(dcfg, "timesteps") else None
            for name, md, mp in (("exact", None, None), ("d fp16", torch.float16, None), ("d bf16", torch.bfloat16, None),
                                 ("d fp16 + p2 fp16", torch.float16, torch.float16), ("d fp16 + p2 bf16", torch.float16, torch.bfloat16)):
                MODE["d"], MODE["p2"] = md, mp
                with torch.enable_grad():
                    zz = z.clone().requires_grad_()
                    B = z.shape[0]
                    tt = O.time_value(t, 1000)
                    p = predictor_forward(wp, pcfg, zz, nm, em, tt)
                    e = scale * tgt(p).sum()
                    g = torch.autograd.grad(e, zz)[0]
                print(ds, t, name, "max|g|=%.3e  err=%.3e" % (float(gref.abs().max()), float((g - gref).abs().max())))

if __name__ == "__main__":
    run()
