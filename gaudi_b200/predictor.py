"""Time-conditioned property predictor with GaUDI's signatures and ``state_dict`` layout.

Mirrors ``E_GCL`` edm/egnn_predictor/gcl.py:157-316, ``EGNN`` edm/egnn_predictor/models.py:492-560 and
``EGNN_predictor`` :390-489.  Parameters only; the forward and the input gradient run in the sm_100a
kernels (``gaudi_b200.runtime``).  ``EGNN_predictor.forward`` is differentiable w.r.t. ``xh`` through a
``torch.autograd.Function`` whose backward is the hand-written input-gradient kernel chain, so the
reference's ``autograd.grad(energy, zs)`` (en_diffusion.py:903) works unchanged for any cond_fn closure.
"""
from __future__ import annotations

import torch
from torch import nn

from . import runtime


class E_GCL(nn.Module):
    def __init__(self, input_nf, output_nf, hidden_nf, edges_in_d=0, nodes_att_dim=0, act_fn=nn.SiLU(),
                 recurrent=True, attention=False, clamp=False, norm_diff=True, tanh=False, coords_range=1,
                 agg="sum"):
        super().__init__()
        if agg != "sum":
            raise NotImplementedError("gaudi_b200 implements agg='sum' (the prediction_args default)")
        if nodes_att_dim != 0 or input_nf != hidden_nf or output_nf != hidden_nf:
            raise NotImplementedError("E_GCL: only the EGNN_predictor configuration is supported")
        self.recurrent = recurrent
        self.attention = attention
        self.norm_diff = norm_diff
        self.agg_type = agg
        self.tanh = tanh
        self.edge_mlp = nn.Sequential(nn.Linear(2 * input_nf + 1 + edges_in_d, hidden_nf), act_fn,
                                      nn.Linear(hidden_nf, hidden_nf), act_fn)
        self.node_mlp = nn.Sequential(nn.Linear(hidden_nf + input_nf + nodes_att_dim, hidden_nf), act_fn,
                                      nn.Linear(hidden_nf, output_nf))
        last = nn.Linear(hidden_nf, 1, bias=False)            # RNG order of gcl.py:205-209
        nn.init.xavier_uniform_(last.weight, gain=0.001)
        parts = [nn.Linear(hidden_nf, hidden_nf), act_fn, last]
        if tanh:
            parts.append(nn.Tanh())
            self.coords_range = coords_range
        self.coord_mlp = nn.Sequential(*parts)
        self.clamp = clamp
        if attention:
            self.att_mlp = nn.Sequential(nn.Linear(hidden_nf, 1), nn.Sigmoid())

    def forward(self, h, edge_index, coord, edge_attr=None, node_attr=None, node_mask=None, edge_mask=None):
        h, coord = runtime.e_gcl_forward(self, h, edge_index, coord, edge_attr, node_mask, edge_mask)
        return h, coord, edge_attr


class EGNN(nn.Module):
    """12 x E_GCL stack of the predictor (edm/egnn_predictor/models.py:492-560)."""

    def __init__(self, in_node_nf, in_edge_nf, hidden_nf, device="cpu", act_fn=nn.SiLU(), n_layers=4,
                 recurrent=True, attention=False, norm_diff=True, out_node_nf=None, tanh=False,
                 coords_range=15, agg="sum"):
        super().__init__()
        if out_node_nf is None:
            out_node_nf = in_node_nf
        self.hidden_nf = hidden_nf
        self.device = device
        self.n_layers = n_layers
        self.coords_range_layer = float(coords_range) / n_layers
        if agg == "mean":
            self.coords_range_layer *= 19
        self.embedding = nn.Linear(in_node_nf, hidden_nf)
        self.embedding_out = nn.Linear(hidden_nf, out_node_nf)
        for i in range(n_layers):
            self.add_module(f"gcl_{i}", E_GCL(hidden_nf, hidden_nf, hidden_nf, edges_in_d=in_edge_nf,
                                              act_fn=act_fn, recurrent=recurrent, attention=attention,
                                              norm_diff=norm_diff, tanh=tanh,
                                              coords_range=self.coords_range_layer, agg=agg))
        self.to(device)

    def forward(self, h, x, edges, edge_attr=None, node_mask=None, edge_mask=None):
        return runtime.pred_egnn_forward(self, h, x, edges, edge_attr, node_mask, edge_mask)


class EGNN_predictor(nn.Module):
    def __init__(self, in_nf=1, out_nf=1, hidden_nf=64, device="cpu", act_fn=torch.nn.SiLU(), n_layers=4,
                 recurrent=True, attention=False, tanh=False, agg="sum", mean=None, std=None,
                 condition_time=False, d=3, coords_range=15):
        super().__init__()
        if d != 3:
            raise NotImplementedError("d must be 3")
        if not recurrent:
            raise NotImplementedError("recurrent=False is never used by GaUDI")
        self.d = d
        in_node_nf = in_nf + 1 if condition_time else in_nf
        self.egnn = EGNN(in_node_nf=in_node_nf, in_edge_nf=1, hidden_nf=hidden_nf, out_node_nf=out_nf,
                         device=device, act_fn=act_fn, n_layers=n_layers, recurrent=recurrent,
                         attention=attention, tanh=tanh, agg=agg, coords_range=coords_range)
        self.mean = mean.to(device) if mean is not None else None
        self.std = std.to(device) if std is not None else None
        self.device = device
        self._edges_dict = {}
        self.condition_time = condition_time
        self.hyper = dict(hidden_nf=hidden_nf, n_layers=n_layers, attention=bool(attention), tanh=bool(tanh),
                          out_nf=out_nf, in_node_nf=in_node_nf, coords_range=float(coords_range))

    def forward(self, xh, node_mask, edge_mask, t=torch.zeros(1)):
        """pred [B, out_nf] (edm/egnn_predictor/models.py:433-457).  Differentiable w.r.t. ``xh`` (the guidance gradient,
        hand-written input-gradient kernels) or -- when ``xh`` needs no gradient and the parameters do, i.e. while training
        the predictor (cond_prediction/train_cond_predictor.py:65-81) -- w.r.t. the parameters."""
        if torch.is_grad_enabled() and not xh.requires_grad and any(p.requires_grad for p in self.parameters()):
            from . import training
            return training.predictor_forward_train(self, xh, node_mask, edge_mask, t)
        return runtime.predictor_forward(self, xh, node_mask, edge_mask, t)

    def unnormalize(self, pred):
        if self.mean is not None:
            pred = pred * self.std + self.mean
        return pred

    def get_adj_matrix(self, n_nodes, batch_size, device):
        key = (n_nodes, batch_size)
        if key not in self._edges_dict:
            base = torch.arange(batch_size, device=device).repeat_interleave(n_nodes * n_nodes) * n_nodes
            i = torch.arange(n_nodes, device=device).repeat_interleave(n_nodes).repeat(batch_size)
            j = torch.arange(n_nodes, device=device).repeat(n_nodes * batch_size)
            self._edges_dict[key] = [base + i, base + j]
        return self._edges_dict[key]
