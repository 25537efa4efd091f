"""Denoiser network modules with GaUDI's constructor signatures and ``state_dict`` layout.

Mirrors (reference file:line): ``GCL`` edm/egnn/egnn_new.py:6-89, ``EquivariantUpdate`` :92-155,
``EquivariantBlock`` :158-235, ``EGNN`` :238-321, ``EGNN_dynamics`` edm/egnn/models.py:8-152.

The modules only *own parameters* (created in the reference's order so that a seeded default
initialisation yields bit-identical weights, and checkpoints load unchanged).  All arithmetic runs in the
hand-written sm_100a kernels behind the C-ABI (``gaudi_b200.runtime``): there is no PyTorch/CPU fallback,
calling a forward on a non-CUDA tensor raises.
"""
from __future__ import annotations

import torch
from torch import nn

from . import runtime


def _mlp(*layers: nn.Module) -> nn.Sequential:
    return nn.Sequential(*layers)


class GCL(nn.Module):
    """Invariant message-passing sub-layer (edge MLP, attention gate, node MLP)."""

    def __init__(self, input_nf, output_nf, hidden_nf, normalization_factor, aggregation_method,
                 edges_in_d=0, nodes_att_dim=0, act_fn=nn.SiLU(), attention=False):
        super().__init__()
        if aggregation_method != "sum":
            raise NotImplementedError("gaudi_b200 implements aggregation_method='sum' (the args_edm default)")
        if nodes_att_dim != 0 or input_nf != hidden_nf or output_nf != hidden_nf:
            raise NotImplementedError("GCL: only the EquivariantBlock configuration is supported")
        self.normalization_factor = normalization_factor
        self.aggregation_method = aggregation_method
        self.attention = attention
        self.edge_mlp = _mlp(nn.Linear(2 * input_nf + edges_in_d, hidden_nf), act_fn,
                             nn.Linear(hidden_nf, hidden_nf), act_fn)
        self.node_mlp = _mlp(nn.Linear(hidden_nf + input_nf + nodes_att_dim, hidden_nf), act_fn,
                             nn.Linear(hidden_nf, output_nf))
        if attention:
            self.att_mlp = _mlp(nn.Linear(hidden_nf, 1), nn.Sigmoid())

    def forward(self, h, edge_index, edge_attr=None, node_attr=None, node_mask=None, edge_mask=None):
        """(h_out, None): same contract as egnn_new.py:75-89 except that the per-edge ``mij`` tensor the
        reference also returns (and every caller discards) is never materialised."""
        return runtime.gcl_forward(self, h, edge_index, edge_attr, node_mask, edge_mask), None


class EquivariantUpdate(nn.Module):
    """Coordinate update x_i += sum_j (x_i-x_j)/(|.|+c) * tanh(phi(h_i,h_j,e_ij)) * range."""

    def __init__(self, hidden_nf, normalization_factor, aggregation_method, edges_in_d=1,
                 act_fn=nn.SiLU(), tanh=False, coords_range=10.0):
        super().__init__()
        if aggregation_method != "sum":
            raise NotImplementedError("gaudi_b200 implements aggregation_method='sum' (the args_edm default)")
        self.tanh = tanh
        self.coords_range = coords_range
        last = nn.Linear(hidden_nf, 1, bias=False)           # created first: keeps the RNG order of :107-108
        nn.init.xavier_uniform_(last.weight, gain=0.001)
        self.coord_mlp = _mlp(nn.Linear(2 * hidden_nf + edges_in_d, hidden_nf), act_fn,
                              nn.Linear(hidden_nf, hidden_nf), act_fn, last)
        self.normalization_factor = normalization_factor
        self.aggregation_method = aggregation_method

    def forward(self, h, coord, edge_index, coord_diff, edge_attr=None, node_mask=None, edge_mask=None):
        return runtime.equiv_update_forward(self, h, coord, edge_index, coord_diff, edge_attr, node_mask, edge_mask)


class EquivariantBlock(nn.Module):
    def __init__(self, hidden_nf, edge_feat_nf=2, device="cpu", act_fn=nn.SiLU(), n_layers=2, attention=True,
                 norm_diff=True, tanh=False, coords_range=15, norm_constant=1, sin_embedding=None,
                 normalization_factor=100, aggregation_method="sum"):
        super().__init__()
        if sin_embedding is not None:
            raise NotImplementedError("sin_embedding is off in args_edm and is not implemented")
        self.hidden_nf = hidden_nf
        self.device = device
        self.n_layers = n_layers
        self.coords_range_layer = float(coords_range)
        self.norm_diff = norm_diff
        self.norm_constant = norm_constant
        self.sin_embedding = None
        self.normalization_factor = normalization_factor
        self.aggregation_method = aggregation_method
        for i in range(n_layers):
            self.add_module(f"gcl_{i}", GCL(hidden_nf, hidden_nf, hidden_nf, edges_in_d=edge_feat_nf,
                                            act_fn=act_fn, attention=attention,
                                            normalization_factor=normalization_factor,
                                            aggregation_method=aggregation_method))
        self.add_module("gcl_equiv", EquivariantUpdate(hidden_nf, edges_in_d=edge_feat_nf, act_fn=nn.SiLU(),
                                                       tanh=tanh, coords_range=self.coords_range_layer,
                                                       normalization_factor=normalization_factor,
                                                       aggregation_method=aggregation_method))
        self.to(device)

    def forward(self, h, x, edge_index, node_mask=None, edge_mask=None, edge_attr=None):
        return runtime.equiv_block_forward(self, h, x, edge_index, node_mask, edge_mask, edge_attr)


class EGNN(nn.Module):
    def __init__(self, in_node_nf, in_edge_nf, hidden_nf, device="cpu", act_fn=nn.SiLU(), n_layers=3,
                 attention=False, norm_diff=True, out_node_nf=None, tanh=False, coords_range=15,
                 norm_constant=1, inv_sublayers=2, sin_embedding=False, normalization_factor=100,
                 aggregation_method="sum"):
        super().__init__()
        if sin_embedding:
            raise NotImplementedError("sin_embedding is off in args_edm and is not implemented")
        if out_node_nf is None:
            out_node_nf = in_node_nf
        self.hidden_nf = hidden_nf
        self.device = device
        self.n_layers = n_layers
        self.coords_range_layer = float(coords_range / n_layers)   # computed, never used (egnn_new.py:264 vs :290)
        self.norm_diff = norm_diff
        self.normalization_factor = normalization_factor
        self.aggregation_method = aggregation_method
        self.sin_embedding = None
        self.embedding = nn.Linear(in_node_nf, hidden_nf)
        self.embedding_out = nn.Linear(hidden_nf, out_node_nf)
        for i in range(n_layers):
            self.add_module(f"e_block_{i}", EquivariantBlock(
                hidden_nf, edge_feat_nf=2, device=device, act_fn=act_fn, n_layers=inv_sublayers,
                attention=attention, norm_diff=norm_diff, tanh=tanh, coords_range=coords_range,
                norm_constant=norm_constant, sin_embedding=None, normalization_factor=normalization_factor,
                aggregation_method=aggregation_method))
        self.to(device)

    def forward(self, h, x, edge_index, node_mask=None, edge_mask=None):
        return runtime.egnn_forward(self, h, x, edge_index, node_mask, edge_mask)


class EGNN_dynamics(nn.Module):
    """Denoiser wrapper: eps = phi(z_t, t).  edm/egnn/models.py:8-152."""

    def __init__(self, in_node_nf, context_node_nf=0, n_dims=3, hidden_nf=64, device="cpu",
                 act_fn=torch.nn.SiLU(), n_layers=4, attention=False, condition_time=True, tanh=False,
                 mode="egnn_dynamics", norm_constant=0, inv_sublayers=2, sin_embedding=False,
                 normalization_factor=100, aggregation_method="sum", coords_range=15):
        super().__init__()
        if mode != "egnn_dynamics":
            raise NotImplementedError("mode 'gnn_dynamics' is never selected by GaUDI and is not implemented")
        if context_node_nf != 0:
            raise NotImplementedError("context conditioning is unused on the guided-sampling path")
        if n_dims != 3:
            raise NotImplementedError("n_dims must be 3")
        self.mode = mode
        if condition_time:
            in_node_nf += 1
        self.egnn = EGNN(in_node_nf=in_node_nf + context_node_nf, in_edge_nf=1, hidden_nf=hidden_nf,
                         device=device, act_fn=act_fn, n_layers=n_layers, attention=attention, tanh=tanh,
                         norm_constant=norm_constant, inv_sublayers=inv_sublayers, sin_embedding=sin_embedding,
                         normalization_factor=normalization_factor, aggregation_method=aggregation_method,
                         coords_range=coords_range)
        self.in_node_nf = in_node_nf
        self.context_node_nf = context_node_nf
        self.device = device
        self.n_dims = n_dims
        self._edges_dict = {}
        self.condition_time = condition_time
        # hyper-parameters the kernels need (the reference keeps them scattered over sub-modules)
        self.hyper = dict(hidden_nf=hidden_nf, n_layers=n_layers, attention=bool(attention), tanh=bool(tanh),
                          norm_constant=float(norm_constant), inv_sublayers=int(inv_sublayers),
                          normalization_factor=float(normalization_factor), coords_range=float(coords_range))

    def forward(self, t, xh, node_mask, edge_mask, context=None):
        raise NotImplementedError

    def wrap_forward(self, node_mask, edge_mask, context):
        return lambda time, state: self._forward(time, state, node_mask, edge_mask, context)

    def unwrap_forward(self):
        return self._forward

    def _forward(self, t, xh, node_mask, edge_mask, context=None):
        if context is not None:
            raise NotImplementedError("context conditioning is unused on the guided-sampling path")
        return runtime.denoiser_forward(self, t, xh, node_mask, edge_mask)

    def get_adj_matrix(self, n_nodes, batch_size, device):
        """Dense (i,j) list incl. self loops, molecule-major (edm/egnn/models.py:154-175) -- vectorised;
        the kernels never read it (they use the compacted CSR built from the masks)."""
        key = (n_nodes, batch_size)
        if key not in self._edges_dict:
            base = torch.arange(batch_size, device=device).repeat_interleave(n_nodes * n_nodes) * n_nodes
            i = torch.arange(n_nodes, device=device).repeat_interleave(n_nodes).repeat(batch_size)
            j = torch.arange(n_nodes, device=device).repeat(n_nodes * batch_size)
            self._edges_dict[key] = [base + i, base + j]
        return self._edges_dict[key]
