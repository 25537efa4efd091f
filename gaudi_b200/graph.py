"""Mask construction and the compacted edge topology the kernels iterate over.

Reference code replaced:
  * ``node2edge_mask`` and the mask loops of ``sample_guidance`` / ``sample_pos_edm``
    (sampling_edm.py:119-125, 137-160, 172-209) -- vectorised, bit-exact 0/1 fp32 results;
  * ``get_adj_matrix`` (edm/egnn/models.py:154-175): the dense (i,j) list is never materialised; instead the
    edges with ``edge_mask != 0`` are kept, in the same molecule-major / row-major order, as a CSR structure
    plus a tiling into GEMM tiles of <=128 edges that never splits the edges of one row node.

Everything here is integer/index work done once per mask set (torch ops on the tensors' own device plus the
``gb_tile_pack`` host helper of the C library).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

TILE = 128


def node2edge_mask(node_mask: torch.Tensor) -> torch.Tensor:
    """[B,N] -> [B,N,N] outer product without the diagonal (sampling_edm.py:119-125)."""
    em = node_mask.unsqueeze(1) * node_mask.unsqueeze(2)
    n = node_mask.size(1)
    return em * (~torch.eye(n, dtype=torch.bool, device=node_mask.device)).unsqueeze(0)


def build_masks(nodesxsample: torch.Tensor, max_nodes: int, orientation: bool, device="cpu"):
    """node_mask [B,N,1], edge_mask [B*N*N,1] exactly as sampling_edm.py:172-209 builds them.

    ``orientation`` (dataset != 'cata') appends one orientation node per ring: node_mask = [m, m] and
    edge_mask = [[ring-ring, I], [I, 0]] (the identity blocks are present even for masked rings).
    """
    nx = torch.as_tensor(nodesxsample).to(device=device, dtype=torch.long).view(-1)
    B = nx.numel()
    nm = (torch.arange(max_nodes, device=device).unsqueeze(0) < nx.unsqueeze(1)).to(torch.float32)
    em = node2edge_mask(nm)
    nm = nm.unsqueeze(2)
    if orientation:
        eye = torch.eye(max_nodes, device=device).unsqueeze(0).expand(B, -1, -1)
        zero = torch.zeros(B, max_nodes, max_nodes, device=device)
        em = torch.cat([torch.cat([em, eye], dim=1), torch.cat([eye, zero], dim=1)], dim=2)
        nm = torch.cat([nm, nm], dim=1)
    return nm.contiguous(), em.reshape(-1, 1).contiguous()


@dataclass
class Topology:
    B: int
    N: int
    n_edges: int
    n_tiles: int
    n_tc: int
    rowptr: torch.Tensor
    erow: torch.Tensor
    ecol: torch.Tensor
    tile_ptr: torch.Tensor
    tc_ptr: torch.Tensor
    tc_node: torch.Tensor
    tc_start: torch.Tensor
    cperm: torch.Tensor
    node_mask: torch.Tensor     # flat fp32 [B*N]
    dense_idx: torch.Tensor = None   # int64 [n_edges]: position of every kept edge in the dense B*N*N list
    colptr: torch.Tensor = None      # int32 [n_nodes+1] CSC pointer (edges grouped by column node)
    cedge: torch.Tensor = None       # int32 [n_edges] edge ids in CSC order


def tile_pack(rowptr_host: np.ndarray, nodes_per_graph: int = 0) -> np.ndarray:
    """Greedy tiling of consecutive nodes (<=128 edges and <=128 nodes per tile); C helper ``gb_tile_pack_graphs``.
    With ``nodes_per_graph`` (the padded N) a tile also keeps  nodes + N * graphs_touched <= gb_stage_rows()  so that the
    edge kernels can stage the node projections of the tile in shared memory."""
    rowptr_host = np.ascontiguousarray(rowptr_host, dtype=np.int32)
    n_nodes = rowptr_host.shape[0] - 1
    out = np.zeros(n_nodes + 1, dtype=np.int32)
    n_tiles = C.c_int(0)
    _lib.check(_lib.lib().gb_tile_pack_graphs(rowptr_host.ctypes.data_as(C.c_void_p), n_nodes, int(nodes_per_graph),
                                              out.ctypes.data_as(C.c_void_p), C.byref(n_tiles)))
    return out[: n_tiles.value + 1].copy()


def build_topology(node_mask: torch.Tensor, edge_mask: torch.Tensor, B: int, N: int) -> Topology:
    dev = node_mask.device
    nm = node_mask.reshape(B * N).to(torch.float32).contiguous()
    em = edge_mask.reshape(B * N, N)
    if not bool(((em == 0) | (em == 1)).all()):
        raise ValueError("edge_mask must contain only 0/1 values")
    valid = em != 0
    counts = valid.sum(1)
    rowptr = torch.zeros(B * N + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(counts, 0)
    flat = valid.reshape(-1).nonzero(as_tuple=False).view(-1)          # dense edge ids, ascending = reference order
    erow = torch.div(flat, N, rounding_mode="floor")
    ecol = torch.div(erow, N, rounding_mode="floor") * N + flat % N
    n_edges = int(flat.numel())
    tile_ptr_h = tile_pack(rowptr.cpu().numpy().astype(np.int32), N)
    n_tiles = tile_ptr_h.shape[0] - 1
    tile_ptr = torch.from_numpy(tile_ptr_h).to(dev)
    # per-tile grouping of the edges by column node (backward column scatter)
    tile_e0 = rowptr[tile_ptr.long()]                                   # first edge of every tile (+ sentinel)
    e_idx = torch.arange(n_edges, device=dev)
    tile_of_e = torch.searchsorted(tile_e0, e_idx, right=True) - 1
    key = tile_of_e * (B * N) + ecol
    order = torch.argsort(key, stable=True)
    cperm = (e_idx - tile_e0[tile_of_e])[order]
    skey = key[order]
    uniq, cnt = torch.unique_consecutive(skey, return_counts=True)
    n_tc = int(uniq.numel())
    tc_node = uniq % (B * N)
    tc_start = torch.zeros(n_tc + 1, dtype=torch.int64, device=dev)
    tc_start[1:] = torch.cumsum(cnt, 0)
    tc_tile = torch.div(uniq, B * N, rounding_mode="floor")
    tc_ptr = torch.searchsorted(tc_tile.contiguous(), torch.arange(n_tiles + 1, device=dev), right=False)
    cedge = torch.argsort(ecol, stable=True)
    colptr = torch.zeros(B * N + 1, dtype=torch.int64, device=dev)
    colptr[1:] = torch.cumsum(torch.bincount(ecol, minlength=B * N), 0)
    i32 = lambda t: t.to(torch.int32).contiguous()
    return Topology(B, N, n_edges, n_tiles, n_tc, i32(rowptr), i32(erow), i32(ecol), i32(tile_ptr), i32(tc_ptr),
                    i32(tc_node), i32(tc_start), i32(cperm), nm, flat, i32(colptr), i32(cedge))
