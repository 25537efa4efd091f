"""CPU oracle of the geometric validity checks -- TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline).

Restates, per molecule and in the reference's own fp32 torch arithmetic:
  positions2adj          utils/helpers.py:164-196
  check_stability        analyze/analyze.py:50-100   (networkx replaced by an explicit BFS with ascending neighbours,
                                                       which is the order nx.from_numpy_array + bfs_edges produce)
  find_triplets_quads    analyze/analyze.py:276-318, angel3 :234-240, angel4 :243-273
  check_angels3/4        analyze/analyze.py:19-47
Pinned against the reference's check_stability on tests/golden/validity_*.npz (tests/golden/make_golden_validity.py).
The ring statistics come from the same extracted table the product uses (gaudi_b200/ring_tables.json, data only).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Tuple

import torch

_TAB = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gaudi_b200", "ring_tables.json")))
FLAG_NAMES = ("orientation_nodes", "dist_stable", "connected", "angels3", "angels4")


def positions2adj(x: torch.Tensor, ring_type: torch.Tensor, tol: float, dataset: str) -> Tuple[torch.Tensor, torch.Tensor]:
    rings, table = _TAB[dataset]["rings"], _TAB[dataset]["distances"]
    xx = x.unsqueeze(2)
    dist = torch.sqrt(torch.sum((xx - xx.transpose(1, 2)) ** 2, 3))
    adj = torch.zeros(dist.shape[0], dist.shape[1], dist.shape[1])
    for b in range(dist.shape[0]):
        for i in range(dist.shape[1]):
            for j in range(i + 1, dist.shape[1]):
                si, sj = rings[int(ring_type[b, i])], rings[int(ring_type[b, j])]
                key = f"{si}-{sj}"
                if key not in table:
                    key = f"{sj}-{si}"
                if key in table and table[key][0] * (1 - tol) < dist[b, i, j] < table[key][1] * (1 + tol):
                    adj[b, i, j] = adj[b, j, i] = 1
    return dist, adj


def angel3(p: torch.Tensor) -> torch.Tensor:
    v1, v2 = p[0] - p[1], p[2] - p[1]
    a = torch.rad2deg(torch.acos(torch.dot(v1, v2) / (torch.norm(v1) * torch.norm(v2))))
    return a if a >= 0 else a + 360


def angel4(p: torch.Tensor) -> torch.Tensor:
    b0, b1, b2 = -1.0 * (p[1] - p[0]), p[2] - p[1], p[3] - p[2]
    b1 = b1 / torch.linalg.norm(b1)
    v = b0 - torch.dot(b0, b1) * b1
    w = b2 - torch.dot(b2, b1) * b1
    return torch.rad2deg(torch.atan2(torch.dot(torch.cross(b1, v, dim=0), w), torch.dot(v, w))).abs()


def _neighbours(adj: torch.Tensor, i: int) -> List[int]:
    return [j for j in range(adj.shape[0]) if adj[i, j] != 0]


def _bfs_edges(adj: torch.Tensor) -> Tuple[List[Tuple[int, int]], int]:
    seen, order, edges = {0}, [0], []
    k = 0
    while k < len(order):
        u = order[k]
        k += 1
        for v in _neighbours(adj, u):
            if v not in seen:
                seen.add(v)
                order.append(v)
                edges.append((u, v))
    return edges, len(order)


def triplets_quads(adj: torch.Tensor, x: torch.Tensor, ring_types: torch.Tensor, dataset: str):
    rings = [_TAB[dataset]["rings"][int(i)] for i in ring_types]
    edges, _ = _bfs_edges(adj)
    trip = []
    for n1, n2 in edges:
        trip += [(n2, n1, n3) for n3 in _neighbours(adj, n1) if n3 != n2]
        trip += [(n1, n2, n3) for n3 in _neighbours(adj, n2) if n3 != n1]
    trip = sorted(set((a, c, b) if a < b else (b, c, a) for a, c, b in trip))
    angels3 = [(rings[t[1]], angel3(x[list(t)])) for t in trip]
    quads = []
    for n1, n2, n3 in [t for t in trip if not 170 < angel3(x[list(t)]) < 190]:
        for n4 in _neighbours(adj, n1):
            if n4 not in (n2, n3) and not 175 < angel3(x[[n4, n1, n2]]) < 185:
                quads.append((n4, n1, n2, n3))
        for n4 in _neighbours(adj, n3):
            if n4 not in (n1, n2) and not 175 < angel3(x[[n2, n3, n4]]) < 185:
                quads.append((n1, n2, n3, n4))
    quads = sorted(set(q if q[0] < q[3] else q[::-1] for q in quads))
    angels4 = [angel4(x[list(q)]) for q in quads]
    return angels3, angels4


def check_stability(positions: torch.Tensor, ring_type: torch.Tensor, tol: float = 0.1, dataset: str = "cata") -> Dict[str, bool]:
    res = {"orientation_nodes": True, "dist_stable": False, "connected": False, "angels3": False, "angels4": False}
    positions = torch.as_tensor(positions, dtype=torch.float32)
    ring_type = torch.as_tensor(ring_type)
    if ring_type.dim() == 2:
        ring_type = ring_type.argmax(1)
    tab = _TAB[dataset]
    if dataset != "cata":
        n = positions.shape[0] // 2
        positions = positions[:n]
        orient = len(tab["rings"]) - 1
        if set(ring_type[n:].tolist()) != {orient} or orient in ring_type[:n].tolist():
            res["orientation_nodes"] = False
            return res
        ring_type = ring_type[:n]
    n = positions.shape[0]
    dist, adj = positions2adj(positions[None], ring_type[None], tol, dataset)
    dist, adj = dist[0], adj[0]
    min_dist = min(r[0] for r in tab["distances"].values())
    if ((dist < min_dist * (1 - tol)) * (1 - torch.eye(n))).bool().any():
        return res
    res["dist_stable"] = True
    if n == 0:
        raise ValueError("null graph")
    if _bfs_edges(adj)[1] != n:
        return res
    res["connected"] = True
    a3, a4 = triplets_quads(adj, positions, ring_type, dataset)
    ok3 = True
    for sym, a in a3:
        ranges = tab["angels3"][sym]                                  # KeyError for a symbol without statistics, as the reference
        ok3 = ok3 and any(bool(lo * (1 - tol) <= a) and bool(a <= hi * (1 + tol)) for lo, hi in ranges)
    res["angels3"] = ok3
    if len(a4) == 0 or dataset == "hetro":
        res["angels4"] = True
    else:
        a4 = torch.stack(a4)
        res["angels4"] = bool(torch.logical_or(tab["angels4"]["180"] * (1 - tol) <= a4, a4 <= tab["angels4"]["0"] * (1 + tol)).all())
    return res


def flags_of(res: Dict[str, bool]) -> List[int]:
    f = [int(res[k]) for k in FLAG_NAMES]
    return f + [int(all(f))]
