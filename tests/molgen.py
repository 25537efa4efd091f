"""Seeded synthetic graphs of rings for the validity tests (shared by tests/golden/make_golden_validity.py and the tests).

Rings are grown as a tree on 60-degree directions at the mid-range centre distance of each ring pair, then perturbed
(Gaussian noise of several scales, occasional stretched / collapsed / detached rings, wrong orientation types), so that
every flag of check_stability takes both values."""
import json
import os

import numpy as np

_TAB = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gaudi_b200", "ring_tables.json")))


def _pair(tab, rings, a, b):
    k = f"{rings[a]}-{rings[b]}"
    if k not in tab:
        k = f"{rings[b]}-{rings[a]}"
    return tab.get(k)


def molecule(rng, dataset, n, noise, defect):
    tab, rings = _TAB[dataset]["distances"], _TAB[dataset]["rings"]
    n_types = len(rings) - (1 if dataset == "hetro" else 0)
    types = [int(rng.integers(0, n_types))]
    pos = [np.zeros(3)]
    tries = 0
    while len(pos) < n and tries < 500:
        tries += 1
        parent = int(rng.integers(0, len(pos)))
        t = int(rng.integers(0, n_types))
        pr = _pair(tab, rings, types[parent], t)
        if pr is None:
            continue
        ang = np.deg2rad(60.0 * rng.integers(0, 6))
        d = 0.5 * (pr[0] + pr[1])
        cand = pos[parent] + d * np.array([np.cos(ang), np.sin(ang), 0.0])
        if min(np.linalg.norm(cand - q) for q in pos) < 1.9:
            continue
        if defect != 6 and any(np.linalg.norm(cand - q) < 3.1 for k, q in enumerate(pos) if k != parent):
            continue                              # tree-like (cata-condensed) unless asked for fused triangles
        pos.append(cand)
        types.append(t)
    pos = np.stack(pos)
    if defect == 1 and len(pos) > 1:          # detach one ring
        pos[-1] += np.array([9.0, 0.0, 0.0])
    elif defect == 2 and len(pos) > 1:        # collapse two rings
        pos[-1] = pos[0] + np.array([0.6, 0.1, 0.0])
    elif defect == 3 and len(pos) > 2:        # buckle out of plane
        pos[:, 2] += rng.normal(0, 0.9, len(pos))
    pos = pos + rng.normal(0, noise, pos.shape)
    # random rotation
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])
    pos = (pos - pos.mean(0)) @ R.T
    types = np.array(types, dtype=np.int64)
    if dataset == "hetro":
        m = len(pos)
        orient = pos + rng.normal(0, 0.5, pos.shape)
        otypes = np.full(m, len(rings) - 1, dtype=np.int64)
        if defect == 4:
            otypes[int(rng.integers(0, m))] = 0
        elif defect == 5:
            types[int(rng.integers(0, m))] = len(rings) - 1
        pos, types = np.concatenate([pos, orient]), np.concatenate([types, otypes])
    return pos.astype(np.float32), types


def batch(seed, dataset, count, max_rings):
    """(x [B,N,3] fp32, ring_type [B,N] int64, node_mask [B,N] fp32) padded like the sampler's output."""
    rng = np.random.default_rng(seed)
    N = max_rings * (2 if dataset == "hetro" else 1)
    x = np.zeros((count, N, 3), np.float32)
    rt = np.zeros((count, N), np.int64)
    nm = np.zeros((count, N), np.float32)
    for b in range(count):
        n = int(rng.integers(1, max_rings + 1)) if rng.random() < 0.3 else max_rings - int(rng.integers(0, 2))
        noise = float(rng.choice([0.0, 0.01, 0.03, 0.08, 0.2]))
        defect = int(rng.choice([0, 0, 0, 0, 0, 0, 1, 2, 3, 4, 5, 6]))
        p, t = molecule(rng, dataset, n, noise, defect)
        m = len(p) // 2 if dataset == "hetro" else len(p)
        if dataset == "hetro":
            x[b, :m], x[b, max_rings:max_rings + m] = p[:m], p[m:]
            rt[b, :m], rt[b, max_rings:max_rings + m] = t[:m], t[m:]
            nm[b, :m] = nm[b, max_rings:max_rings + m] = 1.0
        else:
            x[b, :m], rt[b, :m], nm[b, :m] = p, t, 1.0
    return x, rt, nm
