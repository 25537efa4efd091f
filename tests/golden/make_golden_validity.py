"""Golden flags of the UNMODIFIED reference's check_stability / positions2adj on seeded synthetic ring graphs.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_validity.py

Writes tests/golden/validity_{cata,hetro}.npz (inputs + the five flags per molecule + dist/adj of positions2adj).
rdkit / matplotlib / imageio are stubbed (only the rdkit-based chemistry check, which is not called, needs them)."""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

REF = os.environ.get("GAUDI_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))
for _m in ["rdkit", "rdkit.Chem", "rdkit.Chem.Draw", "rdkit.Chem.rdchem", "matplotlib", "matplotlib.pyplot", "imageio"]:
    sys.modules[_m] = MagicMock()

from analyze.analyze import check_stability  # noqa: E402
from utils.helpers import positions2adj  # noqa: E402
import molgen  # noqa: E402

NAMES = ("orientation_nodes", "dist_stable", "connected", "angels3", "angels4")


def main():
    for ds, seed, count, max_rings in [("cata", 101, 400, 11), ("hetro", 202, 300, 10)]:
        x, rt, nm = molgen.batch(seed, ds, count, max_rings)
        flags = np.zeros((count, 5), np.uint8)
        for b in range(count):
            m = nm[b].astype(bool)
            res = check_stability(torch.from_numpy(x[b][m]), torch.from_numpy(rt[b][m]), tol=0.1, dataset=ds)
            flags[b] = [int(bool(res[k])) for k in NAMES]
        dist, adj = positions2adj(torch.from_numpy(x[:16]), torch.from_numpy(rt[:16]), 0.1, dataset=ds)
        np.savez_compressed(os.path.join(HERE, f"validity_{ds}.npz"), x=x, ring_type=rt, node_mask=nm, flags=flags,
                            dist16=dist.numpy(), adj16=adj.numpy())
        print(ds, "flag means", flags.mean(0), "stable", flags.all(1).mean())


if __name__ == "__main__":
    main()
