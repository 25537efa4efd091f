#!/usr/bin/env python
"""Dump the reference's ring-geometry statistics (data constants, not code) to gaudi_b200/ring_tables.json.

Sources: utils/helpers.py:11-63 (angle quantiles), :96-161 (ring-centre distance ranges), data/aromatic_dataloader.py:31-35
(ring-type vocabularies).  Run in the build container only:  PYTHONDONTWRITEBYTECODE=1 python tools/extract_ring_tables.py
"""
import json
import os
import sys
from unittest.mock import MagicMock

REF = os.environ.get("GAUDI_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
for _m in ["rdkit", "rdkit.Chem", "rdkit.Chem.Draw", "matplotlib", "matplotlib.pyplot", "imageio"]:
    sys.modules[_m] = MagicMock()

from utils import helpers as H  # noqa: E402
from data.aromatic_dataloader import RINGS_LIST  # noqa: E402

out = {}
for ds in ("cata", "hetro"):
    out[ds] = {
        "rings": list(RINGS_LIST[ds]),
        "distances": {k: list(v) for k, v in H.ring_distances[ds].items()},
        "angels3": {sym: [list(r) for r in d.values()] for sym, d in H.angels3_dict[ds].items()},
        "angels4": {k: float(v) for k, v in H.angels4_dict[ds].items()},
    }
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gaudi_b200", "ring_tables.json")
with open(dst, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote", dst)
