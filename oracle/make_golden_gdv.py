"""Generate tests/golden/gdv_*.npz by EXECUTING the reference's own GDV functions (validate.py:16-49, loaded by
oracle/ref_loader.load_gdv).  Run in the dev container only:   python -m oracle.make_golden_gdv"""
from __future__ import annotations

import os

import numpy as np

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [("gdv_p37_n59_d96", 37, 59, 96, 31, False), ("gdv_p64_n40_d24_constcol", 64, 40, 24, 32, True),
         ("gdv_p200_n312_d128", 200, 312, 128, 33, False)]


def inputs(Np, Nn, D, seed, const_col):
    r = np.random.RandomState(seed)
    centre = r.randn(D).astype(np.float32)
    pos = (np.maximum(r.randn(Np, D), 0) + 0.4 * centre).astype(np.float32)          # post-ReLU image embeddings
    neg = (np.maximum(r.randn(Nn, D), 0) - 0.2 * centre + 0.3 * r.randn(Nn, 1)).astype(np.float32)
    if const_col:
        pos[:, 3] = 0.0                      # a dead ReLU feature: StandardScaler leaves it unscaled
        neg[:, 3] = 0.0
        neg[:, 7] = 1.25
    return pos, neg


def main():
    ref = ref_loader.load_gdv()
    os.makedirs(OUT, exist_ok=True)
    for name, Np, Nn, D, seed, cc in CASES:
        pos, neg = inputs(Np, Nn, D, seed, cc)
        zp, zn = ref.z_scored_transform(source_tensor=pos), ref.z_scored_transform(source_tensor=neg)
        payload = dict(pos=pos, neg=neg, intra_pos=np.float64(ref.mean_intra_class_distance(zp)),
                       intra_neg=np.float64(ref.mean_intra_class_distance(zn)),
                       inter=np.float64(ref.mean_inter_class_distance(source=zp, dest=zn)),
                       gdv=np.float64(ref.gdv_calculation(list(pos), list(neg))))   # lists of rows, as validate.py:118-127 builds them
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **payload)
        print(name, {k: float(v) for k, v in payload.items() if k not in ("pos", "neg")})


if __name__ == "__main__":
    main()
