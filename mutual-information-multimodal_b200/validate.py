"""Host-side mirror of the reference's GDV separability metric (validate.py:16-49).

``gdv_calculation(positive_embeddings, negative_embeddings)`` keeps the reference's name, argument meaning (two
collections of embedding rows, validate.py:118-127 builds them as Python lists of numpy vectors) and result (a float),
but the z-scoring and the three pairwise-distance sums run on the GPU through ``mi_gdv`` — no N x N matrix, no
``n_jobs=10`` process pool.  There is no CPU fallback."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _as_matrix(rows, device) -> torch.Tensor:
    if torch.is_tensor(rows):
        return rows.to(device=device, dtype=torch.float32)
    return torch.from_numpy(np.asarray(rows, dtype=np.float32)).to(device)


def gdv_terms(positive_embeddings, negative_embeddings, device=None, precision: str = "strict") -> dict:
    """{gdv, intra_pos, intra_neg, inter} — the terms of validate.py:37-49."""
    dev = torch.device("cuda" if device is None else device)
    out = ops.gdv(_as_matrix(positive_embeddings, dev), _as_matrix(negative_embeddings, dev), precision).cpu()
    return {"gdv": float(out[0]), "intra_pos": float(out[1]), "intra_neg": float(out[2]), "inter": float(out[3])}


def gdv_calculation(positive_embeddings, negative_embeddings, device=None, precision: str = "strict") -> float:
    """Drop-in for validate.py:37-49."""
    return gdv_terms(positive_embeddings, negative_embeddings, device, precision)["gdv"]
