"""CPU oracle (TEST INFRASTRUCTURE) for the Generalised Discrimination Value of validate.py:16-49.

    z      = StandardScaler().fit_transform(class embeddings)        per class, per feature, population std   (:16-21)
    intra  = sum_ij |z_i - z_j| * 2 / (T (T - 1)),   T = N * D         the reference's normaliser counts ELEMENTS (:23-27)
    inter  = sum_ij |p_i - n_j| / (Np D * Nn D)                                                               (:29-34)
    gdv    = ((intra_p + intra_n) / 2 - inter) / sqrt(Np + Nn)                                                (:37-49)

Restated with numpy float64 (sklearn upcasts float32 inputs to float64 for the Euclidean distances).  Pinned against
the reference's own functions: ``tests/golden/gdv_*.npz`` are produced by executing them (``oracle/make_golden_gdv.py``)
and ``tests/test_oracle_gdv.py`` compares.  Only tests/ and bench legs may import this module."""
from __future__ import annotations

import math

import numpy as np


def z_scored_transform(x: np.ndarray) -> np.ndarray:
    """StandardScaler semantics (validate.py:16-21): mean 0, population std 1 per feature; a feature whose variance is
    (numerically) zero is left unscaled (sklearn's _handle_zeros_in_scale)."""
    x = np.asarray(x)
    xd = x.astype(np.float64)
    mean = xd.mean(axis=0)
    var = xd.var(axis=0)
    n = xd.shape[0]
    eps = np.finfo(np.float64).eps
    constant = var <= n * eps * var + (n * mean * eps) ** 2          # sklearn.preprocessing._data._is_constant_feature
    scale = np.where(constant, 1.0, np.sqrt(var))
    return ((xd - mean) / scale).astype(x.dtype if x.dtype.kind == "f" else np.float64)


def _pair_sum(a: np.ndarray, b: np.ndarray, same: bool) -> float:
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    d2 = (a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * (a @ b.T)
    np.maximum(d2, 0.0, out=d2)
    if same:
        np.fill_diagonal(d2, 0.0)
    return float(np.sqrt(d2).sum())


def mean_intra_class_distance(items: np.ndarray) -> float:
    total = items.shape[0] * items.shape[1]                      # validate.py:25 — N * D, not N
    return _pair_sum(items, items, True) * 2 / (total * (total - 1))


def mean_inter_class_distance(source: np.ndarray, dest: np.ndarray) -> float:
    return _pair_sum(source, dest, False) / ((source.shape[0] * source.shape[1]) * (dest.shape[0] * dest.shape[1]))


def gdv_calculation(positive_embeddings, negative_embeddings) -> dict:
    p = z_scored_transform(np.asarray(positive_embeddings))
    n = z_scored_transform(np.asarray(negative_embeddings))
    ip, in_, it = mean_intra_class_distance(p), mean_intra_class_distance(n), mean_inter_class_distance(p, n)
    inv = 1 / math.sqrt(len(positive_embeddings) + len(negative_embeddings))
    return {"gdv": inv * ((ip + in_) / 2 - it), "intra_pos": ip, "intra_neg": in_, "inter": it}
