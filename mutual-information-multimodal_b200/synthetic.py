"""Synthetic "CXR-shaped" embeddings for benchmarks and examples.

The image embedding is post-ReLU (>= 0: ``ResNet256_6_2_1`` ends in ReLU + average pooling, reference
model.py:355-365); the text embedding is tanh-pooled in (-1, 1) (BERT pooler, model.py:76-77).  Positives are
correlated (Y depends on X through a fixed random projection) so that the diagonal of the score matrix is
informative and the softmax moderately peaked.  ``oracle/matrix_oracle.py`` carries an independent copy of
this generator for the tests (``tests/test_host_cpu.py`` asserts that the two agree bit for bit); the
product never imports ``oracle/``.
"""
from __future__ import annotations

import math

import torch


def synthetic_embeddings(B: int, D: int, seed: int = 1234, dup_frac: float = 0.0, device="cpu", bilinear: bool = True):
    """Returns fp32 (X [B, D], Y [B, D], study ids [B] int64, W [D, D] or None); callers round to bf16.
    ``dup_frac`` of the ids are duplicated onto their neighbour to exercise the negatives mask (main_utils.py:105)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    X = torch.relu(torch.randn(B, D, generator=g))
    P0 = torch.randn(D, D, generator=torch.Generator().manual_seed(99)) / math.sqrt(D)
    Y = torch.tanh(0.5 * (X @ P0 + torch.randn(B, D, generator=g)))
    sid = torch.arange(B, dtype=torch.long)
    if dup_frac > 0:
        n_dup = int(B * dup_frac)
        idx = torch.randperm(B - 1, generator=g)[:n_dup]
        sid[idx + 1] = sid[idx]
    W = None
    if bilinear:
        gw = torch.Generator().manual_seed(7)
        W = (torch.eye(D) + 0.1 * torch.randn(D, D, generator=gw) / math.sqrt(D)) / math.sqrt(D)
    return X.to(device), Y.to(device), sid.to(device), (None if W is None else W.to(device))
