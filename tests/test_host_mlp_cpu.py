"""CPU tests (not gpu) of the host-side mirror of the reference's own critic and GDV entry points: FusedMLPCritic is the
reference's nn.Sequential (names, shapes, initialisation, outputs on explicit pair rows), and neither it nor
gdv_calculation has a CPU fallback."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mlp_oracle, ref_loader  # noqa: E402


def test_fused_mlp_critic_is_the_reference_sequential():
    import mi_b200
    torch.manual_seed(7)
    c = mi_b200.FusedMLPCritic(768, (1024, 512))
    assert list(c.state_dict().keys()) == ["0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias"]
    assert c[0].weight.shape == (1024, 1536) and c[2].weight.shape == (512, 1024) and c[4].weight.shape == (1, 512)
    p = mlp_oracle.init_params(768, 1024, 512, seed=7)                 # make_mlp's construction order under the same seed
    for k, t in zip(mlp_oracle.PARAM_NAMES, [c[0].weight, c[0].bias, c[2].weight, c[2].bias, c[4].weight, c[4].bias]):
        assert torch.equal(t.detach(), p[k]), k
    rows = torch.randn(9, 1536)
    torch.testing.assert_close(c(rows), mlp_oracle.mlp_logits_on_rows(rows, p))      # compat path = the module itself
    assert len(list(c.parameters())) == 6                              # optim.Adam(self.mi_discriminator.parameters()), main_utils.py:153


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
def test_reference_checkpoint_loads_into_fused_mlp_critic():
    import mi_b200
    ref = ref_loader.load()
    torch.manual_seed(3)
    net = ref.make_mlp(1536, [1024, 512])                              # main_utils.py:77
    c = mi_b200.FusedMLPCritic(768, (1024, 512))
    c.load_state_dict(net.state_dict())                                # same keys: a reference checkpoint drops in
    rows = torch.randn(5, 1536)
    assert torch.equal(c(rows), net(rows))


def test_no_cpu_fallback_for_mlp_critic_and_gdv():
    import mi_b200
    B, D = 6, 16
    x, y = torch.randn(B, D), torch.randn(B, D)
    c = mi_b200.FusedMLPCritic(D, (32, 16))
    handle = c(mi_b200.create_mi_pairs(x, y, [str(i) for i in range(B)], torch.device("cpu")))
    assert isinstance(handle, mi_b200.ScoreHandle)
    with pytest.raises(mi_b200.MIError):
        mi_b200.dv_bound_loss(handle, B, torch.device("cpu"))
    with pytest.raises(mi_b200.MIError):
        mi_b200.gdv_calculation(np.random.rand(8, 16).astype(np.float32), np.random.rand(9, 16).astype(np.float32), device="cpu")
    with pytest.raises(ValueError):
        mi_b200.FusedMLPCritic(D, (32, 16, 8))                         # the fused path is the two-hidden-layer critic


def test_mlp_and_gdv_workspace_planning_without_gpu():
    from mi_b200 import _lib
    lib = _lib.load()
    assert lib.mi_mlp_critic_workspace_bytes(256, 768, 1024, 512, 0) > 0
    assert lib.mi_mlp_critic_workspace_bytes(256, 768, 1024, 512, 1) > lib.mi_mlp_critic_workspace_bytes(256, 768, 1024, 512, 0)
    assert lib.mi_mlp_critic_workspace_bytes(256, 768, 1024, 520, 0) == 0          # H2 <= 512
    assert lib.mi_mlp_critic_workspace_bytes(256, 770, 1024, 512, 0) == 0          # D % 8
    assert lib.mi_gdv_workspace_bytes(1000, 1200, 768, 1) > 0
    assert lib.mi_gdv_workspace_bytes(1, 1200, 768, 1) == 0                        # at least two samples per class


def _single_pass_emulation(X, Y, sid, p, panel_rows, sample, dtype=torch.float64):
    """Plain-torch restatement of csrc/mlp_critic.cuh::mlp_single_pass (dv): the algebra the kernels implement, panel by
    panel — softmax weights against the maximum logit of a strided SAMPLE of the pairs times 2^-69, every sum kept in those
    reference units, one multiplication by 1 / sum at the end, the positive pairs as a separate panel with weight -1/B."""
    X, Y = X.to(dtype), Y.to(dtype)
    W1, b1, W2, b2, W3, b3 = (p[k].to(dtype) for k in mlp_oracle.PARAM_NAMES)
    B, D = X.shape
    w3 = W3.reshape(-1)
    A = X @ W1[:, :D].T + b1                                   # layer 1 is separable: W1 [x ; y] + b1 = A_i + C_j
    C = Y @ W1[:, D:].T
    ids = torch.as_tensor(sid)
    incl = ids[:, None] != ids[None, :]                        # main_utils.py:105
    ns = min(B, sample)
    st = B // ns
    rows = torch.arange(ns) * st

    def logits(a_rows, c_rows):
        h = torch.relu(a_rows[:, None, :] + c_rows[None, :, :])
        z2 = h @ W2.T + b2
        return h, z2, torch.relu(z2) @ w3 + b3

    m_s = logits(A[rows], C[rows])[2].max()
    scale = 2.0 ** -69
    dW2 = torch.zeros_like(W2); dw3 = torch.zeros_like(w3); db2 = torch.zeros_like(b2)
    dA = torch.zeros_like(A); dC = torch.zeros_like(C)
    rowsum = torch.zeros(B, dtype=dtype); diag = torch.zeros(B, dtype=dtype)
    for r0 in range(0, B, panel_rows):
        r1 = min(B, r0 + panel_rows)
        h, z2, S = logits(A[r0:r1], C)                         # [rr, B, H1], [rr, B, H2], [rr, B]
        g = torch.where(incl[r0:r1], torch.exp(S - m_s) * scale, torch.zeros_like(S))
        rowsum[r0:r1] = g.sum(1)
        diag[r0:r1] = S[torch.arange(r1 - r0), torch.arange(r0, r1)]
        pos = (z2 > 0).to(dtype)
        dz2 = g[..., None] * w3 * pos
        dw3 += (g[..., None] * torch.relu(z2)).sum((0, 1)); db2 += dz2.sum((0, 1))
        dW2 += torch.einsum("ijn,ijk->nk", dz2, h)
        dz1 = (dz2 @ W2) * (h > 0).to(dtype)
        dA[r0:r1] += dz1.sum(1); dC += dz1.sum(0)
    total = rowsum.sum()
    c = 1.0 / total
    for t in (dW2, dw3, db2, dA, dC):
        t *= c
    # the positive pairs (i, i): dL/dlogit = -1/B (mi_critics.py:6)
    h = torch.relu(A + C); z2 = h @ W2.T + b2
    dz2 = (-1.0 / B) * w3 * (z2 > 0).to(dtype)
    dw3 += (-1.0 / B) * torch.relu(z2).sum(0); db2 += dz2.sum(0)
    dW2 += dz2.T @ h
    dz1 = (dz2 @ W2) * (h > 0).to(dtype)
    dA += dz1; dC += dz1
    n_neg = incl.sum().to(dtype)
    lse = m_s + 69.0 * np.log(2.0) + torch.log(total)
    loss = lse - torch.log(n_neg) - diag.mean()               # mi_critics.py:3-12
    return {"loss": loss, "dX": dA @ W1[:, :D], "dY": dC @ W1[:, D:], "dW1": torch.cat([dA.T @ X, dC.T @ Y], 1),
            "db1": dA.sum(0), "dW2": dW2, "db2": db2, "dW3": dw3.reshape(1, -1)}


@pytest.mark.parametrize("B,panel,sample", [(24, 24, 256), (37, 5, 8), (64, 16, 16)])
def test_single_pass_algebra_matches_the_autograd_oracle(B, panel, sample):
    """Reference units + one final 1/sum + a separate positive-pair panel give exactly the gradients autograd takes of
    dv_bound_loss over the reference's pair rows — for one panel, ragged panels and a sample that misses the maximum."""
    from oracle import matrix_oracle as mo
    D, H1, H2 = 12, 40, 24
    X, Y, sid, _ = mo.synthetic_embeddings(B, D, seed=B, dup_frac=0.1, bilinear=False)
    p = mlp_oracle.init_params(D, H1, H2, seed=3, dtype=torch.float64)
    p["W2"] = p["W2"] * 2.0
    p["W3"] = p["W3"] * 6.0
    ref = mlp_oracle.mlp_loss_matrix_form(X.double(), Y.double(), [int(s) for s in sid], p, "dv", dtype=torch.float64)
    got = _single_pass_emulation(X, Y, [int(s) for s in sid], p, panel, sample)
    assert abs(float(got["loss"]) - float(ref["loss"])) < 1e-6 * max(1.0, abs(float(ref["loss"])))      # log N_neg is taken in fp32 (mi_critics.py:10)
    for k in ("dX", "dY", "dW1", "db1", "dW2", "db2", "dW3"):
        torch.testing.assert_close(got[k], ref[k].reshape(got[k].shape).double(), rtol=1e-8, atol=1e-12, msg=k)
