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
