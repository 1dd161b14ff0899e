"""Panel-based distillation term (clipk/distill.py) without a GPU: the oracle's restatement of
DistillClipLoss's distillation term against the reference's own outputs (tests/golden/shells/distill_w1_*), and the host
logic of the panel path (panel loop, scales, gradient GEMM orientation) with the kernel entries emulated."""
import os

import numpy as np
import pytest
import torch

from oracle import cliploss_oracle as O

SHELL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shells", "distill_w1_b9_d24_ll0_gwg0.npz")


@pytest.fixture
def emu():
    from clipk import ops
    from tests.emu_backend import EmuBackend
    ops.set_backend_for_testing(EmuBackend())
    yield
    ops.set_backend_for_testing(None)


def test_oracle_distill_term_matches_reference_golden():
    """The golden holds the reference's (contrastive, distill) losses and the gradients of contrastive + 2 * distill."""
    z = np.load(SHELL)
    I, T, tI, tT = z["r0_image"], z["r0_text"], z["r0_t_image"], z["r0_t_text"]
    s, ts = float(z["scale"]), float(z["r0_t_scale"])
    loss, dI, dT, ds = O.distill_loss_single(I, T, s, tI, tT, ts, grad_output=2.0)
    con = O.clip_loss_single(I, T, s)
    assert abs(loss - float(z["r0_distill_loss"])) <= 1e-12 * abs(loss)
    assert abs(con.loss - float(z["r0_contrastive_loss"])) <= 1e-12
    assert np.abs(con.d_image + dI - z["r0_g_d_image"]).max() <= 1e-12
    assert np.abs(con.d_text + dT - z["r0_g_d_text"]).max() <= 1e-12
    assert abs(con.d_scale + ds - float(z["r0_g_d_scale"])) <= 1e-12


@pytest.mark.parametrize("n,d,dt,panel_bytes", [(37, 40, 24, 256 << 20), (300, 64, 96, 1), (513, 32, 32, 1)])
def test_panel_path_host_logic(n, d, dt, panel_bytes, emu):
    """bf16 features through the panel path (one panel, and several 256-row panels) against the oracle on the same
    bf16 values: loss and dlogit_scale to fp32 accuracy, feature gradients to the rounding of their bf16 output."""
    from clipk import distill
    g = torch.Generator().manual_seed(n)
    feats = [torch.nn.functional.normalize(torch.randn(n, w, generator=g), dim=-1).bfloat16() for w in (d, d, dt, dt)]
    I, T = feats[0].clone().requires_grad_(True), feats[1].clone().requires_grad_(True)
    s = torch.tensor(9.0, requires_grad=True)
    out = distill.fused_distill_term(I, T, s, feats[2], feats[3], 15.0, panel_bytes=panel_bytes)
    (out * 1.7).backward()
    ref = O.distill_loss_single(*(f.float().numpy() for f in feats[:2]), 9.0, *(f.float().numpy() for f in feats[2:]), 15.0,
                                grad_output=1.7)
    assert abs(float(out.detach()) - ref[0]) <= 1e-5 * abs(ref[0])
    assert I.grad.dtype == torch.bfloat16 and I.grad.shape == (n, d)
    assert np.linalg.norm(I.grad.float().numpy() - ref[1]) <= 8e-3 * np.linalg.norm(ref[1])
    assert np.linalg.norm(T.grad.float().numpy() - ref[2]) <= 8e-3 * np.linalg.norm(ref[2])
    assert abs(float(s.grad) - ref[3]) <= 1e-4 * max(abs(ref[3]), 1e-3)


def test_distill_class_switch(emu, monkeypatch):
    """DistillClipLoss takes the panel path (the default; CLIPK_FUSED_DISTILL=0 switches it off) only for calls it covers;
    both paths return the reference's tuple / dictionary."""
    from clipk import DistillClipLoss, distill
    g = torch.Generator().manual_seed(1)
    f = [torch.nn.functional.normalize(torch.randn(20, 16, generator=g), dim=-1) for _ in range(4)]
    s, ts = torch.tensor(8.0), torch.tensor(12.0)
    base = DistillClipLoss()(f[0], f[1], s, f[2], f[3], ts)
    assert not distill.applicable(f[0], f[1], f[2], f[3], 1)          # fp32 outside autocast: stays on the formula
    b16 = [x.bfloat16() for x in f]
    assert distill.applicable(*b16, 1) and not distill.applicable(*b16, 2)
    monkeypatch.setenv("CLIPK_FUSED_DISTILL", "1")
    fused = DistillClipLoss()(b16[0], b16[1], s, b16[2], b16[3], ts, output_dict=True)
    assert set(fused) == {"contrastive_loss", "distill_loss"}
    assert abs(float(fused["contrastive_loss"]) - float(base[0])) <= 2e-2 * float(base[0])     # bf16 vs fp32 features
    assert abs(float(fused["distill_loss"]) - float(base[1])) <= 2e-2 * float(base[1])
