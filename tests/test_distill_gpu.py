"""Panel-based distillation term (clipk/distill.py) on the real kernels against the oracle."""
import numpy as np
import pytest
import torch

from oracle import cliploss_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d,dt,panel_bytes", [(700, 256, 512, 256 << 20), (1000, 128, 128, 1 << 20), (333, 100, 64, 1)])
def test_panel_distill_term_against_oracle(n, d, dt, panel_bytes):
    from clipk import distill
    g = torch.Generator().manual_seed(n)
    feats = [torch.nn.functional.normalize(torch.randn(n, w, generator=g), dim=-1).bfloat16() for w in (d, d, dt, dt)]
    I, T = feats[0].cuda().requires_grad_(True), feats[1].cuda().requires_grad_(True)
    s = torch.tensor(1 / 0.07, device="cuda", requires_grad=True)
    out = distill.fused_distill_term(I, T, s, feats[2].cuda(), feats[3].cuda(), torch.tensor(25.0, device="cuda"),
                                     panel_bytes=panel_bytes)
    (out * 1.7).backward()
    torch.cuda.synchronize()
    ref = O.distill_loss_single(*(f.float().numpy() for f in feats[:2]), 1 / 0.07,
                                *(f.float().numpy() for f in feats[2:]), 25.0, grad_output=1.7)
    assert abs(out.item() - ref[0]) <= 1e-5 * abs(ref[0])
    assert np.linalg.norm(I.grad.float().cpu().numpy() - ref[1]) <= 4e-3 * np.linalg.norm(ref[1])
    assert np.linalg.norm(T.grad.float().cpu().numpy() - ref[2]) <= 4e-3 * np.linalg.norm(ref[2])
    assert abs(s.grad.item() - ref[3]) <= 1e-4 * max(abs(ref[3]), 1e-3)


def test_distill_class_on_the_panel_path(monkeypatch):
    from clipk import DistillClipLoss
    monkeypatch.delenv("CLIPK_FUSED_DISTILL", raising=False)         # the panel path is the default for bf16 features
    g = torch.Generator().manual_seed(5)
    f = [torch.nn.functional.normalize(torch.randn(512, 256, generator=g), dim=-1).bfloat16().cuda() for _ in range(4)]
    s, ts = torch.tensor(1 / 0.07, device="cuda"), torch.tensor(30.0, device="cuda")
    con, dis = DistillClipLoss()(f[0], f[1], s, f[2], f[3], ts)
    monkeypatch.setenv("CLIPK_FUSED_DISTILL", "0")
    con0, dis0 = DistillClipLoss()(*(x.float() for x in f[:2]), s, *(x.float() for x in f[2:]), ts)
    assert abs(con.item() - con0.item()) <= 1e-4 * con0.item() and abs(dis.item() - dis0.item()) <= 1e-4 * dis0.item()
