"""ICT retrieval loss, the logits tile view and the materialising get_logits on a B200 (real kernels)."""
import numpy as np
import pytest
import torch

from tests.test_ict_cpu import reference_ict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d,dtype,tol", [(1024, 256, torch.bfloat16, 4e-3), (700, 128, torch.float32, 2e-5), (300, 100, torch.bfloat16, 4e-3)])
def test_ict_retrieval_loss_matches_the_reference_formula(n, d, dtype, tol):
    """pretrain_ict.py:73-114 in one process: loss, both gradients and the top-k accuracies against the reference's op
    sequence in float64 on the values the kernels saw."""
    from clipk import ict_retrieval_loss
    g = torch.Generator().manual_seed(n)
    q0 = 0.12 * torch.randn(n, d, generator=g)
    c0 = 0.6 * q0 + 0.12 * torch.randn(n, d, generator=g)
    q = q0.cuda().to(dtype).requires_grad_(True)
    c = c0.cuda().to(dtype).requires_grad_(True)
    loss, stats = ict_retrieval_loss(q, c, report_topk_accuracies=(1, 5))
    (loss * 2.0).backward()
    torch.cuda.synchronize()
    ref = reference_ict([q.detach().float().cpu()], [c.detach().float().cpu()], 0, topk=(1, 5))
    assert abs(loss.item() - ref[0]) <= max(tol, 1e-5) * abs(ref[0])
    assert q.grad.dtype == dtype and c.grad.dtype == dtype
    assert (q.grad.double().cpu() / 2 - ref[1]).norm() <= tol * ref[1].norm()
    assert (c.grad.double().cpu() / 2 - ref[2]).norm() <= tol * ref[2].norm()
    assert abs(float(stats["top1_acc"]) - ref[3][1]) <= 0.25 and abs(float(stats["top5_acc"]) - ref[3][5]) <= 0.25


def test_logits_panels_and_get_logits():
    """The tile view of the logits (tcgen05 panels) and the materialising ClipLoss.get_logits (reference loss.py:104-121)
    against torch.matmul in fp32."""
    from clipk import ClipLoss, logits_panels
    g = torch.Generator().manual_seed(4)
    I = torch.nn.functional.normalize(torch.randn(1500, 320, generator=g), dim=-1).cuda()
    T = torch.nn.functional.normalize(torch.randn(1100, 320, generator=g), dim=-1).cuda()
    want = 20.0 * I @ T.T
    for dtype, tol in ((torch.float32, 2e-5), (torch.bfloat16, 1e-2)):
        Ix, Tx = I.to(dtype), T.to(dtype)
        ref = 20.0 * Ix.float() @ Tx.float().T
        got = torch.empty_like(want)
        for r0, panel in logits_panels(Ix, Tx, 20.0, panel_bytes=2 << 20):
            got[r0:r0 + panel.shape[0]] = panel
        assert (got - ref).abs().max() <= 2e-5 * ref.abs().max(), (dtype, float((got - ref).abs().max()))   # fp32 accumulation of the same operand values
        assert (got - want).abs().max() <= tol * want.abs().max(), (dtype, float((got - want).abs().max()))
    # get_logits: same square problem as the reference's W = 1 branch, differentiable
    Iq = I[:1100].clone().requires_grad_(True)
    s = torch.tensor(20.0, device="cuda", requires_grad=True)
    per_image, per_text = ClipLoss().get_logits(Iq, T, s)
    assert torch.allclose(per_image, 20.0 * I[:1100] @ T.T, rtol=1e-4, atol=1e-4)
    assert torch.allclose(per_text, per_image.T, rtol=1e-4, atol=1e-4)      # two separate GEMMs in the reference's W = 1 branch
    per_image.sum().backward()
    assert Iq.grad is not None and s.grad is not None
