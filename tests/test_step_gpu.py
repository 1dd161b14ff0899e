"""The fused step (clipk_step_forward / clipk_step_backward) on one B200: what changed in round 2.

* logit_scale = 100 keeps the single-sweep forward (positives' bound, fwd_bound in csrc/gemm_core.cuh) and falls back to
  the exact form when a positive pair is bad;
* fp32 features under bf16 autocast are cast in the operand pass; row-strided views are read in place;
* the normalising entry issues no standalone normalise / cast / norm kernels;
* several evaluations in flight share the scratch workspace.
Tolerances: BASELINE.json's 2e-3 for bf16 operands, against the fp64 oracle on the operand values.
"""
import numpy as np
import pytest
import torch

from oracle import cliploss_oracle as O
from tests.util import rel

pytestmark = pytest.mark.gpu


def _run(I, T, s, go=1.0, fn=None):
    from clipk import ClipLoss
    S = torch.tensor(s, device="cuda", requires_grad=True)
    loss = (fn or ClipLoss(cache_labels=True))(I, T, S)
    (loss * go).backward()
    torch.cuda.synchronize()
    return loss.item(), I.grad.float().cpu().numpy(), T.grad.float().cpu().numpy(), S.grad.item()


def _check(out, ref, tol, s):
    loss, dI, dT, ds = out
    assert abs(loss - ref.loss) <= tol * abs(ref.loss), (loss, ref.loss)
    assert rel(dI, ref.d_image) <= tol and rel(dT, ref.d_text) <= tol, (rel(dI, ref.d_image), rel(dT, ref.d_text))
    assert abs(ds - ref.d_scale) <= tol * max(abs(ref.d_scale), 1.0 / s), (ds, ref.d_scale)


@pytest.mark.parametrize("n,d", [(8192, 512), (1000, 256)])
def test_single_sweep_at_logit_scale_100(n, d):
    """Trained CLIP models sit at the clamp logit_scale = 100 (training/train.py:470-471): unit-norm bf16 features with
    positives at cos ~0.3 keep the ONE-sweep forward there, and the results stay inside 2e-3."""
    from clipk import ops
    x, t = O.synthetic_features(n, d, seed=3)
    I = torch.from_numpy(x).cuda().bfloat16().requires_grad_(True)
    T = torch.from_numpy(t).cuda().bfloat16().requires_grad_(True)
    out = _run(I, T, 100.0)
    assert ops.last_forward_was_single_sweep() is True
    ref = O.clip_loss_single(I.detach().float().cpu().numpy(), T.detach().float().cpu().numpy(), 100.0)
    _check(out, ref, 2e-3, 100.0)


def test_exact_form_when_a_positive_pair_is_bad():
    """One positive pair with cosine -0.9 at logit_scale = 100: the positives' bound fails, the device takes the exact
    two-sweep form by itself, and the results are still right."""
    from clipk import ops
    x, t = O.synthetic_features(1024, 256, seed=4)
    t[17] = -0.9 * x[17] + 0.1 * t[17]
    t[17] /= np.linalg.norm(t[17])
    I = torch.from_numpy(x).cuda().bfloat16().requires_grad_(True)
    T = torch.from_numpy(t).cuda().bfloat16().requires_grad_(True)
    out = _run(I, T, 100.0)
    assert ops.last_forward_was_single_sweep() is False
    ref = O.clip_loss_single(I.detach().float().cpu().numpy(), T.detach().float().cpu().numpy(), 100.0)
    _check(out, ref, 2e-3, 100.0)
    # the same inputs at the initial scale are bounded by their norms alone
    I.grad = T.grad = None
    _run(I, T, 1 / 0.07)
    assert ops.last_forward_was_single_sweep() is True


def test_autocast_fp32_features_and_strided_views():
    """fp32 features under torch.autocast(bf16) - what open_clip's towers hand to the loss - are cast to
    the bf16 operands inside the operand pass; a row-strided view is read in place."""
    x, t = O.synthetic_features(700, 512, seed=6)
    big = torch.zeros(700, 640, device="cuda")
    big[:, :512] = torch.from_numpy(x).cuda()
    I = big[:, :512].detach().requires_grad_(True)          # stride (640, 1)
    T = torch.from_numpy(t).cuda().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = _run(I, T, 1 / 0.07, go=2.0)
    assert I.grad.dtype == torch.float32 and I.grad.shape == (700, 512)
    xb = I.detach().bfloat16().float().cpu().numpy()
    tb = T.detach().bfloat16().float().cpu().numpy()
    _check(out, O.clip_loss_single(xb, tb, 1 / 0.07, grad_output=2.0), 2e-3, 1 / 0.07)


def test_normalising_entry_is_fused_into_the_operand_and_gradient_passes():
    """fused_normalize_clip_loss on raw bf16 embeddings: values against torch autograd through F.normalize in fp32 on the
    same inputs, and the kernel list holds no standalone normalise / cast / norm / amax kernel."""
    import torch.nn.functional as F
    from torch.profiler import ProfilerActivity, profile
    from clipk import fused_normalize_clip_loss
    g = torch.Generator().manual_seed(9)
    I = (torch.randn(1536, 512, generator=g) * 1.5).cuda().bfloat16().requires_grad_(True)
    T = (torch.randn(1536, 512, generator=g) * 0.4).cuda().bfloat16().requires_grad_(True)
    s = torch.tensor(1 / 0.07, device="cuda", requires_grad=True)
    fused_normalize_clip_loss(I, T, s).backward()           # warm-up (workspace, module load)
    I.grad = T.grad = s.grad = None
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        loss = fused_normalize_clip_loss(I, T, s)
        loss.backward()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    assert any("prep_kernel" in n for n in names) and any("finish_grad_kernel" in n for n in names), names
    for bad in ("normalize_fwd_kernel", "normalize_bwd_kernel", "cast_kernel", "amax_kernel", "norm2_max", "to_f16_kernel<"):
        assert not any(bad in n for n in names), (bad, names)
    I2, T2 = I.detach().float().requires_grad_(True), T.detach().float().requires_grad_(True)
    s2 = torch.tensor(1 / 0.07, device="cuda", requires_grad=True)
    a = s2 * F.normalize(I2, dim=-1) @ F.normalize(T2, dim=-1).T
    lab = torch.arange(1536, device="cuda")
    ref = (F.cross_entropy(a, lab) + F.cross_entropy(a.T, lab)) / 2
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 2e-3 * ref.item()
    # operands rounded to bf16 after the normalisation + bf16 gradients: same budget as the unfused entry's test
    assert rel(I.grad.float().cpu().numpy(), I2.grad.cpu().numpy()) <= 4e-3
    assert rel(T.grad.float().cpu().numpy(), T2.grad.cpu().numpy()) <= 4e-3
    assert abs(s.grad.item() - s2.grad.item()) <= 2e-3 * max(abs(s2.grad.item()), 0.07)


def test_evaluations_in_flight_share_the_scratch():
    """Two forwards before their backwards (gradient accumulation), backwards in the opposite order, and a forward
    without a backward in between: each evaluation keeps what its backward needs, the scratch is only scratch."""
    from clipk import ClipLoss
    mod = ClipLoss(cache_labels=True)
    data = []
    for seed in (1, 2):
        x, t = O.synthetic_features(640, 256, seed=seed)
        data.append((torch.from_numpy(x).cuda().bfloat16().requires_grad_(True),
                     torch.from_numpy(t).cuda().bfloat16().requires_grad_(True),
                     torch.tensor(1 / 0.07, device="cuda", requires_grad=True)))
    losses = [mod(I, T, S) for I, T, S in data]
    with torch.no_grad():
        mod(data[0][0], data[1][1], data[0][2])
    losses[1].backward()
    losses[0].backward()
    torch.cuda.synchronize()
    for (I, T, S), loss in zip(data, losses):
        ref = O.clip_loss_single(I.detach().float().cpu().numpy(), T.detach().float().cpu().numpy(), 1 / 0.07)
        _check((loss.item(), I.grad.float().cpu().numpy(), T.grad.float().cpu().numpy(), S.grad.item()), ref, 2e-3, 1 / 0.07)


def test_step_abi_rejects_bad_arguments():
    """clipk_step_forward validates before it launches: widths, aliasing rules, workspace size."""
    import ctypes
    from clipk import _lib
    lib = _lib.load()
    st = _lib.Step()
    assert lib.clipk_step_forward(ctypes.byref(st)) == -1
    x = torch.zeros(128, 96, device="cuda", dtype=torch.bfloat16)
    f = torch.zeros(4096, device="cuda")
    st.rows, st.cols, st.d = 128, 128, 96
    st.image = st.text = st.x_op = st.y_all = x.data_ptr()
    st.ld_image = st.ld_text = 96
    st.logit_scale = st.stats = st.lse_row = st.lse_col = st.scal = st.workspace = f.data_ptr()
    st.loss_div = 256.0
    assert lib.clipk_step_forward(ctypes.byref(st)) == -2            # d % 64 != 0
    st.d = 64
    st.ld_image = st.ld_text = 96
    st.workspace_bytes = 16
    assert lib.clipk_step_forward(ctypes.byref(st)) == -4            # workspace too small
    assert b"workspace" in lib.clipk_last_error()
