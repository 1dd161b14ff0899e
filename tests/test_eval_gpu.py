"""Evaluation-side contractions on the real kernels (clipk_gemm16 panels + clipk_rank_count) against the oracle.

Ranks are integers and compared exactly wherever the oracle says the order is decided: a logit within `eps` of the
target's may fall on either side when the accumulation order differs (fp32 sums of exact bf16 products here, fp64 in the
oracle), so the kernel's rank must lie inside oracle.target_rank_bounds, and the bounds must coincide for almost all rows."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import eval_oracle as E

pytestmark = pytest.mark.gpu

EVAL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval")


def _check_ranks(q, k, target, ranks, min_decided=0.97):
    logits = q.double().cpu().numpy() @ k.double().cpu().numpy().T
    lo, hi = E.target_rank_bounds(logits, target, eps=2e-6 * np.abs(logits).max())
    r = ranks.cpu().numpy()
    assert ranks.dtype == torch.long and r.shape == lo.shape
    assert np.all((lo <= r) & (r <= hi)), np.nonzero((r < lo) | (r > hi))[0][:10]
    assert np.mean(lo == hi) >= min_decided


@pytest.mark.parametrize("rows,cols,d,dtype", [(1000, 1000, 128, torch.bfloat16), (2500, 2500, 512, torch.bfloat16),
                                              (700, 1300, 96, torch.bfloat16), (900, 900, 64, torch.float32),
                                              (600, 2049, 256, torch.float32), (512, 640, 128, torch.float16)])
def test_target_ranks_against_oracle(rows, cols, d, dtype):
    import clipk
    g = torch.Generator().manual_seed(rows + cols + d)
    k = torch.nn.functional.normalize(torch.randn(cols, d, generator=g), dim=-1)
    z = torch.nn.functional.normalize(torch.randn(rows, d, generator=g), dim=-1)
    target = torch.randint(0, cols, (rows,), generator=g)
    q = torch.nn.functional.normalize(0.15 * k[target] + z, dim=-1)
    q, k = q.cuda().to(dtype), k.cuda().to(dtype)
    ranks = clipk.target_ranks(q, k, target=target.cuda())
    _check_ranks(q, k, target.numpy(), ranks)
    # several small panels give the same answer as one
    assert torch.equal(ranks, clipk.target_ranks(q, k, target=target.cuda(), panel_bytes=1 << 20))
    if rows == cols:                                  # default targets: the diagonal
        # (these targets sit in the middle of the crowd, so more rows have a logit within eps of theirs)
        _check_ranks(q, k, np.arange(rows), clipk.target_ranks(q, k), min_decided=0.8)


def test_ties_follow_a_stable_sort():
    import clipk
    g = torch.Generator().manual_seed(3)
    k = torch.randn(300, 64, generator=g)
    k[200] = k[17]
    k[250] = k[17]
    q = torch.randn(128, 64, generator=g)
    q, k = q.cuda().bfloat16(), k.cuda().bfloat16()
    r = [clipk.target_ranks(q, k, target=torch.full((128,), j, device="cuda")) for j in (17, 200, 250)]
    assert torch.equal(r[1], r[0] + 1) and torch.equal(r[2], r[0] + 2)
    with pytest.raises(IndexError):
        clipk.target_ranks(q, k, target=torch.full((128,), 300, device="cuda"))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(EVAL, "metrics_*.npz"))), ids=lambda p: os.path.basename(p)[:-4])
def test_get_clip_metrics_against_reference_goldens(path):
    """The dictionary of training/train.py:631-648 from the reference's own run (fp32 features on the split-precision
    path).  A rank may differ by one where two fp32 logits are closer than their rounding."""
    import clipk
    z = np.load(path)
    I, T = torch.from_numpy(z["image"]).float().cuda(), torch.from_numpy(z["text"]).float().cuda()
    got = clipk.get_clip_metrics(I, T, torch.tensor(float(z["scale"]), device="cuda"))
    assert sorted(got) == list(z["keys"])
    n = I.shape[0]
    for k, v in zip(z["keys"], z["values"]):
        slack = 2.0 / n if ("mean" in k or "R@" in k) else 1.0
        assert abs(got[k] - v) <= slack, (k, got[k], v)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(EVAL, "zeroshot_*.npz"))), ids=lambda p: os.path.basename(p)[:-4])
def test_zero_shot_accuracy_against_reference_goldens(path):
    import clipk
    z = np.load(path)
    got = clipk.zero_shot_accuracy(torch.from_numpy(z["image"]).cuda(), torch.from_numpy(z["classifier"]).cuda(),
                                   torch.from_numpy(z["target"]).cuda(), tuple(int(k) for k in z["topk"]))
    assert all(abs(a - b) <= 1 for a, b in zip(got, z["correct"])), (got, z["correct"])


def test_validation_loss_matches_training_loss_value():
    import clipk
    from oracle import cliploss_oracle as O
    x, t = O.synthetic_features(500, 256, seed=8)
    I, T = torch.from_numpy(x).cuda().bfloat16(), torch.from_numpy(t).cuda().bfloat16()
    loss = clipk.clip_val_loss(I, T, torch.tensor(1 / 0.07, device="cuda"))
    ref = O.clip_loss_single(I.float().cpu().numpy(), T.float().cpu().numpy(), 1 / 0.07)
    assert not loss.requires_grad and abs(loss.item() - ref.loss) <= 1e-5 * ref.loss
