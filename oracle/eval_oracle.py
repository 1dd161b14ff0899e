"""CPU oracle of the evaluation-side contractions (TEST INFRASTRUCTURE - never imported by the product).

numpy float64 restatement of
    get_clip_metrics   /root/reference/open_CLIP/src/training/train.py:631-648
    accuracy           /root/reference/open_CLIP/src/training/zero_shot.py:36-39   (on 100 * I @ classifier, :54-57)
Parity pinned: tests/golden/eval/*.npz hold outputs of those two reference functions themselves (extracted from the
reference source and executed by tests/golden/make_golden_eval.py); tests/test_oracle_cpu.py checks this file against them.

The reference sorts (argsort / topk) and looks the target up; ties are broken however the sort happens to.  The
restatement uses a STABLE descending order (ties: smaller column index first), which is also what the CUDA path counts.
"""
import numpy as np


def target_ranks(logits, target):
    """Index of column target[i] in a stable descending sort of row i."""
    logits = np.asarray(logits)
    order = np.argsort(-logits, axis=1, kind="stable")           # train.py:640 (argsort, descending)
    return np.where(order == np.asarray(target).reshape(-1, 1))[1]   # train.py:641


def target_rank_bounds(logits, target, eps):
    """(lo, hi): the rank of the target if every logit within eps of it were decided against / in favour of it.  A
    kernel that accumulates in another order must land in [lo, hi]; lo == hi wherever no logit is that close."""
    logits = np.asarray(logits, dtype=np.float64)
    thr = logits[np.arange(logits.shape[0]), np.asarray(target)][:, None]
    lo = (logits > thr + eps).sum(1)
    hi = (logits >= thr - eps).sum(1) - 1                         # minus the target itself
    return lo, hi


def clip_metrics(image_features, text_features, logit_scale):
    """train.py:631-648."""
    image = np.asarray(image_features, dtype=np.float64)
    text = np.asarray(text_features, dtype=np.float64)
    per_image = float(logit_scale) * image @ text.T               # :633
    out = {}
    gt = np.arange(text.shape[0])                                 # :637
    for name, logit in (("image_to_text", per_image), ("text_to_image", per_image.T)):
        preds = target_ranks(logit, gt)
        out[f"{name}_mean_rank"] = preds.mean() + 1                # :643
        out[f"{name}_median_rank"] = np.floor(np.median(preds)) + 1
        for k in (1, 5, 10):
            out[f"{name}_R@{k}"] = np.mean(preds < k)
    return out


def topk_correct(logits, target, topk=(1,)):
    """zero_shot.py:36-39: how many rows have their target among the k largest logits."""
    ranks = target_ranks(logits, target)
    return [float((ranks < k).sum()) for k in topk]
