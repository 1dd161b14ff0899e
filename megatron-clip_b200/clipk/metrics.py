"""Evaluation-side contractions of the reference on the clipk kernels.

The reference evaluates with the same `logit_scale * I @ T^T` contraction as the loss, but ranks instead of reducing:

    get_clip_metrics   training/train.py:631-648   N x N logits on the CPU, argsort of every row, position of the diagonal
    accuracy / run     training/zero_shot.py:36-39, 54-57   100 * I @ classifier, topk, compare with the target
    validation loss    training/train.py:569-577   the loss formula again, per batch

A rank needs no sort: the position of the target in a descending sort of its row is the number of entries that beat it.
`target_ranks` forms the logits panel by panel on the tensor cores (clipk_gemm16: a bounded fp32 panel in HBM, never the
whole N x N matrix, never on the host) and counts per row (clipk_rank_count, one pass over the panel).  Ties are resolved
as a STABLE descending sort would (equal entries with a smaller column index come first); the reference's unstable
argsort / topk leave that order unspecified.

No CPU path: like the loss, these raise without libclipk.so and a compute-capability-10.x device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .ops import fused_clip_loss

PANEL_BYTES = 256 << 20          # fp32 logits held at a time


def target_ranks(queries: torch.Tensor, keys: torch.Tensor, target: torch.Tensor | None = None, diag_offset: int = 0,
                 panel_bytes: int = PANEL_BYTES) -> torch.Tensor:
    """rank[i] = index of column t_i in a stable descending sort of (queries @ keys^T)[i, :], as int64 [rows].

    queries [rows, d], keys [cols, d]: same dtype (bf16, fp16 or fp32) and device.  t_i = target[i] (int64 tensor), or
    diag_offset + i when target is None.  A target outside [0, cols) raises."""
    if queries.dim() != 2 or keys.dim() != 2 or queries.shape[1] != keys.shape[1]:
        raise ValueError("queries and keys must be [rows, dim] and [cols, dim]")
    if queries.dtype != keys.dtype:
        raise TypeError("queries and keys must have the same dtype")
    rows, cols = queries.shape[0], keys.shape[0]
    dev = queries.device
    if rows == 0:
        return torch.empty(0, dtype=torch.long, device=dev)
    if target is not None:
        target = target.to(device=dev, dtype=torch.long).contiguous()
        if target.shape != (rows,):
            raise ValueError("target must hold one column index per query row")
    be = ops._backend()
    cols4 = (cols + 3) // 4 * 4
    keys = keys.detach()
    if cols4 != cols:             # the GEMM writes whole groups of 4 columns: give it zero rows to read for the last one
        keys = torch.nn.functional.pad(keys, (0, 0, 0, cols4 - cols))
    Q, K = be.dense_operand(queries.detach()), be.dense_operand(keys)
    per = max(256, panel_bytes // (4 * cols4) // 256 * 256)
    per = min(per, rows)
    panel = torch.empty(per, cols4, dtype=torch.float32, device=dev)
    greater = torch.empty(rows, dtype=torch.int32, device=dev)
    ties = torch.empty(rows, dtype=torch.int32, device=dev)
    for r0 in range(0, rows, per):
        n = min(per, rows - r0)
        be.logits_panel(Q, K, r0, n, panel)
        be.rank_count(panel, n, cols, target, diag_offset, r0, greater, ties)
    if bool((greater < 0).any()):
        raise IndexError("clipk.target_ranks: a target column lies outside [0, cols)")
    return greater.long() + ties.long()


def logits_panels(image_features: torch.Tensor, text_features: torch.Tensor, logit_scale=1.0, panel_bytes: int = PANEL_BYTES):
    """Tile view of ClipLoss.get_logits (reference loss.py:104-121) for callers that need the logits themselves but not
    all at once: yields (row0, panel) with panel = logit_scale * image_features[row0:row0 + n] @ text_features.T as an
    fp32 [n, cols] tensor made on the tcgen05 mainloop (clipk_gemm16).  The SAME buffer is reused for every panel
    (<= panel_bytes): consume or copy it before advancing.  No autograd (the differentiable consumers - the loss and the
    distillation term - never need the logits in memory)."""
    if image_features.dim() != 2 or text_features.dim() != 2 or image_features.shape[1] != text_features.shape[1]:
        raise ValueError("image_features and text_features must be [rows, dim] and [cols, dim]")
    if image_features.dtype != text_features.dtype:
        raise TypeError("image_features and text_features must have the same dtype")
    rows, cols = image_features.shape[0], text_features.shape[0]
    if rows == 0 or cols == 0:
        return
    be = ops._backend()
    dev = image_features.device
    cols4 = (cols + 3) // 4 * 4
    keys = text_features.detach()
    if cols4 != cols:
        keys = torch.nn.functional.pad(keys, (0, 0, 0, cols4 - cols))
    Q, K = be.dense_operand(image_features.detach()), be.dense_operand(keys)
    # fp32 / fp16 operands are held as scaled fp16 planes: the product carries both power-of-two scales
    mul = float(logit_scale) if not isinstance(logit_scale, torch.Tensor) else logit_scale.detach().float().to(dev)
    if getattr(Q, "keep", None) is not None:
        mul = mul * Q.keep.inv_scale * K.keep.inv_scale
    per = min(max(256, panel_bytes // (4 * cols4) // 256 * 256), rows)
    panel = torch.empty(per, cols4, dtype=torch.float32, device=dev)
    for r0 in range(0, rows, per):
        n = min(per, rows - r0)
        be.logits_panel(Q, K, r0, n, panel)
        view = panel[:n, :cols]
        view.mul_(mul)
        yield r0, view


def _signed(features, logit_scale):
    """A positive logit_scale does not change any order; a negative one reverses it, zero makes every entry tie."""
    s = float(logit_scale)
    if s != s:
        raise ValueError("logit_scale is NaN")
    return s, (features if s >= 0 else -features)


def get_clip_metrics(image_features, text_features, logit_scale):
    """Same dictionary as training/train.py:631-648: {image_to_text, text_to_image} x {mean_rank, median_rank, R@1, R@5,
    R@10}, for the N validation pairs (pair i = row i of both matrices)."""
    metrics = {}
    s, image_signed = _signed(image_features, logit_scale)
    n = image_features.shape[0]
    for name, (q, k) in {"image_to_text": (image_signed, text_features),
                         "text_to_image": (text_features, image_signed)}.items():
        if s == 0.0:
            preds = np.arange(n)                    # all logits equal: the stable order is the column order
        else:
            preds = target_ranks(q, k).cpu().numpy()
        metrics[f"{name}_mean_rank"] = preds.mean() + 1
        metrics[f"{name}_median_rank"] = np.floor(np.median(preds)) + 1
        for top in (1, 5, 10):
            metrics[f"{name}_R@{top}"] = np.mean(preds < top)
    return metrics


def zero_shot_accuracy(image_features, classifier, target, topk=(1,)):
    """[number of rows whose target class is among the k best of image_features @ classifier, for k in topk], as floats:
    what accuracy(100. * image_features @ classifier, target, topk) returns (training/zero_shot.py:36-39, 54-57).
    classifier is [dim, classes] as zero_shot_classifier builds it (zero_shot.py:31)."""
    if classifier.dim() != 2 or classifier.shape[0] != image_features.shape[1]:
        raise ValueError("classifier must be [dim, classes]")
    keys = classifier.t().contiguous().to(image_features.dtype)
    ranks = target_ranks(image_features, keys, target=target)
    return [float((ranks < k).sum().item()) for k in topk]


def clip_val_loss(image_features, text_features, logit_scale):
    """The per-batch validation loss of training/train.py:569-577 (single process, both directions), on the fused
    forward: no logits, no graph."""
    with torch.no_grad():
        return fused_clip_loss(image_features, text_features, logit_scale)
