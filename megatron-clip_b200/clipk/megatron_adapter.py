"""`loss_func` for the Megatron entry point of the reference (pretrain_CLIP.py:115-136) on top of the fused loss.

The reference's `loss_func(text_output, image_output)` is an inlined single-process copy of the contrastive loss: fp32
features, no logit scale, labels = arange, the mean of the two cross-entropies, plus the text->image top-1 accuracy,
and it returns `(total_loss, {"loss": ..., "accuracy": ...})` with both values averaged over the data-parallel group
(`megatron/utils.py:96-105`).  It materialises two b x b logit matrices; this adapter keeps the same signature and
return convention and runs the fused kernels instead:

    from clipk.megatron_adapter import make_loss_func
    loss_func = make_loss_func()                                   # drop-in for pretrain_CLIP.py:115-136
    loss_func = make_loss_func(data_parallel=True, group=mpu.get_data_parallel_group())   # contrast against the DP-global batch

With data_parallel=False (the reference's behaviour) every rank contrasts its own micro-batch only.  With
data_parallel=True the features are gathered over `group` (ClipLoss(local_loss=True, gather_with_grad=True)), which is
what open_clip's training loop does (training/train.py:148) and what SURVEY.md section 0 says the Megatron script lacks.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from .loss import ClipLoss


def _top1_text_accuracy(image_features, text_features, logit_scale):
    """Fraction of texts whose best-matching image is their own (argmax of text_logits == label, pretrain_CLIP.py:127-129),
    from the exact column maxima of the fused forward instead of an argmax over materialised logits."""
    be = ops._backend()
    with torch.no_grad():
        X, Y = be.prepare(image_features.detach()), be.prepare(text_features.detach())
        scale = logit_scale.detach().to(device=image_features.device, dtype=torch.float32).reshape(1).contiguous()
        _, pos, col = be.fwd_both(X, Y, scale, 0, exact=True)
        # the positive comes from the row sweep and the maximum from the column sweep: the same element, rounded in a
        # different order (fp32 inputs: different plane pairs) - compare with a few ulps of slack
        return (pos >= col[0] - 1e-5 * (1.0 + col[0].abs())).float().mean()


def _megatron_data_parallel_group():
    """Megatron's data-parallel group when Megatron is importable and initialised (megatron/utils.py:96-105 averages over
    exactly this group), else None."""
    for modname in ("megatron.core.mpu", "megatron.core.parallel_state", "megatron.mpu"):
        try:
            mod = __import__(modname, fromlist=["get_data_parallel_group"])
            return mod.get_data_parallel_group()
        except Exception:       # not installed, or model parallelism not initialised
            continue
    return None


def make_loss_func(logit_scale=1.0, data_parallel=False, group=None, with_accuracy=True, cast_to_float=True):
    """Returns loss_func(text_output, image_output[, logit_scale]) -> (loss, {"loss": avg, "accuracy": avg}).

    `group` is the data-parallel group the loss / accuracy are averaged over (and, with data_parallel=True, the features
    gathered over).  When it is None the adapter asks Megatron for `mpu.get_data_parallel_group()` - what the reference
    averages over - and only falls back to WORLD when Megatron is not there (open_clip-style pure data parallelism).
    Ranks outside the group (other pipeline stages) simply never call loss_func: no collective here spans WORLD unless
    WORLD is the data-parallel group."""
    state = {"mod": None, "group": group, "resolved": group is not None}

    def dp_group():
        if not state["resolved"]:
            state["group"] = _megatron_data_parallel_group()     # None = WORLD
            state["resolved"] = True
        return state["group"]

    def loss_func(text_output: torch.Tensor, image_output: torch.Tensor, scale=None):
        text_features, image_features = text_output.contiguous(), image_output.contiguous()
        if cast_to_float:                       # the reference casts both to fp32 first (pretrain_CLIP.py:122-123)
            text_features, image_features = text_features.float(), image_features.float()
        s = logit_scale if scale is None else scale
        if not isinstance(s, torch.Tensor):
            s = torch.tensor(float(s), dtype=torch.float32, device=image_features.device)
        distributed = dist.is_available() and dist.is_initialized()
        grp = dp_group() if distributed else None
        if state["mod"] is None:
            if data_parallel and distributed and dist.get_world_size(grp) > 1:
                state["mod"] = ClipLoss(local_loss=True, gather_with_grad=True,
                                        cache_labels=True).set_process_group(grp)
            else:
                state["mod"] = ClipLoss(cache_labels=True)
        total_loss = state["mod"](image_features, text_features, s)
        stats = [total_loss.detach().reshape(1)]
        if with_accuracy:
            stats.append(_top1_text_accuracy(image_features, text_features, s).reshape(1))
        averaged = torch.cat(stats)
        if distributed:
            dist.all_reduce(averaged, group=grp)
            averaged = averaged / dist.get_world_size(grp)
        out = {"loss": averaged[0]}
        if with_accuracy:
            out["accuracy"] = averaged[1]
        return total_loss, out

    return loss_func
