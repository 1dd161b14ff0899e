"""World size 2 and 4 on CPU (gloo): the host-side orchestration of the fused loss - all-gather of the text features,
exchange of the column statistics, mode coefficients, reduce-scatter of the partial text gradient - against golden
outputs of the unmodified reference in all four (local_loss, gather_with_grad) modes.

The CUDA kernels cannot run here, so the kernel entry points are replaced by tests/emu_backend.py (plain torch).
What is under test is everything in clipk/ops.py and clipk/loss.py around them.
"""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import golden_files, load_golden


def _worker(rank, world, path, tmp):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "megatron-clip_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from clipk import ClipLoss, ops
    from tests.emu_backend import EmuBackend
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"file://{tmp}/store", rank=rank, world_size=world)
    ops.set_backend_for_testing(EmuBackend())
    z, W, ranks = load_golden(path)
    g = ranks[rank]
    I = torch.from_numpy(g["image"].astype(np.float32)).requires_grad_(True)
    T = torch.from_numpy(g["text"].astype(np.float32)).requires_grad_(True)
    s = torch.tensor(float(z["scale"]), requires_grad=True)
    mod = ClipLoss(local_loss=bool(z["local_loss"]), gather_with_grad=bool(z["gather_with_grad"]), cache_labels=True,
                   rank=rank, world_size=world)
    loss = mod(I, T, s)
    (loss * float(z["grad_output"])).backward()
    n_logits = I.shape[0] if bool(z["local_loss"]) else I.shape[0] * world
    labels = mod.get_ground_truth(I.device, n_logits)
    np.savez(f"{tmp}/out{rank}.npz", loss=loss.detach().numpy(), d_image=I.grad.numpy(), d_text=T.grad.numpy(),
             d_scale=s.grad.numpy(), labels=labels.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _multi_rank_goldens():
    return [p for p in golden_files() if not os.path.basename(p).startswith("w1_")]


@pytest.mark.parametrize("path", _multi_rank_goldens(), ids=lambda p: os.path.basename(p)[:-4])
def test_world_matches_reference(path):
    z, W, ranks = load_golden(path)
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(W, path, tmp), nprocs=W, join=True)
        outs = [dict(np.load(f"{tmp}/out{r}.npz")) for r in range(W)]
    tol = 2e-5 if float(z["scale"]) < 50 else 3e-4      # the goldens are fp32 (or fp64) runs of the reference
    for r in range(W):
        g, o = ranks[r], outs[r]
        assert np.array_equal(g["labels"], o["labels"]) and o["labels"].dtype == np.int64
        assert abs(float(o["loss"]) - float(g["loss"])) <= tol * abs(float(g["loss"]))
        for k in ("d_image", "d_text"):
            ref = g[k].astype(np.float64)
            assert np.linalg.norm(o[k] - ref) <= tol * np.linalg.norm(ref), (k, r)
        ds_ref = float(g["d_scale"])
        assert abs(float(o["d_scale"]) - ds_ref) <= tol * max(abs(ds_ref), float(z["grad_output"]) / float(z["scale"]))
