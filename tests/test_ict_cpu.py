"""ICT retrieval loss (pretrain_ict.py:73-114) on the emulated kernel entries: host logic against the reference's formula
written out in torch (all-gather whose backward keeps the own chunk, full N x N log_softmax, NLL of the diagonal, times
the data-parallel world size), single process and two gloo ranks."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def reference_ict(queries, contexts, rank, score_scale=1.0, topk=(1, 5)):
    """what pretrain_ict.py's loss_func computes on rank `rank`, given every rank's embeddings (lists of [b, d])"""
    W = len(queries)
    q = [x.clone().double().requires_grad_(r == rank) for r, x in enumerate(queries)]
    c = [x.clone().double().requires_grad_(r == rank) for r, x in enumerate(contexts)]
    scores = torch.cat(q) @ torch.cat(c).T * score_scale
    lsm = F.log_softmax(scores, dim=1)
    n = scores.shape[0]
    loss = F.nll_loss(lsm, torch.arange(n)) * W
    loss.backward()
    order = torch.argsort(scores, dim=1, descending=True, stable=True)
    pos = (order == torch.arange(n)[:, None]).nonzero()[:, 1]
    accs = {k: float((pos < k).float().mean()) * 100 for k in topk}
    return float(loss.detach()), q[rank].grad, c[rank].grad, accs


@pytest.fixture
def emu():
    from clipk import ops
    from tests.emu_backend import EmuBackend
    ops.set_backend_for_testing(EmuBackend())
    yield
    ops.set_backend_for_testing(None)


@pytest.mark.parametrize("d,scaling", [(64, False), (40, True)])
def test_ict_single_process(emu, d, scaling):
    from clipk import ict_retrieval_loss
    g = torch.Generator().manual_seed(2)
    q = (0.3 * torch.randn(30, d, generator=g)).requires_grad_(True)
    c = (q.detach() * 0.5 + 0.3 * torch.randn(30, d, generator=g)).requires_grad_(True)
    loss, stats = ict_retrieval_loss(q, c, retriever_score_scaling=scaling, hidden_size=768, report_topk_accuracies=(1, 5))
    (loss * 3.0).backward()
    ref = reference_ict([q.detach()], [c.detach()], 0, 768 ** -0.5 if scaling else 1.0)
    assert abs(float(loss) - ref[0]) <= 1e-5 * abs(ref[0]) and abs(float(stats["loss"]) - ref[0]) <= 1e-5 * abs(ref[0])
    assert (q.grad / 3.0 - ref[1]).norm() <= 1e-5 * ref[1].norm() and (c.grad / 3.0 - ref[2]).norm() <= 1e-5 * ref[2].norm()
    assert abs(float(stats["top1_acc"]) - ref[3][1]) < 1e-4 and abs(float(stats["top5_acc"]) - ref[3][5]) < 1e-4


def _worker(rank, world, tmp):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "megatron-clip_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from clipk import ict_retrieval_loss, ops
    from tests.emu_backend import EmuBackend
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"file://{tmp}/store", rank=rank, world_size=world)
    ops.set_backend_for_testing(EmuBackend())
    g = torch.Generator().manual_seed(50 + rank)
    q = (0.2 * torch.randn(9, 48, generator=g)).requires_grad_(True)          # scores of order 1: a loss of order 1
    c = (q.detach() * 0.7 + 0.2 * torch.randn(9, 48, generator=g)).requires_grad_(True)
    loss, stats = ict_retrieval_loss(q, c, report_topk_accuracies=(1, 3))
    loss.backward()
    np.savez(f"{tmp}/ict{rank}.npz", loss=loss.detach().numpy(), avg=stats["loss"].numpy(), top1=stats["top1_acc"].numpy(),
             top3=stats["top3_acc"].numpy(), q=q.detach().numpy(), c=c.detach().numpy(), dq=q.grad.numpy(), dc=c.grad.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_ict_two_ranks():
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(2, tmp), nprocs=2, join=True)
        outs = [dict(np.load(f"{tmp}/ict{r}.npz")) for r in range(2)]
    qs, cs = [torch.from_numpy(o["q"]) for o in outs], [torch.from_numpy(o["c"]) for o in outs]
    for r in range(2):
        ref = reference_ict(qs, cs, r, topk=(1, 3))
        o = outs[r]
        assert abs(float(o["loss"]) - ref[0]) <= 1e-5 * abs(ref[0])
        assert abs(float(o["avg"]) - ref[0] / 2) <= 1e-5 * abs(ref[0])       # the logged loss is the un-multiplied mean
        assert (torch.from_numpy(o["dq"]) - ref[1]).norm() <= 1e-5 * ref[1].norm()
        assert (torch.from_numpy(o["dc"]) - ref[2]).norm() <= 1e-5 * ref[2].norm()
        assert abs(float(o["top1"]) - ref[3][1]) < 1e-4 and abs(float(o["top3"]) - ref[3][3]) < 1e-4


def test_logits_panels_tile_view(emu):
    from clipk import logits_panels
    g = torch.Generator().manual_seed(8)
    I, T = torch.randn(50, 32, generator=g), torch.randn(37, 32, generator=g)
    got = torch.zeros(50, 37)
    seen = 0
    for r0, panel in logits_panels(I, T, 3.5, panel_bytes=1):        # smallest panels: several of them
        got[r0:r0 + panel.shape[0]] = panel
        seen += panel.shape[0]
    assert seen == 50 and torch.allclose(got, 3.5 * I @ T.T, rtol=1e-5, atol=1e-5)
