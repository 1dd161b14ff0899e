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


def _worker(rank, world, paths, tmp):
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
    for i, path in enumerate(paths):
        z, W, ranks = load_golden(path)
        g = ranks[rank]
        I = torch.from_numpy(g["image"].astype(np.float32)).requires_grad_(True)
        T = torch.from_numpy(g["text"].astype(np.float32)).requires_grad_(True)
        s = torch.tensor(float(z["scale"]), requires_grad=True)
        mod = ClipLoss(local_loss=bool(z["local_loss"]), gather_with_grad=bool(z["gather_with_grad"]),
                       cache_labels=True, rank=rank, world_size=world)
        loss = mod(I, T, s)
        (loss * float(z["grad_output"])).backward()
        n_logits = I.shape[0] if bool(z["local_loss"]) else I.shape[0] * world
        labels = mod.get_ground_truth(I.device, n_logits)
        np.savez(f"{tmp}/out{i}_{rank}.npz", loss=loss.detach().numpy(), d_image=I.grad.numpy(),
                 d_text=T.grad.numpy(), d_scale=s.grad.numpy(), labels=labels.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _multi_rank_goldens():
    return [p for p in golden_files() if not os.path.basename(p).startswith("w1_")]


@pytest.fixture(scope="module")
def world_outputs():
    """Every golden of one world size in ONE set of gloo processes (a spawn per case costs ~5 s of imports each)."""
    outs = {}
    for W in (2, 4):
        paths = golden_files(world=W)
        with tempfile.TemporaryDirectory() as tmp:
            mp.spawn(_worker, args=(W, paths, tmp), nprocs=W, join=True)
            for i, p in enumerate(paths):
                outs[p] = [dict(np.load(f"{tmp}/out{i}_{r}.npz")) for r in range(W)]
    return outs


@pytest.mark.parametrize("path", _multi_rank_goldens(), ids=lambda p: os.path.basename(p)[:-4])
def test_world_matches_reference(path, world_outputs):
    z, W, ranks = load_golden(path)
    outs = world_outputs[path]
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


def _subgroup_worker(rank, world, path, tmp):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "megatron-clip_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from clipk import ClipLoss, ops
    from clipk.megatron_adapter import make_loss_func
    from tests.emu_backend import EmuBackend
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"file://{tmp}/store", rank=rank, world_size=world)
    ops.set_backend_for_testing(EmuBackend())
    # two data-parallel groups of two ranks inside a world of four, interleaved like Megatron's DP groups at TP=2
    groups = [dist.new_group([0, 2]), dist.new_group([1, 3])]
    group, grank = groups[rank % 2], rank // 2
    z, W, ranks = load_golden(path)
    g = ranks[grank]
    I = torch.from_numpy(g["image"].astype(np.float32)).requires_grad_(True)
    T = torch.from_numpy(g["text"].astype(np.float32)).requires_grad_(True)
    s = torch.tensor(float(z["scale"]), requires_grad=True)
    mod = ClipLoss(local_loss=bool(z["local_loss"]), gather_with_grad=bool(z["gather_with_grad"]),
                   cache_labels=True).set_process_group(group)
    assert (mod.rank, mod.world_size) == (grank, 2)
    loss = mod(I, T, s)
    (loss * float(z["grad_output"])).backward()
    # the Megatron adapter over the same group: local + gather_with_grad, averaged over the group only
    I2, T2 = I.detach().clone().requires_grad_(True), T.detach().clone().requires_grad_(True)
    l2, avg = make_loss_func(logit_scale=float(z["scale"]), data_parallel=True, group=group)(T2, I2)
    np.savez(f"{tmp}/out{rank}.npz", loss=loss.detach().numpy(), d_image=I.grad.numpy(), d_text=T.grad.numpy(),
             d_scale=s.grad.numpy(), adapter_loss=l2.detach().numpy(), adapter_avg=avg["loss"].numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["w2_b12_d64_raw_float32_s100_ll1_gwg1.npz", "w2_b8_d32_unit_float64_s14_ll0_gwg0.npz",
                                  "w2_b33_d128_unit_float32_s14_ll1_gwg1.npz"])
def test_data_parallel_subgroups_of_a_larger_world(name):
    """ClipLoss.set_process_group: two independent 2-rank data-parallel groups inside a 4-rank job (the layout Megatron
    produces with tensor parallelism 2, megatron/core/parallel_state.py) each reproduce the reference's 2-rank golden;
    nothing leaks between the groups."""
    z, W, ranks = load_golden(name)
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_subgroup_worker, args=(4, os.path.join(os.path.dirname(__file__), "golden", name), tmp), nprocs=4,
                 join=True)
        outs = [dict(np.load(f"{tmp}/out{r}.npz")) for r in range(4)]
    tol = 2e-5 if float(z["scale"]) < 50 else 3e-4
    for r in range(4):
        g, o = ranks[r // 2], outs[r]
        assert abs(float(o["loss"]) - float(g["loss"])) <= tol * abs(float(g["loss"]))
        for k in ("d_image", "d_text"):
            ref = g[k].astype(np.float64)
            assert np.linalg.norm(o[k] - ref) <= tol * np.linalg.norm(ref), (k, r)
        ds_ref = float(g["d_scale"])
        assert abs(float(o["d_scale"]) - ds_ref) <= tol * max(abs(ds_ref), float(z["grad_output"]) / float(z["scale"]))
        if bool(z["local_loss"]) and bool(z["gather_with_grad"]):
            assert abs(float(o["adapter_loss"]) - float(g["loss"])) <= tol * abs(float(g["loss"]))
            mean = 0.5 * (float(ranks[0]["loss"]) + float(ranks[1]["loss"]))
            assert abs(float(o["adapter_avg"]) - mean) <= tol * abs(mean)


def _odd_worker(rank, world, tmp, cases):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "megatron-clip_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from clipk import ClipLoss, ops
    from oracle import cliploss_oracle as O
    from tests.emu_backend import EmuBackend
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"file://{tmp}/store", rank=rank, world_size=world)
    ops.set_backend_for_testing(EmuBackend())
    for i, (b, d, dtype, ll, gwg) in enumerate(cases):
        x, t = O.synthetic_features(b, d, seed=40 + i, rank=rank)
        I = torch.from_numpy(x).to(getattr(torch, dtype)).requires_grad_(True)
        T = torch.from_numpy(t).to(getattr(torch, dtype)).requires_grad_(True)
        s = torch.tensor(11.0, requires_grad=True)
        loss = ClipLoss(local_loss=ll, gather_with_grad=gwg, rank=rank, world_size=world)(I, T, s)
        loss.backward()
        assert I.grad.shape == (b, d) and I.grad.dtype == I.dtype and T.grad.dtype == T.dtype
        np.savez(f"{tmp}/odd{i}_{rank}.npz", loss=loss.detach().numpy(), d_image=I.grad.float().numpy(),
                 d_text=T.grad.float().numpy(), d_scale=s.grad.numpy(), image=I.detach().float().numpy(),
                 text=T.detach().float().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_odd_widths_and_fp16_across_ranks():
    """Widths that are not whole K blocks (host zero-padding, gradient columns dropped after the reduce-scatter) and fp16
    features, two ranks, against the float64 oracle on the values the ranks held."""
    from oracle import cliploss_oracle as O
    cases = [(10, 100, "float32", True, True), (9, 72, "float16", False, True), (8, 40, "float32", True, False),
             (6, 200, "bfloat16", False, False)]
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_odd_worker, args=(2, tmp, cases), nprocs=2, join=True)
        for i, (b, d, dtype, ll, gwg) in enumerate(cases):
            outs = [dict(np.load(f"{tmp}/odd{i}_{r}.npz")) for r in range(2)]
            ref = O.clip_loss_world([o["image"] for o in outs], [o["text"] for o in outs], 11.0, ll, gwg)
            tol = {"float32": 1e-5, "float16": 2e-3, "bfloat16": 8e-3}[dtype]       # rounding of the returned gradients
            for r in range(2):
                o, g = outs[r], ref[r]
                assert abs(float(o["loss"]) - g.loss) <= 1e-5 * abs(g.loss)
                for k, want in (("d_image", g.d_image), ("d_text", g.d_text)):
                    assert np.linalg.norm(o[k] - want) <= tol * np.linalg.norm(want), (i, r, k)
                assert abs(float(o["d_scale"]) - g.d_scale) <= 1e-5 * max(abs(g.d_scale), 1 / 11.0)


def test_fused_step_route_across_ranks():
    """bf16 features with d % 64 == 0 take the fused step on several ranks as well (here with the peer-memory collectives
    of the two entries emulated by gloo ones): coefficients of the modes, global-loss all-reduce, rank offsets."""
    from oracle import cliploss_oracle as O
    cases = [(12, 64, "bfloat16", True, True), (8, 128, "bfloat16", False, True), (8, 64, "bfloat16", False, False),
             (8, 64, "bfloat16", True, False)]          # the last one: dI / dT from different softmaxes (two-plane recompute)
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_odd_worker, args=(2, tmp, cases), nprocs=2, join=True)
        for i, (b, d, dtype, ll, gwg) in enumerate(cases):
            outs = [dict(np.load(f"{tmp}/odd{i}_{r}.npz")) for r in range(2)]
            ref = O.clip_loss_world([o["image"] for o in outs], [o["text"] for o in outs], 11.0, ll, gwg)
            for r in range(2):
                o, g = outs[r], ref[r]
                assert abs(float(o["loss"]) - g.loss) <= 1e-5 * abs(g.loss), (i, r)
                for k, want in (("d_image", g.d_image), ("d_text", g.d_text)):
                    assert np.linalg.norm(o[k] - want) <= 8e-3 * np.linalg.norm(want), (i, r, k)
                assert abs(float(o["d_scale"]) - g.d_scale) <= 1e-5 * max(abs(g.d_scale), 1 / 11.0)


def _pipeline_worker(rank, world, tmp):
    import sys
    import types
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "megatron-clip_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from clipk import ops
    from clipk.megatron_adapter import make_loss_func
    from oracle import cliploss_oracle as O
    from tests.emu_backend import EmuBackend
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"file://{tmp}/store", rank=rank, world_size=world)
    ops.set_backend_for_testing(EmuBackend())
    # 4 ranks = 2 pipeline stages x 2 data-parallel ranks; only the LAST stage (ranks 2, 3) computes the loss
    dp_groups = [dist.new_group([0, 1]), dist.new_group([2, 3])]
    my_dp = dp_groups[rank // 2]
    # what `from megatron.core import mpu` would give the adapter
    mpu = types.ModuleType("megatron.core.mpu")
    mpu.get_data_parallel_group = lambda: my_dp
    core = types.ModuleType("megatron.core")
    core.mpu = mpu
    meg = types.ModuleType("megatron")
    meg.core = core
    sys.modules.update({"megatron": meg, "megatron.core": core, "megatron.core.mpu": mpu})
    if rank >= 2:
        x, t = O.synthetic_features(10, 32, seed=5, rank=rank - 2)
        I, T = torch.from_numpy(x).requires_grad_(True), torch.from_numpy(t).requires_grad_(True)
        # group=None: the adapter must find Megatron's data-parallel group by itself; a WORLD collective would hang
        # here, because ranks 0 and 1 never call loss_func
        loss, avg = make_loss_func(logit_scale=7.0, data_parallel=True)(T, I)
        loss.backward()
        np.savez(f"{tmp}/pp{rank}.npz", loss=loss.detach().numpy(), avg=avg["loss"].numpy(), image=x, text=t,
                 d_image=I.grad.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_megatron_adapter_defaults_to_the_data_parallel_group():
    """pretrain_CLIP.py's loss_func averages over mpu.get_data_parallel_group() (megatron/utils.py:96-105) and, with
    pipeline parallelism, runs on the last stage only.  make_loss_func(group=None) must use that group - not WORLD - for
    the gather, the reduce-scatter and the logged average: ranks of the other stage never enter a collective."""
    from oracle import cliploss_oracle as O
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_pipeline_worker, args=(4, tmp), nprocs=4, join=True)
        outs = [dict(np.load(f"{tmp}/pp{r}.npz")) for r in (2, 3)]
    ref = O.clip_loss_world([o["image"] for o in outs], [o["text"] for o in outs], 7.0, True, True)
    mean = 0.5 * (ref[0].loss + ref[1].loss)
    for o, g in zip(outs, ref):
        assert abs(float(o["loss"]) - g.loss) <= 1e-5 * abs(g.loss)
        assert abs(float(o["avg"]) - mean) <= 1e-5 * abs(mean)
        assert np.linalg.norm(o["d_image"] - g.d_image) <= 1e-5 * np.linalg.norm(g.d_image)


def _ddp_worker(rank, world, tmp, feat_dtype):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "megatron-clip_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from torch.nn.parallel import DistributedDataParallel as DDP
    from clipk import ClipLoss, ops
    from tests.ddp_towers import Towers, global_batch
    from tests.emu_backend import EmuBackend
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"file://{tmp}/store", rank=rank, world_size=world)
    ops.set_backend_for_testing(EmuBackend())
    torch.manual_seed(0)
    model = Towers(24, 16, 64, getattr(torch, feat_dtype))
    ddp = DDP(model, static_graph=True)
    loss_mod = ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=rank, world_size=world)
    b = 6
    out = {}
    for it in range(3):                 # static_graph changes DDP's behaviour from the second iteration on
        images, texts = global_batch(world * b, 24, 16, seed=10 + it)
        model.zero_grad()
        i, t, s = ddp(images[rank * b:(rank + 1) * b], texts[rank * b:(rank + 1) * b])
        loss = loss_mod(i, t, s)
        (loss * 64.0).backward()        # an upstream factor, as AMP's GradScaler applies (training/train.py:58-62)
        out[f"loss{it}"] = loss.detach().numpy()
        for k, p in model.named_parameters():
            out[f"{k}@{it}"] = (p.grad / 64.0).numpy()
    np.savez(f"{tmp}/ddp{rank}.npz", **out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("feat_dtype,tol", [("float32", 2e-5), ("bfloat16", 2e-2)])
def test_ddp_static_graph_towers_get_the_single_process_gradient(feat_dtype, tol):
    """SURVEY 7 parity matrix: DDP(static_graph=True)-wrapped towers + the loss's own collectives, upstream gradient scale.
    fp32 features take the general route, bf16 features (width 64) the fused step; both with the kernels emulated."""
    from tests.ddp_towers import Towers, global_batch, reference_grads
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_ddp_worker, args=(2, tmp, feat_dtype), nprocs=2, join=True)
        outs = [dict(np.load(f"{tmp}/ddp{r}.npz")) for r in range(2)]
    torch.manual_seed(0)
    model = Towers(24, 16, 64, getattr(torch, feat_dtype))
    for it in range(3):
        images, texts = global_batch(12, 24, 16, seed=10 + it)
        loss, grads = reference_grads(model, images, texts)
        assert abs(0.5 * (float(outs[0][f"loss{it}"]) + float(outs[1][f"loss{it}"])) - loss) <= tol * abs(loss)
        for k, g in grads.items():
            for r in range(2):
                got = torch.from_numpy(outs[r][f"{k}@{it}"])
                assert (got - g).norm() <= tol * g.norm().clamp_min(1e-3), (it, k, r, float((got - g).norm()), float(g.norm()))
