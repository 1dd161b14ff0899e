"""Multi-GPU parity (NCCL, real kernels): every (local_loss, gather_with_grad) mode against the reference goldens and
against the oracle on bf16 inputs.  Needs >= 2 GPUs on one box; skipped otherwise (run with `gpurun --gpus 2`)."""
import os
import tempfile

import numpy as np
import pytest
import torch

from tests.util import golden_files, load_golden, rel

pytestmark = pytest.mark.gpu


def _worker(rank, world, tmp, case):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "megatron-clip_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from clipk import ClipLoss
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"file://{tmp}/store", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    kind = case["kind"]
    if kind == "golden":
        z, W, ranks = load_golden(case["path"])
        g = ranks[rank]
        x, t = g["image"].astype(np.float32), g["text"].astype(np.float32)
        s, go, ll, gwg, dtype = float(z["scale"]), float(z["grad_output"]), bool(z["local_loss"]), bool(z["gather_with_grad"]), torch.float32
    else:
        from oracle import cliploss_oracle as O
        x, t = O.synthetic_features(case["b"], case["d"], seed=77, rank=rank)
        s, go, ll, gwg, dtype = 1 / 0.07, 1.0, case["ll"], case["gwg"], torch.bfloat16
    I = torch.from_numpy(x).cuda().to(dtype).requires_grad_(True)
    T = torch.from_numpy(t).cuda().to(dtype).requires_grad_(True)
    S = torch.tensor(s, device="cuda", requires_grad=True)
    mod = ClipLoss(local_loss=ll, gather_with_grad=gwg, cache_labels=True, rank=rank, world_size=world)
    loss = mod(I, T, S)
    (loss * go).backward()
    torch.cuda.synchronize()
    np.savez(f"{tmp}/out{rank}.npz", loss=loss.item(), d_image=I.grad.float().cpu().numpy(),
             d_text=T.grad.float().cpu().numpy(), d_scale=S.grad.item(), image=I.detach().float().cpu().numpy(),
             text=T.detach().float().cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def _run(world, case):
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(world, tmp, case), nprocs=world, join=True)
        return [dict(np.load(f"{tmp}/out{r}.npz")) for r in range(world)]


def _need(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("path", [p for p in golden_files(world=2)], ids=lambda p: os.path.basename(p)[:-4])
def test_two_gpus_match_reference_goldens(path):
    _need(2)
    z, W, ranks = load_golden(path)
    outs = _run(2, {"kind": "golden", "path": path})
    tol = 1e-5 if float(z["scale"]) < 50 else 2e-4
    for r in range(2):
        g, o = ranks[r], outs[r]
        assert abs(float(o["loss"]) - float(g["loss"])) <= tol * abs(float(g["loss"]))
        assert rel(o["d_image"], g["d_image"]) <= tol and rel(o["d_text"], g["d_text"]) <= tol
        assert abs(float(o["d_scale"]) - float(g["d_scale"])) <= tol * max(abs(float(g["d_scale"])), float(z["grad_output"]) / float(z["scale"]))


@pytest.mark.parametrize("ll,gwg", [(True, True), (True, False), (False, True), (False, False)])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_bf16_modes_against_oracle(world, ll, gwg):
    _need(world)
    from oracle import cliploss_oracle as O
    outs = _run(world, {"kind": "oracle", "b": 640, "d": 256, "ll": ll, "gwg": gwg})
    ref = O.clip_loss_world([o["image"] for o in outs], [o["text"] for o in outs], 1 / 0.07, ll, gwg)
    for r in range(world):
        o, g = outs[r], ref[r]
        assert abs(float(o["loss"]) - g.loss) <= 2e-3 * abs(g.loss)
        assert rel(o["d_image"], g.d_image) <= 2e-3 and rel(o["d_text"], g.d_text) <= 2e-3
        assert abs(float(o["d_scale"]) - g.d_scale) <= 2e-3 * max(abs(g.d_scale), 0.07)


def _steps_worker(rank, world, tmp, case):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "megatron-clip_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    for k, v in case.get("env", {}).items():
        os.environ[k] = v
    from clipk import ClipLoss, ops
    from oracle import cliploss_oracle as O
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"file://{tmp}/store", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    mod = ClipLoss(local_loss=case.get("ll", True), gather_with_grad=True, cache_labels=True, rank=rank, world_size=world)
    out = {}
    for step in range(case["steps"]):
        x, t = O.synthetic_features(case["b"], case["d"], seed=100 + step, rank=rank)
        I = torch.from_numpy(x).cuda().bfloat16().requires_grad_(True)
        T = torch.from_numpy(t).cuda().bfloat16().requires_grad_(True)
        S = torch.tensor(case.get("s", 1 / 0.07), device="cuda", requires_grad=True)
        if step == 1:
            with torch.no_grad():           # a forward without a backward in between (validation): epochs stay in step
                mod(I, T, S)
        loss = mod(I, T, S)
        loss.backward()
        out[f"loss{step}"] = np.array(loss.item())
        out[f"d_scale{step}"] = np.array(S.grad.item())
        out[f"d_image{step}"] = I.grad.float().cpu().numpy()
        out[f"d_text{step}"] = T.grad.float().cpu().numpy()
        out[f"image{step}"] = I.detach().float().cpu().numpy()
        out[f"text{step}"] = T.detach().float().cpu().numpy()
    torch.cuda.synchronize()
    out["peer_used"] = np.array(len(ops._PEER_CONTEXTS))
    out["single"] = np.array(1 if ops.last_forward_was_single_sweep() else 0)
    np.savez(f"{tmp}/out{rank}.npz", **out)
    dist.barrier()
    dist.destroy_process_group()


def _check_steps(world, case, single=None):
    import torch.multiprocessing as mp
    from oracle import cliploss_oracle as O
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_steps_worker, args=(world, tmp, case), nprocs=world, join=True)
        outs = [dict(np.load(f"{tmp}/out{r}.npz")) for r in range(world)]
    assert all(int(o["peer_used"]) == 1 for o in outs), "the peer-memory path was not taken"
    if single is not None:
        assert all(int(o["single"]) == int(single) for o in outs)
    s = case.get("s", 1 / 0.07)
    for step in range(case["steps"]):
        ref = O.clip_loss_world([o[f"image{step}"] for o in outs], [o[f"text{step}"] for o in outs], s,
                                case.get("ll", True), True)
        for r in range(world):
            o, g = outs[r], ref[r]
            assert abs(float(o[f"loss{step}"]) - g.loss) <= 2e-3 * abs(g.loss), (step, r)
            assert rel(o[f"d_image{step}"], g.d_image) <= 2e-3, (step, r, rel(o[f"d_image{step}"], g.d_image))
            assert rel(o[f"d_text{step}"], g.d_text) <= 2e-3, (step, r, rel(o[f"d_text{step}"], g.d_text))
            assert abs(float(o[f"d_scale{step}"]) - g.d_scale) <= 2e-3 * max(abs(g.d_scale), 1 / s), (step, r)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_memory_step_over_consecutive_steps(world):
    """The fused step between ranks - text rows and statistics pulled from peer memory, column statistics pulled, dY tiles
    stored straight into their owners' slots, flag barriers - over several steps with different inputs and a forward
    without a backward in between: the double-buffered sources / slots and the epochs stay consistent."""
    _need(world)
    _check_steps(world, {"b": 384, "d": 128, "steps": 4}, single=True)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_memory_step_with_several_row_panels(world):
    """A local batch that needs several row panels of the softmax gradient (8 MB panel budget): the dY tiles of the later
    row panels are ADDED into the owners' slots over NVLink."""
    _need(world)
    _check_steps(world, {"b": 1536, "d": 256, "steps": 2, "env": {"CLIPK_PANEL_MB": "8"}})


@pytest.mark.parametrize("world", [2, 8])
def test_peer_memory_step_global_loss_at_scale_100(world):
    """local_loss=False (the loss all-reduced over the ranks) at the clamp logit_scale = 100: the positives' bound travels
    with the gathered statistics, every rank keeps the single sweep."""
    _need(world)
    _check_steps(world, {"b": 640, "d": 256, "steps": 2, "ll": False, "s": 100.0}, single=True)


def _ddp_worker(rank, world, tmp):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "megatron-clip_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from clipk import ClipLoss, ops
    from tests.ddp_towers import Towers, global_batch
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"file://{tmp}/store", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    torch.manual_seed(0)
    model = Towers(96, 48, 128, torch.bfloat16).cuda()
    ddp = DDP(model, device_ids=[rank], static_graph=True)
    loss_mod = ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=rank, world_size=world)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    out = {}
    for it, b in enumerate((256, 256, 384)):        # the last one: cache_labels with a changed num_logits
        images, texts = global_batch(world * b, 96, 48, seed=20 + it)
        opt.zero_grad()
        i, t, s = ddp(images[rank * b:(rank + 1) * b].cuda(), texts[rank * b:(rank + 1) * b].cuda())
        loss = loss_mod(i, t, s)
        scaler.scale(loss).backward()               # training/train.py:58-62
        scaler.unscale_(opt)
        assert loss_mod.prev_num_logits == b and torch.equal(loss_mod.labels[i.device], torch.arange(b, device=i.device) + b * rank)
        out[f"loss{it}"] = np.array(loss.item())
        for k, p in model.named_parameters():
            out[f"{k}@{it}"] = p.grad.float().cpu().numpy()
        scaler.step(opt)
        scaler.update()
    out["peer_used"] = np.array(len(ops._PEER_CONTEXTS))
    np.savez(f"{tmp}/ddp{rank}.npz", **out)
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_static_graph_towers_with_grad_scaler():
    """SURVEY 7 parity matrix on real kernels: DDP(static_graph=True) towers, torch.amp.GradScaler upstream gradient,
    cache_labels across a change of the batch size; two ranks over NCCL + peer memory.  Every rank must end with the
    gradient a single process gets on the global batch with the reference's formula in fp32."""
    _need(2)
    import torch.multiprocessing as mp
    from tests.ddp_towers import Towers, global_batch, reference_grads
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_ddp_worker, args=(2, tmp), nprocs=2, join=True)
        outs = [dict(np.load(f"{tmp}/ddp{r}.npz")) for r in range(2)]
    assert all(int(o["peer_used"]) == 2 for o in outs)          # two shapes (b = 256, 384), both on peer memory
    torch.manual_seed(0)
    model = Towers(96, 48, 128, torch.bfloat16).cuda()
    for it, b in enumerate((256, 256, 384)):
        images, texts = global_batch(2 * b, 96, 48, seed=20 + it)
        loss, grads = reference_grads(model, images.cuda(), texts.cuda())
        assert abs(0.5 * (float(outs[0][f"loss{it}"]) + float(outs[1][f"loss{it}"])) - loss) <= 2e-3 * abs(loss)
        for k, g in grads.items():
            for r in range(2):
                got = torch.from_numpy(outs[r][f"{k}@{it}"]).cuda()
                # bf16 features and bf16 feature gradients on both sides of the loss
                assert (got - g).norm() <= 1e-2 * g.norm().clamp_min(1e-4), (it, k, r, float((got - g).norm()), float(g.norm()))


def _shell_worker(rank, world, tmp, paths):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "megatron-clip_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from clipk import loss as L
    from tests.test_shells_cpu import load
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"file://{tmp}/store", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    for i, path in enumerate(paths):
        z, W, ranks = load(path)
        g = ranks[rank]
        I = torch.from_numpy(g["image"]).cuda().requires_grad_(True)          # fp64, as the goldens were made
        T = torch.from_numpy(g["text"]).cuda().requires_grad_(True)
        s = torch.tensor(float(z["scale"]), dtype=torch.float64, device="cuda", requires_grad=True)
        ll, gwg = bool(z["local_loss"]), bool(z["gather_with_grad"])
        wi, wt = torch.from_numpy(g["w_image"]).cuda(), torch.from_numpy(g["w_text"]).cuda()
        if str(z["what"]) == "gather":
            a, b = L.gather_features(I, T, ll, gwg, rank, world)
        else:
            a, b = L.ClipLoss(local_loss=ll, gather_with_grad=gwg, rank=rank, world_size=world).get_logits(I, T, s)
        f = (a * wi).sum() + (b * wt).sum()
        if f.requires_grad:
            f.backward()
        zero = lambda x: np.zeros(tuple(x.shape)) if x.grad is None else x.grad.cpu().numpy()
        np.savez(f"{tmp}/shell{i}_{rank}.npz", a=a.detach().cpu().numpy(), b=b.detach().cpu().numpy(), g_d_image=zero(I),
                 g_d_text=zero(T), g_d_scale=zero(s))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_features_and_get_logits_over_nccl():
    """The public gather_features (reference loss.py:20-64: values AND which gradient reaches which rank in the four
    modes) and the materialising ClipLoss.get_logits (loss.py:104-121) on two GPUs over NCCL, against outputs of the
    unmodified reference (tests/golden/shells, made under gloo in fp64)."""
    _need(2)
    import glob
    import torch.multiprocessing as mp
    from tests.test_shells_cpu import SHELLS, load
    paths = sorted(glob.glob(os.path.join(SHELLS, "gather_w2_*.npz")) + glob.glob(os.path.join(SHELLS, "logits_w2_*.npz")))
    assert len(paths) == 8
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_shell_worker, args=(2, tmp, paths), nprocs=2, join=True)
        for i, path in enumerate(paths):
            z, W, ranks = load(path)
            names = ("all_image", "all_text") if str(z["what"]) == "gather" else ("per_image", "per_text")
            for r in range(2):
                o, g = dict(np.load(f"{tmp}/shell{i}_{r}.npz")), ranks[r]
                for mine, ref in (("a", names[0]), ("b", names[1]), ("g_d_image", "g_d_image"), ("g_d_text", "g_d_text"),
                                  ("g_d_scale", "g_d_scale")):
                    if ref in g:
                        assert np.allclose(o[mine], g[ref], rtol=1e-9, atol=1e-11), (os.path.basename(path), r, ref)
