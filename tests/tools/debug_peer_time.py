"""Manual experiment (torchrun, >= 2 GPUs): backward with NCCL reduce-scatter vs fused peer reduce, and the peer barrier."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch, torch.distributed as dist
from clipk import ops
from oracle import cliploss_oracle as O
rank = int(os.environ["RANK"]); W = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N, d = 32768, 512
b = N // W
x, t = O.synthetic_features(b, d, seed=1234, rank=rank)
I = torch.from_numpy(x).cuda().bfloat16(); Tl = torch.from_numpy(t).cuda().bfloat16()
T = ops._all_gather_rows(Tl, W)
be = ops._backend()
X, Y = be.prepare(I), be.prepare(T)
sc = torch.tensor([1 / 0.07], device=dev)
off = rank * b
parts = torch.empty(1, 3, N, device=dev)
rs, pos, _ = be.fwd_both(X, Y, sc, off, col_out=parts[0])
gp = ops._all_gather_rows(parts, W)
lr_, lc, sums = be.finalize(rs, pos, gp, off)
Xg, Yg = be.prepare_grad(X), be.prepare_grad(Y)
gs = torch.tensor([1.0 / (2 * b)], device=dev)
os.environ["CLIPK_PEER"] = "1"
peer = ops._peer_state(b, d, rank, W, None, dev)
def nccl_path():
    dX, dY = be.bwd(X, Y, Xg, Yg, sc, off, lr_, lc, 1.0, 1.0, gs, True, True)
    return ops._reduce_scatter_rows(dY, W)
def peer_path():
    be.peer_barrier(peer)
    dX = be.bwd_peer(X, Y, Xg, Yg, sc, off, lr_, lc, 1.0, 1.0, gs, peer)
    be.peer_barrier(peer)
    return be.reduce_slots(peer, torch.float32)
def a2a_path():
    dX, dY = be.bwd(X, Y, Xg, Yg, sc, off, lr_, lc, 1.0, 1.0, gs, True, True)
    return ops._reduce_scatter_rows_a2a(be, dY, W, torch.float32)
def tm(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
a = nccl_path(); p = peer_path(); torch.cuda.synchronize()
err = float((a - p).norm() / a.norm())
a2 = a2a_path(); torch.cuda.synchronize()
err2 = float((a - a2).norm() / a.norm())
t_n = tm(nccl_path); t_p = tm(peer_path); t_b = tm(lambda: be.peer_barrier(peer), 50); t_a = tm(a2a_path)
t_k = tm(lambda: be.bwd(X, Y, Xg, Yg, sc, off, lr_, lc, 1.0, 1.0, gs, True, True))
if rank == 0:
    print(f"W={W} b={b}: bwd only {t_k:.3f} ms | bwd + NCCL reduce_scatter {t_n:.3f} ms | all_to_all + sum_slots {t_a:.3f} ms (rel diff {err2:.1e}) | fused peer reduce {t_p:.3f} ms | peer barrier {t_b*1e3:.1f} us | rel diff {err:.2e}", file=sys.stderr)
dist.destroy_process_group()
