"""Manual experiment: time the backward C entry alone (GRAD + pair kernels), optionally one of them."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
from clipk import ops
from oracle import cliploss_oracle as O
b = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = 512
x, t = O.synthetic_features(b, d, seed=1234)
I = torch.from_numpy(x).cuda().bfloat16(); T = torch.from_numpy(t).cuda().bfloat16()
be = ops._backend()
X, Y = be.prepare(I), be.prepare(T)
sc = torch.tensor([1 / 0.07], device="cuda")
rs, pos = be.fwd_stats(X, Y, sc, 0, True)
parts = torch.empty(1, 3, b, device="cuda"); be.fwd_stats(Y, X, sc, 0, False, out=parts[0])
lr, lcl, sums = be.finalize(rs, pos, parts, 0)
Xg, Yg = be.prepare_grad(X), be.prepare_grad(Y)
gs = torch.tensor([1.0 / (2 * b)], device="cuda")
def run(dx, dy):
    for _ in range(2): be.bwd(X, Y, Xg, Yg, sc, 0, lr, lcl, 1.0, 1.0, gs, dx, dy)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): be.bwd(X, Y, Xg, Yg, sc, 0, lr, lcl, 1.0, 1.0, gs, dx, dy)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
print(f"DBG={os.environ.get('CLIPK_DBG','0')} b={b}: bwd(dX+dY) {run(True, True):.3f} ms   bwd(dX only) {run(True, False):.3f} ms   bwd(dY only) {run(False, True):.3f} ms")
