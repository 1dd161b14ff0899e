"""Manual check: clipk_fwd_both against two clipk_fwd_stats calls and against torch, plus timing."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
from clipk import ops
from oracle import cliploss_oracle as O
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cols = int(sys.argv[2]) if len(sys.argv) > 2 else rows
d = int(sys.argv[3]) if len(sys.argv) > 3 else 512
s = float(sys.argv[4]) if len(sys.argv) > 4 else 1 / 0.07
off = int(sys.argv[5]) if len(sys.argv) > 5 else 0
n = max(rows, cols)
x, t = O.synthetic_features(n, d, seed=3)
I = torch.from_numpy(x[:rows]).cuda().bfloat16().contiguous(); T = torch.from_numpy(t[:cols]).cuda().bfloat16().contiguous()
be = ops._backend()
X, Y = be.prepare(I), be.prepare(T)
sc = torch.tensor([s], device="cuda")
rs, pos, cs = be.fwd_both(X, Y, sc, off)
torch.cuda.synchronize()
rs2, pos2 = be.fwd_stats(X, Y, sc, off, True)
cs2, _ = be.fwd_stats(Y, X, sc, 0, False)
torch.cuda.synchronize()
lse = lambda st: st[0] + st[1].log()
ex = lambda st: st[2] / st[1]
print(f"rows={rows} cols={cols} d={d} s={s:.2f}")
print("  lse_row   max abs diff vs 2-sweep:", float((lse(rs) - lse(rs2)).abs().max()), " E_row:", float((ex(rs) - ex(rs2)).abs().max()))
print("  lse_col   max abs diff vs 2-sweep:", float((lse(cs) - lse(cs2)).abs().max()), " E_col:", float((ex(cs) - ex(cs2)).abs().max()))
print("  pos       max abs diff vs 2-sweep:", float((pos - pos2).abs().max()))
if rows * cols <= 8192 * 8192:
    S = (I.float() @ T.float().T * s).double()
    print("  vs torch fp64: lse_row", float((lse(rs).double() - torch.logsumexp(S, 1)).abs().max()),
          "lse_col", float((lse(cs).double() - torch.logsumexp(S, 0)).abs().max()),
          "E_col", float((ex(cs).double() - (torch.softmax(S, 0) * S).sum(0)).abs().max()))
def tm(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print(f"  fwd_both {tm(lambda: be.fwd_both(X, Y, sc, off)):.3f} ms   two fwd_stats {tm(lambda: (be.fwd_stats(X, Y, sc, off, True), be.fwd_stats(Y, X, sc, 0, False))):.3f} ms")
