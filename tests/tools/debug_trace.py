"""Manual experiment: clock64 trace of CTA (0,0) for one GRAD launch and one pair launch of the backward."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
from clipk import ops, _lib
from oracle import cliploss_oracle as O
b, d = 4736, 512
x, t = O.synthetic_features(b, d, seed=1234)
I = torch.from_numpy(x).cuda().bfloat16(); T = torch.from_numpy(t).cuda().bfloat16()
be = ops._backend(); lib = _lib.load()
X, Y = be.prepare(I), be.prepare(T)
sc = torch.tensor([1 / 0.07], device="cuda")
rs, pos = be.fwd_stats(X, Y, sc, 0, True)
parts = torch.empty(1, 3, b, device="cuda"); be.fwd_stats(Y, X, sc, 0, False, out=parts[0])
lr, lcl, sums = be.finalize(rs, pos, parts, 0)
Xg, Yg = be.prepare_grad(X), be.prepare_grad(Y)
gs = torch.tensor([1.0 / (2 * b)], device="cuda")
for _ in range(3): be.bwd(X, Y, Xg, Yg, sc, 0, lr, lcl, 1.0, 1.0, gs, True, True)
torch.cuda.synchronize()
tr = torch.zeros(3, 3, 64, dtype=torch.int64, device="cuda")
lib.clipk_debug_set_trace(tr.data_ptr())
be.bwd(X, Y, Xg, Yg, sc, 0, lr, lcl, 1.0, 1.0, gs, True, True)
rs, pos = be.fwd_stats(X, Y, sc, 0, True)
torch.cuda.synchronize()
lib.clipk_debug_set_trace(None)
names = {0: "STATS", 1: "GRAD", 2: "OUT(pair)"}
roles = {0: "producer", 1: "mma", 2: "epilogue(warp2)"}
for m in (1, 2, 0):
    allv = [x for r in range(3) for x in tr[m, r].tolist() if x != 0]
    base = min(allv)
    for r in range(3):
        v = [x - base for x in tr[m, r].tolist() if x != 0]
        print(f"{names[m]:10s} {roles[r]:16s} n={len(v):2d}", v[:44])
