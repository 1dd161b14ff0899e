"""Manual experiment: phase timeline of CTA 0 in the persistent backward kernel."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "megatron-clip_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from clipk import ops, _lib
from oracle import cliploss_oracle as O
b, d = 32768, 512
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
for _ in range(2): be.bwd(X, Y, Xg, Yg, sc, 0, lr, lcl, 1.0, 1.0, gs, True, True)
torch.cuda.synchronize()
tr = torch.zeros(3 * 3 * 64, dtype=torch.int64, device="cuda")
lib.clipk_debug_set_trace(tr.data_ptr())
be.bwd(X, Y, Xg, Yg, sc, 0, lr, lcl, 1.0, 1.0, gs, True, True)
torch.cuda.synchronize()
lib.clipk_debug_set_trace(None)
v = tr[:60].tolist()
base = v[0]
print("phase: [start, after-barrier-wait, arrive]  (cycles since start); wait = time producer waited for the grid barrier")
prev_arr = 0
for ph in range(12):
    s0, s1, s2 = (v[3 * ph] - base, v[3 * ph + 1] - base, v[3 * ph + 2] - base)
    print(f"ph {ph:2d}: start {s0:8d}  go {s1:8d} (wait {s1 - s0:6d})  arrive {s2:8d}  work {s2 - s1:6d}")
