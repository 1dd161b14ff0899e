"""Manual experiment: per-role clock64 timeline of CTA 0 in the dataflow backward kernel."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
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
TN = 512
tr = torch.zeros(3 * TN, dtype=torch.int64, device="cuda")
lib.clipk_debug_set_trace(tr.data_ptr())
be.bwd(X, Y, Xg, Yg, sc, 0, lr, lcl, 1.0, 1.0, gs, True, True)
torch.cuda.synchronize()
lib.clipk_debug_set_trace(None)
v = tr.tolist()
base = min(x for x in v if x > 0)
names = ["producer (tile/job: start, issued)", "mma (per tile: start, acc free, first data, committed)",
         "epilogue w2 (per tile: start, acc full, acc released, done)"]
per = [2, 4, 4]
nshow = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for role in range(3):
    print(names[role])
    s = [x - base for x in v[role * TN:(role + 1) * TN] if x > 0]
    for i in range(0, min(len(s), per[role] * nshow), per[role]):
        row = s[i:i + per[role]]
        print("  ", " ".join(f"{x:9d}" for x in row), "  d=", " ".join(f"{b - a:7d}" for a, b in zip(row, row[1:])))
