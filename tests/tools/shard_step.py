"""One rank's kernels at a given block shape, for ncu: python tests/tools/shard_step.py <rows> <cols> <d> [reps]
(the forward sweep, the recompute sweeps and the gradient GEMMs of a [rows x cols] block through the individual entries
- the same kernels the fused step launches; no collectives, one GPU)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "megatron-clip_b200"))
from clipk import ops  # noqa: E402
from oracle import cliploss_oracle as O  # noqa: E402

rows, cols, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
be = ops._backend()
x, _ = O.synthetic_features(rows, d, seed=1)
_, t = O.synthetic_features(cols, d, seed=1)
X = be.prepare(torch.from_numpy(x).cuda().bfloat16())
Y = be.prepare(torch.from_numpy(t).cuda().bfloat16())
sc = torch.tensor([1 / 0.07], device="cuda")
gs = torch.tensor([1.0 / (2 * rows)], device="cuda")
for _ in range(reps):
    parts = torch.empty(1, 3, cols, dtype=torch.float32, device="cuda")
    rs, pos, _ = be.fwd_both(X, Y, sc, 0, col_out=parts[0])
    lse_row, lse_col, sums = be.finalize(rs, pos, parts, 0)
    Xg, Yg = be.prepare_grad(X), be.prepare_grad(Y)
    dX, dY = be.bwd(X, Y, Xg, Yg, sc, 0, lse_row, lse_col, 1.0, 1.0, gs, True, True)
torch.cuda.synchronize()
print("ok", float(sums[0]), float(dX.abs().sum()), float(dY.abs().sum()))
if os.environ.get("SHARD_TIME"):
    def ev(fn, n=10):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    parts = torch.empty(1, 3, cols, dtype=torch.float32, device="cuda")
    t_f = ev(lambda: be.fwd_both(X, Y, sc, 0, col_out=parts[0]))
    t_b = ev(lambda: be.bwd(X, Y, Xg, Yg, sc, 0, lse_row, lse_col, 1.0, 1.0, gs, True, True))
    print(f"timing us: fwd_both {t_f:.1f}  bwd {t_b:.1f}")
