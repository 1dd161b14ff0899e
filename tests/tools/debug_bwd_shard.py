"""Manual experiment: fwd+bwd kernel time for one rank's shard (rows x N) without collectives."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
from clipk import ops
from oracle import cliploss_oracle as O
rows = int(sys.argv[1]); N = int(sys.argv[2]); d = int(sys.argv[3]) if len(sys.argv) > 3 else 512
x, t = O.synthetic_features(N, d, seed=1234)
I = torch.from_numpy(x[:rows]).cuda().bfloat16(); T = torch.from_numpy(t).cuda().bfloat16()
be = ops._backend()
sc = torch.tensor([1 / 0.07], device="cuda")
gs = torch.tensor([1.0 / (2 * rows)], device="cuda")
def step():
    X, Y = be.prepare(I), be.prepare(T)
    parts = torch.empty(1, 3, N, device="cuda")
    rs, pos, _ = be.fwd_both(X, Y, sc, 0, col_out=parts[0])
    lr, lcl, sums = be.finalize(rs, pos, parts, 0)
    Xg, Yg = be.prepare_grad(X), be.prepare_grad(Y)
    return be.bwd(X, Y, Xg, Yg, sc, 0, lr, lcl, 1.0, 1.0, gs, True, True)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
import time
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"CPU launch time per step: {(t1 - t0) / 10 * 1e3:.3f} ms")
print(f"rows={rows} N={N} d={d} PERSISTENT={os.environ.get('CLIPK_PERSISTENT','1')} PANEL_MB={os.environ.get('CLIPK_PANEL_MB','48')}: "
      f"fwd+bwd kernels {ms:.3f} ms  ({6.0*rows*N*d/(ms*1e-3)/1e12:.0f} TF/s F_alg)")
