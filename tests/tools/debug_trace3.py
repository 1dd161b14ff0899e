"""Manual experiment: chip-wide timeline of the dataflow backward (globaltimer stamps of every pair's MMA thread)."""
import sys, os
os.environ["CLIPK_DBG"] = str(int(os.environ.get("CLIPK_DBG", "0")) | 1024)
os.environ["CLIPK_PERSISTENT"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch, numpy as np
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
NC, NP = 74, int(sys.argv[1]) if len(sys.argv) > 1 else 56
tr = torch.zeros(NC * NP * 4 + 4096, dtype=torch.int64, device="cuda")
lib.clipk_debug_set_trace(tr.data_ptr())
be.bwd(X, Y, Xg, Yg, sc, 0, lr, lcl, 1.0, 1.0, gs, True, True)
torch.cuda.synchronize()
lib.clipk_debug_set_trace(None)
v = tr[:NC * NP * 4].cpu().numpy().reshape(NC, NP, 4).astype(np.float64)
t0 = v[v > 0].min()
v = np.where(v > 0, (v - t0) / 1e3, np.nan)   # us
print("panel: start(min..max)  grad_done(min..max)  job_start(min..max)  job_done(min..max) | mean grad us, mean job us")
for q in range(min(NP, 16)):
    a = v[:, q, :]
    f = lambda k: f"{np.nanmin(a[:, k]):8.1f}..{np.nanmax(a[:, k]):8.1f}"
    print(f"{q:3d}: {f(0)}  {f(1)}  {f(2)}  {f(3)} | {np.nanmean(a[:,1]-a[:,0]):6.1f} {np.nanmean(a[:,3]-a[:,2]):6.1f}"
          f"  job max {np.nanmax(a[:,3]-a[:,2]):6.1f} min {np.nanmin(a[:,3]-a[:,2]):6.1f}")
print("total us:", np.nanmax(v))
# per-cluster view of panel 5
q = 5
for c in range(0, NC, 6):
    print(f"cluster {c:2d} panel {q}: " + " ".join(f"{v[c, q, k]:8.1f}" for k in range(4)), " next start", f"{v[c, q + 1, 0]:8.1f}")
