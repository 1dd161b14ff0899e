"""Manual experiment: chunk-level clock64 stamps of epilogue warp 2 of CTA 0 (CLIPK_DBG bit 2048), forward or backward."""
import sys, os
os.environ["CLIPK_DBG"] = str(int(os.environ.get("CLIPK_DBG", "0")) | 2048)
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
from clipk import ops, _lib
from oracle import cliploss_oracle as O
b, d = 32768, 512
which = sys.argv[1] if len(sys.argv) > 1 else "bwd"
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
def run():
    if which == "bwd":
        be.bwd(X, Y, Xg, Yg, sc, 0, lr, lcl, 1.0, 1.0, gs, True, True)
    else:
        be.fwd_stats(X, Y, sc, 0, True)
for _ in range(2): run()
torch.cuda.synchronize()
TN = 512
tr = torch.zeros(3 * TN, dtype=torch.int64, device="cuda")
lib.clipk_debug_set_trace(tr.data_ptr())
run()
torch.cuda.synchronize()
lib.clipk_debug_set_trace(None)
v = tr.tolist()
m = [x for x in v[TN:2 * TN] if x > 0]
print("mma per tile: start | acc free, first data, committed (deltas)")
for i in range(0, min(len(m), 4 * 40), 4):
    row = m[i:i + 4]
    print("  ", f"{row[0] - m[0]:9d}", " ".join(f"{b - a:6d}" for a, b in zip(row, row[1:])))
s = [x for x in v[2 * TN:3 * TN] if x > 0]
NS = 5 if which == "bwd" else 4
print("epilogue w2 per tile: start | accfull, released, done, (flag)   (deltas)")
for i in range(0, min(len(s), NS * 40), NS):
    row = s[i:i + NS]
    print("  ", f"{row[0] - s[0]:9d}", " ".join(f"{b - a:6d}" for a, b in zip(row, row[1:])))
