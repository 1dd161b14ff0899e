"""Manual timing (not collected by pytest): ClipLoss fwd / bwd ms at a given shape, CUDA events, L2 flushed."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
from clipk import ClipLoss
from oracle import cliploss_oracle as O
b = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
x, t = O.synthetic_features(b, d, seed=1234)
I = torch.from_numpy(x).cuda().bfloat16().requires_grad_(True)
T = torch.from_numpy(t).cuda().bfloat16().requires_grad_(True)
S = torch.tensor(1 / 0.07, device="cuda", requires_grad=True)
mod = ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def step():
    I.grad = T.grad = S.grad = None
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    flush.fill_(1)
    e[0].record(); loss = mod(I, T, S); e[1].record(); loss.backward(); e[2].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
for _ in range(3): step()
f = b_ = 0.0
n = 8
for _ in range(n):
    a, c = step(); f += a; b_ += c
print(f"b={b} d={d} PANEL_MB={os.environ.get('CLIPK_PANEL_MB','48')}: fwd {f/n:.3f} ms  bwd {b_/n:.3f} ms  total {(f+b_)/n:.3f} ms  "
      f"F_alg TF/s {6.0*b*b*d/((f+b_)/n*1e-3)/1e12:.1f}")
