"""Manual profiling target (not collected by pytest): a few ClipLoss fwd+bwd steps at the bench shape."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
from clipk import ClipLoss
from oracle import cliploss_oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
x, t = O.synthetic_features(N, d, seed=1234)
I = torch.from_numpy(x).cuda().bfloat16().requires_grad_(True)
T = torch.from_numpy(t).cuda().bfloat16().requires_grad_(True)
S = torch.tensor(1 / 0.07, device="cuda", requires_grad=True)
mod = ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True)
for _ in range(steps):
    I.grad = T.grad = S.grad = None
    loss = mod(I, T, S)
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
