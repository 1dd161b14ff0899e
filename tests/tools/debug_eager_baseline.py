"""Manual measurement (not collected by pytest): the reference's own op sequence (two matmuls with the scale on the
left operand, two mean cross-entropies, autograd) on the GPU through PyTorch eager - the "existing Blackwell kernels"
(cuBLAS + ATen) that the fused path replaces.  Same synthetic inputs as bench.py."""
import sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
from oracle import cliploss_oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
x, t = O.synthetic_features(N, d, seed=1234)
I = torch.from_numpy(x).cuda().bfloat16().requires_grad_(True)
T = torch.from_numpy(t).cuda().bfloat16().requires_grad_(True)
S = torch.tensor(1 / 0.07, device="cuda", requires_grad=True)
labels = torch.arange(N, device="cuda")
def step():
    I.grad = T.grad = S.grad = None
    a = S * I @ T.T
    b = S * T @ I.T
    loss = (torch.nn.functional.cross_entropy(a, labels) + torch.nn.functional.cross_entropy(b, labels)) / 2
    loss.backward()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
torch.cuda.reset_peak_memory_stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): loss = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({"what": "torch eager ClipLoss fwd+bwd (reference op sequence) on B200, bf16", "N": N, "d": d,
                  "ms_per_step": ms, "samples_per_s": N / (ms * 1e-3), "loss": loss.item(),
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
