"""Kernel timeline of a few steps (torch.profiler / CUPTI), to see what runs - and what idles - around the library's own
kernels: python tests/tools/step_timeline.py <rows> <d> [steps]   (one GPU, the fused step at W = 1)"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "megatron-clip_b200"))
from clipk import ClipLoss  # noqa: E402
from oracle import cliploss_oracle as O  # noqa: E402

rows, d = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
x, t = O.synthetic_features(rows, d, seed=1)
I = torch.from_numpy(x).cuda().bfloat16().requires_grad_(True)
T = torch.from_numpy(t).cuda().bfloat16().requires_grad_(True)
S = torch.tensor(1 / 0.07, device="cuda", requires_grad=True)
mod = ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True)


def step():
    I.grad = T.grad = S.grad = None
    loss = mod(I, T, S)
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
prev_end = t0
for e in ev:
    gap = e.time_range.start - prev_end
    print(f"{(e.time_range.start - t0):10.1f} us  gap {gap:7.1f}  dur {e.time_range.end - e.time_range.start:8.1f}  {e.name[:90]}")
    prev_end = e.time_range.end
