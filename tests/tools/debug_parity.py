"""Manual diagnostic (not collected by pytest): print kernel-vs-oracle errors for a few shapes."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
from oracle import cliploss_oracle as O
from tests.util import make_inputs, rel
from clipk import ClipLoss

for (b, d, dtype, s, kind) in [(128, 256, torch.bfloat16, 1.0, "unit"), (16, 1024, torch.bfloat16, 14.2857, "unit"),
                               (256, 512, torch.bfloat16, 14.2857, "unit"), (256, 512, torch.float32, 14.2857, "unit"),
                               (300, 256, torch.float32, 100.0, "unit"), (200, 192, torch.bfloat16, 14.2857, "raw"),
                               (2048, 256, torch.bfloat16, 14.2857, "unit"), (4300, 64, torch.bfloat16, 14.2857, "unit")]:
    x, t = make_inputs(b, d, seed=b + d, kind=kind)
    I = torch.from_numpy(x).cuda().to(dtype).requires_grad_(True)
    T = torch.from_numpy(t).cuda().to(dtype).requires_grad_(True)
    S = torch.tensor(s, device="cuda", requires_grad=True)
    loss = ClipLoss()(I, T, S)
    loss.backward()
    torch.cuda.synchronize()
    ref = O.clip_loss_single(I.detach().float().cpu().numpy(), T.detach().float().cpu().numpy(), s)
    dI = I.grad.float().cpu().numpy(); dT = T.grad.float().cpu().numpy()
    print(f"b={b} d={d} {dtype} s={s} {kind}: loss {loss.item():.6f} ref {ref.loss:.6f} rel {abs(loss.item()-ref.loss)/abs(ref.loss):.2e} | "
          f"dI {rel(dI, ref.d_image):.2e} dT {rel(dT, ref.d_text):.2e} | ds {S.grad.item():.6e} ref {ref.d_scale:.6e} "
          f"| maxdI {np.abs(dI-ref.d_image).max()/np.abs(ref.d_image).max():.2e}")
