import sys, json, torch, numpy as np
sys.path.insert(0, 'megatron-clip_b200'); sys.path.insert(0, '.')
from clipk import ClipLoss
from oracle import cliploss_oracle as O
b = int(sys.argv[1]); d = int(sys.argv[2])
x, t = O.synthetic_features(b, d, seed=21)
I = torch.from_numpy(x).cuda().bfloat16().requires_grad_(True)
T = torch.from_numpy(t).cuda().bfloat16().requires_grad_(True)
S = torch.tensor(1 / 0.07, device='cuda', requires_grad=True)
ClipLoss()(I, T, S).backward(); torch.cuda.synchronize()
ref = O.clip_loss_single(I.detach().float().cpu().numpy(), T.detach().float().cpu().numpy(), 1 / 0.07)
r = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
print(b, d, json.dumps([r(I.grad.float().cpu().numpy(), ref.d_image), r(T.grad.float().cpu().numpy(), ref.d_text)]))
