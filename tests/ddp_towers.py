"""Toy two-tower model + the comparison the DDP tests make (tests only).

DistributedDataParallel(static_graph=True) around towers whose loss is clipk.ClipLoss(local_loss=True,
gather_with_grad=True) - the reference's training configuration (training/main.py:283-293, training/train.py:148) - must
give every rank the gradient a single process gets on the global batch with the reference's formula: the reduce-scatter
inside the loss returns the other ranks' contributions to this rank's features, DDP's average does the rest."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Towers(nn.Module):
    def __init__(self, d_img, d_txt, d_out, feat_dtype):
        super().__init__()
        self.img = nn.Linear(d_img, d_out)
        self.txt = nn.Linear(d_txt, d_out)
        self.logit_scale = nn.Parameter(torch.tensor(2.0))
        self.feat_dtype = feat_dtype

    def forward(self, images, texts):
        i = F.normalize(self.img(images), dim=-1).to(self.feat_dtype)
        t = F.normalize(self.txt(texts), dim=-1).to(self.feat_dtype)
        return i, t, self.logit_scale.exp()


def global_batch(n, d_img, d_txt, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d_img, generator=g), torch.randn(n, d_txt, generator=g)


def reference_grads(model, images, texts):
    """single process, global batch, the reference's op sequence (loss.py:117-119, 135-138) in fp32"""
    model.zero_grad()
    i, t, s = model(images, texts)
    a = s * i.float() @ t.float().T
    lab = torch.arange(a.shape[0], device=a.device)
    loss = (F.cross_entropy(a, lab) + F.cross_entropy(a.T, lab)) / 2
    loss.backward()
    return float(loss.detach()), {k: p.grad.detach().clone() for k, p in model.named_parameters()}
