"""Stand-in towers for the end-to-end step of BASELINE.json configs[4] / SURVEY 8 f-2 (TEST INFRASTRUCTURE).

The towers are out of scope of this repository: the reference's own `open_clip.CLIP` cannot travel to
the GPU box, so the step test uses this plain-PyTorch two-tower model with the reference's ViT-B/16 hyper-parameters
(open_CLIP/src/open_clip/model_configs/ViT-B-16.json: embed 512; vision 224 px, patch 16, width 768, 12 layers; text
context 77, vocab 49408, width 512, 8 heads, 12 layers) and the same interface as `CLIP.forward`
(open_clip/model.py:232-241): L2-normalised image features, L2-normalised text features taken at the end-of-text token
(the largest token id of each sequence), and `logit_scale.exp()` with logit_scale initialised to log(1 / 0.07).
Only the loss is under test; everything here is library code (torch.nn, scaled_dot_product_attention).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

VIT_B_16 = dict(embed_dim=512, image_size=224, patch=16, vision_width=768, vision_layers=12, vision_heads=12,
                context=77, vocab=49408, text_width=512, text_heads=8, text_layers=12)
TINY = dict(embed_dim=64, image_size=32, patch=8, vision_width=96, vision_layers=2, vision_heads=4,
            context=12, vocab=200, text_width=64, text_heads=4, text_layers=2)


class Block(nn.Module):
    def __init__(self, width, heads, causal):
        super().__init__()
        self.heads, self.causal = heads, causal
        self.ln1, self.ln2 = nn.LayerNorm(width), nn.LayerNorm(width)
        self.qkv, self.out = nn.Linear(width, 3 * width), nn.Linear(width, width)
        self.up, self.down = nn.Linear(width, 4 * width), nn.Linear(4 * width, width)

    def forward(self, x):                                   # [batch, tokens, width]
        b, n, w = x.shape
        q, k, v = self.qkv(self.ln1(x)).view(b, n, 3, self.heads, w // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, is_causal=self.causal)
        x = x + self.out(a.transpose(1, 2).reshape(b, n, w))
        return x + self.down(F.gelu(self.up(self.ln2(x))))


class TwoTowerClip(nn.Module):
    def __init__(self, embed_dim, image_size, patch, vision_width, vision_layers, vision_heads, context, vocab,
                 text_width, text_heads, text_layers):
        super().__init__()
        grid = image_size // patch
        self.patchify = nn.Conv2d(3, vision_width, patch, patch, bias=False)
        self.cls = nn.Parameter(vision_width ** -0.5 * torch.randn(vision_width))
        self.vis_pos = nn.Parameter(vision_width ** -0.5 * torch.randn(grid * grid + 1, vision_width))
        self.vis_pre, self.vis_post = nn.LayerNorm(vision_width), nn.LayerNorm(vision_width)
        self.vis_blocks = nn.ModuleList(Block(vision_width, vision_heads, False) for _ in range(vision_layers))
        self.vis_proj = nn.Parameter(vision_width ** -0.5 * torch.randn(vision_width, embed_dim))
        self.tok = nn.Embedding(vocab, text_width)
        self.txt_pos = nn.Parameter(0.01 * torch.randn(context, text_width))
        self.txt_blocks = nn.ModuleList(Block(text_width, text_heads, True) for _ in range(text_layers))
        self.txt_final = nn.LayerNorm(text_width)
        self.txt_proj = nn.Parameter(text_width ** -0.5 * torch.randn(text_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))

    def encode_image(self, image):
        x = self.patchify(image).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls.expand(x.shape[0], 1, -1).to(x.dtype), x], dim=1) + self.vis_pos.to(x.dtype)
        x = self.vis_pre(x)
        for blk in self.vis_blocks:
            x = blk(x)
        return self.vis_post(x[:, 0]) @ self.vis_proj

    def encode_text(self, text):
        x = self.tok(text) + self.txt_pos
        for blk in self.txt_blocks:
            x = blk(x)
        x = self.txt_final(x)
        return x[torch.arange(x.shape[0], device=x.device), text.argmax(dim=-1)] @ self.txt_proj

    def forward(self, image, text):
        return (F.normalize(self.encode_image(image), dim=-1), F.normalize(self.encode_text(text), dim=-1),
                self.logit_scale.exp())


def synthetic_batch(batch, cfg, seed, device):
    """Images and token rows shaped like the reference's synthetic pipeline (training/data.py:464-482):
    randn images, random tokens between begin (vocab-2) and end-of-text (vocab-1) markers."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, cfg["image_size"], cfg["image_size"], generator=g)
    tokens = torch.randint(1, cfg["vocab"] - 2, (batch, cfg["context"]), generator=g)
    tokens[:, 0] = cfg["vocab"] - 2
    lengths = torch.randint(3, cfg["context"] + 1, (batch,), generator=g)
    for i, n in enumerate(lengths.tolist()):
        tokens[i, n - 1] = cfg["vocab"] - 1
        tokens[i, n:] = 0
    return images.to(device), tokens.to(device)


def reference_step_loss(image_features, text_features, logit_scale):
    """The reference's single-process training loss: loss.py:118-119, 135-138 (world_size == 1)."""
    per_image = logit_scale * image_features @ text_features.T
    per_text = logit_scale * text_features @ image_features.T
    labels = torch.arange(per_image.shape[0], device=per_image.device)
    return (F.cross_entropy(per_image, labels) + F.cross_entropy(per_text, labels)) / 2


def train_step(model, images, tokens, loss_fn, lr=0.05):
    """forward, loss, backward, one SGD step; returns (loss value, {parameter name: gradient})."""
    opt = torch.optim.SGD(model.parameters(), lr=lr)
    opt.zero_grad(set_to_none=True)
    image_features, text_features, scale = model(images, tokens)
    loss = loss_fn(image_features, text_features, scale)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    opt.step()
    return float(loss.detach()), grads
