"""Generate tests/golden/shells/*.npz from the UNMODIFIED reference: the callers around the fused path.

    gather_features                      open_CLIP/src/open_clip/loss.py:20-64   (values and gradient flow, 4 modes)
    ClipLoss.get_logits                  loss.py:104-121
    CoCaLoss.forward                     loss.py:143-183
    DistillClipLoss.forward              loss.py:186-221

Run in the build container (needs /root/reference):  python tests/golden/make_golden_shells.py
Same isolated import of loss.py as make_golden.py; world size 2 runs under gloo on CPU.  Inputs are float64 so that the
fixtures pin the formulas, not a rounding order.
"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import load_reference_loss, make_inputs  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shells")


def weights(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64)


def run_rank(rank, world, case, tmp, ret):
    L = load_reference_loss()
    if world > 1:
        dist.init_process_group("gloo", init_method=f"file://{tmp}/store", rank=rank, world_size=world)
    b, d = case["b"], case["d"]
    I, T = make_inputs(b, d, case["seed"], rank, "unit", torch.float64)
    tI, tT = make_inputs(b, d, case["seed"] + 100, rank, "unit", torch.float64)      # teacher features (Distill)
    for x in (I, T):
        x.requires_grad_(True)
    s = torch.tensor(case["scale"], dtype=torch.float64, requires_grad=True)
    kw = dict(local_loss=case["local_loss"], gather_with_grad=case["gather_with_grad"], cache_labels=True, rank=rank,
              world_size=world)
    out = dict(image=I.detach().numpy(), text=T.detach().numpy(), t_image=tI.numpy(), t_text=tT.numpy())

    def grads(prefix):
        for n, x in (("d_image", I), ("d_text", T), ("d_scale", s)):
            out[prefix + n] = np.zeros(x.shape) if x.grad is None else x.grad.numpy().copy()
            x.grad = None

    if case["what"] == "gather":
        all_i, all_t = L.gather_features(I, T, case["local_loss"], case["gather_with_grad"], rank, world)
        out["all_image"], out["all_text"] = all_i.detach().numpy(), all_t.detach().numpy()
        # gradient flow: a rank-dependent linear functional of the gathered tensors
        wi, wt = weights(all_i.shape, 900 + rank), weights(all_t.shape, 950 + rank)
        f = (all_i * wi).sum() + (all_t * wt).sum()
        if f.requires_grad:
            f.backward()
        grads("g_")
        out["w_image"], out["w_text"] = wi.numpy(), wt.numpy()
    elif case["what"] == "logits":
        mod = L.ClipLoss(**kw)
        per_image, per_text = mod.get_logits(I, T, s)
        out["per_image"], out["per_text"] = per_image.detach().numpy(), per_text.detach().numpy()
        wi, wt = weights(per_image.shape, 900 + rank), weights(per_text.shape, 950 + rank)
        ((per_image * wi).sum() + (per_text * wt).sum()).backward()
        grads("g_")
        out["w_image"], out["w_text"] = wi.numpy(), wt.numpy()
    elif case["what"] == "coca":
        mod = L.CoCaLoss(caption_loss_weight=2.0, clip_loss_weight=0.5, pad_id=0, **kw)
        g = torch.Generator().manual_seed(case["seed"] + 7 + rank)
        cap_logits = torch.randn(b, 6, 11, generator=g, dtype=torch.float64, requires_grad=True)   # [b, seq, vocab]
        cap_labels = torch.randint(0, 11, (b, 6), generator=g)                                     # 0 = pad, ignored
        clip_loss, caption_loss = mod(I, T, cap_logits, cap_labels, s)
        (clip_loss * 1.5 + caption_loss).backward()
        grads("g_")
        out.update(cap_logits=cap_logits.detach().numpy(), cap_labels=cap_labels.numpy(),
                   clip_loss=clip_loss.detach().numpy(), caption_loss=caption_loss.detach().numpy(),
                   g_cap_logits=cap_logits.grad.numpy())
        d_out = mod(I, T, cap_logits, cap_labels, s, output_dict=True)
        out["dict_keys"] = np.array(sorted(d_out))
    elif case["what"] == "distill":
        mod = L.DistillClipLoss(**kw)
        ts = torch.tensor(case["scale"] * 1.7, dtype=torch.float64)
        contrastive, distill = mod(I, T, s, tI, tT, ts)
        (contrastive + 2.0 * distill).backward()
        grads("g_")
        out.update(contrastive_loss=contrastive.detach().numpy(), distill_loss=distill.detach().numpy(),
                   t_scale=ts.numpy())
        d_out = mod(I, T, s, tI, tT, ts, output_dict=True)
        out["dict_keys"] = np.array(sorted(d_out))
    else:
        raise ValueError(case["what"])
    if world > 1:
        np.savez(f"{tmp}/rank{rank}.npz", **out)
        dist.barrier()
        dist.destroy_process_group()
    else:
        ret.update(out)


CASES = []
MODES = [(False, False), (False, True), (True, False), (True, True)]
for (ll, gwg) in MODES:
    CASES.append(dict(what="gather", world=2, b=5, d=16, seed=31, scale=1 / 0.07, local_loss=ll, gather_with_grad=gwg))
    CASES.append(dict(what="logits", world=2, b=6, d=16, seed=37, scale=1 / 0.07, local_loss=ll, gather_with_grad=gwg))
    CASES.append(dict(what="distill", world=2, b=7, d=24, seed=41, scale=1 / 0.07, local_loss=ll, gather_with_grad=gwg))
    CASES.append(dict(what="coca", world=2, b=6, d=24, seed=43, scale=1 / 0.07, local_loss=ll, gather_with_grad=gwg))
for what in ("logits", "distill", "coca"):
    CASES.append(dict(what=what, world=1, b=9, d=24, seed=47, scale=1 / 0.07, local_loss=False, gather_with_grad=False))


def case_name(c):
    return f"{c['what']}_w{c['world']}_b{c['b']}_d{c['d']}_ll{int(c['local_loss'])}_gwg{int(c['gather_with_grad'])}"


def main():
    torch.set_num_threads(1)
    os.makedirs(OUT, exist_ok=True)
    for c in CASES:
        W = c["world"]
        if W == 1:
            ret = {}
            run_rank(0, 1, c, None, ret)
            ranks = [ret]
        else:
            with tempfile.TemporaryDirectory() as tmp:
                mp.spawn(run_rank, args=(W, c, tmp, None), nprocs=W, join=True)
                ranks = [dict(np.load(f"{tmp}/rank{r}.npz")) for r in range(W)]
        flat = {k: np.array(v) for k, v in c.items()}
        for r, o in enumerate(ranks):
            for k, v in o.items():
                flat[f"r{r}_{k}"] = v
        path = os.path.join(OUT, case_name(c) + ".npz")
        np.savez_compressed(path, **flat)
        print("wrote", path)


if __name__ == "__main__":
    main()
