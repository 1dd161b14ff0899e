"""Generate tests/golden/*.npz from the UNMODIFIED reference open_CLIP/src/open_clip/loss.py.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so its outputs are committed as small fixtures.

`import open_clip` fails here (ftfy/webdataset/... missing), so loss.py is loaded in isolation: a stub package
named open_clip whose __path__ is the reference directory, then tprofiler and loss by file location.
World sizes > 1 run the reference under torch.distributed with the gloo backend on CPU.
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REF = os.environ.get("CLIPK_REF_DIR", "/root/reference") + "/open_CLIP/src/open_clip"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference_loss():
    pkg = types.ModuleType("open_clip")
    pkg.__path__ = [REF]
    sys.modules["open_clip"] = pkg
    for n in ("tprofiler", "loss"):
        spec = importlib.util.spec_from_file_location("open_clip." + n, f"{REF}/{n}.py")
        m = importlib.util.module_from_spec(spec)
        sys.modules["open_clip." + n] = m
        spec.loader.exec_module(m)
    return sys.modules["open_clip.loss"]


def make_inputs(b, d, seed, rank, kind, dtype):
    g = torch.Generator().manual_seed(seed + rank)
    x = torch.randn(b, d, generator=g, dtype=torch.float64)
    z = torch.randn(b, d, generator=g, dtype=torch.float64)
    if kind == "unit":          # the benchmark's synthetic distribution
        x = torch.nn.functional.normalize(x, dim=-1)
        z = torch.nn.functional.normalize(z, dim=-1)
        t = torch.nn.functional.normalize(0.3 * x + 0.954 * z, dim=-1)
    elif kind == "raw":         # un-normalised: the API does not require unit norm
        x = 0.25 * x
        t = 0.25 * z
    else:
        raise ValueError(kind)
    return x.to(dtype), t.to(dtype)


def run_rank(rank, world, case, tmp, ret):
    L = load_reference_loss()
    if world > 1:
        dist.init_process_group("gloo", init_method=f"file://{tmp}/store", rank=rank, world_size=world)
    dtype = getattr(torch, case["dtype"])
    I, T = make_inputs(case["b"], case["d"], case["seed"], rank, case["kind"], dtype)
    I.requires_grad_(True)
    T.requires_grad_(True)
    s = torch.tensor(case["scale"], dtype=dtype if dtype == torch.float64 else torch.float32, requires_grad=True)
    mod = L.ClipLoss(local_loss=case["local_loss"], gather_with_grad=case["gather_with_grad"],
                     cache_labels=True, rank=rank, world_size=world)
    loss = mod(I, T, s)
    (loss * case["grad_output"]).backward()
    n_logits = case["b"] if (world == 1 or case["local_loss"]) else case["b"] * world
    labels = mod.get_ground_truth(I.device, n_logits)
    out = dict(image=I.detach().numpy(), text=T.detach().numpy(), loss=loss.detach().numpy(),
               d_image=I.grad.numpy(), d_text=T.grad.numpy(), d_scale=s.grad.numpy(), labels=labels.numpy())
    if world > 1:
        np.savez(f"{tmp}/rank{rank}.npz", **out)
        dist.barrier()
        dist.destroy_process_group()
    else:
        ret.update(out)


CASES = []
for W in (1, 2, 4):
    modes = [(False, False)] if W == 1 else [(False, False), (False, True), (True, False), (True, True)]
    for (ll, gwg) in modes:
        CASES.append(dict(world=W, b=8, d=32, seed=11, kind="unit", dtype="float64", scale=1 / 0.07,
                          local_loss=ll, gather_with_grad=gwg, grad_output=1.0))
        CASES.append(dict(world=W, b=12, d=64, seed=23, kind="raw", dtype="float32", scale=100.0,
                          local_loss=ll, gather_with_grad=gwg, grad_output=3.0))
# C1 of BASELINE.json: CPU, single process, batch 256, d=512, fp32
CASES.append(dict(world=1, b=256, d=512, seed=1234, kind="unit", dtype="float32", scale=1 / 0.07,
                  local_loss=False, gather_with_grad=False, grad_output=1.0))
# ragged single-process shapes
CASES.append(dict(world=1, b=100, d=128, seed=5, kind="unit", dtype="float32", scale=1 / 0.07,
                  local_loss=False, gather_with_grad=False, grad_output=1.0))
CASES.append(dict(world=2, b=33, d=128, seed=7, kind="unit", dtype="float32", scale=1 / 0.07,
                  local_loss=True, gather_with_grad=True, grad_output=1.0))


def case_name(c):
    return (f"w{c['world']}_b{c['b']}_d{c['d']}_{c['kind']}_{c['dtype']}_s{int(round(c['scale']))}"
            f"_ll{int(c['local_loss'])}_gwg{int(c['gather_with_grad'])}")


def main():
    torch.set_num_threads(1)
    for c in CASES:
        W = c["world"]
        meta = {k: np.array(v) for k, v in c.items()}
        if W == 1:
            ret = {}
            run_rank(0, 1, c, None, ret)
            ranks = [ret]
        else:
            with tempfile.TemporaryDirectory() as tmp:
                mp.spawn(run_rank, args=(W, c, tmp, None), nprocs=W, join=True)
                ranks = [dict(np.load(f"{tmp}/rank{r}.npz")) for r in range(W)]
        flat = dict(meta)
        for r, o in enumerate(ranks):
            for k, v in o.items():
                flat[f"r{r}_{k}"] = v
        path = os.path.join(OUT, case_name(c) + ".npz")
        np.savez_compressed(path, **flat)
        print("wrote", path, "loss(rank0)=", float(ranks[0]["loss"]))


if __name__ == "__main__":
    main()
