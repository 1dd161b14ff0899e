"""Generate tests/golden/eval/*.npz from the UNMODIFIED reference's evaluation functions.

    get_clip_metrics   open_CLIP/src/training/train.py:631-648
    accuracy           open_CLIP/src/training/zero_shot.py:36-39

Run in the build container (needs /root/reference):  python tests/golden/make_golden_eval.py
Their modules import wandb / open_clip / tqdm machinery that is not installed here, so the two functions are cut out of
the reference source with `ast` and executed as they are, with torch and numpy in scope.
"""
import ast
import os

import numpy as np
import torch

REF = os.environ.get("CLIPK_REF_DIR", "/root/reference") + "/open_CLIP/src/training"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "eval")


def reference_function(path, name):
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def features(n, d, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    x = torch.nn.functional.normalize(torch.randn(n, d, generator=g, dtype=torch.float64), dim=-1)
    z = torch.nn.functional.normalize(torch.randn(n, d, generator=g, dtype=torch.float64), dim=-1)
    t = torch.nn.functional.normalize(0.12 * x + z, dim=-1)      # weak pairing: ranks spread over many values
    return x.to(dtype), t.to(dtype)


def main():
    os.makedirs(OUT, exist_ok=True)
    get_clip_metrics = reference_function(f"{REF}/train.py", "get_clip_metrics")
    accuracy = reference_function(f"{REF}/zero_shot.py", "accuracy")
    for (n, d, seed, dtype, scale) in [(64, 32, 1, torch.float64, 14.2857), (333, 64, 2, torch.float32, 100.0),
                                       (1000, 128, 3, torch.float32, 1.0)]:
        I, T = features(n, d, seed, dtype)
        m = get_clip_metrics(I, T, torch.tensor(scale, dtype=dtype))
        path = os.path.join(OUT, f"metrics_n{n}_d{d}_{str(dtype).split('.')[-1]}.npz")
        np.savez_compressed(path, image=I.numpy(), text=T.numpy(), scale=np.array(scale),
                            keys=np.array(sorted(m)), values=np.array([float(m[k]) for k in sorted(m)]))
        print("wrote", path, {k: round(float(v), 4) for k, v in m.items()})
    for (n, d, classes, seed) in [(200, 64, 50, 5), (512, 128, 1000, 6)]:
        I, _ = features(n, d, seed, torch.float32)
        g = torch.Generator().manual_seed(seed + 50)
        classifier = torch.nn.functional.normalize(torch.randn(d, classes, generator=g), dim=0)
        target = torch.randint(0, classes, (n,), generator=g)
        # make about a third of the rows easy, so that top-1 / top-5 are neither 0 nor n
        I[: n // 3] = torch.nn.functional.normalize(I[: n // 3] + 0.5 * classifier.T[target[: n // 3]], dim=-1)
        logits = 100. * I @ classifier                              # zero_shot.py:57
        acc = accuracy(logits, target, topk=(1, 5))
        path = os.path.join(OUT, f"zeroshot_n{n}_d{d}_c{classes}.npz")
        np.savez_compressed(path, image=I.numpy(), classifier=classifier.numpy(), target=target.numpy(),
                            topk=np.array([1, 5]), correct=np.array(acc))
        print("wrote", path, acc)


if __name__ == "__main__":
    main()
