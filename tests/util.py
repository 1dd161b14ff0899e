"""Shared helpers for the parity tests (tests only)."""
import glob
import os

import numpy as np

from oracle import cliploss_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load_golden(name_or_path):
    path = name_or_path if os.path.isabs(name_or_path) else os.path.join(GOLDEN, name_or_path)
    z = dict(np.load(path))
    W = int(z["world"])
    ranks = [{k[len(f"r{r}_"):]: v for k, v in z.items() if k.startswith(f"r{r}_")} for r in range(W)]
    return z, W, ranks


def golden_files(world=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        if world is None or os.path.basename(p).startswith(f"w{world}_"):
            out.append(p)
    return out


def rel(a, b):
    """|a - b|_2 / |b|_2 for arrays, |a-b|/|b| for scalars."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


def make_inputs(b, d, seed, kind="unit", rank=0):
    if kind == "unit":
        return O.synthetic_features(b, d, seed, rank)
    rng = np.random.default_rng(seed + rank)
    return (0.25 * rng.standard_normal((b, d))).astype(np.float32), (0.25 * rng.standard_normal((b, d))).astype(np.float32)
