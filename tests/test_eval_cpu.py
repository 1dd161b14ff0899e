"""Evaluation-side contractions without a GPU: the oracle against outputs of the reference's own
get_clip_metrics / accuracy (tests/golden/eval, made by tests/golden/make_golden_eval.py), and the host logic of
clipk.metrics (panel loop, target handling, tie and sign conventions) with the two kernel calls emulated."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import eval_oracle as E

EVAL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval")
METRIC_FILES = sorted(glob.glob(os.path.join(EVAL, "metrics_*.npz")))
ZEROSHOT_FILES = sorted(glob.glob(os.path.join(EVAL, "zeroshot_*.npz")))


@pytest.fixture
def emu():
    from clipk import ops
    from tests.emu_backend import EmuBackend
    ops.set_backend_for_testing(EmuBackend())
    yield
    ops.set_backend_for_testing(None)


def test_fixtures_present():
    assert len(METRIC_FILES) == 3 and len(ZEROSHOT_FILES) == 2


@pytest.mark.parametrize("path", METRIC_FILES, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_metrics_match_reference(path):
    z = np.load(path)
    got = E.clip_metrics(z["image"], z["text"], float(z["scale"]))
    assert sorted(got) == list(z["keys"])
    n = z["image"].shape[0]
    for k, v in zip(z["keys"], z["values"]):
        # the reference ranks fp32 logits, the oracle fp64 ones: a pair of logits closer than fp32 rounding may swap,
        # which moves one rank by one
        slack = 0.0 if z["image"].dtype == np.float64 else (2.0 / n if "mean" in k or "R@" in k else 1.0)
        assert abs(got[k] - v) <= slack, (k, got[k], v)


@pytest.mark.parametrize("path", ZEROSHOT_FILES, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_topk_matches_reference(path):
    z = np.load(path)
    logits = 100.0 * z["image"].astype(np.float64) @ z["classifier"].astype(np.float64)
    got = E.topk_correct(logits, z["target"], tuple(z["topk"]))
    assert all(abs(a - b) <= 1 for a, b in zip(got, z["correct"])), (got, z["correct"])


@pytest.mark.parametrize("path", METRIC_FILES, ids=lambda p: os.path.basename(p)[:-4])
def test_get_clip_metrics_host_logic(path, emu):
    import clipk
    z = np.load(path)
    I, T = torch.from_numpy(z["image"]).float(), torch.from_numpy(z["text"]).float()
    got = clipk.get_clip_metrics(I, T, torch.tensor(float(z["scale"])))
    ref = E.clip_metrics(I.numpy(), T.numpy(), float(z["scale"]))
    assert sorted(got) == list(z["keys"])
    for k in ref:
        assert got[k] == ref[k], k
    # several panels give the same ranks as one
    one = clipk.target_ranks(I, T)
    many = clipk.target_ranks(I, T, panel_bytes=1)
    assert torch.equal(one, many) and one.dtype == torch.long


@pytest.mark.parametrize("path", ZEROSHOT_FILES, ids=lambda p: os.path.basename(p)[:-4])
def test_zero_shot_accuracy_host_logic(path, emu):
    import clipk
    z = np.load(path)
    got = clipk.zero_shot_accuracy(torch.from_numpy(z["image"]), torch.from_numpy(z["classifier"]),
                                   torch.from_numpy(z["target"]), tuple(int(k) for k in z["topk"]))
    assert all(abs(a - b) <= 1 for a, b in zip(got, z["correct"])), (got, z["correct"])
    assert all(isinstance(a, float) for a in got)


def test_ties_sign_and_errors(emu):
    import clipk
    g = torch.Generator().manual_seed(0)
    keys = torch.randn(7, 8, generator=g)
    keys[5] = keys[2]                                   # columns 2 and 5 tie in every row
    q = torch.randn(4, 8, generator=g)
    r2 = clipk.target_ranks(q, keys, target=torch.tensor([2, 2, 2, 2]))
    r5 = clipk.target_ranks(q, keys, target=torch.tensor([5, 5, 5, 5]))
    assert torch.equal(r5, r2 + 1)                      # stable: the smaller column index comes first
    logits = (q.double() @ keys.double().T).numpy()
    assert np.array_equal(r2.numpy(), E.target_ranks(logits, [2] * 4))
    assert np.array_equal(r5.numpy(), E.target_ranks(logits, [5] * 4))
    # 7 columns is not a multiple of 4: the padding columns are never counted
    assert int(clipk.target_ranks(q, keys, target=torch.tensor([0, 1, 3, 6])).max()) <= 6
    with pytest.raises(IndexError):
        clipk.target_ranks(q, keys, target=torch.tensor([0, 1, 2, 7]))
    with pytest.raises(TypeError):
        clipk.target_ranks(q, keys.double())
    # a negative logit_scale reverses every ranking, zero makes everything tie (stable order = column order)
    I, T = torch.randn(9, 8, generator=g), torch.randn(9, 8, generator=g)
    neg = clipk.get_clip_metrics(I, T, torch.tensor(-3.0))
    ref = E.clip_metrics(I.numpy(), T.numpy(), -3.0)
    assert all(neg[k] == ref[k] for k in ref)
    zero = clipk.get_clip_metrics(I, T, 0.0)
    ref0 = E.clip_metrics(I.numpy(), T.numpy(), 0.0)
    assert all(zero[k] == ref0[k] for k in ref0)


def test_no_cpu_fallback_for_metrics():
    import clipk
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Exception):
        clipk.target_ranks(torch.randn(4, 64), torch.randn(8, 64))
