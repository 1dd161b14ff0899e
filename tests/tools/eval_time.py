"""Manual measurement (not collected by pytest): retrieval ranks of N validation pairs, clipk.target_ranks against the
reference's way (logits, argsort, where - training/train.py:631-648) run on the same GPU in row chunks.
Prints one JSON line.   python tests/tools/eval_time.py [N] [d]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
import clipk

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
g = torch.Generator(device="cuda").manual_seed(1)
T = torch.nn.functional.normalize(torch.randn(N, d, device="cuda", generator=g), dim=-1)
I = torch.nn.functional.normalize(0.2 * T + torch.nn.functional.normalize(torch.randn(N, d, device="cuda", generator=g), dim=-1), dim=-1)
I, T = I.bfloat16(), T.bfloat16()


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def eager(chunk=4096):
    gt = torch.arange(N, device="cuda").view(-1, 1)
    preds = []
    for r0 in range(0, N, chunk):
        logits = I[r0:r0 + chunk].float() @ T.float().T
        ranking = torch.argsort(logits, descending=True)
        preds.append(torch.where(ranking == gt[r0:r0 + chunk])[1])
    return torch.cat(preds)


l0 = clipk.gpu_launches()
ms, ranks = timed(lambda: clipk.target_ranks(I, T), 5)
launches = (clipk.gpu_launches() - l0) // 6
ms_e, ranks_e = timed(eager, 1)
agree = float((ranks == ranks_e).float().mean())
logit_bytes = 4.0 * N * N
print(json.dumps({"what": "image->text retrieval ranks", "N": N, "d": d, "dtype": "bf16", "clipk_ms": round(ms, 3),
                  "clipk_launches": launches, "logits_GBps_write_plus_read": round(2 * logit_bytes / ms / 1e6, 1),
                  "gemm_tflops": round(2.0 * N * N * d / ms / 1e9, 1), "torch_argsort_chunks_ms": round(ms_e, 3),
                  "ranks_equal_frac": round(agree, 5), "R@1": float((ranks < 1).float().mean())}))
