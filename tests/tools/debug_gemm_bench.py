"""Manual microbenchmark (not collected by pytest): mainloop throughput by operand layout and shape."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
import torch
from clipk import _lib
lib = _lib.load()
st = lambda: torch.cuda.current_stream().cuda_stream

def run(M, N, K, a_mn, b_mn, f16, reps=10):
    dt = torch.float16 if f16 else torch.bfloat16
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(dt)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(dt)
    D = torch.empty(M, N, device="cuda")
    def go():
        _lib.check(lib.clipk_gemm16(A.data_ptr(), B.data_ptr(), D.data_ptr(), M, N, K, A.stride(0), B.stride(0), N,
                                    a_mn, b_mn, f16, 0, st()), "gemm")
    go(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ctas = ((M + 127) // 128) * ((N + 255) // 256)
    tf = 2.0 * M * N * K / ms / 1e9
    print(f"M={M:6d} N={N:6d} K={K:6d} a_mn={a_mn} b_mn={b_mn} f16={f16}: {ms*1e3:8.1f} us  {tf:7.1f} TF/s  ctas={ctas:5d}  "
          f"per-active-SM-equiv={tf*148/min(ctas,148):7.1f}")

for (a_mn, b_mn, f16) in [(0, 0, 0), (0, 1, 1), (1, 1, 1), (0, 0, 1)]:
    run(4096, 512, 4096, a_mn, b_mn, f16)
for (a_mn, b_mn, f16) in [(0, 0, 0), (0, 1, 1), (1, 1, 1)]:
    run(128 * 74, 512, 4096, a_mn, b_mn, f16)      # exactly 148 CTAs
    run(128 * 74, 512, 16384, a_mn, b_mn, f16)
run(16384, 16384, 512, 0, 0, 0)
run(16384, 16384, 4096, 0, 0, 0)
run(16384, 16384, 4096, 0, 1, 1)
run(16384, 16384, 4096, 1, 1, 1)
