import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "megatron-clip_b200"))
import torch
from clipk import _lib
lib = _lib.load()
out = torch.zeros(8 * 32 * 16, dtype=torch.int32, device="cuda")
_lib.check(lib.clipk_debug_tmem_layout(out.data_ptr(), torch.cuda.current_stream().cuda_stream), "layout")
torch.cuda.synchronize()
o = out.cpu().view(4, 2, 32, 16)
bad = 0
for w in range(4):
    for h in range(2):
        for t in range(32):
            for k in range(16):
                g, c = t // 4, 2 * (t % 4)
                lane = w * 32 + h * 16 + g + (8 if (k % 4) >= 2 else 0)
                col = 8 * (k // 4) + c + (k % 2)
                if int(o[w, h, t, k]) != lane * 1000 + col:
                    bad += 1
                    if bad < 10: print("mismatch", w, h, t, k, int(o[w, h, t, k]), "expected", lane * 1000 + col)
print("thread 5 of warp 1 half 0:", o[1, 0, 5].tolist())
print("layout as assumed" if bad == 0 else f"{bad} mismatches")
