import sys, time, os
import numpy as np, torch
sys.path.insert(0, ".")
import bliss_b200
from bliss_b200 import engine as E
n_in, Be = 7938000, 192
stride = (n_in + 63) // 64 * 64 + 64
pinned = torch.empty(Be * stride, dtype=torch.float32, pin_memory=True)
g = torch.Generator(); g.manual_seed(1)
blk = (torch.rand(stride, generator=g) - 0.5) * 0.4
for i in range(Be):
    pinned[i * stride:(i + 1) * stride] = torch.roll(blk, i * 977)
ptrs = [pinned.data_ptr() + 4 * i * stride for i in range(Be)]
lens = [n_in] * Be
eng = bliss_b200.Engine(0)
out = np.zeros(Be, dtype=bliss_b200.RESULT_DTYPE)
eng.analyze_host_ptrs(E.FMT_F32, ptrs, lens, out=out)
os.environ["BLX_TRACE"] = "1"
t0 = time.perf_counter()
eng.analyze_host_ptrs(E.FMT_F32, ptrs, lens, out=out)
print("total ms", (time.perf_counter() - t0) * 1e3)
