import sys, time, ctypes
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import bliss_b200
from bliss_b200 import engine as E
eng = bliss_b200.Engine(0)
n30 = 30 * 44100
B = 1024
stride = (n30 + 63) // 64 * 64 + 64
buf = (torch.rand(B * stride, device="cuda") - 0.5) * 0.3
so = (ctypes.c_int64 * B)(*[i * stride for i in range(B)])
sl = (ctypes.c_int64 * B)(*([n30] * B))
d = torch.zeros(B, dtype=torch.float32, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for prof in (False, True):
    eng.profile(prof)
    for reps in (10, 50):
        for _ in range(5): eng.spectral_device(E.FMT_F32, buf.data_ptr(), so, sl, d.data_ptr(), stream=st.cuda_stream)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(reps): eng.spectral_device(E.FMT_F32, buf.data_ptr(), so, sl, d.data_ptr(), stream=st.cuda_stream)
        ev1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print("prof", prof, "reps", reps, "gpu ms/pass %.4f" % (ev0.elapsed_time(ev1) / reps), "host enqueue ms/pass %.4f" % ((t1 - t0) * 1e3 / reps))
    if prof: print(eng.profile_read())
