"""Time of distance_nearest_device for row slabs of different heights against 1 M columns (what a rank of a
1/2/4/8-GPU run computes)."""
import sys
import torch
sys.path.insert(0, ".")
import bliss_b200
eng = bliss_b200.Engine(0)
n = 1 << 20
g = torch.Generator(device="cuda"); g.manual_seed(1)
v = torch.randn((n, 4), generator=g, device="cuda") * torch.tensor([8.0, 6.0, 10.0, 12.0], device="cuda")
ts = torch.cuda.Stream()  # a NULL stream would mean "the engine's own" to the C-ABI
torch.cuda.synchronize()
torch.cuda.set_stream(ts)
st = ts.cuda_stream
for rows in (n, n // 2, n // 4, n // 8, n // 32):
    idx = torch.empty(rows, dtype=torch.int32, device="cuda"); dst = torch.empty(rows, dtype=torch.float32, device="cuda")
    eng.distance_nearest_device(v.data_ptr(), n, 0, min(rows, 4096), idx.data_ptr(), dst.data_ptr(), 0, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.distance_nearest_device(v.data_ptr(), n, 0, rows, idx.data_ptr(), dst.data_ptr(), 0, stream=st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rows {rows:8d}: {ms:8.2f} ms  {rows * n / ms / 1e9:7.2f} T pairs/s")
