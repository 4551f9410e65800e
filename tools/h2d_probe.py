"""Pinned host-to-device copy rate of this box for 360 MB (the values of BASELINE config 2): the floor of the host-buffer call.
    python tools/h2d_probe.py        # B200 box of this round: 53.6 - 55.0 GB/s"""
import torch, time
x = torch.empty(360_000_000 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("pinned H2D 360 MB: %.2f ms = %.1f GB/s" % (ms, 0.36 / ms * 1e3))
