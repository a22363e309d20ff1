#!/usr/bin/env python
"""Raw host->device and device->host copy rates from page-locked memory on this box: the ceiling of the
end-to-end ingest rate (6.22 MB of pixels in and 0.98 MB of thumbnail + preview out per 1080p image)."""
import torch

dev = torch.device("cuda", 0)
n = 8 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: d.copy_(h, non_blocking=True))
print(f"H2D alone: {n / ms / 1e6:.1f} GB/s")
ms = timed(lambda: h.copy_(d, non_blocking=True))
print(f"D2H alone: {n / ms / 1e6:.1f} GB/s")
h2 = torch.empty(n // 6, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n // 6, dtype=torch.uint8, device=dev)


def both():
    e = torch.cuda.Event()
    e.record()
    s_in.wait_event(e)
    s_out.wait_event(e)
    with torch.cuda.stream(s_in):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s_out):
        h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_in)
    torch.cuda.current_stream().wait_stream(s_out)


ms = timed(both)
print(f"H2D with a concurrent D2H of 1/6 the size (the ingest mix): {n / ms / 1e6:.1f} GB/s H2D "
      f"= {n / ms / 1e6 / 6.2208e-3 / 1e3:.2f} k 1080p images/s")
