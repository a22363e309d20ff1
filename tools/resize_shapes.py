#!/usr/bin/env python
"""Resize throughput over the image shapes of BASELINE configs 1, 3 and 5 (256^2 .. 4096^2, 4K), both vertical passes (scatter = the default for downscales), ~6 GB of input per shape, outputs 256x256 u8 + f32."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ics_b200  # noqa: E402,F401
from ics_b200 import engine  # noqa: E402

dev = torch.device("cuda", 0)
engine.init(0)
g = torch.Generator(device=dev).manual_seed(3)
for (H, W) in [(256, 256), (512, 512), (1024, 1024), (1080, 1920), (2048, 2048), (2160, 3840), (4096, 4096)]:
    L = H * W * 3
    n = max(148 * 2, min(20000, int(6e9) // L))
    data = torch.empty(n * L, dtype=torch.uint8, device=dev)
    data.random_(0, 256, generator=g)
    off = torch.arange(n, dtype=torch.int64, device=dev) * L
    plan = engine.get_plan(H, W, 256, 256)
    thumb = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device=dev)
    prev = torch.empty((n, 3, 256, 256), dtype=torch.float32, device=dev)
    out = []
    for name, vs_ in {"scatter": "1", "gather ": "0"}.items():
        os.environ["B2_RESIZE_VSCAT"] = vs_
        fn = lambda: plan.run(data, off, thumb=thumb, preview=prev)  # noqa: E731
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        b = n * (L + 256 * 256 * 3 * 5)
        out.append(f"{name} {ms:8.3f} ms {b / ms / 1e6:6.0f} GB/s {n / ms:7.1f} k img/s")
    print(f"{H:>4}x{W:<4} n={n:<6} " + "   ".join(out))
    del data, thumb, prev
