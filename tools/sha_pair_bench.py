#!/usr/bin/env python
"""Per-message SHA-256 speed for batches that do not fill the GPU: one lane per message (sha256_lanes_kernel)
against the warp-pair kernel (sha256_pair_kernel).  Device-resident messages, CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ics_b200  # noqa: E402,F401
from ics_b200 import engine  # noqa: E402

dev = torch.device("cuda", 0)
engine.init(0)
g = torch.Generator(device=dev).manual_seed(2)
for n, L in [(50, 2 << 20), (256, 6220800), (120, 50331648), (1250, 24883200), (4096, 1 << 20), (9472, 1 << 20),
             (18944, 1 << 20)]:
    data = torch.empty(n * L, dtype=torch.uint8, device=dev)
    for lo in range(0, n * L, 1 << 30):
        data[lo:lo + (1 << 30)].random_(0, 256, generator=g)
    off = torch.arange(n, dtype=torch.int64, device=dev) * L
    ln = torch.full((n,), L, dtype=torch.int64, device=dev)
    out = {}
    res = []
    for name, pair in (("lanes", "0"), ("pair", "1")):
        os.environ["B2_SHA_PAIR"] = pair
        d = engine.sha256_device(data, off, ln)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d = engine.sha256_device(data, off, ln)
        e1.record()
        torch.cuda.synchronize()
        out[name] = d
        ms = e0.elapsed_time(e1)
        res.append(f"{name} {ms:9.2f} ms {n * L / ms / 1e6:7.1f} GB/s {L / ms / 1e3:6.1f} MB/s per message")
    assert torch.equal(out["lanes"], out["pair"])
    print(f"{n:>6} x {L / 1e6:6.1f} MB ({(n + 31) // 32:>3} warps)  " + "   ".join(res), flush=True)
    del data
