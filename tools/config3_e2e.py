#!/usr/bin/env python
"""BASELINE config 3 end to end on ONE GPU: a listing of mixed-size images (256^2 .. 4096^2, 1080p and 4K among
them) in page-locked HOST memory -> MixedShapeIngest (one native ingest stream per shape class, all in flight at
once, one dedupe over the listing) -> digests, dedupe flags, thumbnails and previews in host memory.

    python tools/config3_e2e.py [images_per_shape]

A single listing has nothing to hide its tail behind: the 50 MB images need ~1.15 s in their hash lanes whatever else
happens (tools/config3_mixed.py).  With two listings in flight the tail of one hides under the copies of the next.
Thumbnails and previews stay in the per-shape page-locked blocks (MixedResult.thumb(i) addresses them by listing
position); only the 32-byte digests are put into listing order on the host."""
import hashlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402

import ics_b200  # noqa: E402,F401
from ics_b200 import engine  # noqa: E402
from ics_b200.pipeline import MixedShapeIngest  # noqa: E402

SHAPES = [(256, 256), (512, 512), (1024, 1024), (1080, 1920), (2048, 2048), (2160, 3840), (4096, 4096)]


def main():
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    dev = torch.device("cuda", 0)
    engine.init(0)
    g = torch.Generator(device=dev).manual_seed(3)
    n = per * len(SHAPES)
    groups, total = {}, 0
    for k, (h, w) in enumerate(SHAPES):
        L = h * w * 3
        host = torch.empty((per, L), dtype=torch.uint8, pin_memory=True)
        for lo in range(0, per, 16):
            blk = torch.empty((min(16, per - lo), L), dtype=torch.uint8, device=dev)
            blk.random_(0, 256, generator=g)
            host[lo:lo + blk.shape[0]].copy_(blk)
        if per > 4:
            host[per - 1].copy_(host[0])                              # one duplicate per shape class
        groups[(h, w)] = (host, list(range(k, n, len(SHAPES))))       # interleaved listing
        total += per * L
    torch.cuda.synchronize()
    mixed = [MixedShapeIngest({s: per for s in SHAPES}, chunk_bytes=1 << 30) for _ in range(2)]
    mixed[0].run(groups)                                              # warm-up
    mixed[1].run(groups)
    one = None
    for _ in range(3):
        t0 = time.perf_counter()
        res = mixed[0].run(groups)
        dt = time.perf_counter() - t0
        one = dt if one is None else min(one, dt)
    for shape in SHAPES:                                              # every shape class on its own
        t0 = time.perf_counter()
        mixed[0].pipes[shape].run(groups[shape][0])
        dt = time.perf_counter() - t0
        print(f"    {shape[0]:>4}x{shape[1]:<4} alone: {per} images in {dt * 1e3:7.1f} ms = {per * shape[0] * shape[1] * 3 / dt / 1e9:5.1f} GB/s")
    reps = 6                                                          # two listings in flight: submit i+1, then result i
    t0 = time.perf_counter()
    mixed[0].submit(groups)
    for i in range(1, reps):
        mixed[i % 2].submit(groups)
        res = mixed[(i - 1) % 2].result()
    res = mixed[(reps - 1) % 2].result()
    piped = (time.perf_counter() - t0) / reps
    print(f"config 3 listing: {n} images, {total / 1e9:.1f} GB in page-locked host memory ({per} of each of {len(SHAPES)} shapes)")
    print(f"  one listing alone      {one * 1e3:.0f} ms = {n / one / 1e3:.2f} k images/s = {total / one / 1e9:.1f} GB/s of H2D")
    print(f"  two listings in flight {piped * 1e3:.0f} ms per listing = {n / piped / 1e3:.2f} k images/s = "
          f"{total / piped / 1e9:.1f} GB/s of H2D; D2H {res.d2h_bytes / 1e9:.2f} GB; stats {res.stats}")
    assert res.stats["created"] == n - (len(SHAPES) if per > 4 else 0)
    for shape, (host, pos) in groups.items():                         # sampled parity, every shape class
        i = 1 % per
        buf = host[i].numpy()
        assert bytes(res.digests[pos[i]]) == hashlib.sha256(buf.tobytes()).digest()
        want = np.asarray(Image.fromarray(buf.reshape(*shape, 3), "RGB").resize((256, 256), Image.BILINEAR))
        assert np.array_equal(res.thumb(pos[i]), want)
    print("  sampled digests == hashlib, sampled thumbnails == Pillow")
    for m in mixed:
        m.close()


if __name__ == "__main__":
    main()
