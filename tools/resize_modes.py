#!/usr/bin/env python
"""Time resize_bands_kernel with the gather-form and the scatter-form vertical pass, alone and beside the hash.

    python tools/resize_modes.py [n_images]

Alone: 2 368 images 1920x1080x3 -> 256x256 u8 + f32 (17.06 GB algorithmic), mean of 5 launches.
Beside the hash: 18 944 x 1 MiB messages (one hash warp per SM sub-partition, ~24 ms) on a high-priority stream
while five resize launches run on a second stream: the time of the pair is what the ingest step sees.
CUDA events, never under ncu.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ics_b200  # noqa: E402,F401
from ics_b200 import engine  # noqa: E402

MODES = {"gather": "0", "scatter": "1"}


def set_mode(name):
    os.environ["B2_RESIZE_VSCAT"] = MODES[name]


def main():
    dev = torch.device("cuda", 0)
    engine.init(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
    H, W = 1080, 1920
    L = H * W * 3
    g = torch.Generator(device=dev).manual_seed(1)
    data = torch.empty(n * L, dtype=torch.uint8, device=dev)
    data.random_(0, 256, generator=g)
    off = torch.arange(n, dtype=torch.int64, device=dev) * L
    plan = engine.get_plan(H, W, 256, 256)
    thumb = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device=dev)
    prev = torch.empty((n, 3, 256, 256), dtype=torch.float32, device=dev)
    nb = n * (L + 256 * 256 * 3 * 5)

    hn, hl = 18944, 1 << 20
    hdata = torch.empty(hn * hl, dtype=torch.uint8, device=dev)
    hdata.random_(0, 256, generator=g)
    hoff = torch.arange(hn, dtype=torch.int64, device=dev) * hl
    hlen = torch.full((hn,), hl, dtype=torch.int64, device=dev)
    hout = torch.empty((hn, 32), dtype=torch.uint8, device=dev)
    hstream = torch.cuda.Stream(dev, priority=-1)
    side = torch.cuda.Stream(dev)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    ref = None
    only = os.environ.get("B2_MODES")
    for name in MODES:
        if only and name not in only.split(","):
            continue
        set_mode(name)
        run = lambda: plan.run(data, off, thumb=thumb, preview=prev)  # noqa: E731
        run()
        torch.cuda.synchronize()
        if ref is None:
            ref = (thumb.clone(), prev.clone())
        else:
            assert torch.equal(thumb, ref[0]) and torch.equal(prev, ref[1]), name
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        # beside the hash
        pair = []
        for _ in range(3):
            main = torch.cuda.current_stream()
            e0, e1 = ev(), ev()
            e0.record(main)
            hstream.wait_event(e0)
            side.wait_event(e0)
            with torch.cuda.stream(hstream):
                engine.sha256_device(hdata, hoff, hlen, None, hout)
            with torch.cuda.stream(side):
                for _ in range(5):
                    run()
            main.wait_stream(hstream)
            main.wait_stream(side)
            e1.record(main)
            torch.cuda.synchronize()
            pair.append(e0.elapsed_time(e1))
        print(f"{name:12s} alone {ms:.3f} ms = {nb / ms / 1e6:.0f} GB/s   hash + 5 resize: {min(pair):.2f} ms", flush=True)
    e0, e1 = ev(), ev()
    e0.record()
    engine.sha256_device(hdata, hoff, hlen, None, hout)
    e1.record()
    torch.cuda.synchronize()
    print(f"hash alone {e0.elapsed_time(e1):.2f} ms")


if __name__ == "__main__":
    main()
