#!/usr/bin/env python
"""Launch each hot kernel on a full-occupancy input (warm-up + measured), for ncu:

    python tools/profile_kernels.py && \
    ncu --set full --clock-control none --import-source on \
        -k regex:'sha256_lanes|resize_bands|tally_slab' -c 6 -o gpurun_out/prof python tools/profile_kernels.py

Sizes keep device memory small enough for ncu's save/restore between replay passes: the hash
gets 18 944 messages of 1 MiB (one warp per SM sub-partition, same per-block work as 1080p
images), the resize 2 368 images of 1920x1080x3, the tally BASELINE config 4 (100 M rows).
Prints CUDA-event timings of the measured launches (never taken under ncu).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ics_b200  # noqa: E402,F401
from ics_b200 import engine  # noqa: E402


def timed(fn, reps=1):
    """Mean time of `reps` back-to-back launches (after one warm-up).  Short kernels need reps > 1: a single
    launch between two events also measures the host's launch latency."""
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda", 0)
    engine.init(0)
    which = sys.argv[1:] or ["sha", "resize", "tally"]
    g = torch.Generator(device=dev).manual_seed(1)
    if "sha" in which:
        n, L = [int(x) for x in os.environ.get("B2_PROF_SHA", "18944,1048576").split(",")]
        data = torch.empty(n * L, dtype=torch.uint8, device=dev)
        data.random_(0, 256, generator=g)
        off = torch.arange(n, dtype=torch.int64, device=dev) * L
        ln = torch.full((n,), L, dtype=torch.int64, device=dev)
        out = torch.empty((n, 32), dtype=torch.uint8, device=dev)
        ms = timed(lambda: engine.sha256_device(data, off, ln, None, out))
        print(f"sha256: {n} x {L} B  {ms:.3f} ms  {n * L / ms / 1e6:.1f} GB/s")
        del data
    if "resize" in which:
        n, H, W = 2368, 1080, 1920
        L = H * W * 3
        data = torch.empty(n * L, dtype=torch.uint8, device=dev)
        data.random_(0, 256, generator=g)
        off = torch.arange(n, dtype=torch.int64, device=dev) * L
        plan = engine.get_plan(H, W, 256, 256)
        thumb = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device=dev)
        prev = torch.empty((n, 3, 256, 256), dtype=torch.float32, device=dev)
        ms = timed(lambda: plan.run(data, off, thumb=thumb, preview=prev))
        b = n * (L + 256 * 256 * 3 * 5)
        print(f"resize: {n} x 1080p  {ms:.3f} ms  {b / ms / 1e6:.1f} GB/s")
        del data
    if "tally" in which:
        N, k, r = 1_000_000, 50, 100
        rows = N * r
        img = (torch.arange(rows, device=dev, dtype=torch.int64) // r).to(torch.int32)
        true_cls = torch.randint(0, k, (N,), device=dev, generator=g)
        pick = torch.rand(rows, device=dev, generator=g) < 0.7
        uni = torch.randint(0, k, (rows,), device=dev, generator=g)
        cls = torch.where(pick, true_cls[img.long()], uni).to(torch.uint8)
        act = (torch.rand(rows, device=dev, generator=g) < 0.95).to(torch.uint8)
        counts = torch.empty((N, k), dtype=torch.int32, device=dev)
        part = torch.empty(k + 7, dtype=torch.int64, device=dev)
        reps = int(os.environ.get("B2_PROF_REPS", "20"))
        ms = timed(lambda: engine.label_tally_device(img, cls, act, N, k, 0, True, counts, part), reps=reps)
        b = 6 * rows + 4 * N * k
        print(f"tally: {rows} rows  {ms:.3f} ms  {b / ms / 1e6:.1f} GB/s")
        hist = torch.empty(1024, dtype=torch.int64, device=dev)
        ms = timed(lambda: engine.label_tally_device(img, cls, act, N, k, 0, True, counts, part, hist), reps=reps)
        print(f"tally + agreement histogram: {rows} rows  {ms:.3f} ms  {b / ms / 1e6:.1f} GB/s")


if __name__ == "__main__":
    main()
