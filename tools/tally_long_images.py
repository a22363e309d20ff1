#!/usr/bin/env python
"""Sorted-mode tally against rows per image: 100 M rows, k = 50.  With more rows per image than a slab (132)
the lanes of a warp share images and their atomics collide; this sweep shows what that costs.  (Round 1 used it
to compare against the MATCH.ANY warp kernel, which lost on every line and was removed:
profiles/r1_tally_history.md.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ics_b200  # noqa: E402,F401
from ics_b200 import engine  # noqa: E402

dev = torch.device("cuda", 0)
engine.init(0)
rows, k = 100_000_000, 50
g = torch.Generator(device=dev).manual_seed(1)
uni = torch.randint(0, k, (rows,), device=dev, generator=g, dtype=torch.int64).to(torch.uint8)
act = (torch.rand(rows, device=dev, generator=g) < 0.95).to(torch.uint8)
for per in (100, 132, 300, 1000, 10_000, 1_000_000):
    N = (rows + per - 1) // per
    img = (torch.arange(rows, device=dev, dtype=torch.int64) // per).to(torch.int32)
    for skew in (0.0, 0.7):                       # share of an image's rows that carry its "true" class
        cls = uni
        if skew:
            true_cls = torch.randint(0, k, (N,), device=dev, generator=g).to(torch.uint8)
            pick = torch.rand(rows, device=dev, generator=g) < skew
            cls = torch.where(pick, true_cls[img.long()], uni)
        counts = torch.empty((N, k), dtype=torch.int32, device=dev)
        part = torch.empty(k + 7, dtype=torch.int64, device=dev)
        fn = lambda: engine.label_tally_device(img, cls, act, N, k, 0, True, counts, part)  # noqa: E731
        line, ref = [], None
        for inc in ("0", "1"):                    # plain ATOMS.ADD of 0/1 vs predicated ATOMS.POPC.INC
            os.environ["B2_TALLY_INC"] = inc
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            ok = int(part[k + 1].item()) == int(act.sum().item())
            same = True if ref is None else bool(torch.equal(ref, counts))
            ref = counts.clone()
            line.append(f"inc={inc}: {ms:.3f} ms {(6 * rows + 4 * N * k) / ms / 1e6:5.0f} GB/s ok={ok and same}")
        print(f"rows/image {per:>8}  class skew {skew:.1f}:  " + "   ".join(line))
