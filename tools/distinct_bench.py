#!/usr/bin/env python
"""Bulk COUNT(DISTINCT id_img) WHERE id_con = ? AND ativo for every annotator at config-4
scale: 100 M rows sorted by (annotator, image), 10 000 annotators, ~1.2 rows per (annotator, image) pair."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ics_b200  # noqa: E402,F401
from ics_b200 import engine  # noqa: E402

dev = torch.device("cuda", 0)
engine.init(0)
rows, n_ann, n_img = 100_000_000, 10_000, 1_000_000
g = torch.Generator(device=dev).manual_seed(4)
ann = (torch.arange(rows, device=dev, dtype=torch.int64) * n_ann // rows).to(torch.int32)
img = torch.randint(0, n_img, (rows,), device=dev, generator=g, dtype=torch.int32)
key = ann.to(torch.int64) * n_img + img
order = torch.sort(key).indices
ann, img = ann[order].contiguous(), img[order].contiguous()
del key, order
act = (torch.rand(rows, device=dev, generator=g) < 0.95).to(torch.uint8)
fn = lambda: engine.distinct_images_per_annotator_device(ann, img, act, n_ann)  # noqa: E731
out = fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
keep = act.bool()
pairs = torch.unique(ann[keep].to(torch.int64) * n_img + img[keep])
want = torch.bincount(pairs // n_img, minlength=n_ann)
print(f"distinct images per annotator: {ms:.3f} ms  {rows / ms / 1e6:.1f} G rows/s  {9 * rows / ms / 1e6:.0f} GB/s  "
      f"equal={bool(torch.equal(out.to(torch.int64), want))}")
