#!/usr/bin/env python
"""BASELINE config 5 on ONE of its eight GPUs: 1 250 synthetic 3840x2160x3 images (20 % duplicates by the config's
rule) in page-locked host memory -> digests + dedupe + thumbnails through `b2_ingest_stream_*`, plus the label tally
of this GPU's share of rows.  One batch, nothing to hide its tail behind: the worst case for the hash latency
(one lane needs ~0.5 s for a 24.9 MB image)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ics_b200  # noqa: E402,F401
from ics_b200 import engine  # noqa: E402
from ics_b200.pipeline import IngestPipeline  # noqa: E402

H, W, n, n_unique = 2160, 3840, 1250, 1000
L = H * W * 3
dev = torch.device("cuda", 0)
engine.init(0)
g = torch.Generator(device=dev).manual_seed(5)
host = torch.empty((n, L), dtype=torch.uint8, pin_memory=True)
for lo in range(0, n_unique, 50):
    blk = torch.empty((50, L), dtype=torch.uint8, device=dev)
    blk.random_(0, 256, generator=g)
    host[lo:lo + 50].copy_(blk)
torch.cuda.synchronize()
for i in range(n_unique, n):                                   # images >= n_unique are byte copies (config 5's rule)
    host[i].copy_(host[(i * 2654435761) % n_unique])
for chunk in ([int(x) for x in sys.argv[1:]] or (32, 64, 128)):          # chunk sizes to try
    pipe = IngestPipeline(H, W, n, chunk_images=chunk, want_preview=True)
    pipe.run(host)                                             # warm-up
    t0 = time.perf_counter()
    res = pipe.run(host)
    dt = time.perf_counter() - t0
    print(f"chunk {chunk:>3}: {n} 4K images in {dt * 1e3:.0f} ms = {n / dt:.0f} images/s = {n * L / dt / 1e9:.1f} GB/s of H2D; "
          f"stats {res.stats}")
    pipe.close()
    del pipe
    torch.cuda.empty_cache()
assert res.stats == {"processed": n, "created": n_unique, "updated": n - n_unique}
