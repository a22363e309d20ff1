#!/usr/bin/env python
"""Device dictionary encoding of `classificacoes` rows at config-4 scale: 100 M rows (64-char hex id_img +
16-byte id_opc + ativo = 81 B per row) against 1 M stored digests and 50 options -> the tally's SoA arrays,
then the stable sort by image and the tally itself."""
import os
import sys
import uuid

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ics_b200  # noqa: E402,F401
from ics_b200 import engine, labels  # noqa: E402

dev = torch.device("cuda", 0)
engine.init(0)
n_images, k, rows = 1_000_000, 50, 100_000_000
rng = np.random.default_rng(5)
keys = engine.sort_digests(rng.integers(0, 256, size=(n_images, 32), dtype=np.uint8))
enc = labels.DeviceLabelEncoder([], [uuid.UUID(int=i + 1) for i in range(k)])
enc.d_image_keys = torch.from_numpy(keys).to(dev)
g = torch.Generator(device=dev).manual_seed(2)
hexs = torch.empty((rows, 64), dtype=torch.uint8, device=dev)
for lo in range(0, rows, 10_000_000):                      # rows arrive in table order here, shuffled below
    pick = torch.randint(0, n_images, (10_000_000,), device=dev, generator=g)
    hexs[lo:lo + 10_000_000] = engine.digest_hex_device(enc.d_image_keys[pick].contiguous())
opc = enc.d_option_keys[torch.randint(0, k, (rows,), device=dev, generator=g)].contiguous()
act = (torch.rand(rows, device=dev, generator=g) < 0.95).to(torch.uint8)


def timed(fn, reps=3):
    out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


ms, (img, cls, a, unknown) = timed(lambda: enc.encode_columns(hexs, opc, act, sort=False))
print(f"encode: {ms:.2f} ms  {rows / ms / 1e6:.2f} G rows/s  {rows * 87 / ms / 1e6:.0f} GB/s (81 B in + 6 B out per row)  unknown={unknown.tolist()}")
ms_sort, order = timed(lambda: torch.sort(img, stable=True).indices)
img_s, cls_s, act_s = img[order].contiguous(), cls[order].contiguous(), a[order].contiguous()
print(f"stable sort by image (torch.sort, 100 M int32 keys): {ms_sort:.2f} ms")
ms_t, (counts, part) = timed(lambda: engine.label_tally_device(img_s, cls_s, act_s, n_images, k), reps=10)
ms_u, (counts2, part2) = timed(lambda: engine.label_tally_device(img, cls, a, n_images, k, 0, False), reps=3)
print(f"tally of the encoded rows: sorted-mode {ms_t:.3f} ms, any-order {ms_u:.3f} ms, equal={bool(torch.equal(counts, counts2))}")
