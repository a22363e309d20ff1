#!/usr/bin/env python
"""BASELINE config 3 on ONE GPU: a slice of mixed-size images (256^2 .. 4096^2, 1080p and 4K among them), resident
in HBM -> SHA-256 of every image (one launch, lanes ordered by length), one resize launch per shape class writing
into the batch's output slots, dedupe.  Hash and resizes run on two streams as in the ingest step of bench.py.

    python tools/config3_mixed.py [images_per_shape]

What bounds it: one SHA-256 lane hashes ~48 MB/s, so the 50 MB images need ~1.05 s whatever the batch holds; the
batch would have to be ~870 GB to hide that behind other lanes' work.  Sampled images are checked against hashlib
and Pillow (not timed).
"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402

import ics_b200  # noqa: E402,F401
from ics_b200 import engine  # noqa: E402

SHAPES = [(256, 256), (512, 512), (1024, 1024), (1080, 1920), (2048, 2048), (2160, 3840), (4096, 4096)]


def main():
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    dev = torch.device("cuda", 0)
    engine.init(0)
    g = torch.Generator(device=dev).manual_seed(3)
    # images interleaved by shape (as a directory listing would be), each start 16-byte aligned
    shapes = [SHAPES[i % len(SHAPES)] for i in range(per * len(SHAPES))]
    lengths = torch.tensor([h * w * 3 for h, w in shapes], dtype=torch.int64)
    offsets = torch.zeros_like(lengths)
    offsets[1:] = torch.cumsum((lengths[:-1] + 15) // 16 * 16, 0)
    total = int(offsets[-1] + lengths[-1])
    n = len(shapes)
    data = torch.empty(total, dtype=torch.uint8, device=dev)
    for lo in range(0, total, 1 << 30):
        data[lo:lo + (1 << 30)].random_(0, 256, generator=g)
    off_d, len_d = offsets.to(dev), lengths.to(dev)
    order = torch.argsort(len_d, descending=True).to(torch.int32)
    digests = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    thumbs = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device=dev)
    previews = torch.empty((n, 3, 256, 256), dtype=torch.float32, device=dev)
    classes = []
    for (h, w) in SHAPES:
        idx = torch.tensor([i for i, s in enumerate(shapes) if s == (h, w)], dtype=torch.int64)
        classes.append((engine.get_plan(h, w, 256, 256), off_d[idx.to(dev)].contiguous(), idx.to(torch.int32).to(dev)))
    hstream = torch.cuda.Stream(dev, priority=-1)
    side = torch.cuda.Stream(dev)

    def step():
        main_s = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main_s)
        hstream.wait_event(fork)
        side.wait_event(fork)
        with torch.cuda.stream(hstream):
            engine.sha256_device(data, off_d, len_d, order, digests)
        with torch.cuda.stream(side):
            for plan, offs, slots in classes:
                plan.run(data, offs, thumb=thumbs, preview=previews, out_slot=slots)
        main_s.wait_stream(hstream)
        main_s.wait_stream(side)
        return engine.dedupe_device(digests)

    def timed(fn, reps=2):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms = timed(step)
    ms_hash = timed(lambda: engine.sha256_device(data, off_d, len_d, order, digests))

    def resizes():
        for plan, offs, slots in classes:
            plan.run(data, offs, thumb=thumbs, preview=previews, out_slot=slots)
    ms_resize = timed(resizes)
    out_bytes = n * 256 * 256 * 3 * 5
    print(f"config 3 slice: {n} images, {total / 1e9:.1f} GB ({per} of each of {len(SHAPES)} shapes)")
    print(f"  step (hash || {len(SHAPES)} resizes, dedupe): {ms:.1f} ms = {n / ms:.2f} k images/s = {total / ms / 1e6:.0f} GB/s of pixels")
    print(f"  hash alone   {ms_hash:.1f} ms = {total / ms_hash / 1e6:.0f} GB/s (longest lane: {max(lengths) / 1e6:.1f} MB)")
    print(f"  resize alone {ms_resize:.1f} ms = {(total + out_bytes) / ms_resize / 1e6:.0f} GB/s")
    for i in (0, 3, 6, n - 1):                              # parity of sampled images, every shape class touched
        h, w = shapes[i]
        host = data[int(offsets[i]):int(offsets[i] + lengths[i])].cpu().numpy()
        assert bytes(digests[i].cpu().numpy()).hex() == hashlib.sha256(host.tobytes()).hexdigest(), i
        want = np.asarray(Image.fromarray(host.reshape(h, w, 3), "RGB").resize((256, 256), Image.BILINEAR))
        assert np.array_equal(thumbs[i].cpu().numpy(), want), i
    print("  sampled digests == hashlib, sampled thumbnails == Pillow")


if __name__ == "__main__":
    main()
