import sys, os
sys.path.insert(0, os.getcwd())
import torch, ics_b200
from ics_b200 import engine
dev = torch.device("cuda", 0); engine.init(0)
N, k, r = 1_000_000, 50, 100
rows = N * r
g = torch.Generator(device=dev).manual_seed(1)
img = (torch.arange(rows, device=dev, dtype=torch.int64) // r).to(torch.int32)
cls = torch.randint(0, k, (rows,), device=dev, generator=g, dtype=torch.int64).to(torch.uint8)
act = (torch.rand(rows, device=dev, generator=g) < 0.95).to(torch.uint8)
perm = torch.randperm(rows, device=dev, generator=g)
img_s, cls_s, act_s = img[perm].contiguous(), cls[perm].contiguous(), act[perm].contiguous()
counts = torch.empty((N, k), dtype=torch.int32, device=dev); part = torch.empty(k + 7, dtype=torch.int64, device=dev)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timed(lambda: engine.label_tally_device(img_s, cls_s, act_s, N, k, 0, False, counts, part))
print(f"scatter (shuffled rows): {ms:.3f} ms  {rows/ms/1e6:.2f} G rows/s")
ref = counts.clone()
ms = timed(lambda: engine.label_tally_device(img, cls, act, N, k, 0, True, counts, part), 20)
print(f"sorted: {ms:.3f} ms; equal={bool(torch.equal(ref, counts))}")
ms = timed(lambda: torch.sort(img_s))
print(f"torch.sort of 100M int32 keys: {ms:.3f} ms")
