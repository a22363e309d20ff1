#!/usr/bin/env python
"""Benchmark of the ingest + label-aggregation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference] [--only a,b,...]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic input.

  `value` (images/s)   BASELINE config 2 shape, inputs resident in HBM: `images_per_gpu` synthetic 1920x1080x3 images
      -> 256x256 uint8 thumbnail + float32 CHW preview (HBM bound) on a second stream beside the SHA-256 of every image
      (INT32-ALU bound: one warp per SM sub-partition), then the dedupe decision over ALL ranks' digests
      (`b2_dedupe_global`: NCCL all-gather inside the C ABI + the listing-position rule).  20 % of the images are byte
      copies of earlier ones of the GLOBAL listing, so at N > 1 most duplicates live on another rank than their first
      occurrence; `created` and every first-occurrence position are checked at every N.
  `e2e` (images/s)     the same metric through the host-buffer C ABI (`b2_ingest_ring_*`): page-locked HOST buffers in,
      host results out, H2D and D2H inside the timed region; the same images per step at every N.  `e2e.raw_h2d` is the
      plain-copy ceiling measured in the same run (all ranks copying at once).
  `configs`            the other BASELINE configs, each with parity checks and its own e2e / roofline figures:
      c1  1 000 x 512^2 images + 10 k label rows (the CPU-runnable case), both arms timed;
      c3  mixed 256^2..4096^2 listing sharded by bytes (`dist.shard_by_bytes`) through the ring, global dedupe;
      c4  ONE 100 M-row table (1 M images, k = 50), same rows at every N, sharded by image range, tally + NCCL
          all-reduce captured in a CUDA graph; all k+7 partials + the agreement histogram equal NumPy's on rank 0,
          so kappa is bit-identical at 1/2/4/8 GPUs;
      c5  3840x2160 images (1 250 per GPU = 10 k on 8), 20 % global duplicates, thumbnails + previews + label tally,
          end to end from page-locked host memory.
  `labels` (rows/s)    config 4 shape PER GPU (100 M rows each: the HBM-roofline point of the tally kernel).

`roofline` is for the dominant kernel of the ingest step (the hash: bound by the INT32 ALU pipe, its HBM fraction is
reported beside it); `kernels` lists every kernel's own roofline.  `cpu_baseline` / `--impl reference` time the oracle
(hashlib + Pillow + NumPy: the reference's own host libraries) on the box's host cores.  Inputs are larger than L2 (no
flush needed): stated in `config`.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# The streaming ingest overlaps tens of long-lived hash kernels with copies and resizes: it needs the 32 hardware work
# queues (default 8; measured on the config-3 stream: 8 -> 13.8 GB/s, 32 -> PCIe bound).  libb2ingest.so asks for them
# when it is loaded; said here too because the variable is read when the CUDA context is created.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

IMG_H, IMG_W = 1080, 1920
IMG_BYTES = IMG_H * IMG_W * 3                     # 6 220 800
OUT = 256
THUMB_BYTES = OUT * OUT * 3
PREVIEW_BYTES = OUT * OUT * 3 * 4
LABEL_IMAGES, LABEL_K, LABEL_RATERS = 1_000_000, 50, 100
FULL_IMAGES_PER_GPU = 148 * 4 * 32                # one hash warp per SM sub-partition: 18 944 images = 117.8 GB
SEED = 0xB200
LABEL_SEED = 0xF1E155
MULT = 2654435761                                 # duplicate rule of SURVEY 8(d): g >= U copies (g * MULT) mod U
C4_BLOCK_IMAGES = 1000                            # config-4 rows are generated per block of images (any rank, any N)
C5_H, C5_W = 2160, 3840
C5_PER_GPU = 1250


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0: float, t1: float):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.25:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- synthetic label tables (host, NumPy)
def c4_block(block: int, k: int = LABEL_K, raters: int = LABEL_RATERS, images: int = C4_BLOCK_IMAGES):
    """Rows of images [block*images, (block+1)*images) of THE config-4 table: a counter-based generator keyed by the
    block, so every rank builds exactly its slice of the same 100 M-row table whatever the GPU count (SURVEY 8(d):
    image_idx = row // 100, the image's "true" class w.p. 0.7 else uniform, active w.p. 0.95)."""
    import numpy as np
    rng = np.random.Generator(np.random.Philox(key=[LABEL_SEED, block]))
    rows = images * raters
    true_cls = rng.integers(0, k, size=images, dtype=np.int64)
    pick = rng.random(rows, dtype=np.float32) < 0.7
    uni = rng.integers(0, k, size=rows, dtype=np.uint8)
    cls = np.where(pick, np.repeat(true_cls, raters).astype(np.uint8), uni)
    act = (rng.random(rows, dtype=np.float32) < 0.95).astype(np.uint8)
    img = (np.arange(rows, dtype=np.int64) // raters + block * images).astype(np.int32)
    return img, cls, act


def c4_table(block_lo: int, block_hi: int, threads: int = 8):
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        parts = list(ex.map(c4_block, range(block_lo, block_hi)))
    return (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]),
            np.concatenate([p[2] for p in parts]))


def numpy_tally_partials(img, cls, act, image_base: int, n_images: int, k: int):
    """NumPy statement of b2_label_tally's integer outputs (bincount + sums): the checker of config 4."""
    import numpy as np
    keep = act != 0
    flat = (img[keep].astype(np.int64) - image_base) * k + cls[keep]
    counts = np.bincount(flat, minlength=n_images * k).reshape(n_images, k)
    n_i = counts.sum(axis=1)
    s2_i = (counts * counts).sum(axis=1)
    part = list(counts.sum(axis=0)) + [int(s2_i.sum()), int(n_i.sum()), int((n_i >= 1).sum()), int((n_i >= 2).sum()),
                                       int((n_i * (n_i - 1)).sum()), int(img.size), 0]
    hist = np.zeros(1024, dtype=np.int64)
    m = (n_i >= 2) & (n_i < 1024)
    np.add.at(hist, n_i[m], (s2_i - n_i)[m])
    hist[0] = int((n_i >= 1024).sum())
    return np.array(part, dtype=np.int64), hist


# ------------------------------------------------------------------------------- CPU arm
def cpu_ingest_sample(n_images: int, threads: int, h: int = IMG_H, w: int = IMG_W, seed: int = SEED):
    """The oracle's ingest path (hashlib.sha256 + dict dedupe + Pillow BILINEAR + float32 preview)
    on `n_images` synthetic h x w images with `threads` host threads.  Returns images/s."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import dedupe_batch, preview_f32, sha256_hex, thumbnail_u8

    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(min(n_images, 16), h, w, 3), dtype=np.uint8)
    imgs = [base[i % len(base)] for i in range(n_images)]

    def one(im):
        hx = sha256_hex(im.data)                     # the "file bytes" = the raw RGB buffer
        t = thumbnail_u8(im, OUT, OUT)
        return hx, t, preview_f32(t)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        out = list(ex.map(one, imgs))
    dedupe_batch([o[0] for o in out])
    dt = time.perf_counter() - t0
    return n_images / dt, dt


def cpu_label_sample(rows: int, threads: int, raters: int = LABEL_RATERS):
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import fleiss_kappa, fleiss_partials, label_tally, synth_label_rows

    n_images = rows // raters
    img, cls, act = synth_label_rows(n_images, LABEL_K, raters)
    shards = max(1, min(threads, 16))
    bounds = [(n_images * s // shards, n_images * (s + 1) // shards) for s in range(shards)]

    def one(b):
        lo, hi = b
        r0, r1 = lo * raters, hi * raters
        return label_tally(img[r0:r1] - lo, cls[r0:r1], act[r0:r1], hi - lo, LABEL_K)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=shards) as ex:
        parts = list(ex.map(one, bounds))
    counts = np.concatenate(parts)
    p = fleiss_partials(counts)
    fleiss_kappa(p["class_totals"], p["S2"], p["R"], n_images, raters)
    dt = time.perf_counter() - t0
    return rows / dt, dt


def cpu_config1(threads: int):
    """BASELINE config 1 in full on the host cores: 1 000 x 512^2 images (hash + thumbnail) + 10 k label rows."""
    v_img, dt_img = cpu_ingest_sample(1000, threads, 512, 512)
    v_lab, dt_lab = cpu_label_sample(10_000, 1, raters=10)
    return {"images_per_s": v_img, "ingest_s": dt_img, "label_rows_per_s": v_lab, "label_s": dt_lab, "threads": threads}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (the oracle port:
    the reference is pure Python on hashlib + Pillow and cannot be imported here) on all host
    cores; each step is a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_img = 64 * cores                                        # bounded sample: ~1.5 s per step on every core
    for _ in range(args.warmup):
        cpu_ingest_sample(min(n_img, cores), cores)
    tot_img = tot_t = 0.0
    for _ in range(args.steps):
        v, dt = cpu_ingest_sample(n_img, cores)
        tot_img += n_img
        tot_t += dt
    value = tot_img / tot_t
    lab_v, lab_dt = cpu_label_sample(10_000_000, cores)
    sample = f"{n_img} synthetic 1920x1080x3 images per step x {args.steps} steps; labels: 10M rows once"
    line = {
        "impl": "reference", "metric": "ingest images/s", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32", "data": "synthetic",
        # the graft arm's config (the workload both arms are measured on); what this arm actually ran per step — a bounded
        # sample of it in host memory — is stated in cpu_baseline.sample
        "config": workload_config(args.gpus, args.images_per_gpu or FULL_IMAGES_PER_GPU, RESIDENT_INPUTS),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "labels": {"value": lab_v, "unit": "rows/s", "cores": min(cores, 16), "sample": "10M rows, N=100k, k=50"},
        "configs": {"c1": cpu_config1(cores)},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


RESIDENT_INPUTS = ("resident in HBM, generated on device (content keyed by listing position), "
                   "20 % duplicates of the GLOBAL listing")


def workload_config(n_gpus, images_per_gpu, residency):
    return {
        "workload": "BASELINE config 2 shape: synthetic 1920x1080x3 RGB images, SHA-256 + dedupe + 256x256 "
                    "uint8 thumbnail + float32 CHW preview",
        "images_per_gpu_per_step": images_per_gpu, "image_bytes": IMG_BYTES, "n_gpus": n_gpus,
        "inputs": residency, "l2": "inputs larger than L2 (no flush needed)",
        "labels_workload": f"BASELINE config 4 shape per GPU: {LABEL_IMAGES * LABEL_RATERS} rows, "
                           f"{LABEL_IMAGES} images, k={LABEL_K}, clustered by image",
    }


# ------------------------------------------------------------------------------- GPU arm
def run_graft(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from PIL import Image as PILImage

    import ics_b200
    from ics_b200 import engine, hostapi
    from ics_b200 import dist as b2dist
    from ics_b200 import labels as b2labels
    from ics_b200._lib import B2_AGREE_BINS
    from ics_b200.pipeline import IngestRing

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    only = set(x for x in args.only.split(",") if x)

    def want(name):
        return not only or name in only

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    engine.init(local_rank)
    comm = b2dist.comm()                                     # b2_comm (NCCL inside the C ABI); None at N = 1
    hbm_peak, peak_src = peaks()
    # host side of the end-to-end runs: page-locked buffers on the NUMA node next to this rank's GPU
    # (undone before the CPU baseline, which uses every host core)
    affinity0 = os.sched_getaffinity(0)
    numa_node = engine.bind_host_thread_to_gpu_node(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x: float, op="max") -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}[op])
        return float(t.item())

    def all_ok(ok: bool) -> bool:
        return reduce_ranks(1.0 if ok else 0.0, "min") > 0.5

    gen = torch.Generator(device=dev)

    def fill_image(dst: "torch.Tensor", content_id: int):
        """Image content = i.i.d. uniform bytes keyed by (SEED, content id): any rank regenerates any image."""
        gen.manual_seed((SEED << 32) + int(content_id))
        dst.random_(0, 256, generator=gen)

    def source_of(g, n_unique):
        """Content id of global listing position g under the duplicate rule."""
        return np.where(g < n_unique, g, (g * MULT) % n_unique)

    launches = {"n": 0}
    per_step = {"ms": []}

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        launches["n"] = 0
        t0 = time.time()
        e0.record()
        out = None
        for i in range(steps):
            out = fn()
            marks[i].record()
        e1.record()
        barrier()
        t1 = time.time()
        per_step["ms"] = [round(a.elapsed_time(b), 3) for a, b in zip([e0] + marks[:-1], marks)]
        return reduce_ranks(e0.elapsed_time(e1)), out, (t0, t1)

    def kernel_ms(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    warmup = max(3, args.warmup)                             # timing rule: at least three untimed steps
    out = {}                                                 # pieces of the JSON line

    # =========================================================================================================
    # (1) config 2 shape, resident in HBM: hash || resize, then the global dedupe decision
    # =========================================================================================================
    if want("resident"):
        free_b, _ = torch.cuda.mem_get_info()
        n_img = args.images_per_gpu or FULL_IMAGES_PER_GPU
        budget = int(free_b * 0.80) - 4 * (1 << 30)
        n_img = max(32, min(n_img, budget // IMG_BYTES) // 32 * 32)
        n_img = int(reduce_ranks(float(n_img), "min"))
        total = n_img * world
        n_unique = total - total // 5                        # 20 % duplicates of the GLOBAL listing (config 5's rule)
        g_np = np.arange(n_img, dtype=np.int64) * world + rank          # listing position of local image i
        src_np = source_of(g_np, n_unique)
        data = torch.empty((n_img, IMG_BYTES), dtype=torch.uint8, device=dev)
        for i in range(n_img):
            fill_image(data[i], src_np[i])
        offsets = torch.arange(n_img, dtype=torch.int64, device=dev) * IMG_BYTES
        lengths = torch.full((n_img,), IMG_BYTES, dtype=torch.int64, device=dev)
        flat = data.view(-1)
        digests = torch.empty((n_img, 32), dtype=torch.uint8, device=dev)
        thumbs = torch.empty((n_img, OUT, OUT, 3), dtype=torch.uint8, device=dev)
        previews = torch.empty((n_img, 3, OUT, OUT), dtype=torch.float32, device=dev)
        plan = engine.get_plan(IMG_H, IMG_W, OUT, OUT, local_rank)
        seq = torch.from_numpy(g_np.astype(np.int32)).to(dev)
        side = torch.cuda.Stream(dev)                        # resize (default = lowest priority)
        hstream = torch.cuda.Stream(dev, priority=-1)        # hash (higher priority: its CTAs are placed first)

        def ingest_step():
            main = torch.cuda.current_stream()
            if args.overlap:
                # The hash keeps one warp per SM sub-partition busy on the INT32 ALU pipe for the whole step and
                # leaves HBM and most issue slots idle: the HBM-bound resize runs beside it on a second stream.
                fork = torch.cuda.Event()
                fork.record(main)
                hstream.wait_event(fork)
                side.wait_event(fork)
                with torch.cuda.stream(hstream):
                    engine.sha256_device(flat, offsets, lengths, None, digests)
                with torch.cuda.stream(side):
                    plan.run(flat, offsets, thumb=thumbs, preview=previews)
                main.wait_stream(hstream)
                main.wait_stream(side)
            else:
                plan.run(flat, offsets, thumb=thumbs, preview=previews)
                engine.sha256_device(flat, offsets, lengths, None, digests)
            res = b2dist.global_dedupe(digests, seq, n_max=n_img)
            launches["n"] += 4 if world == 1 else 5          # hash, resize, dedupe insert + resolve (+ slice at N > 1)
            return res

        ms_ingest, (is_new, first_seq, last_seq, counts), window = timed(ingest_step, args.steps, warmup)
        ingest_launches = launches["n"]
        ingest_step_ms = list(per_step["ms"])
        counts_h = counts.cpu().tolist()
        value = total * args.steps / (ms_ingest / 1e3)
        clocks = sampler.summary(*window) if rank == 0 else None

        ms_sha = kernel_ms(lambda: engine.sha256_device(flat, offsets, lengths, None, digests))
        ms_resize = kernel_ms(lambda: plan.run(flat, offsets, thumb=thumbs, preview=previews))
        ms_dedupe = kernel_ms(lambda: engine.dedupe_device(digests))

        # parity of the timed outputs: (a) every first-occurrence position and every is_new flag of this rank against
        # the duplicate rule (what the sequential loop of webdav_sync.py:311-400 yields on this listing: the first
        # occurrence of a copy is its source, which precedes it); (b) global counts; (c) sampled digests / thumbnails
        # against the reference's own host libraries called directly (hashlib, Pillow).
        want_first = torch.from_numpy(src_np).to(dev)
        ok_first = bool(torch.equal(first_seq, want_first))
        ok_new = bool(torch.equal(is_new.bool(), torch.from_numpy(g_np < n_unique).to(dev)))
        ok_counts = counts_h == [total, n_unique, total - n_unique]
        idx = [0, n_img // 2, n_img - 1]
        ok_sample = True
        hexes = engine.hex_strings(engine.digest_hex_device(digests[idx].contiguous()))
        for j, i in enumerate(idx):
            host = data[i].cpu().numpy()
            ok_sample &= hexes[j] == hashlib.sha256(host.tobytes()).hexdigest()
            wantt = np.asarray(PILImage.fromarray(host.reshape(IMG_H, IMG_W, 3), "RGB").resize((OUT, OUT), PILImage.BILINEAR))
            ok_sample &= bool(np.array_equal(thumbs[i].cpu().numpy(), wantt))
        parity = {"ok": all_ok(ok_first and ok_new and ok_counts and ok_sample),
                  "first_occurrence_positions_checked": total, "is_new_flags_checked": total,
                  "cross_rank_duplicates": int(((g_np >= n_unique) & (src_np % world != rank)).sum()) if world > 1 else 0,
                  "dedupe_counts": counts_h, "expected_counts": [total, n_unique, total - n_unique],
                  "sampled_images_vs_hashlib_pillow": len(idx) * world}
        out["resident"] = dict(value=value, ms=ms_ingest, n_img=n_img, launches=ingest_launches, step_ms=ingest_step_ms,
                               clocks=clocks, ms_sha=ms_sha, ms_resize=ms_resize, ms_dedupe=ms_dedupe, parity=parity)
        del data, flat, thumbs, previews, digests
        torch.cuda.empty_cache()

    # =========================================================================================================
    # (2) end to end through the ring: page-locked host buffers, configs 2, 3, 5 (and 1)
    # =========================================================================================================
    ring = None
    e2e_any = any(want(x) for x in ("e2e", "c3", "c5", "c1"))
    if e2e_any:
        with open("/proc/meminfo") as f:
            mem = {ln.split(":")[0]: int(ln.split()[1]) * 1024 for ln in f if ":" in ln}
        host_avail = mem.get("MemAvailable", 64 << 30) // world      # this rank's share of free host memory
        free_b, _ = torch.cuda.mem_get_info()
        ring_bytes = min(args.ring_gb << 30, int(free_b * 0.55))
        max_imgs = max(args.e2e_images, C5_PER_GPU, args.c3_images, 1000)
        ring = IngestRing(ring_bytes=ring_bytes, chunk_bytes=args.chunk_mb << 20, max_listings=args.listings,
                          max_images=max_imgs, out_h=OUT, out_w=OUT, want_preview=True, device=local_rank)

    def make_host_listing(shapes, content_ids):
        """Page-locked host buffer holding the listing's images back to back (generated on the device image by
        image, copied out).  Returns (buffer, per-image uint8 views)."""
        sizes = [h * w * 3 for h, w in shapes]
        buf = hostapi.pinned_empty((sum(sizes),), np.uint8)
        tbuf = torch.from_numpy(buf)
        big = torch.empty(max(sizes), dtype=torch.uint8, device=dev)
        o = 0
        views = []
        for sz, cid in zip(sizes, content_ids):
            fill_image(big[:sz], cid)
            tbuf[o:o + sz].copy_(big[:sz], non_blocking=True)
            views.append(buf[o:o + sz])
            o += sz
        torch.cuda.synchronize()
        del big
        return buf, views

    def stream_listings(views, shapes, n_listings, inflight):
        """Submit the same listing n_listings times, `inflight` at once; returns (total ms incl. fill and drain, the
        completion time of every listing, last result)."""
        hw = np.asarray(shapes, dtype=np.uint32)
        ring.prepare(len(views), inflight)                   # page-locked result buffers: allocated outside the timed region
        pending, done_t, res = [], [], None
        barrier()
        t0 = time.perf_counter()
        for i in range(n_listings):
            if len(pending) == inflight:
                res = ring.result(pending.pop(0))
                done_t.append(time.perf_counter())
            pending.append(ring.submit(views, hw))
        while pending:
            res = ring.result(pending.pop(0))
            done_t.append(time.perf_counter())
        t1 = time.perf_counter()
        return 1e3 * (t1 - t0), done_t, res

    def check_listing(res, views, shapes, sample):
        ok = True
        for i in sample:
            ok &= bytes(res.digests[i]) == hashlib.sha256(views[i].tobytes()).digest()
            h, w = shapes[i]
            wantt = np.asarray(PILImage.fromarray(views[i].reshape(h, w, 3), "RGB").resize((OUT, OUT), PILImage.BILINEAR))
            ok &= bool(np.array_equal(res.thumbs[i], wantt))
            ok &= bool(np.allclose(res.previews[i], wantt.astype(np.float32).transpose(2, 0, 1) / 255.0, rtol=1e-5, atol=1e-7))
        return bool(ok)

    def global_counts(res, positions, n_max, n_unique_global, total_global):
        """Cross-rank dedupe decision over the digests the ring produced (b2_dedupe_global) + its check."""
        d = torch.from_numpy(np.ascontiguousarray(res.digests)).to(dev)
        s = torch.from_numpy(positions.astype(np.int32)).to(dev)
        is_new, first, _, counts = b2dist.global_dedupe(d, s, n_max=n_max)
        c = counts.cpu().tolist()
        want_first = torch.from_numpy(source_of(positions, n_unique_global)).to(dev)
        ok = c == [total_global, n_unique_global, total_global - n_unique_global] and bool(torch.equal(first, want_first))
        return c, bool(ok)

    # ---- config 2 shape: the headline e2e
    if want("e2e"):
        e2e_n = args.e2e_images
        shapes2 = [(IMG_H, IMG_W)] * e2e_n
        total2 = e2e_n * world
        nu2 = total2 - total2 // 5
        pos2 = np.arange(e2e_n, dtype=np.int64) * world + rank
        buf2, views2 = make_host_listing(shapes2, source_of(pos2, nu2))
        # raw-copy ceilings: every rank copies its listing H2D at once, nothing else running — alone, and with the
        # D2H share of this workload (thumbnails + previews = 0.98 MB out per 6.22 MB in) going the other way
        n_raw = min(buf2.size, 8 << 30)
        dst = torch.empty(n_raw, dtype=torch.uint8, device=dev)
        src_t = torch.from_numpy(buf2)[:n_raw]
        n_out = int(n_raw * (THUMB_BYTES + PREVIEW_BYTES) / IMG_BYTES)
        d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
        h_out = torch.from_numpy(hostapi.pinned_empty((n_out,), np.uint8))
        s_out = torch.cuda.Stream(dev)

        def raw_copy(with_d2h, reps=6):
            for r in range(2 + reps):
                if r == 2:
                    barrier()
                    ev0 = torch.cuda.Event(enable_timing=True)
                    ev0.record()
                    s_out.wait_event(ev0)
                dst.copy_(src_t, non_blocking=True)
                if with_d2h:
                    with torch.cuda.stream(s_out):
                        h_out.copy_(d_out, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s_out)
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            barrier()
            return n_raw * reps * world / reduce_ranks(ev0.elapsed_time(ev1)) / 1e6

        raw_gbs = raw_copy(False)
        raw_mix_gbs = raw_copy(True)
        del dst, d_out, h_out
        torch.cuda.empty_cache()

        e2e_steps = max(10, args.steps)
        stream_listings(views2, shapes2, 3, min(3, args.listings))       # warm-up
        ms_total, done_t, res = stream_listings(views2, shapes2, e2e_steps, min(3, args.listings))
        ms_e2e = reduce_ranks(ms_total)
        e2e_value = e2e_n * world * e2e_steps / (ms_e2e / 1e3)
        ok2 = check_listing(res, views2, shapes2, [0, e2e_n - 1])
        c2counts, ok2g = global_counts(res, pos2, e2e_n, nu2, total2)
        st = ring.stats()
        out["e2e"] = {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": res.h2d_bytes, "d2h_bytes_per_step": res.d2h_bytes,
                      "images_per_step_per_gpu": e2e_n, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                      "h2d_gbs": res.h2d_bytes * world * e2e_steps / ms_e2e / 1e6,
                      "gpu_launches_per_step": res.kernel_launches, "api": "b2_ingest_ring_submit/wait (C ABI, host pointers)",
                      "pipelining": f"{min(3, args.listings)} listings in flight, {args.chunk_mb} MiB chunks, "
                                    f"{ring_bytes >> 30} GiB device ring (stalls so far: {st['stalls']})",
                      "raw_h2d": {"aggregate_gbs": raw_gbs, "per_gpu_gbs": raw_gbs / world,
                                  "with_d2h_mix_aggregate_gbs": raw_mix_gbs, "with_d2h_mix_per_gpu_gbs": raw_mix_gbs / world,
                                  "how": "every rank copies its page-locked listing to the device at once, nothing else running "
                                         "(cudaMemcpyAsync, CUDA events, max over ranks); `with_d2h_mix`: the same while thumbnail + "
                                         "preview sized buffers travel device -> host on a second stream"},
                      "frac_of_raw_h2d": (res.h2d_bytes * world * e2e_steps / ms_e2e / 1e6) / raw_gbs,
                      "frac_of_raw_h2d_with_d2h_mix": (res.h2d_bytes * world * e2e_steps / ms_e2e / 1e6) / raw_mix_gbs,
                      "limiter": "the host side of the copies (PCIe + host memory): the plain-copy ceilings of this box are under raw_h2d",
                      "host_numa_node_rank0": numa_node,
                      "parity": {"ok": all_ok(ok2 and ok2g), "dedupe_counts_global": c2counts}}
        del buf2, views2, src_t

    # ---- config 1: 1 000 x 512^2 + 10 k rows
    if want("c1"):
        n1 = 1000
        shapes1 = [(512, 512)] * n1
        buf1, views1 = make_host_listing(shapes1, np.arange(n1))
        stream_listings(views1, shapes1, 2, 2)
        ms1, _, res1 = stream_listings(views1, shapes1, 5, min(3, args.listings))
        ok1 = check_listing(res1, views1, shapes1, [0, 999]) and res1.stats == {"processed": n1, "created": n1, "updated": 0}
        rng1 = np.random.default_rng(1)
        img1 = np.repeat(np.arange(1000, dtype=np.int32), 10)
        cls1 = rng1.integers(0, LABEL_K, 10_000).astype(np.uint8)
        act1 = (rng1.random(10_000) < 0.95).astype(np.uint8)
        t0 = time.perf_counter()
        for _ in range(20):
            t1 = b2labels.label_tally(img1, cls1, act1, 1000, LABEL_K)
        dt1 = (time.perf_counter() - t0) / 20
        p1, h1 = numpy_tally_partials(img1, cls1, act1, 0, 1000, LABEL_K)
        ok1 &= [int(x) for x in t1.class_totals] == p1[:LABEL_K].tolist() and t1.S2 == int(p1[LABEL_K]) and \
            np.array_equal(t1.agree_hist, h1)
        out["c1"] = {"workload": "BASELINE config 1: 1 000 x 512x512x3 images (hash + dedupe + thumbnail + preview) + 10 k label rows",
                     "e2e_images_per_s": n1 * 5 / (ms1 / 1e3), "ms_per_listing": ms1 / 5, "label_rows_per_s_e2e": 10_000 / dt1,
                     "label_call_ms": dt1 * 1e3, "kappa_general": t1.kappa_general(), "per_gpu": True,
                     "parity": {"ok": bool(ok1)}}
        del buf1, views1

    # ---- config 3: mixed 256^2 .. 4096^2 listing sharded by bytes
    if want("c3"):
        per_gpu = args.c3_images
        total3 = per_gpu * world
        nu3 = total3 - total3 // 5
        perm = np.random.Generator(np.random.Philox(key=[SEED, 3])).permutation(nu3)
        pos_all = np.arange(total3, dtype=np.int64)
        cid_all = source_of(pos_all, nu3)
        side_all = 256 << (perm[cid_all] % 5)                # the size belongs to the content (a copy has its source's size)
        len_all = side_all.astype(np.int64) ** 2 * 3
        shards = b2dist.shard_by_bytes(len_all, world)
        mine = shards[rank]
        n_max3 = max(len(s) for s in shards)
        shapes3 = [(int(s), int(s)) for s in side_all[mine]]
        buf3, views3 = make_host_listing(shapes3, cid_all[mine])
        bytes3 = int(len_all[mine].sum())
        infl = args.listings
        n_list = max(infl + 8, int(args.c3_seconds * 50e9 / max(bytes3, 1)))
        stream_listings(views3, shapes3, 2, 2)
        ms3, done3, res3 = stream_listings(views3, shapes3, n_list, infl)
        ms3 = reduce_ranks(ms3)
        k_feed = n_list - infl                               # completions observed while listings were still being submitted
        steady_s = done3[k_feed - 1] - done3[0]
        steady_rate = reduce_ranks((k_feed - 1) * bytes3 / steady_s / 1e9, "min")
        small = [i for i, s in enumerate(shapes3) if s[0] <= 1024][:2] + [i for i, s in enumerate(shapes3) if s[0] == 4096][:1]
        ok3 = check_listing(res3, views3, shapes3, small)
        c3counts, ok3g = global_counts(res3, mine, n_max3, nu3, total3)
        st = ring.stats()
        tot_bytes = reduce_ranks(float(bytes3), "sum")
        out["c3"] = {"workload": f"BASELINE config 3 shape: {total3} mixed-size square images (side 256*2^j, j=0..4), 20 % global "
                                 f"duplicates, sharded by bytes over {world} GPU(s) (dist.shard_by_bytes), streamed through the ring",
                     "images_per_listing_per_gpu": [len(s) for s in shards], "bytes_per_listing_per_gpu": bytes3,
                     "listings": n_list, "listings_in_flight": infl,
                     "e2e_images_per_s": total3 * n_list / (ms3 / 1e3), "e2e_h2d_gbs": tot_bytes * n_list / ms3 / 1e6,
                     "e2e_h2d_gbs_per_gpu": tot_bytes * n_list / ms3 / 1e6 / world,
                     "steady_h2d_gbs_per_gpu_min": steady_rate,
                     "note": "e2e_* include the fill and the drain of the pipeline (a 4096^2 image hashes for ~0.8 s); steady_* is the "
                             "completion rate of listings while later listings were still being submitted (pipeline full at both ends)",
                     "ring_stalls": st["stalls"],
                     "parity": {"ok": all_ok(ok3 and ok3g), "dedupe_counts_global": c3counts,
                                "expected_counts": [total3, nu3, total3 - nu3], "sampled_vs_hashlib_pillow": len(small) * world}}
        del buf3, views3

    # ---- config 5: 4K images, 20 % global duplicates, thumbnails + previews + tally
    if want("c5"):
        per5 = min(C5_PER_GPU, max(64, int(host_avail * 0.5) // (C5_H * C5_W * 3)))
        per5 = int(reduce_ranks(float(per5), "min"))
        total5 = per5 * world
        nu5 = total5 - total5 // 5
        pos5 = np.arange(per5, dtype=np.int64) * world + rank
        shapes5 = [(C5_H, C5_W)] * per5
        buf5, views5 = make_host_listing(shapes5, source_of(pos5, nu5))
        bytes5 = per5 * C5_H * C5_W * 3
        stream_listings(views5[:64], shapes5[:64], 2, 2)
        n5 = max(3, args.c5_listings)
        ms5, done5, res5 = stream_listings(views5, shapes5, n5, min(3, args.listings))
        ms5 = reduce_ranks(ms5)
        k5f = n5 - min(3, args.listings)
        steady5 = reduce_ranks((k5f - 1) * bytes5 / (done5[k5f - 1] - done5[0]) / 1e9, "min") if k5f >= 2 else None
        ok5 = check_listing(res5, views5, shapes5, [0, per5 - 1])
        c5counts, ok5g = global_counts(res5, pos5, per5, nu5, total5)
        # the label tally of the rows that reference the unique images (100 per image), sharded by image range
        lo5, hi5 = b2dist.shard_range(nu5, rank, world)
        rng5 = np.random.Generator(np.random.Philox(key=[LABEL_SEED, 5]))
        cls5_all = rng5.integers(0, LABEL_K, nu5 * 100).astype(np.uint8)
        act5_all = (rng5.random(nu5 * 100) < 0.95).astype(np.uint8)
        img5_all = np.repeat(np.arange(nu5, dtype=np.int32), 100)
        r0, r1 = lo5 * 100, hi5 * 100
        vec5 = torch.zeros(LABEL_K + 7 + B2_AGREE_BINS, dtype=torch.int64, device=dev)
        if hi5 > lo5:
            engine.label_tally_device(torch.from_numpy(img5_all[r0:r1]).to(dev), torch.from_numpy(cls5_all[r0:r1]).to(dev),
                                      torch.from_numpy(act5_all[r0:r1]).to(dev), hi5 - lo5, LABEL_K, lo5, True, None,
                                      vec5[:LABEL_K + 7], vec5[LABEL_K + 7:])
        b2dist.allreduce_partials(vec5)
        v5 = vec5.cpu().numpy()
        p5, h5 = numpy_tally_partials(img5_all, cls5_all, act5_all, 0, nu5, LABEL_K)
        ok5t = np.array_equal(v5[:LABEL_K + 6], p5[:LABEL_K + 6]) and np.array_equal(v5[LABEL_K + 7:], h5)
        k5 = b2labels.fleiss_kappa_from_hist(v5[:LABEL_K], int(v5[LABEL_K + 1]), int(v5[LABEL_K + 3]), v5[LABEL_K + 7:])
        out["c5"] = {"workload": f"BASELINE config 5 shape: {total5} images of 3840x2160x3 ({per5} per GPU; 10 000 on 8 GPUs), 20 % "
                                 "global duplicates, digests + dedupe + thumbnails + previews + label tally, end to end from "
                                 "page-locked host memory",
                     "images_per_gpu": per5, "scaled_down_for_host_memory": per5 < C5_PER_GPU, "listings": n5,
                     "e2e_images_per_s": total5 * n5 / (ms5 / 1e3), "e2e_h2d_gbs_per_gpu": bytes5 * n5 / ms5 / 1e6,
                     "steady_h2d_gbs_per_gpu_min": steady5,
                     "ms_per_listing": ms5 / n5, "note": "e2e_* include fill and drain (a 4K image hashes for ~0.41 s); steady_* as in c3",
                     "label_rows": int(nu5 * 100), "kappa_general": k5,
                     "parity": {"ok": all_ok(ok5 and ok5g and bool(ok5t)), "dedupe_counts_global": c5counts,
                                "expected_counts": [total5, nu5, total5 - nu5], "tally_equals_numpy": bool(ok5t)}}
        del buf5, views5

    if ring is not None:
        ring.close()
        torch.cuda.empty_cache()

    # =========================================================================================================
    # (3) labels: config 4 as stated (ONE table, sharded) and the per-GPU 100 M-row roofline point
    # =========================================================================================================
    vec_len = LABEL_K + 7 + B2_AGREE_BINS
    if want("c4"):
        n_blocks = LABEL_IMAGES // C4_BLOCK_IMAGES
        b_lo, b_hi = b2dist.shard_range(n_blocks, rank, world)
        img_lo, img_hi = b_lo * C4_BLOCK_IMAGES, b_hi * C4_BLOCK_IMAGES
        threads = max(1, (os.cpu_count() or 8) // world)
        h_img, h_cls, h_act = c4_table(b_lo, b_hi, threads)
        d_img, d_cls, d_act = (torch.from_numpy(a).to(dev) for a in (h_img, h_cls, h_act))
        rows_local = h_img.size
        c_counts = torch.empty((img_hi - img_lo, LABEL_K), dtype=torch.int32, device=dev)
        vec = torch.empty(vec_len, dtype=torch.int64, device=dev)

        def c4_step_nccl():
            engine.label_tally_device(d_img, d_cls, d_act, img_hi - img_lo, LABEL_K, img_lo, True, c_counts,
                                      vec[:LABEL_K + 7], vec[LABEL_K + 7:])
            b2dist.allreduce_partials(vec)                   # b2_allreduce_i64 (NCCL) on the same stream

        fused = world > 1
        if fused:
            comm.enable_peer_reduce(vec_len)                 # mailboxes in every rank's HBM, mapped by its peers (CUDA IPC)

        def c4_step_fused():
            # ONE kernel: the slab tally whose last CTA all-reduces partials + histogram over NVLink peer memory
            comm.label_tally_reduce(d_img, d_cls, d_act, img_hi - img_lo, LABEL_K, img_lo, c_counts, vec)

        # `inner` steps captured in ONE CUDA graph: at 12.5 M rows per GPU the tally is ~30 us, launch gaps would
        # dominate otherwise
        inner = 20
        outer = max(args.steps, 5)

        def run_graphed(step):
            cap_stream = torch.cuda.Stream(dev)
            graph, graphed, err = torch.cuda.CUDAGraph(), False, None
            step()
            torch.cuda.synchronize()
            try:
                with torch.cuda.stream(cap_stream):
                    step()
                    cap_stream.synchronize()
                    with torch.cuda.graph(graph, stream=cap_stream):
                        for _ in range(inner):
                            step()
                graphed = True
            except Exception as e:  # noqa: BLE001 - report and fall back to plain launches
                err = repr(e)[:200]
                torch.cuda.synchronize()

            def outer_step():
                if graphed:
                    graph.replay()
                else:
                    for _ in range(inner):
                        step()

            ms, _, _ = timed(outer_step, outer, 3)
            torch.cuda.synchronize()
            return ms, graphed, err, graph

        ms_nccl, graphed_nccl, graph_err, _g1 = run_graphed(c4_step_nccl)
        got_nccl = vec.cpu().numpy()
        if fused:
            ms_c4, graphed, graph_err2, _g2 = run_graphed(c4_step_fused)
            graph_err = graph_err or graph_err2
            peer_timeout = comm.peer_timed_out()
        else:
            ms_c4, graphed, peer_timeout = ms_nccl, graphed_nccl, False
        torch.cuda.synchronize()
        got = vec.cpu().numpy()
        ms_tally_local = kernel_ms(lambda: engine.label_tally_device(d_img, d_cls, d_act, img_hi - img_lo, LABEL_K, img_lo, True,
                                                                     c_counts, vec[:LABEL_K + 7], vec[LABEL_K + 7:]), reps=20)
        # the checker: NumPy over the WHOLE table on rank 0 (the same table at every N)
        ok4 = True
        cpu_rows_per_s = None
        if rank == 0:
            if world > 1:
                f_img, f_cls, f_act = c4_table(0, n_blocks, os.cpu_count() or 8)
            else:
                f_img, f_cls, f_act = h_img, h_cls, h_act
            t0 = time.perf_counter()
            p_ref, h_ref = numpy_tally_partials(f_img, f_cls, f_act, 0, LABEL_IMAGES, LABEL_K)
            cpu_rows_per_s = f_img.size / (time.perf_counter() - t0)
            ok4 = bool(np.array_equal(got[:LABEL_K + 7], p_ref) and np.array_equal(got[LABEL_K + 7:], h_ref))
            ok4 = ok4 and bool(np.array_equal(got, got_nccl)) and not peer_timeout
            del f_img, f_cls, f_act
        digest = hashlib.sha256(got.tobytes()).hexdigest()
        kappa_g = b2labels.fleiss_kappa_from_hist(got[:LABEL_K], int(got[LABEL_K + 1]), int(got[LABEL_K + 3]), got[LABEL_K + 7:])
        all_const = int(got[LABEL_K + 1]) == LABEL_IMAGES * LABEL_RATERS
        rows_total = LABEL_IMAGES * LABEL_RATERS
        c4_bytes = 6 * rows_local + 4 * (img_hi - img_lo) * LABEL_K
        out["c4"] = {"workload": f"BASELINE config 4 as stated: ONE table of {rows_total} rows ({LABEL_IMAGES} images, k={LABEL_K}, "
                                 f"~95 active ratings per image), the same rows at every N, sharded by image range over {world} GPU(s); "
                                 "tally + b2_allreduce_i64 (NCCL) of k+7 partials and the 1024-bin agreement histogram",
                     "rows_per_gpu": rows_local, "value": rows_total * outer * inner / (ms_c4 / 1e3), "unit": "rows/s",
                     "us_per_step": 1e3 * ms_c4 / (outer * inner), "steps": outer * inner, "cuda_graph": graphed,
                     "collective": ("fused into the tally kernel: its last CTA all-reduces the 8.6 KB vector over NVLink peer memory "
                                    "(b2_label_tally_reduce; CUDA IPC mailboxes, no NCCL call)") if fused else None,
                     "us_per_step_tally_then_nccl_allreduce": 1e3 * ms_nccl / (outer * inner) if fused else None,
                     "tally_kernel_us": 1e3 * ms_tally_local,
                     "roofline": {"bound": "hbm", "achieved": c4_bytes / (ms_tally_local / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                  "frac": c4_bytes / (ms_tally_local / 1e3) / 1e9 / hbm_peak, "algorithmic_bytes": c4_bytes},
                     "kappa_general": kappa_g, "kappa_general_repr": repr(kappa_g),
                     "kappa_constant_n": None if not all_const else ics_b200.fleiss_kappa(
                         got[:LABEL_K], int(got[LABEL_K]), int(got[LABEL_K + 1]), LABEL_IMAGES, LABEL_RATERS),
                     "kappa_note": "rows are active w.p. 0.95, so n_i varies: the general-n kappa (mean of P_i, from the integer "
                                   "agreement histogram of the same pass) is the meaningful one; the constant-n form is reported only "
                                   "when every image has exactly 100 active ratings",
                     "partials_sha256": digest, "numpy_rows_per_s_1_thread": cpu_rows_per_s,
                     "parity": {"ok": all_ok(ok4), "checked": "all k+7 partials and 1024 histogram bins equal NumPy's over the whole "
                                                              "table (rank 0) and the NCCL path's; identical integers at every N => identical kappa"}}
        if graph_err:
            out["c4"]["cuda_graph_error"] = graph_err
        del d_img, d_cls, d_act, c_counts, _g1
        torch.cuda.empty_cache()

    if want("labels"):
        rows = LABEL_IMAGES * LABEL_RATERS
        gl = torch.Generator(device=dev).manual_seed(LABEL_SEED + rank)
        l_img = (torch.arange(rows, device=dev, dtype=torch.int64) // LABEL_RATERS).to(torch.int32)
        true_cls = torch.randint(0, LABEL_K, (LABEL_IMAGES,), device=dev, generator=gl)
        pick = torch.rand(rows, device=dev, generator=gl) < 0.7
        uni = torch.randint(0, LABEL_K, (rows,), device=dev, generator=gl)
        l_cls = torch.where(pick, true_cls[l_img.long()], uni).to(torch.uint8)
        del pick, uni
        l_act = (torch.rand(rows, device=dev, generator=gl) < 0.95).to(torch.uint8)
        l_counts = torch.empty((LABEL_IMAGES, LABEL_K), dtype=torch.int32, device=dev)
        l_vecs = [torch.empty(vec_len, dtype=torch.int64, device=dev) for _ in range(2)]
        l_vec = l_vecs[0]
        l_local = torch.empty(vec_len, dtype=torch.int64, device=dev)
        comm_stream = torch.cuda.Stream(dev)
        if world > 1:
            comm.enable_peer_reduce(vec_len)
        tallied = [torch.cuda.Event() for _ in range(2)]
        reduced = [torch.cuda.Event() for _ in range(2)]
        flip = {"i": 0}

        def label_step():
            # tally on the main stream, the all-reduce of its partials on a second stream: the ~15 us NCCL latency of step i
            # hides under the tally of step i+1 (two partial buffers); every step still produces its all-reduced vector
            i = flip["i"]
            flip["i"] ^= 1
            main = torch.cuda.current_stream()
            main.wait_event(reduced[i])                       # buffer i: its previous all-reduce is done
            engine.label_tally_device(l_img, l_cls, l_act, LABEL_IMAGES, LABEL_K, 0, True, l_counts,
                                      l_vecs[i][:LABEL_K + 7], l_vecs[i][LABEL_K + 7:])
            if world > 1:
                tallied[i].record(main)
                comm_stream.wait_event(tallied[i])
                with torch.cuda.stream(comm_stream):
                    comm.peer_allreduce_i64(l_vecs[i])        # one-CTA kernel over NVLink peer memory (no NCCL call)
                    reduced[i].record(comm_stream)
            launches["n"] += 1 if world == 1 else 2

        def label_drain():
            torch.cuda.current_stream().wait_stream(comm_stream)

        label_steps = max(args.steps, 20)
        for _ in range(3):
            label_step()
        label_drain()
        barrier()
        le0, le1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches["n"] = 0
        le0.record()
        for _ in range(label_steps):
            label_step()
        label_drain()
        le1.record()
        barrier()
        ms_labels = reduce_ranks(le0.elapsed_time(le1))
        label_launches = launches["n"]
        l_vec = l_vecs[flip["i"] ^ 1]                          # the vector of the last step
        rows_per_s = rows * world * label_steps / (ms_labels / 1e3)
        engine.label_tally_device(l_img, l_cls, l_act, LABEL_IMAGES, LABEL_K, 0, True, l_counts,
                                  l_local[:LABEL_K + 7], l_local[LABEL_K + 7:])
        local = l_local.cpu().numpy()
        ph = l_vec.cpu().numpy()
        engine.check_tally(local[:LABEL_K + 7], LABEL_K, rows)
        label_ok = int(local[LABEL_K + 1]) == int(l_act.sum().item())
        kappa_g = b2labels.fleiss_kappa_from_hist(ph[:LABEL_K], int(ph[LABEL_K + 1]), int(ph[LABEL_K + 3]), ph[LABEL_K + 7:])
        ms_tally = kernel_ms(lambda: engine.label_tally_device(l_img, l_cls, l_act, LABEL_IMAGES, LABEL_K, 0, True,
                                                               l_counts, l_local[:LABEL_K + 7], l_local[LABEL_K + 7:]), reps=20)
        ms_tally_nohist = kernel_ms(lambda: engine.label_tally_device(l_img, l_cls, l_act, LABEL_IMAGES, LABEL_K, 0, True,
                                                                      l_counts, l_local[:LABEL_K + 7]), reps=20)
        # the same rows in random order (a heap scan instead of an index scan on id_img): any-order path
        perm = torch.randperm(rows, device=dev, generator=gl)
        s_img, s_cls, s_act = l_img[perm].contiguous(), l_cls[perm].contiguous(), l_act[perm].contiguous()
        del perm
        s_vec = torch.empty(vec_len, dtype=torch.int64, device=dev)
        ms_scatter = kernel_ms(lambda: engine.label_tally_device(s_img, s_cls, s_act, LABEL_IMAGES, LABEL_K, 0, False,
                                                                 l_counts, s_vec[:LABEL_K + 7], s_vec[LABEL_K + 7:]), reps=5)
        sv = s_vec.cpu().numpy()
        scatter_ok = bool(np.array_equal(sv[:LABEL_K + 5], local[:LABEL_K + 5]) and np.array_equal(sv[LABEL_K + 7:], local[LABEL_K + 7:]))
        del s_img, s_cls, s_act
        # labels e2e: rows in pinned host memory -> device -> tally -> partials back on the host
        h_img, h_cls, h_act = (l_img.cpu().pin_memory(), l_cls.cpu().pin_memory(), l_act.cpu().pin_memory())
        n_img_np, n_cls_np, n_act_np = h_img.numpy(), h_cls.numpy(), h_act.numpy()
        hist_np = np.zeros(B2_AGREE_BINS, dtype=np.int64)

        def label_e2e_step():
            _, p = b2labels.label_tally_host(n_img_np, n_cls_np, n_act_np, LABEL_IMAGES, LABEL_K, want_counts=False,
                                             device=local_rank, agree_hist=hist_np)
            if world > 1:
                t = torch.from_numpy(np.concatenate([p, hist_np])).to(dev)
                b2dist.allreduce_partials(t)
                p = t.cpu().numpy()
            return p

        ms_le2e, _, _ = timed(label_e2e_step, 5, 2)
        tally_bytes = 6 * rows + 4 * LABEL_IMAGES * LABEL_K
        out["labels"] = dict(rows=rows, rows_per_s=rows_per_s, steps=label_steps, ms=ms_labels, launches=label_launches,
                             ms_tally=ms_tally, ms_tally_nohist=ms_tally_nohist, ms_scatter=ms_scatter, scatter_ok=scatter_ok,
                             label_ok=label_ok, kappa_g=kappa_g, e2e=rows * world * 5 / (ms_le2e / 1e3), tally_bytes=tally_bytes)
        del l_img, l_cls, l_act, l_counts

    os.sched_setaffinity(0, affinity0)
    if rank == 0:
        sampler.stop()
        print(json.dumps(build_line(args, out, world, warmup, hbm_peak, peak_src, torch, dev)), flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def build_line(args, out, world, warmup, hbm_peak, peak_src, torch, dev):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            ncu = json.load(f)
    except (OSError, ValueError):
        ncu = {}

    def roof(bytes_, ms, kernel=None, note=None):
        ach = bytes_ / (ms / 1e3) / 1e9
        d = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
             "traffic": None, "algorithmic_bytes": bytes_, "peak_source": peak_src, "ms_per_launch": ms}
        m = ncu.get(kernel)
        if m:
            d["traffic_scaled_from_ncu"] = bytes_ * m["dram_bytes"] / m["algorithmic_bytes"]
            d["traffic_source"] = (f"ncu dram bytes / algorithmic bytes = {m['dram_bytes'] / m['algorithmic_bytes']:.4f} "
                                   f"on {m['workload']} (profiles/traffic.json), scaled to this launch — not a counter of this run")
            d["pipes_ncu"] = {"alu_pct": m["alu_pipe_pct"], "fma_pct": m["fma_pipe_pct"], "issue_slots_pct": m["issue_slots_pct"]}
        if note:
            d["note"] = note
        return d

    line = {"metric": "ingest images/s", "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32", "data": "synthetic"}
    res = out.get("resident")
    kernels = {}
    if res:
        n_img, clocks = res["n_img"], res["clocks"]
        sha_bytes = n_img * (IMG_BYTES + 32)
        resize_bytes = n_img * (IMG_BYTES + THUMB_BYTES + PREVIEW_BYTES)
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        # the hash's real bound: 1 290 ALU-pipe warp instructions per 64-byte block and warp (SASS count; ncu: 1 416
        # instructions per block, 91 % of them SHF/LOP3/IADD3/PRMT) against 0.5 ALU warp instructions per clock and SM
        # sub-partition at the SM clock sampled during the run
        alu_ach = sha_bytes / 64.0 / 32.0 * 1290.0 / (res["ms_sha"] / 1e3) / 1e9
        alu_peak = sms * 4 * 0.5 * sm_mhz * 1e6 / 1e9
        hbm_ach = sha_bytes / (res["ms_sha"] / 1e3) / 1e9
        warps = (n_img + 31) // 32
        issue_ach = sha_bytes / 64.0 / 32.0 * 1416.0 / (res["ms_sha"] / 1e3) / 1e9
        issue_peak = warps * 0.5 * sm_mhz * 1e6 / 1e9
        m = ncu.get("sha256_lanes_kernel") or {}
        line.update({
            "value": res["value"], "ms_per_step": res["ms"] / args.steps,
            "config": workload_config(world, n_img, RESIDENT_INPUTS),
            "clocks": clocks, "gpu_launches": res["launches"], "step_ms_rank0": res["step_ms"],
            "roofline": {
                "kernel": "sha256_lanes_kernel", "bound": "int32 ALU pipe", "achieved": alu_ach, "peak": alu_peak,
                "unit": "G warp-instructions/s", "frac": alu_ach / alu_peak, "alu_instructions_per_64B_block": 1290,
                "hbm_achieved_gbs": hbm_ach, "hbm_peak_gbs": hbm_peak, "hbm_frac": hbm_ach / hbm_peak, "peak_source": peak_src,
                "algorithmic_bytes": sha_bytes, "ms_per_launch": res["ms_sha"],
                "traffic": None,
                "traffic_scaled_from_ncu": sha_bytes * m["dram_bytes"] / m["algorithmic_bytes"] if m else None,
                "lone_warp_issue": {"bound": "issue rate of a lone warp (1 instruction / 2 clocks)", "achieved": issue_ach,
                                    "peak": issue_peak, "frac": issue_ach / issue_peak, "resident_warps": warps},
                "note": "dominant kernel of the ingest step; SHA-256 is bound by the INT32 ALU pipe, not by HBM (hbm_frac is "
                        "for reference); the HBM-bound kernels are under `kernels`"},
            "parity": res["parity"],
        })
        kernels["sha256_lanes_kernel"] = roof(sha_bytes, res["ms_sha"], "sha256_lanes_kernel", note="ALU bound: see roofline")
        kernels["resize_pairs_kernel"] = roof(resize_bytes, res["ms_resize"], "resize_pairs_kernel",
                                              note="timed alone; two output columns per thread (1080p -> 256 wide); the band kernel "
                                                   "(B2_RESIZE_NO_PAIRS=1) measures 0.745 on the same launch")
        kernels["dedupe (insert+resolve)"] = {"ms_per_launch": res["ms_dedupe"], "digests": n_img}
    lab = out.get("labels")
    if lab:
        kernels["tally_slab_kernel"] = roof(lab["tally_bytes"], lab["ms_tally"], "tally_slab_kernel",
                                            note="with the agreement histogram (general-n kappa) in the same pass; without it: "
                                                 f"{lab['ms_tally_nohist']:.4f} ms")
        line["labels"] = {"value": lab["rows_per_s"], "unit": "rows/s", "rows_per_gpu_per_step": lab["rows"], "steps": lab["steps"],
                          "ms_per_step": lab["ms"] / lab["steps"], "gpu_launches": lab["launches"],
                          "roofline": kernels["tally_slab_kernel"], "kappa_general": lab["kappa_g"],
                          "partials_ok": bool(lab["label_ok"]),
                          "collective": "b2_peer_allreduce_i64 (one CTA, NVLink peer memory) of k+7+1024 int64 on a second stream: the "
                                        "all-reduce of step i overlaps the tally of step i+1 (two partial buffers)" if world > 1 else None,
                          "shuffled_rows": {"value": lab["rows"] / (lab["ms_scatter"] / 1e3), "unit": "rows/s per GPU",
                                            "ms_per_launch": lab["ms_scatter"],
                                            "path": "memset + tally_scatter_kernel (global RED.ADD) + fleiss_partials_kernel",
                                            "same_partials_as_sorted": lab["scatter_ok"]},
                          "e2e": {"value": lab["e2e"], "unit": "rows/s", "rows_per_step": lab["rows"],
                                  "h2d_bytes_per_step": 6 * lab["rows"], "d2h_bytes_per_step": 8 * (LABEL_K + 7 + 1024)}}
    if kernels:
        line["kernels"] = kernels
    if "e2e" in out:
        line["e2e"] = out["e2e"]
    line["configs"] = {k: out[k] for k in ("c1", "c3", "c4", "c5") if k in out}
    if "value" not in line:                                  # a partial run (--only ...): keep the line well-formed
        line.update({"value": None, "ms_per_step": None, "config": {"workload": "partial run: " + args.only}})

    cores = os.cpu_count() or 1
    if not args.only or "cpu" in args.only:
        cpu_n = 512 * cores                                   # ~10-15 s of work on every core
        cpu_v, cpu_dt = cpu_ingest_sample(cpu_n, cores)
        cpu_1, _ = cpu_ingest_sample(48, 1)                   # the reference loop is sequential per image (webdav_sync.py:311)
        line["cpu_baseline"] = {"value": cpu_v, "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"{cpu_n} synthetic 1920x1080x3 images, hashlib+Pillow+NumPy oracle, "
                                          f"{cores} threads, {cpu_dt:.1f} s",
                                "single_thread_images_per_s": cpu_1}
        if "c1" in line["configs"]:
            line["configs"]["c1"]["cpu"] = cpu_config1(cores)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["graft", "reference"], default="graft")
    ap.add_argument("--images-per-gpu", type=int, default=0, help="default: 18944 (one hash warp per SM sub-partition)")
    ap.add_argument("--e2e-images", type=int, default=2048, help="1080p images per listing and GPU of the e2e run (same at every N)")
    ap.add_argument("--c3-images", type=int, default=840, help="images per GPU of one config-3 listing")
    ap.add_argument("--c3-seconds", type=float, default=3.0, help="target length of the config-3 stream")
    ap.add_argument("--c5-listings", type=int, default=6)
    ap.add_argument("--ring-gb", type=int, default=64, help="device staging ring of the e2e runs")
    ap.add_argument("--chunk-mb", type=int, default=1024)
    ap.add_argument("--listings", type=int, default=8, help="listings in flight (ring slots)")
    ap.add_argument("--only", default="", help="comma list of: resident,e2e,c1,c3,c4,c5,labels,cpu (default: all)")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="run resize and hash back to back instead of on two streams (default: concurrently)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_graft(args)


if __name__ == "__main__":
    main()
