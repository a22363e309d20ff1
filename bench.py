#!/usr/bin/env python
"""Benchmark of the ingest + label-aggregation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of
synthetic input:

  ingest step  (headline `value`, images/s): `images_per_gpu` synthetic 1920x1080x3 images
      (BASELINE config 2 shape) resident in HBM -> 256x256 uint8 thumbnail + float32 CHW preview
      of every image (HBM bound) on a second stream beside the SHA-256 of every image (INT32-ALU
      bound: one warp per SM sub-partition), then the dedupe decision over the digests (+ digest
      all-gather at N > 1).
  label step   (`labels.value`, rows/s): BASELINE config 4 shape per GPU — 100 M rows, 1 M images,
      k = 50, clustered by image -> count matrix + integer Fleiss partials (+ all-reduce at N > 1).

`e2e` is the same ingest metric through the host-buffer C ABI (`b2_ingest_stream_*`: pinned HOST buffers
in, host results out; H2D and D2H inside the timed region).  `roofline` is for the dominant kernel of the ingest step (the
hash); `kernels` lists every kernel's own roofline.  `cpu_baseline` / `--impl reference` time the
oracle (hashlib + Pillow + NumPy: the reference's own host libraries) on the box's host cores.
Inputs are larger than L2 (no flush needed): stated in `config`.
"""
from __future__ import annotations

import argparse
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

IMG_H, IMG_W = 1080, 1920
IMG_BYTES = IMG_H * IMG_W * 3                     # 6 220 800
OUT = 256
THUMB_BYTES = OUT * OUT * 3
PREVIEW_BYTES = OUT * OUT * 3 * 4
LABEL_IMAGES, LABEL_K, LABEL_RATERS = 1_000_000, 50, 100
FULL_IMAGES_PER_GPU = 148 * 4 * 32                # one hash warp per SM sub-partition: 18 944 images = 117.8 GB


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


# ------------------------------------------------------------------------------- CPU arm
def cpu_ingest_sample(n_images: int, threads: int, seed: int = 0xB200):
    """The oracle's ingest path (hashlib.sha256 + dict dedupe + Pillow BILINEAR + float32 preview)
    on `n_images` synthetic 1080p images with `threads` host threads.  Returns images/s."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import dedupe_batch, preview_f32, sha256_hex, thumbnail_u8

    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(min(n_images, 16), IMG_H, IMG_W, 3), dtype=np.uint8)
    imgs = [base[i % len(base)] for i in range(n_images)]

    def one(im):
        h = sha256_hex(im.data)                      # the "file bytes" = the raw RGB buffer
        t = thumbnail_u8(im, OUT, OUT)
        return h, t, preview_f32(t)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        out = list(ex.map(one, imgs))
    dedupe_batch([o[0] for o in out])
    dt = time.perf_counter() - t0
    return n_images / dt, dt


def cpu_label_sample(rows: int, threads: int):
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import fleiss_kappa, fleiss_partials, label_tally, synth_label_rows

    n_images = rows // LABEL_RATERS
    img, cls, act = synth_label_rows(n_images, LABEL_K, LABEL_RATERS)
    shards = max(1, min(threads, 16))
    bounds = [(n_images * s // shards, n_images * (s + 1) // shards) for s in range(shards)]

    def one(b):
        lo, hi = b
        r0, r1 = lo * LABEL_RATERS, hi * LABEL_RATERS
        return label_tally(img[r0:r1] - lo, cls[r0:r1], act[r0:r1], hi - lo, LABEL_K)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=shards) as ex:
        parts = list(ex.map(one, bounds))
    counts = np.concatenate(parts)
    p = fleiss_partials(counts)
    fleiss_kappa(p["class_totals"], p["S2"], p["R"], n_images, LABEL_RATERS)
    dt = time.perf_counter() - t0
    return rows / dt, dt


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
        "config": workload_config(args.gpus, n_img, "host memory; bounded sample"),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "labels": {"value": lab_v, "unit": "rows/s", "cores": min(cores, 16), "sample": "10M rows, N=100k, k=50"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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

    import ics_b200
    from ics_b200 import engine
    from ics_b200 import dist as b2dist
    from ics_b200.pipeline import IngestPipeline

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    engine.init(local_rank)
    hbm_peak, peak_src = peaks()
    # host side of the end-to-end runs: page-locked buffers on the NUMA node next to this rank's GPU
    # (undone before the CPU baseline, which uses every host core)
    affinity0 = os.sched_getaffinity(0)
    numa_node = engine.bind_host_thread_to_gpu_node(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- synthetic inputs, generated on device ----------------
    free_b, _ = torch.cuda.mem_get_info()
    n_img = args.images_per_gpu or FULL_IMAGES_PER_GPU
    budget = int(free_b * 0.80) - 4 * (1 << 30)
    n_img = max(32, min(n_img, budget // IMG_BYTES) // 32 * 32)
    gen = torch.Generator(device=dev).manual_seed(0xB200 + rank)
    data = torch.empty((n_img, IMG_BYTES), dtype=torch.uint8, device=dev)
    for lo in range(0, n_img, 256):
        data[lo:lo + 256].random_(0, 256, generator=gen)
    n_dup = n_img // 5                                       # 20 % duplicates (config 5's rule)
    n_unique = n_img - n_dup
    src = (torch.arange(n_unique, n_img, device=dev, dtype=torch.int64) * 2654435761) % n_unique
    for lo in range(0, n_dup, 256):
        data[n_unique + lo:n_unique + lo + 256] = data[src[lo:lo + 256]]
    offsets = torch.arange(n_img, dtype=torch.int64, device=dev) * IMG_BYTES
    lengths = torch.full((n_img,), IMG_BYTES, dtype=torch.int64, device=dev)
    flat = data.view(-1)
    digests = torch.empty((n_img, 32), dtype=torch.uint8, device=dev)
    thumbs = torch.empty((n_img, OUT, OUT, 3), dtype=torch.uint8, device=dev)
    previews = torch.empty((n_img, 3, OUT, OUT), dtype=torch.float32, device=dev)
    plan = engine.get_plan(IMG_H, IMG_W, OUT, OUT, local_rank)
    global_index = (torch.arange(n_img, dtype=torch.int32, device=dev) * world + rank)   # image g = i*G + rank
    side = torch.cuda.Stream(dev)                            # resize (default = lowest priority)
    hstream = torch.cuda.Stream(dev, priority=-1)            # hash (higher priority: its CTAs are placed first)
    launches = {"n": 0}
    per_step = {"ms": []}

    def ingest_step():
        main = torch.cuda.current_stream()
        if args.overlap:
            # The hash keeps one warp per SM sub-partition busy on the INT32 ALU pipe for the whole step and
            # leaves HBM and most issue slots idle: the HBM-bound resize runs beside it on a second stream.
            # The hash stream has the higher priority, so its 592 one-warp CTAs are placed first (four per SM,
            # one per sub-partition) and the resize CTAs fill in around them.
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
        if world > 1:
            is_new, counts = b2dist.global_dedupe(digests, global_index)
        else:
            is_new, _, _, counts = engine.dedupe_device(digests)
        launches["n"] += 4
        return is_new, counts

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
        return max_over_ranks(e0.elapsed_time(e1)), out, (t0, t1)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    warmup = max(3, args.warmup)                             # timing rule: at least three untimed steps
    ms_ingest, (is_new, counts), window = timed(ingest_step, args.steps, warmup)
    ingest_launches = launches["n"]
    ingest_step_ms = list(per_step["ms"])
    counts_h = counts.cpu().tolist()
    total_images = n_img * world * args.steps
    value = total_images / (ms_ingest / 1e3)
    clocks = sampler.summary(*window) if rank == 0 else None

    # per-kernel timings (each alone on the current stream, CUDA events around `reps` launches)
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

    ms_sha = kernel_ms(lambda: engine.sha256_device(flat, offsets, lengths, None, digests))
    ms_resize = kernel_ms(lambda: plan.run(flat, offsets, thumb=thumbs, preview=previews))
    ms_dedupe = kernel_ms(lambda: engine.dedupe_device(digests))

    # sanity check of the timed outputs (sampled images; not timed) against the reference's own host libraries,
    # called directly: hashlib.sha256(...).hexdigest() (webdav_sync.py:59) and Pillow's BILINEAR resize.  (oracle/
    # is imported by the CPU-baseline leg only.)
    parity = None
    if rank == 0:
        import hashlib
        from PIL import Image as PILImage
        idx = [0, n_img // 2, n_img - 1]
        ok = True
        hexes = engine.hex_strings(engine.digest_hex_device(digests[idx].contiguous()))
        for j, i in enumerate(idx):
            host = data[i].cpu().numpy()
            ok &= hexes[j] == hashlib.sha256(host.tobytes()).hexdigest()
            want = np.asarray(PILImage.fromarray(host.reshape(IMG_H, IMG_W, 3), "RGB").resize((OUT, OUT), PILImage.BILINEAR))
            ok &= bool(np.array_equal(thumbs[i].cpu().numpy(), want))
        exp_created = n_unique * world if world == 1 else None
        ok &= counts_h[0] == n_img * world and (exp_created is None or counts_h[1] == exp_created)
        parity = {"sampled_images": len(idx), "ok": bool(ok), "dedupe_counts": counts_h}

    # ---------------- end to end: pinned host buffers through the pipeline ----------------
    # page-locked host memory per rank: the batch + two pipelines' result buffers (~8.3 MB per image)
    e2e_n = min(args.e2e_images if world <= 2 else min(args.e2e_images, 2048), n_img)
    host_images = torch.empty((e2e_n, IMG_BYTES), dtype=torch.uint8, pin_memory=True)
    host_images.copy_(data[:e2e_n])
    torch.cuda.synchronize()
    ref_digests, ref_thumbs = digests[:e2e_n].clone(), thumbs[:8].clone()
    del data, flat, thumbs, previews                          # make room: the pipeline stages the batch on device
    torch.cuda.empty_cache()
    # two pipelines: batch i+1 is submitted before result i is read, as a streaming service would
    pipes = [IngestPipeline(IMG_H, IMG_W, e2e_n, chunk_images=args.e2e_chunk, device=local_rank) for _ in range(2)]
    res = {}
    e2e_steps = max(10, args.steps)                          # the last batch's hash tail (~130 ms) is not hidden by a next batch: amortise it

    def e2e_run(steps):
        pipes[0].submit(host_images)
        for i in range(steps):
            if i + 1 < steps:
                pipes[(i + 1) & 1].submit(host_images)
            res["r"] = pipes[i & 1].result()                 # host read of step i's digests/flags/stats/thumbnails

    e2e_run(2)                                               # warm-up
    barrier()
    t_e2e0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    e2e_run(e2e_steps)
    ev1.record()
    barrier()
    ms_e2e = max_over_ranks(max(ev0.elapsed_time(ev1), 1e3 * (time.perf_counter() - t_e2e0)))
    e2e_value = e2e_n * world * e2e_steps / (ms_e2e / 1e3)
    e2e_ok = bool(torch.equal(res["r"].digests.to(dev), ref_digests)) and \
        bool(torch.equal(res["r"].thumbs[:8].to(dev), ref_thumbs))
    h2d, d2h = res["r"].h2d_bytes, res["r"].d2h_bytes
    e2e_launches = pipes[0].kernel_launches
    del pipes, host_images
    torch.cuda.empty_cache()

    # ---------------- labels ----------------
    rows = LABEL_IMAGES * LABEL_RATERS
    gl = torch.Generator(device=dev).manual_seed(0xF1E155 + rank)
    l_img = (torch.arange(rows, device=dev, dtype=torch.int64) // LABEL_RATERS).to(torch.int32)
    true_cls = torch.randint(0, LABEL_K, (LABEL_IMAGES,), device=dev, generator=gl)
    pick = torch.rand(rows, device=dev, generator=gl) < 0.7
    uni = torch.randint(0, LABEL_K, (rows,), device=dev, generator=gl)
    l_cls = torch.where(pick, true_cls[l_img.long()], uni).to(torch.uint8)
    del pick, uni
    l_act = (torch.rand(rows, device=dev, generator=gl) < 0.95).to(torch.uint8)
    l_counts = torch.empty((LABEL_IMAGES, LABEL_K), dtype=torch.int32, device=dev)
    l_part = torch.empty(LABEL_K + 7, dtype=torch.int64, device=dev)

    def label_step():
        engine.label_tally_device(l_img, l_cls, l_act, LABEL_IMAGES, LABEL_K, 0, True, l_counts, l_part)
        out = l_part.clone()
        b2dist.allreduce_partials(out)
        launches["n"] += 1
        return out

    label_steps = max(args.steps, 20)
    ms_labels, part, _ = timed(label_step, label_steps, max(args.warmup, 3))
    label_launches = launches["n"]
    rows_per_s = rows * world * label_steps / (ms_labels / 1e3)
    ph = part.cpu().numpy()
    local = l_part.cpu().numpy()
    engine.check_tally(local, LABEL_K, rows)
    label_ok = int(local[LABEL_K + 1]) == int(l_act.sum().item())
    kappa = ics_b200.fleiss_kappa(ph[:LABEL_K], int(ph[LABEL_K]), int(ph[LABEL_K + 1]),
                                  LABEL_IMAGES * world, LABEL_RATERS)
    ms_tally = kernel_ms(lambda: engine.label_tally_device(l_img, l_cls, l_act, LABEL_IMAGES, LABEL_K, 0, True,
                                                           l_counts, l_part), reps=20)
    # the same rows in random order (a heap scan instead of an index scan on id_img): any-order path =
    # zeroed matrix + global RED.ADD (L2-atomic bound) + partials pass; reported for information
    perm = torch.randperm(rows, device=dev, generator=gl)
    s_img, s_cls, s_act = l_img[perm].contiguous(), l_cls[perm].contiguous(), l_act[perm].contiguous()
    del perm
    ms_scatter = kernel_ms(lambda: engine.label_tally_device(s_img, s_cls, s_act, LABEL_IMAGES, LABEL_K, 0, False,
                                                             l_counts, l_part), reps=5)
    scatter_ok = bool(torch.equal(l_part[:LABEL_K + 5].cpu(), torch.from_numpy(local[:LABEL_K + 5])))
    del s_img, s_cls, s_act
    # labels e2e: rows in pinned host memory -> device -> tally -> partials back on the host
    e_rows = rows                                             # the whole config-4 table of this GPU: 600 MB of host memory
    h_img, h_cls, h_act = (l_img[:e_rows].cpu().pin_memory(), l_cls[:e_rows].cpu().pin_memory(),
                           l_act[:e_rows].cpu().pin_memory())

    from ics_b200 import labels as b2labels
    n_img_np, n_cls_np, n_act_np = h_img.numpy(), h_cls.numpy(), h_act.numpy()   # views of the pinned buffers

    def label_e2e_step():
        # one C-ABI call with host pointers: stage, tally, read the partials back (b2_label_tally_host)
        _, p = b2labels.label_tally_host(n_img_np, n_cls_np, n_act_np, e_rows // LABEL_RATERS, LABEL_K,
                                         want_counts=False, device=local_rank)
        if world > 1:
            t = torch.from_numpy(p).to(dev)
            b2dist.allreduce_partials(t)
            p = t.cpu().numpy()
        return p

    ms_le2e, _, _ = timed(label_e2e_step, 5, 2)
    label_e2e = e_rows * world * 5 / (ms_le2e / 1e3)

    os.sched_setaffinity(0, affinity0)
    if rank == 0:
        sampler.stop()
        sha_bytes = n_img * (IMG_BYTES + 32)
        resize_bytes = n_img * (IMG_BYTES + THUMB_BYTES + PREVIEW_BYTES)
        tally_bytes = 6 * rows + 4 * LABEL_IMAGES * LABEL_K

        # DRAM traffic per launch: the ratio measured by ncu on the profiling workload (profiles/traffic.json,
        # committed with the ncu summary it comes from), scaled to this launch's algorithmic bytes.
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
                d["traffic"] = bytes_ * m["dram_bytes"] / m["algorithmic_bytes"]
                d["traffic_source"] = (f"ncu dram bytes / algorithmic bytes = {m['dram_bytes'] / m['algorithmic_bytes']:.4f} "
                                       f"on {m['workload']} (profiles/traffic.json), scaled to this launch")
                d["pipes_ncu"] = {"alu_pct": m["alu_pipe_pct"], "fma_pct": m["fma_pipe_pct"], "issue_slots_pct": m["issue_slots_pct"]}
            if note:
                d["note"] = note
            return d

        def int_alu_roofline(bytes_, ms, clk):
            # the hash's real bound: 1 290 ALU-pipe warp instructions per 64-byte block per warp (SASS count and
            # ncu: 1 416 instructions per block, 91 % of them SHF/LOP3/IADD3/PRMT), against 0.5 ALU warp
            # instructions per clock and SM sub-partition (16 lanes) at the SM clock sampled during the run
            sm_mhz = (clk or {}).get("sm_mhz") or 1965.0
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            achieved = bytes_ / 64.0 / 32.0 * 1290.0 / (ms / 1e3) / 1e9
            peak = sms * 4 * 0.5 * sm_mhz * 1e6 / 1e9
            return {"bound": "int32 ALU pipe", "achieved": achieved, "peak": peak, "unit": "G warp-instructions/s",
                    "frac": achieved / peak, "alu_instructions_per_64B_block": 1290}

        def lone_warp_issue_roofline(bytes_, ms, clk, warps):
            # With one resident warp per SM sub-partition (all the 1080p images that fit HBM allow) the hash is bound by
            # the rate at which ONE warp can issue: a 32-thread instruction takes two clocks to dispatch whatever pipe it
            # goes to, so 1 416 instructions per 64-byte block cost 2 832 clocks per block and warp.
            sm_mhz = (clk or {}).get("sm_mhz") or 1965.0
            achieved = bytes_ / 64.0 / 32.0 * 1416.0 / (ms / 1e3) / 1e9
            peak = warps * 0.5 * sm_mhz * 1e6 / 1e9
            return {"bound": "issue rate of a lone warp (1 instruction / 2 clocks)", "achieved": achieved, "peak": peak,
                    "unit": "G warp-instructions/s", "frac": achieved / peak, "instructions_per_64B_block": 1416,
                    "resident_warps": warps}

        cores = os.cpu_count() or 1
        cpu_n = 512 * cores                                    # ~10-15 s of work on every core
        cpu_v, cpu_dt = cpu_ingest_sample(cpu_n, cores)
        cpu_1, _ = cpu_ingest_sample(48, 1)                    # the reference loop is sequential per image (webdav_sync.py:311)
        line = {
            "metric": "ingest images/s", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_ingest / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32",
            "data": "synthetic",
            "config": workload_config(world, n_img, "resident in HBM, generated on device (seeded), 20% duplicates"),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "images_per_step": e2e_n, "chunk_images": args.e2e_chunk, "steps": e2e_steps,
                    "gpu_launches_per_step": e2e_launches, "pipelining": "2 batches in flight (submit i+1 before result i)",
                    "host_numa_node_rank0": numa_node,
                    "matches_device_path": e2e_ok},
            "gpu_launches": ingest_launches, "step_ms_rank0": ingest_step_ms,
            "roofline": dict(roof(sha_bytes, ms_sha, "sha256_lanes_kernel",
                             note="dominant kernel of the ingest step (79 % of it); sha256 is bound by the INT32 ALU pipe "
                                  "(1 290 ALU instructions per 64-byte block, pipe 90 % busy under ncu), not by HBM: "
                                  "frac of HBM peak is reported for reference, the HBM-bound kernels are under `kernels`"),
                             int_alu=int_alu_roofline(sha_bytes, ms_sha, clocks),
                             lone_warp_issue=lone_warp_issue_roofline(sha_bytes, ms_sha, clocks, (n_img + 31) // 32)),
            "kernels": {
                "sha256_lanes_kernel": roof(sha_bytes, ms_sha, "sha256_lanes_kernel"),
                "resize_bands_kernel": roof(resize_bytes, ms_resize, "resize_bands_kernel",
                                            note="timed alone; IDP.4A horizontal pass on a window de-interleaved in registers, "
                                                 "scatter-form vertical pass in registers, three CTAs per SM"),
                "dedupe (insert+resolve)": {"ms_per_launch": ms_dedupe, "digests": n_img},
                "tally_slab_kernel": roof(tally_bytes, ms_tally, "tally_slab_kernel"),
            },
            "cpu_baseline": {"value": cpu_v, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{cpu_n} synthetic 1920x1080x3 images, hashlib+Pillow+NumPy oracle, "
                                       f"{cores} threads, {cpu_dt:.1f} s",
                             "single_thread_images_per_s": cpu_1},
            "labels": {"value": rows_per_s, "unit": "rows/s", "rows_per_gpu_per_step": rows, "steps": label_steps,
                       "ms_per_step": ms_labels / label_steps, "gpu_launches": label_launches,
                       "roofline": roof(tally_bytes, ms_tally, "tally_slab_kernel"), "kappa": kappa, "partials_ok": bool(label_ok),
                       "shuffled_rows": {"value": rows / (ms_scatter / 1e3), "unit": "rows/s per GPU", "ms_per_launch": ms_scatter,
                                         "path": "memset + tally_scatter_kernel (global RED.ADD) + fleiss_partials_kernel",
                                         "same_partials_as_sorted": scatter_ok},
                       "e2e": {"value": label_e2e, "unit": "rows/s", "rows_per_step": e_rows,
                               "h2d_bytes_per_step": 6 * e_rows, "d2h_bytes_per_step": 8 * (LABEL_K + 7)}},
            "parity": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["graft", "reference"], default="graft")
    ap.add_argument("--images-per-gpu", type=int, default=0, help="default: 18944 (one hash warp per SM sub-partition)")
    ap.add_argument("--e2e-images", type=int, default=4096)
    ap.add_argument("--e2e-chunk", type=int, default=256)
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="run resize and hash back to back instead of on two streams (default: concurrently)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_graft(args)


if __name__ == "__main__":
    main()
