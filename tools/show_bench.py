#!/usr/bin/env python
"""Print the headline numbers of a bench.py JSON line (one per file)."""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads([ln for ln in open(path) if ln.startswith("{")][-1])
    print(f"== {path}: N={d['n_gpus']}")
    if d.get("value"):
        r = d["roofline"]
        print(f"  resident {d['value']:.0f} img/s  {d['ms_per_step']:.1f} ms/step  parity {d['parity']['ok']} counts {d['parity']['dedupe_counts']}"
              f" xrank-dups {d['parity']['cross_rank_duplicates']}  hash ALU frac {r['frac']:.3f} hbm {r['hbm_frac']:.3f}")
    for k, v in (d.get("kernels") or {}).items():
        if "frac" in v:
            print(f"  kernel {k}: {v['ms_per_launch']:.4f} ms  {v['achieved']:.0f} GB/s  frac {v['frac']:.3f}")
    if "e2e" in d:
        e = d["e2e"]
        print(f"  e2e {e['value']:.0f} img/s  {e['h2d_gbs']:.1f} GB/s H2D  raw ceiling {e['raw_h2d']['aggregate_gbs']:.1f} / mix {e['raw_h2d'].get('with_d2h_mix_aggregate_gbs', 0):.1f} GB/s"
              f"  frac {e['frac_of_raw_h2d']:.3f}  parity {e['parity']}")
    if "labels" in d:
        la = d["labels"]
        print(f"  labels {la['value'] / 1e9:.1f} G rows/s  {la['ms_per_step']:.4f} ms/step  kappa_g {la['kappa_general']:.6f}  e2e {la['e2e']['value'] / 1e9:.2f} G rows/s"
              f"  shuffled {la['shuffled_rows']['value'] / 1e9:.1f} G rows/s ok {la['shuffled_rows']['same_partials_as_sorted']}")
    for k, c in (d.get("configs") or {}).items():
        keys = [x for x in ("e2e_images_per_s", "e2e_h2d_gbs_per_gpu", "steady_h2d_gbs_per_gpu_min", "value", "us_per_step", "tally_kernel_us",
                            "kappa_general_repr", "partials_sha256", "cuda_graph", "ring_stalls", "label_rows_per_s_e2e") if k_in(c, x)] if False else \
            [x for x in ("e2e_images_per_s", "e2e_h2d_gbs_per_gpu", "steady_h2d_gbs_per_gpu_min", "value", "us_per_step", "tally_kernel_us",
                         "kappa_general_repr", "partials_sha256", "cuda_graph", "ring_stalls", "label_rows_per_s_e2e") if x in c]
        print(f"  {k}: " + "  ".join(f"{x}={c[x]:.4g}" if isinstance(c[x], float) else f"{x}={c[x]}" for x in keys) + f"  parity {c['parity']['ok']}")
        if "cpu" in c:
            print(f"      cpu: {c['cpu']}")
    if "cpu_baseline" in d:
        print(f"  cpu_baseline {d['cpu_baseline']['value']:.0f} img/s on {d['cpu_baseline']['cores']} cores; 1 thread {d['cpu_baseline']['single_thread_images_per_s']:.1f}")
