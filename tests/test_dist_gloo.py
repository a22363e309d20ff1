"""world_size-2 gloo run of the multi-GPU host logic (dist.py) on CPU tensors: the exchanges carry integers
only, so what is checked here (sharding covers every unit once; all-reduced partials and agreement histogram equal
the single-process ones; UNEQUAL shards are padded with invalid entries and the listing-position rule resolves
duplicates that live on different ranks exactly as the sequential reference loop does) is what b2_dedupe_global /
b2_allreduce_i64 do with NCCL on the GPU box (tests/test_comm_gpu.py runs those on two GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import (agreement_hist, dedupe_batch, fleiss_kappa, fleiss_kappa_general, fleiss_partials, label_tally,
                    synth_duplicate_map, synth_label_rows)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from ics_b200 import dist as d, labels
        n_images, k, n_r = 101, 6, 9
        img, cls, act = synth_label_rows(n_images, k, n_r)
        lo, hi, r0, r1 = d.shard_rows_by_image(img, n_images, rank, ws)
        # this rank's partials (CPU stand-in for the device tally: same integer definition)
        counts = label_tally(img[r0:r1] - lo, cls[r0:r1], act[r0:r1], hi - lo, k)
        p = fleiss_partials(counts)
        vec = torch.tensor(list(p["class_totals"]) + [p["S2"], p["R"], p["n_rated"], p["n_pairs_images"],
                                                      p["pairs"], r1 - r0, 0] + agreement_hist(counts).tolist(),
                           dtype=torch.int64)
        d.allreduce_partials(vec)
        # size-aware sharding of a mixed-size listing: computed independently on every rank, must be the same split
        rng = np.random.default_rng(3)
        lengths = ((256 << rng.integers(0, 5, 500)).astype(np.int64) ** 2) * 3
        mine = d.shard_by_bytes(lengths, ws)[rank]
        owned = torch.zeros(500, dtype=torch.int64)
        owned[torch.from_numpy(mine)] = 1
        nbytes = torch.tensor([int(lengths[mine].sum())], dtype=torch.int64)
        both = [torch.zeros(1, dtype=torch.int64) for _ in range(ws)]
        dist.all_reduce(owned)                       # every image owned by exactly one rank
        dist.all_gather(both, nbytes)
        assert bool((owned == 1).all())
        assert abs(int(both[0]) - int(both[1])) <= int(lengths.max())
        # cross-rank dedupe of that listing: 20 % of the entries are byte copies of earlier ones (config 5's rule), the
        # shards have different sizes, most duplicates live on another rank than their first occurrence
        digests = _listing_digests(500)
        all_d, all_s, valid, (lo, hi) = d.gather_shards(torch.from_numpy(digests[mine]), torch.from_numpy(mine))
        assert len(mine) != 250 and all_d.shape[0] == 2 * max(len(mine), 500 - len(mine))
        assert int(valid.sum()) == 500 and torch.equal(all_s[lo:hi], torch.from_numpy(mine))
        # the resolution every rank runs on the gathered arrays (on the device: b2_dedupe keyed on seq, pads invalid);
        # stated here with the sequential oracle: valid entries in listing order
        keep = valid.numpy().astype(bool)
        order = np.argsort(all_s.numpy()[keep], kind="stable")
        hexes = [bytes(x).hex() for x in all_d.numpy()[keep][order]]
        is_new, first, stats = dedupe_batch(hexes)
        if rank == 0:
            kappa = labels.fleiss_kappa(vec[:k].numpy(), int(vec[k]), int(vec[k + 1]), n_images, n_r)
            kg = labels.fleiss_kappa_from_hist(vec[:k].numpy(), int(vec[k + 1]), int(vec[k + 3]), vec[k + 7:].numpy())
            out.put((vec.tolist(), stats, [bool(x) for x in is_new], kappa, kg))
    finally:
        dist.destroy_process_group()


def _listing_digests(n):
    src = synth_duplicate_map(n, n - n // 5)
    base = np.random.default_rng(11).integers(0, 256, size=(n, 32), dtype=np.uint8)
    return base[src]


@pytest.mark.timeout(120)
def test_world_size_2_partials_and_gather():
    ws, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    vec, stats, is_new, kappa, kg = q.get(timeout=100)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    n_images, k, n_r = 101, 6, 9
    img, cls, act = synth_label_rows(n_images, k, n_r)
    full = fleiss_partials(label_tally(img, cls, act, n_images, k))
    assert vec[:k] == list(full["class_totals"])
    assert vec[k:k + 5] == [full["S2"], full["R"], full["n_rated"], full["n_pairs_images"], full["pairs"]]
    assert vec[k + 5] == len(img)
    counts = label_tally(img, cls, act, n_images, k)
    assert vec[k + 7:] == agreement_hist(counts).tolist()
    # the integer route agrees bit for bit with a single process
    assert kappa == fleiss_kappa(full["class_totals"], full["S2"], full["R"], n_images, n_r)
    assert abs(kg - fleiss_kappa_general(counts)) <= 1e-12 * abs(kg)
    # cross-rank duplicates: the gathered, padded shards resolve exactly as the sequential loop over the listing
    want_new, _, want_stats = dedupe_batch([bytes(x).hex() for x in _listing_digests(500)])
    assert stats == want_stats == {"processed": 500, "created": 400, "updated": 100} and is_new == want_new
