"""world_size-2 gloo run of the multi-GPU plumbing (dist.py) on CPU tensors: the two collectives
carry integers only, so what is checked here (sharding covers every unit once; all-reduced
partials equal the single-process partials; gathered digests come back in rank order) is
exactly what NCCL does on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fleiss_kappa, fleiss_partials, label_tally, synth_label_rows


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
                                                      p["pairs"], r1 - r0, 0], dtype=torch.int64)
        d.allreduce_partials(vec)
        dig = torch.full((3, 32), rank, dtype=torch.uint8)
        gi = torch.arange(3, dtype=torch.int32) + 3 * rank
        all_d, all_i = d.allgather_digests(dig, gi)
        # size-aware sharding of a mixed-size listing: computed independently on every rank, must be the same split
        rng = np.random.default_rng(3)
        lengths = ((256 << (rng.permutation(500) % 5)).astype(np.int64) ** 2) * 3
        mine = d.shard_by_bytes(lengths, ws)[rank]
        owned = torch.zeros(500, dtype=torch.int64)
        owned[torch.from_numpy(mine)] = 1
        nbytes = torch.tensor([int(lengths[mine].sum())], dtype=torch.int64)
        both = [torch.zeros(1, dtype=torch.int64) for _ in range(ws)]
        dist.all_reduce(owned)                       # every image owned by exactly one rank
        dist.all_gather(both, nbytes)
        assert bool((owned == 1).all())
        assert abs(int(both[0]) - int(both[1])) <= int(lengths.max())
        if rank == 0:
            kappa = labels.fleiss_kappa(vec[:k].numpy(), int(vec[k]), int(vec[k + 1]), n_images, n_r)
            out.put((vec.tolist(), all_d[:, 0].tolist(), all_i.tolist(), kappa))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_partials_and_gather():
    ws, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    vec, dig0, gi, kappa = q.get(timeout=100)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    n_images, k, n_r = 101, 6, 9
    img, cls, act = synth_label_rows(n_images, k, n_r)
    full = fleiss_partials(label_tally(img, cls, act, n_images, k))
    assert vec[:k] == list(full["class_totals"])
    assert vec[k:k + 5] == [full["S2"], full["R"], full["n_rated"], full["n_pairs_images"], full["pairs"]]
    assert vec[k + 5] == len(img)
    assert dig0 == [0, 0, 0, 1, 1, 1] and gi == [0, 1, 2, 3, 4, 5]
    # general-n data, but the integer route must still agree bit for bit with a single process
    assert kappa == fleiss_kappa(full["class_totals"], full["S2"], full["R"], n_images, n_r)
