"""Two ranks, two GPUs, NCCL inside the C ABI (csrc/comm.cu) — no torch.distributed anywhere: the unique id travels
through a file (`Comm.from_file`, the rendezvous INTEGRATION.md shows), PyTorch only provides the device buffers.
Checked against the sequential oracle over the WHOLE listing: unequal shards (dist.shard_by_bytes), duplicates whose
first occurrence lives on the other rank, a stored digest, an invalid entry; and the integer all-reduce of tally
partials + agreement histogram against the single-process tally.  Needs >= 2 GPUs (skipped otherwise; run with
`gpurun --gpus 2 -- python -m pytest tests/test_comm_gpu.py -m gpu`)."""
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

RANK = r'''
import sys, json
import numpy as np
import torch
sys.path.insert(0, %(root)r)
import ics_b200
from ics_b200 import dist as d, engine, labels
from ics_b200.hostapi import sort_digests
from oracle import (agreement_hist, dedupe_batch, fleiss_kappa_general, fleiss_partials, label_tally, synth_duplicate_map,
                    synth_label_rows)
rank, world, path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
torch.cuda.set_device(rank)
engine.init(rank)
c = d.Comm.from_file(path, rank, world, rank)
d.set_comm(c)
assert c.nccl_version() >= 21800

# ---- cross-rank dedupe of a sharded mixed-size listing
n = 1003
rng = np.random.default_rng(21)
lengths = ((256 << rng.integers(0, 5, n)).astype(np.int64) ** 2) * 3
shards = d.shard_by_bytes(lengths, world)
mine = shards[rank]
n_max = max(len(s) for s in shards)
assert len(set(len(s) for s in shards)) > 1                      # unequal shards
src = synth_duplicate_map(n, n - n // 5)
all_digests = np.random.default_rng(5).integers(0, 256, size=(n, 32), dtype=np.uint8)[src]
valid_all = np.ones(n, dtype=np.uint8)
valid_all[[3, 500, 1002]] = 0
stored = all_digests[[10, 700]]
hexes = [bytes(x).hex() if v else None for x, v in zip(all_digests, valid_all)]
want_new, want_first, want_stats = dedupe_batch(hexes, {bytes(x).hex() for x in stored})
dig = torch.from_numpy(all_digests[mine]).cuda()
seq = torch.from_numpy(mine.astype(np.int32)).cuda()
val = torch.from_numpy(valid_all[mine]).cuda()
table = torch.from_numpy(sort_digests(stored)).cuda()
is_new, first_seq, last_seq, counts = d.global_dedupe(dig, seq, n_max, valid=val, existing_sorted=table)
torch.cuda.synchronize()
assert counts.cpu().tolist() == [want_stats["processed"], want_stats["created"], want_stats["updated"]], counts
assert is_new.cpu().numpy().astype(bool).tolist() == [want_new[i] for i in mine]
assert first_seq.cpu().tolist() == [want_first[i] for i in mine]
last = {}
for i, h in enumerate(hexes):
    if h is not None:
        last[h] = i
assert last_seq.cpu().tolist() == [last[hexes[i]] if hexes[i] is not None else -1 for i in mine]
# n_max exchanged instead of given: same answer
is_new2, _, _, counts2 = d.global_dedupe(dig, seq, None, valid=val, existing_sorted=table)
assert torch.equal(is_new2, is_new) and torch.equal(counts2, counts)
# plain all-gather of equal-sized digest blocks comes back in rank order
blk = torch.full((4, 32), rank, dtype=torch.uint8, device="cuda")
got = c.allgather_digests(blk)
torch.cuda.synchronize()
assert got[:, 0].cpu().tolist() == [r for r in range(world) for _ in range(4)]

# ---- label partials + agreement histogram: shard by image range, all-reduce, compare with one process
n_images, k, n_r = 9001, 50, 12
img, cls, act = synth_label_rows(n_images, k, n_r, p_active=0.8)
lo, hi, r0, r1 = d.shard_rows_by_image(img, n_images, rank, world)
buf = torch.empty(k + 7 + 1024, dtype=torch.int64, device="cuda")
dcounts, _ = engine.label_tally_device(torch.from_numpy(img[r0:r1]).cuda(), torch.from_numpy(cls[r0:r1]).cuda(),
                                       torch.from_numpy(act[r0:r1]).cuda(), hi - lo, k, lo, True, None, buf[:k + 7], buf[k + 7:])
d.allreduce_partials(buf)
p = buf.cpu().numpy()
counts_ref = label_tally(img, cls, act, n_images, k)
full = fleiss_partials(counts_ref)
assert p[:k].tolist() == list(full["class_totals"])
assert p[k:k + 5].tolist() == [full["S2"], full["R"], full["n_rated"], full["n_pairs_images"], full["pairs"]]
assert int(p[k + 5]) == img.size and int(p[k + 6]) == 0
assert np.array_equal(p[k + 7:], agreement_hist(counts_ref))
assert np.array_equal(dcounts.cpu().numpy(), counts_ref[lo:hi])
kg = labels.fleiss_kappa_from_hist(p[:k], int(p[k + 1]), int(p[k + 3]), p[k + 7:])
assert abs(kg - fleiss_kappa_general(counts_ref)) <= 1e-12 * abs(kg)

# ---- the same all-reduce over NVLink peer memory: standalone and fused into the tally kernel; many epochs (the
# mailboxes alternate between two parities), a second vector size, and the NCCL result as the reference
c.enable_peer_reduce(k + 7 + 1024)
want = torch.from_numpy(p.copy()).cuda()
for it in range(25):
    v = torch.arange(37, dtype=torch.int64, device="cuda") * (rank + 1) + it
    c.peer_allreduce_i64(v)
    assert v.cpu().tolist() == [sum(i * (r + 1) + it for r in range(world)) for i in range(37)], it
    buf2 = torch.empty(k + 7 + 1024, dtype=torch.int64, device="cuda")
    dc2 = torch.empty((hi - lo, k), dtype=torch.int32, device="cuda")
    c.label_tally_reduce(torch.from_numpy(img[r0:r1]).cuda(), torch.from_numpy(cls[r0:r1]).cuda(),
                         torch.from_numpy(act[r0:r1]).cuda(), hi - lo, k, lo, dc2, buf2)
    assert torch.equal(buf2, want), it
    assert torch.equal(dc2, dcounts)
# any-order rows: falls back to the scatter tally + NCCL, same integers except the two order-check entries
perm = np.random.default_rng(3).permutation(r1 - r0)
buf3 = torch.empty(k + 7 + 1024, dtype=torch.int64, device="cuda")
c.label_tally_reduce(torch.from_numpy(img[r0:r1][perm]).cuda(), torch.from_numpy(cls[r0:r1][perm]).cuda(),
                     torch.from_numpy(act[r0:r1][perm]).cuda(), hi - lo, k, lo, dc2, buf3, sorted_by_image=False)
assert torch.equal(buf3[:k + 6], want[:k + 6]) and torch.equal(buf3[k + 7:], want[k + 7:])
torch.cuda.synchronize()
assert not c.peer_timed_out()
c.close()
print(json.dumps({"rank": rank, "kappa": repr(kg), "counts": counts.cpu().tolist()}))
'''


@pytest.mark.timeout(300)
def test_two_ranks_nccl_in_the_c_abi():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "nccl_id")
        procs = [subprocess.Popen([sys.executable, "-c", RANK % {"root": ROOT}, str(r), "2", path], cwd=ROOT,
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
        outs = [p.communicate(timeout=280) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-3000:]
    lines = [so.strip().splitlines()[-1] for so, _ in outs]
    import json
    a, b = (json.loads(x) for x in lines)
    assert a["kappa"] == b["kappa"] and a["counts"] == b["counts"]      # bit-identical on both ranks


def test_comm_errors_are_loud():
    import ctypes as C

    from ics_b200._lib import lib
    h = C.c_void_p()
    assert lib.b2_comm_init(0, 2, 2, None, C.byref(h)) == -1             # rank out of range / null id
    assert lib.b2_allreduce_i64(None, None, C.c_uint64(1), None) == -1
    assert lib.b2_comm_destroy(None) == 0
    assert lib.b2_dedupe_global_workspace_bytes(2, 100) > 2 * 100 * 32
