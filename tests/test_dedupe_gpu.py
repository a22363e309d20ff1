"""Parity, dedupe decision: device resolution against the sequential oracle
(webdav_sync.py:311-400 restated) — bit-exact flags, first/last occurrence and counts, with
duplicates, skipped entries, a pre-existing table and a global arrival order."""
import numpy as np
import pytest
import torch

from ics_b200 import engine
from ics_b200.ingest import hash_and_dedupe
from oracle import dedupe_batch, sha256_hex, synth_duplicate_map

pytestmark = pytest.mark.gpu


def _case(n, n_unique, p_skip, n_existing, seed):
    rng = np.random.default_rng(seed)
    base = [rng.integers(0, 256, size=int(rng.integers(0, 200)), dtype=np.uint8).tobytes() + bytes([i % 256, i // 256])
            for i in range(n_unique)]
    src = rng.integers(0, n_unique, size=n)
    datas = [base[s] if rng.random() >= p_skip else None for s in src]
    existing = {sha256_hex(base[i]) for i in rng.choice(n_unique, size=min(n_existing, n_unique), replace=False)}
    existing |= {sha256_hex(b"never seen %d" % i) for i in range(5)}
    return datas, existing


@pytest.mark.parametrize("n,n_unique,p_skip,n_existing", [(50, 40, 0.1, 5), (1000, 300, 0.05, 50), (7, 1, 0.0, 0),
                                                          (5000, 4000, 0.0, 0), (64, 64, 0.5, 64)])
def test_against_sequential_oracle(n, n_unique, p_skip, n_existing):
    datas, existing = _case(n, n_unique, p_skip, n_existing, seed=n)
    dec = hash_and_dedupe(datas, existing)
    hashes = [sha256_hex(d) if d is not None else None for d in datas]
    is_new, first, stats = dedupe_batch(hashes, existing)
    assert dec.hashes == hashes
    assert dec.is_new == is_new
    assert dec.first_index == first
    assert dec.stats == stats
    last = {}
    for i, h in enumerate(hashes):
        if h:
            last[h] = i
    assert dec.last_index == [last[h] if h else -1 for h in hashes]


def test_empty_and_all_skipped():
    assert hash_and_dedupe([]).stats == {"processed": 0, "created": 0, "updated": 0}
    d = hash_and_dedupe([None, None])
    assert d.stats == {"processed": 0, "created": 0, "updated": 0} and d.first_index == [-1, -1]


def test_config5_duplicate_rule_counts():
    """C5 duplicate rule at full count (10 000 images, 8 000 unique), on small stand-in contents:
    created = 8000, updated = 2000, and images >= 8000 are never new."""
    n, nu = 10_000, 8_000
    src = synth_duplicate_map(n, nu)
    rng = np.random.default_rng(5)
    contents = rng.integers(0, 256, size=(nu, 48), dtype=np.uint8)
    blob = torch.from_numpy(np.ascontiguousarray(contents[src])).cuda()
    off = torch.arange(n, dtype=torch.int64, device="cuda") * 48
    ln = torch.full((n,), 48, dtype=torch.int64, device="cuda")
    dig = engine.sha256_device(blob.view(-1), off, ln)
    is_new, first, last, counts = engine.dedupe_device(dig)
    assert counts.cpu().tolist() == [n, nu, n - nu]
    is_new = is_new.cpu().numpy()
    first = first.cpu().numpy()
    uniq_first = {}
    for i, s in enumerate(src):
        uniq_first.setdefault(int(s), i)
    assert np.array_equal(first, np.array([uniq_first[int(s)] for s in src]))
    assert np.array_equal(is_new.astype(bool), first == np.arange(n))


def test_global_order_via_seq():
    """Multi-GPU form: digests gathered in rank order but resolved by GLOBAL image index."""
    rng = np.random.default_rng(9)
    contents = rng.integers(0, 256, size=(20, 32), dtype=np.uint8)
    src = rng.integers(0, 20, size=200)
    gidx = rng.permutation(200).astype(np.int32)          # arrival order != storage order
    dig = torch.from_numpy(np.ascontiguousarray(contents[src])).cuda()
    is_new, first, last, counts = engine.dedupe_device(dig, seq=torch.from_numpy(gidx).cuda())
    is_new = is_new.cpu().numpy().astype(bool)
    for s in range(20):
        members = np.where(src == s)[0]
        if len(members) == 0:
            continue
        winner = members[np.argmin(gidx[members])]
        assert is_new[winner] and is_new[members].sum() == 1
    assert counts.cpu().tolist() == [200, len(set(src)), 200 - len(set(src))]


def test_lookup_sorted():
    rng = np.random.default_rng(10)
    table = rng.integers(0, 256, size=(500, 32), dtype=np.uint8)
    s = engine.sort_digests(table)
    q = np.concatenate([s[[0, 499, 250, 3]], rng.integers(0, 256, size=(4, 32), dtype=np.uint8)])
    out = engine.lookup_sorted_device(torch.from_numpy(q).cuda(), torch.from_numpy(s).cuda()).cpu().tolist()
    assert out == [0, 499, 250, 3, -1, -1, -1, -1]
    assert engine.lookup_sorted_device(torch.from_numpy(q).cuda(), None).cpu().tolist() == [-1] * 8


def test_config3_scale_two_million_digests_with_existing_table():
    """Config 3 scale after the all-gather (1 M images over 8 GPUs -> every rank resolves all of them; 2 M here):
    random 32-byte digests, 25 % duplicates, arrival order given by a permuted global index, 100 k of the
    distinct digests already stored.  Oracle: NumPy group-by on the digest bytes."""
    rng = np.random.default_rng(77)
    n, nu = 2_000_000, 1_500_000
    uniq = rng.integers(0, 256, size=(nu, 32), dtype=np.uint8)
    src = np.concatenate([np.arange(nu), rng.integers(0, nu, size=n - nu)])
    rng.shuffle(src)
    seq = rng.permutation(n).astype(np.int32)
    stored_ids = rng.choice(nu, size=100_000, replace=False)
    table = engine.sort_digests(uniq[stored_ids])
    dig = torch.from_numpy(np.ascontiguousarray(uniq[src])).cuda()
    is_new, first, last, counts = engine.dedupe_device(dig, seq=torch.from_numpy(seq).cuda(),
                                                       existing_sorted=torch.from_numpy(table).cuda())
    is_new, first, last = is_new.cpu().numpy().astype(bool), first.cpu().numpy(), last.cpu().numpy()
    # oracle: per source id, the member with the smallest / largest seq
    order = np.lexsort((seq, src))
    s_sorted = src[order]
    starts = np.r_[0, np.flatnonzero(s_sorted[1:] != s_sorted[:-1]) + 1]
    ends = np.r_[starts[1:], n]
    first_of = np.empty(nu, dtype=np.int64)
    last_of = np.empty(nu, dtype=np.int64)
    present = s_sorted[starts]
    first_of[present] = order[starts]
    last_of[present] = order[ends - 1]
    assert np.array_equal(first, first_of[src]) and np.array_equal(last, last_of[src])
    stored = np.zeros(nu, dtype=bool)
    stored[stored_ids] = True
    want_new = (first_of[src] == np.arange(n)) & ~stored[src]
    assert np.array_equal(is_new, want_new)
    assert counts.cpu().tolist() == [n, int(want_new.sum()), n - int(want_new.sum())]


def test_first_and_last_index_are_optional():
    """include/b2ingest.h: d_first_index / d_last_index may each be NULL."""
    import ctypes as C

    from ics_b200._lib import check, lib
    rng = np.random.default_rng(2)
    base = rng.integers(0, 256, size=(40, 32), dtype=np.uint8)
    dig = torch.from_numpy(base[rng.integers(0, 40, 300)]).cuda()
    is_new, first, last, counts = engine.dedupe_device(dig)
    new2 = torch.empty_like(is_new)
    counts2 = torch.empty_like(counts)
    ws = torch.empty(int(lib.b2_dedupe_workspace_bytes(300)) // 8, dtype=torch.int64, device="cuda")
    check(lib.b2_dedupe(dig.data_ptr(), None, None, 300, None, C.c_uint64(0), new2.data_ptr(), None, None, counts2.data_ptr(),
                        ws.data_ptr(), ws.numel() * 8, torch.cuda.current_stream().cuda_stream))
    assert torch.equal(new2, is_new) and torch.equal(counts2, counts)
