"""Parity, content hash: the CUDA kernel (through the C ABI) against hashlib — the reference's
own arithmetic (webdav_sync.py:59) — bit-exact, on the FIPS known answers, every padding
boundary, ragged / empty / unaligned inputs and both load paths."""
import hashlib
import os

import numpy as np
import pytest
import torch

import ics_b200
from ics_b200 import engine
from oracle import sha256_hex, synth_image

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["iadd3", "imad", "pair"])
def sha_path(request, monkeypatch):
    """The three kernels: one lane per message with the adds as ptxas schedules them (what one warp per
    sub-partition gets) or with two-input adds on the FMA pipe (what larger batches get), and the warp-pair
    kernel (schedule warp + rounds warp per 32 messages; what batches that leave sub-partitions idle get)."""
    monkeypatch.setenv("B2_SHA_PAIR", "1" if request.param == "pair" else "0")
    monkeypatch.setenv("B2_SHA_VARIANT", {"iadd3": "0", "imad": "2", "pair": "0"}[request.param])
    return request.param


def test_known_answers(sha_kat, sha_path):
    msgs = [k["msg_ascii"].encode() * k["repeat"] for k in sha_kat["kat"]]
    assert ics_b200.hash_batch(msgs) == [k["hex"] for k in sha_kat["kat"]]


def test_padding_boundaries(sha_kat, sha_path):
    rng = np.random.default_rng(11)
    msgs = [rng.integers(0, 256, size=n, dtype=np.uint8).tobytes() for n in sha_kat["boundary_lengths"]]
    assert ics_b200.hash_batch(msgs) == [sha256_hex(m) for m in msgs]


def test_ragged_batch_random_lengths(sha_path):
    rng = np.random.default_rng(12)
    lens = list(rng.integers(0, 5000, size=300)) + [0, 0, 1, 65536, 100_003]
    msgs = [rng.integers(0, 256, size=int(n), dtype=np.uint8).tobytes() for n in lens]
    got = ics_b200.hash_batch(msgs)
    assert got == [hashlib.sha256(m).hexdigest() for m in msgs]
    assert all(len(h) == 64 and h == h.lower() for h in got)


def test_empty_batch_and_single():
    assert ics_b200.hash_batch([]) == []
    assert ics_b200.hash_batch([b""]) == ["e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"]


def test_unaligned_message_starts(sha_path):
    """Device-level call with message starts that are NOT 16-byte aligned (byte-load path)."""
    rng = np.random.default_rng(13)
    blob = rng.integers(0, 256, size=20_000, dtype=np.uint8)
    offs = np.array([1, 3, 1000, 4097, 9999, 16], dtype=np.int64)
    lens = np.array([777, 64, 0, 5000, 129, 55], dtype=np.int64)
    d = engine.sha256_device(torch.from_numpy(blob).cuda(), torch.from_numpy(offs).cuda(), torch.from_numpy(lens).cuda())
    got = d.cpu().numpy()
    for i, (o, l) in enumerate(zip(offs, lens)):
        assert bytes(got[i]) == hashlib.sha256(blob[o:o + l].tobytes()).digest()


def test_config1_images_512(sha_path):
    """BASELINE config 1 shape: 512x512x3 synthetic images (a sample of 40 keeps the oracle fast)."""
    imgs = [synth_image(g, 512, 512).tobytes() for g in range(40)]
    assert ics_b200.hash_batch(imgs) == [sha256_hex(b) for b in imgs]


def test_order_permutation_is_transparent():
    rng = np.random.default_rng(14)
    msgs = [rng.integers(0, 256, size=int(n), dtype=np.uint8).tobytes() for n in rng.integers(0, 3000, size=100)]
    p = engine.PackedMessages(msgs)
    data, off, ln, order = p.to_device()
    a = engine.sha256_device(data, off, ln, order).cpu().numpy()
    b = engine.sha256_device(data, off, ln, None).cpu().numpy()
    perm = torch.randperm(100).to(torch.int32).cuda()
    c = engine.sha256_device(data, off, ln, perm).cpu().numpy()
    assert np.array_equal(a, b) and np.array_equal(a, c)


def test_full_size_property_duplicates_and_checksum():
    """Full BASELINE config-2 image size (1920x1080x3) generated on device: a byte-copy hashes
    identically, a one-bit flip does not, and a sampled image matches hashlib."""
    n, L = 96, 1920 * 1080 * 3
    g = torch.Generator(device="cuda").manual_seed(0xB200)
    data = torch.randint(0, 256, (n, L), dtype=torch.uint8, device="cuda", generator=g)
    data[n - 1] = data[0]
    data[n - 2] = data[1]
    data[n - 2, L // 2] ^= 1
    off = (torch.arange(n, dtype=torch.int64, device="cuda") * L)
    ln = torch.full((n,), L, dtype=torch.int64, device="cuda")
    dig = engine.sha256_device(data.view(-1), off, ln).cpu().numpy()
    assert bytes(dig[n - 1]) == bytes(dig[0])
    assert bytes(dig[n - 2]) != bytes(dig[1])
    for i in (0, 1, n - 2, 50):
        assert bytes(dig[i]) == hashlib.sha256(data[i].cpu().numpy().tobytes()).digest()
    hexes = engine.hex_strings(engine.digest_hex_device(torch.from_numpy(dig).cuda()))
    assert hexes[50] == hashlib.sha256(data[50].cpu().numpy().tobytes()).hexdigest()


def test_host_entry_points_match_device_path_and_oracle():
    """b2_sha256_host / b2_dedupe_host (host pointers only) against hashlib and the sequential oracle, incl.
    empty messages, non-bytes buffers, skipped entries and a table of stored digests."""
    import hashlib
    from oracle import dedupe_batch
    rng = np.random.default_rng(11)
    msgs = [bytes(rng.integers(0, 256, size=int(n), dtype=np.uint8)) for n in
            [0, 1, 55, 56, 63, 64, 65, 119, 120, 4096, 100_003, 0, 1 << 20]]
    msgs += [msgs[4], bytearray(msgs[9]), memoryview(msgs[10])]
    digests, hexes = engine.sha256_host(msgs)
    want = [hashlib.sha256(bytes(m)).hexdigest() for m in msgs]
    assert hexes == want and [bytes(d).hex() for d in digests] == want
    d2, none = engine.sha256_host(msgs, want_hex=False)
    assert none is None and np.array_equal(d2, digests)
    assert engine.sha256_host([]) [1] == []
    valid = np.ones(len(msgs), dtype=np.uint8)
    valid[[2, 7]] = 0
    stored = {want[1], want[12], hashlib.sha256(b"elsewhere").hexdigest()}
    table = engine.sort_digests(np.frombuffer(bytes.fromhex("".join(sorted(stored))), dtype=np.uint8).reshape(-1, 32))
    is_new, first, last, counts = engine.dedupe_host(digests, valid, table)
    hashes = [h if v else None for h, v in zip(want, valid)]
    o_new, o_first, o_stats = dedupe_batch(hashes, stored)
    assert [bool(x) for x in is_new] == o_new and first.tolist() == o_first
    assert dict(zip(("processed", "created", "updated"), counts)) == o_stats
