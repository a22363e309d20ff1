"""Re-entrancy: the reference reaches the path from up to five threads at once (initial
sync + its WebDAV and Activity workers, the scheduler's two loops), each with its own session, plus request
threads.  Five threads hammer every entry point concurrently; every result must still match the oracle."""
import threading

import numpy as np
import pytest
import torch

import ics_b200
from ics_b200 import engine, labels
from ics_b200.pipeline import IngestPipeline
from oracle import dedupe_batch, label_tally, sha256_hex, synth_image, synth_label_rows, thumbnail_u8

pytestmark = pytest.mark.gpu


def _worker(tid: int, rounds: int, errors: list):
    try:
        rng = np.random.default_rng(100 + tid)
        shape = [(64, 80), (96, 128), (128, 160), (100, 52), (256, 256)][tid % 5]
        images = [synth_image(1000 * tid + g, *shape) for g in range(12)]
        images.append(images[3].copy())
        want_hashes = [sha256_hex(im.tobytes()) for im in images]
        want_new, want_first, want_stats = dedupe_batch(want_hashes)
        want_thumb = thumbnail_u8(images[5], 256, 256)
        n_img, k = 700 + 13 * tid, [3, 50, 64, 129, 17][tid % 5]
        img, cls, act = synth_label_rows(n_img, k, 9 + tid)
        want_counts = label_tally(img, cls, act, n_img, k)
        pipe = None
        if True:                                            # any shape: the ring pads image starts itself
            pipe = IngestPipeline(shape[0], shape[1], len(images), chunk_images=4)
            host = torch.empty((len(images), shape[0] * shape[1] * 3), dtype=torch.uint8, pin_memory=True)
            host.copy_(torch.from_numpy(np.stack([im.reshape(-1) for im in images])))
        for _ in range(rounds):
            blobs = [bytes(rng.integers(0, 256, size=int(rng.integers(0, 5000)), dtype=np.uint8)) for _ in range(7)]
            assert ics_b200.hash_batch(blobs) == [sha256_hex(b) for b in blobs]
            res = ics_b200.ingest_batch([im.tobytes() for im in images], decoded_rgb=images)
            assert res.decision.hashes == want_hashes and res.decision.is_new == want_new
            assert res.decision.first_index == want_first and res.decision.stats == want_stats
            assert np.array_equal(res.thumbs[5], want_thumb)
            t = labels.label_tally(img, cls, act, n_img, k)                     # host-pointer entry point
            assert np.array_equal(t.counts, want_counts)
            d = [torch.from_numpy(a).cuda() for a in (img, cls, act)]
            c2, p2 = engine.label_tally_device(d[0], d[1], d[2], n_img, k)      # device-pointer entry point
            assert np.array_equal(c2.cpu().numpy(), want_counts)
            if pipe is not None:
                r = pipe.run(host)
                assert [bytes(x).hex() for x in r.digests] == want_hashes and r.stats == want_stats
                assert np.array_equal(r.thumbs[5], want_thumb)
            with pytest.raises(ics_b200.B2Error):                               # errors stay thread-local
                labels.label_tally(img[::-1].copy(), cls, act, n_img, k, sorted_by_image=True)
    except BaseException as e:  # noqa: BLE001 - reported by the main thread
        errors.append((tid, repr(e)))


def test_five_threads_every_entry_point():
    errors: list = []
    threads = [threading.Thread(target=_worker, args=(t, 6, errors)) for t in range(5)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in threads), "a worker hung"
    assert not errors, errors


def test_concurrent_single_file_hashes_share_launches():
    """The reference's call sites hash ONE file per call (webdav_sync.py:59, activity_api_sync.py:798, images.py:62) from
    up to five service threads: concurrent calls are merged into shared device launches, results stay per caller."""
    import hashlib

    from ics_b200 import hostapi
    hostapi.hash_batch([b"warm-up"])
    c = hostapi._coalescers[hostapi.init(None)]
    calls0, launches0 = c.calls, c.launches
    errors: list = []

    def worker(t):
        try:
            rng = np.random.default_rng(t)
            for i in range(40):
                blob = bytes(rng.integers(0, 256, size=int(rng.integers(1, 300_000)), dtype=np.uint8))
                assert hostapi.hash_batch([blob]) == [hashlib.sha256(blob).hexdigest()]
        except BaseException as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(5)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert c.calls - calls0 == 200
    assert c.launches - launches0 < 200                     # some launches carried several callers' files
    with pytest.raises(Exception):
        hostapi.hash_batch([None])                          # an error reaches the caller that caused it
    assert hostapi.hash_batch([b"abc"]) == ["ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"]


def test_coalesced_hashes_keep_errors_with_their_caller():
    """A caller that hands in garbage fails alone: calls merged into the same launch still get their digests."""
    import hashlib

    from ics_b200 import hostapi
    hostapi.hash_batch([b"warm-up"])
    results, errors = {}, {}
    gate = threading.Barrier(4)

    def good(i):
        gate.wait()
        for j in range(30):
            blob = bytes([i, j]) * 1000
            results[(i, j)] = hostapi.hash_batch([blob]) == [hashlib.sha256(blob).hexdigest()]

    def bad():
        gate.wait()
        for _ in range(30):
            try:
                hostapi.hash_batch([None])
                errors["no-raise"] = True
            except Exception:  # noqa: BLE001
                errors["raised"] = errors.get("raised", 0) + 1

    threads = [threading.Thread(target=good, args=(i,)) for i in range(3)] + [threading.Thread(target=bad)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert len(results) == 90 and all(results.values())
    assert errors == {"raised": 30}
