"""Download / decode feeder: host logic only, no GPU.  The fake client sleeps instead
of doing HTTP, so overlap and the in-flight bound are observable."""
import io
import threading
import time

import numpy as np
import pytest
from PIL import Image

from ics_b200.feeder import DownloadDecodeFeeder, decode_rgb, image_metadata


def _png(seed, h=24, w=32, mode="RGB"):
    rng = np.random.default_rng(seed)
    arr = rng.integers(0, 256, (h, w, 3 if mode == "RGB" else 4), dtype=np.uint8)
    buf = io.BytesIO()
    Image.fromarray(arr, mode).save(buf, "PNG")
    return buf.getvalue(), arr


class _Client:
    def __init__(self, files, delay=0.0, fail=()):
        self.files, self.delay, self.fail = files, delay, set(fail)
        self.lock, self.active, self.peak_active, self.calls = threading.Lock(), 0, 0, []

    def fetch(self, info):
        with self.lock:
            self.active += 1
            self.peak_active = max(self.peak_active, self.active)
            self.calls.append(info["name"])
        try:
            time.sleep(self.delay)
            if info["name"] in self.fail:
                raise ConnectionError("boom")
            return self.files[info["name"]]
        finally:
            with self.lock:
                self.active -= 1


def _listing(n):
    files, arrays = {}, {}
    for i in range(n):
        files[f"img{i:03d}.png"], arrays[f"img{i:03d}.png"] = _png(i)
    return [{"name": k, "path": "/set/" + k} for k in files], files, arrays


def test_order_content_and_decode():
    infos, files, arrays = _listing(23)
    client = _Client(files, delay=0.002)
    feeder = DownloadDecodeFeeder(client.fetch, batch_size=5, download_workers=8, decode=True)
    seen = []
    for b in feeder.batches(infos):
        assert len(b.infos) == len(b.datas) == len(b.metadata) == len(b.rgb) <= 5
        for info, data, meta, rgb in zip(b.infos, b.datas, b.metadata, b.rgb):
            assert data == files[info["name"]]
            assert meta == {"width": 32, "height": 24, "format": "PNG", "mode": "RGB"}
            assert rgb.dtype == np.uint8 and rgb.flags["C_CONTIGUOUS"] and np.array_equal(rgb, arrays[info["name"]])
            seen.append(info["name"])
    assert seen == [i["name"] for i in infos]                       # listing order, every image exactly once
    assert sorted(client.calls) == sorted(seen)


def test_failed_download_invalid_entry_and_undecodable_file():
    infos, files, _ = _listing(6)
    files["img002.png"] = b"not an image at all"
    infos[4]["name"] = "img004.txt"                                  # fails the extension filter: never fetched
    client = _Client(files, fail={"img001.png"})
    feeder = DownloadDecodeFeeder(client.fetch, validate=lambda i: i["name"].endswith(".png"), batch_size=6,
                                  download_workers=3, decode=True)
    (b,) = list(feeder.batches(infos))
    assert [d is None for d in b.datas] == [False, True, False, False, True, False]
    assert b.metadata[1] == {} and b.rgb[1] is None                  # failed download: (None, {}) as webdav_sync.py:320
    assert b.datas[2] == b"not an image at all" and b.metadata[2] == {} and b.rgb[2] is None
    assert "img004.txt" not in client.calls


def test_rgba_and_palette_files_decode_to_rgb():
    data, arr = _png(9, mode="RGBA")
    assert image_metadata(data)["mode"] == "RGBA"
    assert np.array_equal(decode_rgb(data), arr[..., :3])
    pal = io.BytesIO()
    Image.fromarray(arr[..., :3], "RGB").convert("P").save(pal, "PNG")
    out = decode_rgb(pal.getvalue())
    assert out.shape == (24, 32, 3) and out.dtype == np.uint8


def test_downloads_overlap_and_next_batch_is_prefetched():
    infos, files, _ = _listing(32)
    client = _Client(files, delay=0.02)
    t0 = time.perf_counter()
    feeder = DownloadDecodeFeeder(client.fetch, batch_size=8, download_workers=8, prefetch_batches=1, want_metadata=False)
    consumed = 0
    for b in feeder.batches(infos):
        time.sleep(0.02)                                             # the consumer's device call + table writes
        consumed += len(b.datas)
    dt = time.perf_counter() - t0
    assert consumed == 32 and client.peak_active > 1
    assert dt < 0.55                                                 # sequential: 0.64 s of sleeps + 4 x 0.02 s of consumer = 0.72 s


@pytest.mark.parametrize("prefetch", [0, 1, 3])
def test_inflight_files_are_bounded(prefetch):
    infos, files, _ = _listing(40)
    client = _Client(files, delay=0.001)
    feeder = DownloadDecodeFeeder(client.fetch, batch_size=4, download_workers=16, prefetch_batches=prefetch,
                                  want_metadata=False)
    n = sum(len(b.datas) for b in feeder.batches(infos))
    assert n == 40
    assert feeder.peak_inflight_files <= (prefetch + 1) * 4


def test_empty_listing_and_bad_arguments():
    assert list(DownloadDecodeFeeder(lambda i: b"").batches([])) == []
    with pytest.raises(ValueError):
        DownloadDecodeFeeder(lambda i: b"", batch_size=0)
