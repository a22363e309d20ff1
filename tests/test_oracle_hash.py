"""Oracle pinning, content hash: hashlib (the reference's own arithmetic) and the pure-Python
FIPS 180-4 restatement against published known answers."""
import hashlib

import numpy as np
import pytest

from oracle import sha256_hex, sha256_restated, sha256_digest


def _msg(k):
    return k["msg_ascii"].encode("ascii") * k["repeat"]


def test_kat_hashlib(sha_kat):
    for k in sha_kat["kat"]:
        assert sha256_hex(_msg(k)) == k["hex"]


def test_kat_restated(sha_kat):
    for k in sha_kat["kat"]:
        if k["repeat"] > 1000:
            continue                      # pure-Python loops: small cases only
        assert sha256_restated(_msg(k)) == k["hex"]


def test_restated_matches_hashlib_on_padding_boundaries(sha_kat):
    rng = np.random.default_rng(1)
    for n in sha_kat["boundary_lengths"]:
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert sha256_restated(data) == hashlib.sha256(data).hexdigest() == sha256_hex(data)
        assert sha256_digest(data).hex() == sha256_hex(data)


def test_reference_singles(ref_ingest):
    """Hashes the reference's own _download_and_process_image produced (golden) == oracle."""
    import base64
    for info, single in zip(ref_ingest["infos"], ref_ingest["singles"]):
        if single["hash"] is None:
            assert ref_ingest["failures"].get(info["path"]) is not None
            continue
        data = base64.b64decode(ref_ingest["files"][info["path"]])
        assert sha256_hex(data) == single["hash"]
        assert len(single["hash"]) == 64 and single["hash"] == single["hash"].lower()
