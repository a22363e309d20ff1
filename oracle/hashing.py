"""Oracle: content hash.  TEST INFRASTRUCTURE — see oracle/__init__.py.

Reference call sites (all the same expression, over the raw downloaded file bytes):
  app/services/webdav_sync.py:49-59   WebDAVSync._calculate_hash_from_bytes
  app/services/activity_api_sync.py:798
  app/api/routes/images.py:62
"""
from __future__ import annotations

import hashlib
import struct
from typing import Union

BytesLike = Union[bytes, bytearray, memoryview]


def sha256_hex(data: BytesLike) -> str:
    """Exactly the reference's expression (webdav_sync.py:59): lowercase 64-char hex."""
    return hashlib.sha256(data).hexdigest()


def sha256_digest(data: BytesLike) -> bytes:
    """32 raw digest bytes (what the device kernel writes before hex encoding)."""
    return hashlib.sha256(data).digest()


# ---------------------------------------------------------------------------------------
# Pure-Python restatement of FIPS 180-4 section 6.2 (SHA-256).  The algorithm lives in a
# third-party dependency of the reference (CPython hashlib -> OpenSSL EVP_sha256; the
# reference pins no version, the container has OpenSSL 3.0.13), so the published standard
# is restated here.  Small inputs only (pure-Python loops); checked against hashlib in
# tests/test_oracle_hash.py.  The CUDA kernel follows the same structure: big-endian word
# load, 48-word message schedule, 64 rounds, Merkle-Damgard padding with a 64-bit
# big-endian bit length.
# ---------------------------------------------------------------------------------------
_K = [
    0x428A2F98, 0x71374491, 0xB5C0FBCF, 0xE9B5DBA5, 0x3956C25B, 0x59F111F1, 0x923F82A4, 0xAB1C5ED5,
    0xD807AA98, 0x12835B01, 0x243185BE, 0x550C7DC3, 0x72BE5D74, 0x80DEB1FE, 0x9BDC06A7, 0xC19BF174,
    0xE49B69C1, 0xEFBE4786, 0x0FC19DC6, 0x240CA1CC, 0x2DE92C6F, 0x4A7484AA, 0x5CB0A9DC, 0x76F988DA,
    0x983E5152, 0xA831C66D, 0xB00327C8, 0xBF597FC7, 0xC6E00BF3, 0xD5A79147, 0x06CA6351, 0x14292967,
    0x27B70A85, 0x2E1B2138, 0x4D2C6DFC, 0x53380D13, 0x650A7354, 0x766A0ABB, 0x81C2C92E, 0x92722C85,
    0xA2BFE8A1, 0xA81A664B, 0xC24B8B70, 0xC76C51A3, 0xD192E819, 0xD6990624, 0xF40E3585, 0x106AA070,
    0x19A4C116, 0x1E376C08, 0x2748774C, 0x34B0BCB5, 0x391C0CB3, 0x4ED8AA4A, 0x5B9CCA4F, 0x682E6FF3,
    0x748F82EE, 0x78A5636F, 0x84C87814, 0x8CC70208, 0x90BEFFFA, 0xA4506CEB, 0xBEF9A3F7, 0xC67178F2,
]
_H0 = [0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19]
_M = 0xFFFFFFFF


def _rotr(x: int, n: int) -> int:
    return ((x >> n) | (x << (32 - n))) & _M


def _compress(state, block: bytes):
    w = list(struct.unpack(">16I", block))
    for t in range(16, 64):
        s0 = _rotr(w[t - 15], 7) ^ _rotr(w[t - 15], 18) ^ (w[t - 15] >> 3)
        s1 = _rotr(w[t - 2], 17) ^ _rotr(w[t - 2], 19) ^ (w[t - 2] >> 10)
        w.append((w[t - 16] + s0 + w[t - 7] + s1) & _M)
    a, b, c, d, e, f, g, h = state
    for t in range(64):
        big1 = _rotr(e, 6) ^ _rotr(e, 11) ^ _rotr(e, 25)
        ch = (e & f) ^ (~e & g)
        t1 = (h + big1 + ch + _K[t] + w[t]) & _M
        big0 = _rotr(a, 2) ^ _rotr(a, 13) ^ _rotr(a, 22)
        maj = (a & b) ^ (a & c) ^ (b & c)
        t2 = (big0 + maj) & _M
        h, g, f, e, d, c, b, a = g, f, e, (d + t1) & _M, c, b, a, (t1 + t2) & _M
    return [(s + v) & _M for s, v in zip(state, (a, b, c, d, e, f, g, h))]


def sha256_restated(data: BytesLike) -> str:
    data = bytes(data)
    bitlen = len(data) * 8
    padded = data + b"\x80" + b"\x00" * ((55 - len(data)) % 64) + struct.pack(">Q", bitlen)
    state = list(_H0)
    for off in range(0, len(padded), 64):
        state = _compress(state, padded[off:off + 64])
    return "".join(f"{v:08x}" for v in state)
