"""dictionary-encoding of `classificacoes` rows on the device
(b2_encode_label_rows) against the dict-based oracle (parity unpinned by the reference: it never builds these
arrays; the rows of the first test ARE the reference's own fixture rows), then straight into the tally."""
import hashlib
import uuid

import numpy as np
import pytest
import torch

from ics_b200 import engine, labels
from oracle import encode_label_rows, label_tally

pytestmark = pytest.mark.gpu


def _check(rows, image_hashes, option_ids):
    enc = labels.DeviceLabelEncoder(image_hashes, option_ids)
    img, cls, act, unknown = enc.encode(rows, sort=False)
    w_img, w_cls, w_act = encode_label_rows(rows, image_hashes, option_ids)
    assert np.array_equal(img.cpu().numpy(), w_img)
    assert np.array_equal(cls.cpu().numpy(), w_cls)
    assert np.array_equal(act.cpu().numpy(), w_act)
    assert unknown.cpu().tolist() == [int((w_img < 0).sum()), int((w_cls == 255).sum())]
    return enc, (w_img, w_cls, w_act)


def test_reference_fixture_rows(ref_labels):
    rows = ref_labels["classificacoes"]
    hashes = sorted({r["id_img"] for r in rows})
    options = sorted({r["id_opc"] for r in rows})
    _check(rows, hashes, options)
    _check(rows, hashes[1:], options[:-1])                  # one image and one option missing from the dictionaries
    _check([], hashes, options)
    _check(rows, [], [])


def test_synthetic_rows_unknown_and_malformed_keys_then_tally():
    rng = np.random.default_rng(21)
    n_images, k, n_rows = 5000, 37, 200_000
    hashes = [hashlib.sha256(b"img%d" % i).hexdigest() for i in range(n_images)]
    options = [uuid.UUID(bytes=bytes(rng.integers(0, 256, size=16, dtype=np.uint8))) for _ in range(k)]
    stranger = hashlib.sha256(b"not stored").hexdigest()
    rows = []
    for r in range(n_rows):
        h = hashes[int(rng.integers(0, n_images))]
        o = options[int(rng.integers(0, k))]
        u = rng.random()
        if u < 0.01:
            h = stranger                                    # valid key, not in the table
        elif u < 0.02:
            h = h.upper()                                   # the primary key is lowercase hex: no match
        elif u < 0.03:
            h = h[:63] + "g"                                # not hex
        elif u < 0.04:
            o = uuid.UUID(bytes=bytes(rng.integers(0, 256, size=16, dtype=np.uint8)))
        rows.append({"id_img": h, "id_opc": str(o) if r % 2 else o, "ativo": bool(rng.random() < 0.9)})
    enc, (w_img, w_cls, w_act) = _check(rows, hashes, options)
    # rows with both keys known, ordered by image on the device, straight into the sorted-mode tally
    img, cls, act, _ = enc.encode(rows, sort=True)
    keep = (img >= 0) & (cls != 255)
    counts, partials = engine.label_tally_device(img[keep].contiguous(), cls[keep].contiguous(), act[keep].contiguous(),
                                                 n_images, k)
    p = partials.cpu().numpy()
    engine.check_tally(p, k, int(keep.sum().item()))
    ok = (w_img >= 0) & (w_cls != 255)
    assert np.array_equal(counts.cpu().numpy(), label_tally(w_img[ok], w_cls[ok], w_act[ok], n_images, k))


def test_encode_columns_large_property():
    """2 M rows against a 1 M-key table: every row's key is the hex form of the table entry it must map to."""
    rng = np.random.default_rng(5)
    n_images, rows = 1_000_000, 2_000_000
    keys = engine.sort_digests(rng.integers(0, 256, size=(n_images, 32), dtype=np.uint8))
    enc = labels.DeviceLabelEncoder([], [uuid.UUID(int=i + 1) for i in range(50)])
    enc.d_image_keys = torch.from_numpy(keys).cuda()
    pick = torch.randint(0, n_images, (rows,), device="cuda")
    digests = enc.d_image_keys[pick].contiguous()
    hexs = engine.digest_hex_device(digests)
    opt = torch.randint(0, 50, (rows,), device="cuda")
    opc = enc.d_option_keys[opt].contiguous()
    act = torch.ones(rows, dtype=torch.uint8, device="cuda")
    img, cls, a, unknown = enc.encode_columns(hexs, opc, act, sort=False)
    assert unknown.cpu().tolist() == [0, 0]
    got = enc.d_image_keys[img.long()]
    assert torch.equal(got, digests)                       # duplicates in a random table map to an equal key
    assert torch.equal(cls.long(), opt)
