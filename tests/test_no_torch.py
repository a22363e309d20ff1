"""The host-pointer layer and the service mirrors work in a process where PyTorch cannot be imported (the
reference service has no tensor library in its requirements).  Run in a subprocess with ``import torch`` blocked."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = r'''
import sys
sys.modules["torch"] = None                      # any "import torch" now raises ImportError
sys.path.insert(0, %r)
import ics_b200
from ics_b200.services.webdav_sync import WebDAVSync            # noqa: F401
from ics_b200.services.activity_api_sync import ActivityAPISync  # noqa: F401
from ics_b200.crud import classificacao_crud                     # noqa: F401
from ics_b200.api.routes import images                           # noqa: F401
''' % ROOT


def _run(body: str):
    return subprocess.run([sys.executable, "-c", PRELUDE + body], capture_output=True, text=True, cwd=ROOT, timeout=300)


def test_package_imports_without_torch_and_fails_loudly_without_gpu():
    r = _run(r'''
loaded = [m for m, v in sys.modules.items() if v is not None and (m == "torch" or m.startswith("torch."))]
assert not loaded, loaded
try:
    out = ics_b200.hash_batch([b"abc"])
    assert out == ["ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"], out
    print("GPU")
except ics_b200.B2Error as e:
    assert "no CPU fallback" in str(e), str(e)
    print("NOGPU")
''')
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() in ("GPU", "NOGPU")


@pytest.mark.gpu
def test_service_path_end_to_end_without_torch():
    r = _run(r'''
import numpy as np
from oracle import dedupe_batch, label_tally, sha256_hex, synth_image, synth_label_rows, thumbnail_u8, preview_f32
imgs = [synth_image(g, *shape) for g, shape in enumerate([(128, 160), (300, 257), (128, 160), (64, 64), (1080, 1920)])]
imgs.append(imgs[1].copy())
datas = [im.tobytes() for im in imgs] + [None]
res = ics_b200.ingest_batch(datas, decoded_rgb=imgs + [None], existing_hashes=[sha256_hex(datas[3])])
hashes = [sha256_hex(d) if d is not None else None for d in datas]
is_new, first, stats = dedupe_batch(hashes, {hashes[3]})
assert res.decision.hashes == hashes and res.decision.is_new == is_new and res.decision.first_index == first
assert res.decision.stats == stats
for i, im in enumerate(imgs):
    want = thumbnail_u8(im, 256, 256)
    assert np.array_equal(res.thumbs[i], want), i
    assert np.allclose(res.previews[i], preview_f32(want), rtol=1e-5, atol=1e-7)
t, p = ics_b200.thumbnails(imgs[:2], 100, 60, want_preview=False)
assert p is None and np.array_equal(t[1], thumbnail_u8(imgs[1], 100, 60))
img, cls, act = synth_label_rows(1000, 50, 10)
tally = ics_b200.label_tally(img, cls, act, 1000, 50)
assert np.array_equal(tally.counts, label_tally(img, cls, act, 1000, 50))
assert abs(tally.kappa_general()) < 1.0
# bulk distinct-image counts (crud mirror) and the streaming ring, both without PyTorch
import uuid
users = [str(uuid.UUID(int=c + 1)) for c in range(4)]
class Db:
    classificacoes = [{"id_con": u, "id_img": "%064x" % (i % 5), "id_opc": "o", "ativo": (i + c) % 3 != 0}
                      for c, u in enumerate(users) for i in range(9)]
got = classificacao_crud.contagem_classificacoes_todos(Db())
for u in users:
    assert got[u] == len({r["id_img"] for r in Db.classificacoes if r["id_con"] == u and r["ativo"]}), (u, got)
from ics_b200.pipeline import IngestRing
ring = IngestRing(ring_bytes=64 << 20, max_listings=2, max_images=8, out_h=256, out_w=256)
res = ring.result(ring.submit(imgs))
assert [bytes(d).hex() for d in res.digests] == hashes[:6]
assert np.array_equal(res.thumbs[4], thumbnail_u8(imgs[4], 256, 256))
assert res.stats == {"processed": 6, "created": 5, "updated": 1}
ring.close()
loaded = [m for m, v in sys.modules.items() if v is not None and (m == "torch" or m.startswith("torch."))]
assert not loaded, loaded
print("OK")
''')
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().endswith("OK")
