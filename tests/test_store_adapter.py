"""The persistence seam under concurrent writers, the SQLAlchemy adapter and the thumbnail write-back (ADVICE r1;
reference: webdav_sync.py:355-369, :402-424, activity_api_sync.py:875-887; SURVEY 8(f) rank 2 iii) — on the golden
generator's own stub session, i.e. the session the reference's unmodified functions ran against when the golden
vectors were made.  Every test runs twice: with the hashing / resizing done by the library on the GPU (``-m gpu``) and,
on the CPU-only container, with those two calls replaced by the oracle so that the HOST logic is covered there too."""
import pytest

from conftest import dump_rows, ingest_scenario
from ics_b200.services.activity_api_sync import ActivityAPISync
from ics_b200.services.webdav_sync import WebDAVSync
from ics_b200.store import DictImageStore

KEYS = ("nome_img", "caminho_img", "existe_no_nextcloud", "id_cnj", "image_meta", "nextcloud_meta", "sync_method",
        "first_seen")


class _Clock:
    def __init__(self):
        from datetime import datetime, timedelta, timezone
        self.t, self.dt = datetime(2025, 1, 1, tzinfo=timezone.utc), timedelta(seconds=1)

    def __call__(self):
        self.t += self.dt
        return self.t


@pytest.fixture(params=["oracle-on-cpu", pytest.param("library-on-gpu", marks=pytest.mark.gpu)])
def hashing(request, monkeypatch):
    """'library-on-gpu': the product as it is.  'oracle-on-cpu': hash_and_dedupe / hash_batch / thumbnails of the two
    service modules swapped for the oracle (test infrastructure only) so the apply logic runs without a GPU."""
    if request.param == "library-on-gpu":
        return request.param
    import numpy as np

    from ics_b200 import hostapi
    from ics_b200.ingest import DedupeDecision
    from ics_b200.services import webdav_sync
    from oracle import dedupe_batch, sha256_hex, thumbnail_u8

    def fake_hash_and_dedupe(datas, existing_hashes=None, device=None):
        hashes = [sha256_hex(d) if d is not None else None for d in datas]
        present = sorted({h for h in hashes if h})
        existing = set(existing_hashes(present)) if callable(existing_hashes) else set(existing_hashes or ())
        is_new, first, stats = dedupe_batch(hashes, existing)
        last = [max((j for j, x in enumerate(hashes) if x == h), default=-1) if h else -1 for h in hashes]
        return DedupeDecision(hashes, is_new, first, last, stats)

    monkeypatch.setattr(webdav_sync, "hash_and_dedupe", fake_hash_and_dedupe)
    monkeypatch.setattr(hostapi, "hash_batch", lambda datas, device=None: [sha256_hex(d) for d in datas])
    monkeypatch.setattr(hostapi, "thumbnails", lambda images, oh=256, ow=256, want_preview=True, **kw:
                        (np.stack([thumbnail_u8(im, oh, ow) for im in images]), None))
    return request.param


def _orm():
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_reference_golden.py")
    spec = importlib.util.spec_from_file_location("golden_stub_orm", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                               # defines the stub classes; main() is not run
    Imagem = mod.make_model("Imagem", ["content_hash", "nome_img", "caminho_img", "metadados", "existe_no_nextcloud",
                                       "data_proc", "data_sinc", "id_cnj"], "content_hash")
    Conjunto = mod.make_model("ConjuntoImagens", ["id_cnj", "nome_conj", "caminho_conj", "file_id", "imagens_sincronizadas",
                                                  "existe_no_nextcloud", "data_proc", "data_sinc"], "id_cnj")
    return mod, Imagem, Conjunto


def test_sqlalchemy_adapter_replays_the_reference_batches(ref_ingest, hashing):
    """The same replay as test_process_image_batch_replay, but through SqlAlchemyImageStore on the stub session."""
    from ics_b200.integration.sqlalchemy_store import SqlAlchemyImageStore
    mod, Imagem, Conjunto = _orm()
    session = mod.Session()
    store = SqlAlchemyImageStore(session, Imagem, Conjunto)
    files, infos, client = ingest_scenario(ref_ingest)
    sync = WebDAVSync(client, store, now=_Clock())
    for b in ref_ingest["webdav_batches"]:
        stats = sync._process_image_batch([infos[i] for i in b["indices"]], "/set1", b["conjunto_id"])
        store.commit()
        assert stats == b["stats"]
        rows = {r.content_hash: {c: getattr(r, c) for c in ("content_hash", "nome_img", "caminho_img", "metadados",
                                                            "existe_no_nextcloud", "data_proc", "data_sinc", "id_cnj")}
                for r in session.tables[Imagem]}
        got, want = dump_rows(rows), b["table_after"]
        assert set(got) == set(want)
        for h in want:
            for key in KEYS:
                assert got[h][key] == want[h][key], (h, key)


def test_concurrent_insert_takes_the_merge_branch(ref_ingest, hashing):
    """Another session inserts one of the batch's hashes between the IN lookup and the insert: the flush raises
    IntegrityError, the adapter turns it into DuplicateKeyError, the service rolls back, re-reads and merges
    (name / path / existe / data_sinc), counts the image as updated and carries on with the rest of the batch."""
    from ics_b200.integration.sqlalchemy_store import SqlAlchemyImageStore
    mod, Imagem, Conjunto = _orm()
    session = mod.Session()
    files, infos, client = ingest_scenario(ref_ingest)
    b = ref_ingest["webdav_batches"][0]
    batch = [infos[i] for i in b["indices"]]
    victim = next(s["hash"] for i, s in enumerate(ref_ingest["singles"]) if i in b["indices"] and s["valid"] and s["hash"])

    class Racy(SqlAlchemyImageStore):
        def get_many(self, hashes):                           # the lookup sees an empty table ...
            out = super().get_many(hashes)
            session.tables.setdefault(Imagem, []).append(      # ... then the other writer commits the same content
                Imagem(content_hash=victim, nome_img="other-writer.png", caminho_img="/elsewhere/other-writer.png",
                       metadados={"nextcloud": {"file_id": "x"}}, existe_no_nextcloud=False, data_proc="t0", data_sinc="t0",
                       id_cnj="other"))
            return out

    store = Racy(session, Imagem, Conjunto)
    stats = WebDAVSync(client, store, now=_Clock())._process_image_batch(batch, "/set1", b["conjunto_id"])
    assert stats["processed"] == b["stats"]["processed"]
    assert stats["created"] == b["stats"]["created"] - 1 and stats["updated"] == b["stats"]["updated"] + 1
    row = next(r for r in session.tables[Imagem] if r.content_hash == victim)
    assert row.id_cnj == "other" and row.data_proc == "t0"            # identity columns stay the first writer's
    assert row.existe_no_nextcloud is True and row.nome_img != "other-writer.png" and row.data_sinc != "t0"
    assert len([r for r in session.tables[Imagem] if r.content_hash == victim]) == 1


def test_vanished_row_and_store_errors_do_not_abort_the_batch(ref_ingest, hashing):
    files, infos, client = ingest_scenario(ref_ingest)
    b0, b1 = ref_ingest["webdav_batches"][0], ref_ingest["webdav_batches"][1]

    class Flaky(DictImageStore):
        calls = 1                                                 # armed after the first batch

        def update(self, content_hash, fields):
            Flaky.calls += 1
            if Flaky.calls == 1:
                raise RuntimeError("connection reset")            # any error: log, rollback, next image (:421-424)
            return super().update(content_hash, fields)

    store = Flaky()
    sync = WebDAVSync(client, store, now=_Clock())
    sync._process_image_batch([infos[i] for i in b0["indices"]], "/set1", b0["conjunto_id"])
    gone = next(iter(store.rows))
    lookup = store.get_many

    def get_many_then_delete(hashes):                             # a row present at lookup time disappears before the apply step
        out = lookup(hashes)
        store.rows.pop(gone, None)
        return out

    store.get_many = get_many_then_delete
    Flaky.calls = 0                                               # the next update raises once
    stats = sync._process_image_batch([infos[i] for i in b0["indices"]], "/set1", b0["conjunto_id"])
    assert stats["created"] == 1                                  # the vanished row is inserted again, not a crash
    assert stats["processed"] == b0["stats"]["processed"] - 1     # the image whose update raised is skipped, the rest applied
    assert gone in store.rows


def test_activity_duplicate_key_merges_and_returns_true(ref_ingest, hashing):
    files, infos, client = ingest_scenario(ref_ingest)
    call = next(c for c in ref_ingest["activity"]["calls"] if c["ok"])
    store = DictImageStore()
    lookup = store.get

    def get_then_race(h, _state={"n": 0}):                        # first lookup: absent; then another writer inserts it
        _state["n"] += 1
        if _state["n"] == 1:
            assert lookup(h) is None
            store.rows[h] = {"content_hash": h, "nome_img": "w", "caminho_img": "/w", "metadados": {}, "existe_no_nextcloud": False,
                             "data_proc": "t0", "data_sinc": "t0", "id_cnj": "other"}
            return None
        return lookup(h)

    store.get = get_then_race
    assert ActivityAPISync(client, store, now=_Clock())._process_new_image(infos[call["index"]]) is True
    (row,) = store.rows.values()
    assert row["id_cnj"] == "other" and row["existe_no_nextcloud"] is True and row["nome_img"] == infos[call["index"]]["name"]


def test_thumbnails_written_back_in_arrival_order(ref_ingest, hashing):
    """with ``store_thumbnails`` the sync decodes through the feeder, resizes the images it
    INSERTS in one device call and writes them back through the store; metadados['image']['thumb'] references them."""
    import hashlib
    import io

    import numpy as np
    from PIL import Image

    from oracle import thumbnail_u8
    files, infos, client = ingest_scenario(ref_ingest)
    client.list_folder = lambda folder, depth=1: infos
    store = DictImageStore()
    stats = WebDAVSync(client, store, download_workers=4, store_thumbnails=True).sync_images_in_folder("/set1", "cid")
    assert stats["images_created"] == len(store.rows) > 0
    by_hash = {hashlib.sha256(data).hexdigest(): data for data in files.values()}
    n_thumbs = 0
    for h, row in store.rows.items():
        ref = row["metadados"]["image"].get("thumb")
        try:
            rgb = np.ascontiguousarray(np.asarray(Image.open(io.BytesIO(by_hash[h])).convert("RGB"), dtype=np.uint8))
        except Exception:
            assert ref is None and h not in store.thumbs          # undecodable file: stored, hashed, no thumbnail
            continue
        assert ref == f"thumbnails/{h}"
        assert np.array_equal(store.thumbs[h], thumbnail_u8(rgb, 256, 256))
        n_thumbs += 1
    assert n_thumbs > 0 and set(store.thumbs) <= set(store.rows)
    # a second pass over the same listing only updates: no thumbnail is recomputed, references stay
    before = {h: t.copy() for h, t in store.thumbs.items()}
    stats2 = WebDAVSync(client, store, download_workers=4, store_thumbnails=True).sync_images_in_folder("/set1", "cid")
    assert stats2["images_created"] == 0 and all(np.array_equal(before[h], store.thumbs[h]) for h in before)
