"""Oracle pinning, ingest: the restated batch logic against outputs of the reference's own
functions (tests/golden/reference_ingest.json)."""
import base64

from conftest import dump_rows, ingest_scenario
from oracle import buscar_por_hash, dedupe_batch, image_metadata, process_image_batch, sha256_hex, validate_image


def test_validate_and_metadata(ref_ingest):
    files, infos, _ = ingest_scenario(ref_ingest)
    for info, single in zip(infos, ref_ingest["singles"]):
        assert validate_image(info) == single["valid"]
        if single["hash"] is not None:
            assert image_metadata(files[info["path"]]) == single["metadata"]


def test_process_image_batch_replay(ref_ingest):
    files, infos, client = ingest_scenario(ref_ingest)
    table = {}
    for b in ref_ingest["webdav_batches"]:
        batch = [infos[i] for i in b["indices"]]
        stats = process_image_batch(batch, lambda p: client.get_file(p).content, table,
                                    conjunto_id=b["conjunto_id"], now_iso=f"t{len(table)}")
        assert stats == b["stats"]
        got = dump_rows(table)
        want = b["table_after"]
        assert set(got) == set(want)
        for h in want:
            for key in ("nome_img", "caminho_img", "existe_no_nextcloud", "id_cnj", "image_meta",
                        "nextcloud_meta", "sync_method"):
                assert got[h][key] == want[h][key], (h, key)


def test_dedupe_batch_matches_reference_counts(ref_ingest):
    files, infos, client = ingest_scenario(ref_ingest)
    existing = set()
    for b in ref_ingest["webdav_batches"]:
        hashes = []
        for i in b["indices"]:
            info = infos[i]
            if not validate_image(info) or info["path"] in ref_ingest["failures"]:
                hashes.append(None)
            else:
                hashes.append(sha256_hex(files[info["path"]]))
        is_new, first, stats = dedupe_batch(hashes, existing)
        assert stats == b["stats"]
        existing |= {h for h in hashes if h}
        assert existing == set(b["table_after"])


def test_upload_lookup(ref_ingest):
    table = {h: {"content_hash": h, "nome_img": r["nome_img"], "caminho_img": r["caminho_img"]}
             for h, r in ref_ingest["webdav_batches"][-1]["table_after"].items()}
    ups = [(u["content_type"], base64.b64decode(u["data"])) for u in ref_ingest["upload_lookup"]["uploads"]]
    assert buscar_por_hash(ups, table) == ref_ingest["upload_lookup"]["response"]
