"""Host-side logic that needs no GPU: packing, digest ordering, sharding, encoders, kappa
arithmetic, and the per-user CRUD mirrors against the reference's own outputs."""
import uuid

import numpy as np
import pytest

import ics_b200
from ics_b200 import engine, labels
from ics_b200.crud import classificacao_crud
from ics_b200.dist import shard_range, shard_rows_by_image


def test_packed_messages_layout():
    datas = [b"", b"a" * 5, b"b" * 16, b"c" * 100, b"d" * 17]
    p = engine.PackedMessages(datas, pin=False)
    off, ln = p.offsets.numpy(), p.lengths.numpy()
    assert list(ln) == [0, 5, 16, 100, 17]
    assert all(o % 16 == 0 for o in off)
    host = p.data.numpy()
    for d, o, l in zip(datas, off, ln):
        assert host[o:o + l].tobytes() == d
    for a, b in zip(off[:-1], off[1:]):
        assert b >= a
    order = p.order.numpy()
    assert sorted(order) == list(range(5)) and list(ln[order]) == sorted(ln, reverse=True)


def test_sort_digests_is_memcmp_order():
    rng = np.random.default_rng(3)
    d = rng.integers(0, 256, size=(200, 32), dtype=np.uint8)
    d[5] = d[17]
    s = engine.sort_digests(d)
    as_bytes = [bytes(r) for r in s]
    assert as_bytes == sorted(bytes(r) for r in d)


def test_shard_range_partitions():
    for n in (0, 1, 7, 100, 1_000_003):
        for ws in (1, 2, 4, 8):
            parts = [shard_range(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_rows_by_image():
    img = np.repeat(np.arange(10, dtype=np.int32), 3)
    covered = []
    for r in range(4):
        lo, hi, r0, r1 = shard_rows_by_image(img, 10, r, 4)
        assert np.all((img[r0:r1] >= lo) & (img[r0:r1] < hi))
        covered.append((r0, r1))
    assert covered[0][0] == 0 and covered[-1][1] == len(img)
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


def test_fleiss_kappa_host_arithmetic(fleiss71):
    c = np.array(fleiss71["table"], dtype=np.int64)
    n, N = fleiss71["n_raters"], c.shape[0]
    k = labels.fleiss_kappa(c.sum(0), int((c * c).sum()), int(c.sum()), N, n)
    assert round(k, 3) == fleiss71["kappa"]
    n_i = c.sum(1)
    sum_pi = float((((c * c).sum(1) - n_i) / (n_i * (n_i - 1))).sum())
    assert abs(labels.fleiss_kappa_general(c.sum(0), int(c.sum()), sum_pi, N) - k) < 1e-12


def test_label_encoder():
    enc = labels.LabelEncoder(["h0", "h1", "h2"], ["o0", "o1"])
    rows = [{"id_img": "h2", "id_opc": "o1", "ativo": True}, {"id_img": "h0", "id_opc": "o0", "ativo": False},
            {"id_img": "h2", "id_opc": "o0", "ativo": True}]
    img, cls, act = enc.encode(rows)
    assert list(img) == [0, 2, 2] and list(cls) == [0, 1, 0] and list(act) == [0, 1, 1]


class _Db:
    def __init__(self, rows):
        self.classificacoes = rows


def test_obter_classificacoes_imagens_vs_reference(ref_labels):
    db = _Db(ref_labels["classificacoes"])
    for case in ref_labels["group_by_image"]:
        imgs = [{"content_hash": h} for h in case["images"]]
        got = classificacao_crud.obter_classificacoes_imagens(db, case["id_con"], imgs)
        assert {h: [c["id_cla"] for c in lst] for h, lst in got.items()} == case["result"]


def test_obter_contagem_vs_reference(ref_labels):
    db = _Db(ref_labels["classificacoes"])
    for case in ref_labels["distinct_count"]:
        if case["id_con"] is None:
            continue
        assert classificacao_crud.obter_contagem_classificacoes(db, case["id_con"]) == {"total": case["total"]}
    assert classificacao_crud.obter_contagem_classificacoes(db, "not-a-uuid") == {"total": 0}


def test_agrupar_historico_vs_reference(ref_labels):
    """a10: the grouping loop of listar_historico_usuario on the reference's own page (row dicts rebuilt from the
    committed fixture: hash, option text, option id, image path)."""
    from ics_b200.crud.classificacao_crud import agrupar_historico
    h = ref_labels["history"]
    path_of = {it["content_hash"]: it["url_img"][len("/nextcloud/images/"):].replace("%20", " ") for it in h["items"]}
    page = [({"data_criado": i}, {"content_hash": ch, "nome_img": "n", "caminho_img": "/" + path_of[ch]},
             {"texto": texto, "id_opc": id_opc}, None, {"id_amb": "amb", "titulo_amb": "Ambiente A"})
            for i, (ch, texto, id_opc) in enumerate(h["joined"])]
    got = agrupar_historico(page)
    assert [{k: it[k] for k in ("content_hash", "ids_opcoes", "opcao_escolhida", "url_img")} for it in got] == h["items"]
    assert all(it["id_amb"] == "amb" and "opcoes_lista" not in it for it in got)
    assert agrupar_historico(page, id_amb="x")[0]["id_amb"] == "x" and agrupar_historico([]) == []


def test_calcular_delta_classificacao_vs_reference(ref_labels):
    """a11: set deltas + the progress-counter rule of criar_ou_atualizar_classificacao, replayed on the states the
    reference went through."""
    from ics_b200.crud.classificacao_crud import calcular_delta_classificacao
    for d in ref_labels["delta"]:
        inativar, criar, reativar, total_novas, delta = calcular_delta_classificacao(
            d["before_active"], d["before_inactive"], d["wanted"])
        assert total_novas == d["total_novas"] and delta == d["counter_delta"]
        after_active = (set(d["before_active"]) - inativar) | criar | reativar
        after_inactive = (set(d["before_inactive"]) - reativar) | inativar
        assert sorted(after_active) == d["after_active"] and sorted(after_inactive) == d["after_inactive"]


def test_shard_by_bytes_balances_a_config3_listing():
    """config 3 images are squares of side 256 * 2^(pi(g) mod 5); shards must balance BYTES, not
    counts, cover every image once, keep listing order inside a shard and be the same on every rank."""
    import numpy as np

    from ics_b200.dist import shard_by_bytes

    rng = np.random.default_rng(0xB200)
    sides = 256 << (rng.permutation(10_000) % 5)
    lengths = (sides.astype(np.int64) ** 2) * 3
    for ws in (1, 2, 3, 4, 8):
        shards = shard_by_bytes(lengths, ws)
        again = shard_by_bytes(lengths.tolist(), ws)
        assert len(shards) == ws and all(np.array_equal(a, b) for a, b in zip(shards, again))
        allidx = np.concatenate(shards)
        assert np.array_equal(np.sort(allidx), np.arange(lengths.shape[0]))          # every image exactly once
        assert all(np.all(np.diff(s) > 0) for s in shards)                           # listing order inside a shard
        totals = np.array([lengths[s].sum() for s in shards])
        assert totals.max() - totals.min() <= lengths.max()                          # within one (largest) image
        if ws > 1:
            naive = np.array([lengths[r::ws].sum() for r in range(ws)])              # g mod G
            assert totals.max() <= naive.max()
    assert [list(s) for s in shard_by_bytes([5, 5, 5], 2)] == [[0, 2], [1]]            # ties: lowest rank, listing order
    assert [list(s) for s in shard_by_bytes([], 3)] == [[], [], []]
