"""Oracle pinning, labels: per-user paths against the reference's own outputs; tally + kappa
against the published Fleiss (1971) example (parity unpinned by the reference: it has no
tally and no kappa)."""
import uuid

import numpy as np

from oracle import (classification_delta, distinct_image_count, fleiss_kappa, fleiss_kappa_general,
                    fleiss_partials, group_by_image, history_grouping, label_tally, synth_label_rows)


def _rows(ref_labels):
    return [dict(r, id_con=uuid.UUID(r["id_con"])) for r in ref_labels["classificacoes"]]


def test_group_by_image(ref_labels):
    rows = _rows(ref_labels)
    for case in ref_labels["group_by_image"]:
        try:
            u = uuid.UUID(case["id_con"])
        except ValueError:
            assert case["result"] == {}
            continue
        got = group_by_image(rows, u, case["images"])
        assert {h: [c["id_cla"] for c in lst] for h, lst in got.items()} == case["result"]


def test_distinct_count(ref_labels):
    rows = _rows(ref_labels)
    for case in ref_labels["distinct_count"]:
        if case["id_con"] is None:
            assert case["total"] == 0
        else:
            assert distinct_image_count(rows, uuid.UUID(case["id_con"])) == case["total"]


def test_history_grouping(ref_labels):
    got = history_grouping([tuple(j) for j in ref_labels["history"]["joined"]])
    want = ref_labels["history"]["items"]
    assert len(got) == len(want)
    # reference quirk: "total" is query.count() = joined ROWS of the page query, not grouped items
    assert ref_labels["history"]["total"] == len(ref_labels["history"]["joined"])
    for g, w in zip(got, want):
        assert g["content_hash"] == w["content_hash"]
        assert g["ids_opcoes"] == w["ids_opcoes"]
        assert g["opcao_escolhida"] == w["opcao_escolhida"]


def test_classification_delta(ref_labels):
    for d in ref_labels["delta"]:
        inativar, criar, reativar, novas, inc = classification_delta(
            d["before_active"], d["before_inactive"], d["wanted"])
        assert novas == d["total_novas"]
        assert int(inc) == d["counter_delta"]
        after_active = (set(d["before_active"]) - inativar) | criar | reativar
        assert sorted(after_active) == d["after_active"]


def test_fleiss_1971(fleiss71):
    counts = np.array(fleiss71["table"], dtype=np.int32)
    p = fleiss_partials(counts)
    n, N = fleiss71["n_raters"], counts.shape[0]
    assert p["R"] == n * N
    kappa = fleiss_kappa(p["class_totals"], p["S2"], p["R"], N, n)
    assert round(kappa, 3) == fleiss71["kappa"]
    assert abs(fleiss_kappa_general(counts) - kappa) < 1e-12
    p_bar = (p["S2"] - p["R"]) / (N * n * (n - 1))
    assert round(p_bar, 3) == fleiss71["P_bar"]


def test_tally_matches_naive_loop():
    img, cls, act = synth_label_rows(50, 7, 9)
    counts = label_tally(img, cls, act, 50, 7)
    naive = np.zeros((50, 7), dtype=np.int32)
    for i, c, a in zip(img, cls, act):
        if a:
            naive[i, c] += 1
    assert np.array_equal(counts, naive)
    p = fleiss_partials(counts)
    assert p["R"] == int(act.sum()) and p["S2"] == int((naive.astype(np.int64) ** 2).sum())


def test_encode_label_rows_on_reference_fixture_rows(ref_labels):
    """The dictionary encoder of the oracle against the product's host encoder on the reference's own fixture
    rows (with sorted dictionaries both define the same dense indices), plus unknown keys."""
    from ics_b200 import labels
    from oracle import encode_label_rows
    rows = ref_labels["classificacoes"]
    hashes = sorted({r["id_img"] for r in rows})
    options = sorted({r["id_opc"] for r in rows})
    img, cls, act = encode_label_rows(rows, hashes, options)
    order = np.argsort(img, kind="stable")
    h_img, h_cls, h_act = labels.LabelEncoder(hashes, options).encode(rows)
    assert np.array_equal(img[order], h_img) and np.array_equal(cls[order], h_cls) and np.array_equal(act[order], h_act)
    img2, cls2, _ = encode_label_rows(rows, hashes[1:], options[1:])
    assert (img2 == -1).sum() == sum(r["id_img"] == hashes[0] for r in rows)
    assert (cls2 == 255).sum() == sum(r["id_opc"] == options[0] for r in rows)


def test_agreement_histogram_is_the_integer_form_of_sum_pi():
    """The histogram the tally pass accumulates (graft-defined) against the textbook definition: sum_i P_i over images
    with n_i >= 2 equals sum_n bin[n] / (n (n - 1)), so the general-n kappa computed from INTEGERS (host code of the
    product, no GPU involved) equals the count-matrix formula; shards add up bin by bin."""
    import numpy as np

    from ics_b200.labels import fleiss_kappa_from_hist, sum_pi_from_hist
    from oracle import agreement_hist, fleiss_kappa_general
    rng = np.random.default_rng(8)
    counts = rng.integers(0, 7, size=(5000, 12)).astype(np.int32)
    counts[::5] = 0
    counts[1::5, 1:] = 0                                               # single-class images
    counts[3] = 0
    counts[3, 0] = 1                                                   # n_i = 1: contributes nothing
    h = agreement_hist(counts)
    c = counts.astype(np.int64)
    n_i = c.sum(1)
    m = n_i >= 2
    want = float(((c[m] ** 2).sum(1) - n_i[m]).astype(np.float64).__truediv__((n_i[m] * (n_i[m] - 1)).astype(np.float64)).sum())
    assert h[0] == 0 and h[1] == 0 and abs(sum_pi_from_hist(h) - want) <= 1e-9 * want
    kg = fleiss_kappa_from_hist(c.sum(0), int(n_i.sum()), int(m.sum()), h)
    assert abs(kg - fleiss_kappa_general(counts)) <= 1e-12 * abs(kg)
    assert np.array_equal(agreement_hist(counts[:2000]) + agreement_hist(counts[2000:]), h)
    big = np.zeros((2, 3), dtype=np.int32)
    big[0] = (600, 500, 0)                                             # 1 100 ratings: past the last bin
    big[1] = (2, 1, 0)
    hb = agreement_hist(big)
    assert hb[0] == 1 and hb[3] == 2 * 2 + 1 - 3
    import pytest
    with pytest.raises(ValueError):
        sum_pi_from_hist(hb)
