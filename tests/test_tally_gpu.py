"""Parity, label tally + Fleiss partials (parity unpinned by the reference: it has neither;
oracle = NumPy bincount + integer partials, pinned on Fleiss 1971).  Bit-exact counts and
partials; kappa identical because it is computed from the same integers."""
import numpy as np
import pytest
import torch

import ics_b200
from ics_b200 import engine, labels
from oracle import fleiss_kappa, fleiss_kappa_general, fleiss_partials, label_tally, synth_label_rows

pytestmark = pytest.mark.gpu


def _compare(img, cls, act, n_images, k, sorted_by_image, image_base=0):
    res = labels.label_tally(img, cls, act, n_images, k, sorted_by_image=sorted_by_image, image_base=image_base)
    want = label_tally(np.asarray(img, np.int64) - image_base, cls, act, n_images, k)
    assert np.array_equal(res.counts, want)
    p = fleiss_partials(want)
    assert np.array_equal(res.class_totals, p["class_totals"])
    assert (res.S2, res.R, res.n_rated, res.n_pairs_images, res.pairs) == (
        p["S2"], p["R"], p["n_rated"], p["n_pairs_images"], p["pairs"])
    return res, p


@pytest.mark.parametrize("n_images,k,n_raters", [(1000, 50, 10), (1, 1, 1), (37, 3, 5), (5000, 50, 100),
                                                 (300, 255, 7), (100_000, 50, 20), (2049, 17, 33), (40, 50, 3000)])
@pytest.mark.parametrize("mode", ["sorted", "sorted-2-stage-ring", "sorted-tiny-window", "sorted-aggregated-inc", "scatter"])
def test_synthetic_rows(n_images, k, n_raters, mode, monkeypatch):
    """The sorted-mode (slab) kernel in its default geometry, with a 2-deep ring at three CTAs per SM and with
    a 4-image counter window (every stage re-scanned window by window), and the any-order kernel, against the
    oracle, bit-exact."""
    if mode == "sorted-2-stage-ring":
        monkeypatch.setenv("B2_TALLY_STAGES", "2")
        monkeypatch.setenv("B2_TALLY_CTAS", "3")
    if mode == "sorted-tiny-window":
        monkeypatch.setenv("B2_TALLY_TILE_LOG2", "2")
    if mode == "sorted-aggregated-inc":                       # the ATOMS.POPC.INC variant (automatic for long images)
        monkeypatch.setenv("B2_TALLY_INC", "1")
    img, cls, act = synth_label_rows(n_images, k, n_raters, shuffled=(mode == "scatter"))
    res, p = _compare(img, cls, act, n_images, k, sorted_by_image=(mode != "scatter"))
    if n_raters > 1 and p["R"] > 0:
        # constant-n formula on the same integers: bit-identical float64
        assert res.kappa(n_images, n_raters) == fleiss_kappa(p["class_totals"], p["S2"], p["R"], n_images, n_raters)


@pytest.fixture(params=["default", "2-stage-ring", "tiny-window"])
def tally_path(request, monkeypatch):
    if request.param == "2-stage-ring":
        monkeypatch.setenv("B2_TALLY_STAGES", "2")
        monkeypatch.setenv("B2_TALLY_CTAS", "3")
    if request.param == "tiny-window":
        monkeypatch.setenv("B2_TALLY_TILE_LOG2", "2")
    return request.param


def test_config1_rows():
    """BASELINE config 1: 10 000 label rows, N = 1000 images, k = 50, n = 10."""
    img, cls, act = synth_label_rows(1000, 50, 10)
    assert len(img) == 10_000
    _compare(img, cls, act, 1000, 50, True)


def test_fleiss_1971_through_device(fleiss71):
    table = np.array(fleiss71["table"], dtype=np.int64)
    img = np.repeat(np.arange(table.shape[0]), table.sum(1)).astype(np.int32)
    cls = np.concatenate([np.repeat(np.arange(table.shape[1]), row) for row in table]).astype(np.uint8)
    res, _ = _compare(img, cls, np.ones_like(cls), table.shape[0], table.shape[1], True)
    assert round(res.kappa(table.shape[0], fleiss71["n_raters"]), 3) == fleiss71["kappa"]


def test_edge_cases_sorted(tally_path):
    # no rows at all
    res, _ = _compare(np.zeros(0, np.int32), np.zeros(0, np.uint8), np.zeros(0, np.uint8), 777, 5, True)
    assert res.R == 0 and not res.counts.any()
    # all inactive
    img, cls, act = synth_label_rows(100, 4, 6)
    _compare(img, cls, np.zeros_like(act), 100, 4, True)
    # images without rows at the start, in the middle and at the end; ragged row counts
    rng = np.random.default_rng(2)
    per = rng.integers(0, 40, size=3000)
    per[:700] = 0
    per[1500:2300] = 0
    per[-100:] = 0
    img = np.repeat(np.arange(3000), per).astype(np.int32)
    cls = rng.integers(0, 9, size=img.size).astype(np.uint8)
    act = (rng.random(img.size) < 0.8).astype(np.uint8)
    _compare(img, cls, act, 3000, 9, True)
    # one image holding every row (more rows than a CTA's nominal share)
    img = np.full(200_000, 3, np.int32)
    cls = rng.integers(0, 50, size=img.size).astype(np.uint8)
    _compare(img, cls, np.ones_like(cls), 10, 50, True)
    # non-zero image_base (a shard of the image range)
    img, cls, act = synth_label_rows(500, 50, 12)
    _compare(img + 1000, cls, act, 500, 50, True, image_base=1000)
    # sparse: few rated images far apart (gaps much longer than any counter window), 1-3 rows each
    ids = np.sort(rng.choice(2_000_000, size=20_000, replace=False))
    img = np.repeat(ids, rng.integers(1, 4, size=ids.size)).astype(np.int32)
    cls = rng.integers(0, 4, size=img.size).astype(np.uint8)
    _compare(img, cls, np.ones_like(cls), 2_000_000, 4, True)
    # dense: one row per image (a stage of rows spans thousands of images), k = 200 (small counter window)
    img = np.arange(30_000, dtype=np.int32)
    cls = rng.integers(0, 200, size=img.size).astype(np.uint8)
    act = (rng.random(img.size) < 0.9).astype(np.uint8)
    _compare(img, cls, act, 30_000, 200, True)


@pytest.mark.parametrize("seed", range(12))
def test_random_shapes_all_sorted_kernels(seed, monkeypatch):
    """Randomised row distributions (empty images, one-row images, very long images, gaps, odd k, non-zero
    image_base, ragged table ends) through the sorted-mode kernel in three geometries, twice each: bit-exact
    and repeatable."""
    rng = np.random.default_rng(1000 + seed)
    n_images = int(rng.integers(1, 6000))
    k = int(rng.choice([1, 2, 7, 31, 32, 33, 50, 64, 65, 128, 129, 200, 256]))
    kind = seed % 4
    if kind == 0:
        per = rng.integers(0, 60, size=n_images)
    elif kind == 1:
        per = (rng.random(n_images) < 0.02) * rng.integers(1, 5000, size=n_images)      # few, long images; big gaps
    elif kind == 2:
        per = np.ones(n_images, dtype=np.int64)                                        # one row per image
    else:
        per = rng.integers(90, 140, size=n_images)                                     # around one slab per image
    base = int(rng.integers(0, 1 << 20))
    img = (np.repeat(np.arange(n_images), per) + base).astype(np.int32)
    cls = rng.integers(0, k, size=img.size).astype(np.uint8)
    act = (rng.random(img.size) < 0.9).astype(np.uint8)
    want = label_tally(img.astype(np.int64) - base, cls, act, n_images, k)
    for env in ({}, {"B2_TALLY_STAGES": "2", "B2_TALLY_CTAS": "3"}, {"B2_TALLY_TILE_LOG2": "1"}, {"B2_TALLY_INC": "1"},
                {"B2_TALLY_INC": "0"}):
        for name in ("B2_TALLY_STAGES", "B2_TALLY_CTAS", "B2_TALLY_TILE_LOG2", "B2_TALLY_INC"):
            monkeypatch.delenv(name, raising=False)
        for name, v in env.items():
            monkeypatch.setenv(name, v)
        for _ in range(2):
            res = labels.label_tally(img, cls, act, n_images, k, sorted_by_image=True, image_base=base)
            assert np.array_equal(res.counts, want), (seed, env, n_images, k, kind)
            assert res.R == int(act.sum())


def test_unsorted_rows_are_rejected_in_sorted_mode(tally_path):
    img, cls, act = synth_label_rows(2000, 10, 8, shuffled=True)
    with pytest.raises(ics_b200.B2Error) as e:
        labels.label_tally(img, cls, act, 2000, 10, sorted_by_image=True)
    assert e.value.code == -3


def test_out_of_range_rows_are_rejected(tally_path):
    img, cls, act = synth_label_rows(100, 10, 8)
    bad = cls.copy()
    bad[17] = 10
    for mode in (True, False):
        with pytest.raises(ics_b200.B2Error) as e:
            labels.label_tally(img, bad, act, 100, 10, sorted_by_image=mode)
        assert e.value.code == -1
    img2 = img.copy()
    img2[-1] = 100
    with pytest.raises(ics_b200.B2Error):
        labels.label_tally(img2, cls, act, 100, 10, sorted_by_image=True)


def test_partials_from_counts_and_general_kappa():
    rng = np.random.default_rng(4)
    counts = rng.integers(0, 6, size=(70_001, 50)).astype(np.int32)
    counts[::7] = 0
    d = torch.from_numpy(counts).cuda()
    partials, sum_pi = engine.fleiss_partials_device(d, want_sum_pi=True)
    partials2, sum_pi2 = engine.fleiss_partials_device(d, want_sum_pi=True)
    p = engine.partials_dict(partials.cpu().numpy(), 50)
    want = fleiss_partials(counts)
    assert np.array_equal(p["class_totals"], want["class_totals"])
    for key in ("S2", "R", "n_rated", "n_pairs_images", "pairs"):
        assert p[key] == want[key]
    assert float(sum_pi) == float(sum_pi2)                     # fixed reduction order: reproducible
    k_dev = labels.fleiss_kappa_general(p["class_totals"], p["R"], float(sum_pi), p["n_pairs_images"])
    assert abs(k_dev - fleiss_kappa_general(counts)) <= 1e-12 * max(1.0, abs(k_dev))


def test_sharded_partials_sum_to_global():
    """Mode M1 on one GPU: tally two image shards separately; integer partials add up to the
    single-shard result, so kappa is identical for any GPU count."""
    from ics_b200.dist import shard_rows_by_image
    n_images, k, n_r = 20_000, 50, 30
    img, cls, act = synth_label_rows(n_images, k, n_r)
    whole = labels.label_tally(img, cls, act, n_images, k)
    tot = None
    for r in range(4):
        lo, hi, r0, r1 = shard_rows_by_image(img, n_images, r, 4)
        part = labels.label_tally(img[r0:r1], cls[r0:r1], act[r0:r1], hi - lo, k, image_base=lo)
        assert np.array_equal(part.counts, whole.counts[lo:hi])
        vec = np.concatenate([part.class_totals, [part.S2, part.R, part.n_rated, part.n_pairs_images, part.pairs]])
        tot = vec if tot is None else tot + vec
    assert np.array_equal(tot[:k], whole.class_totals)
    assert list(tot[k:]) == [whole.S2, whole.R, whole.n_rated, whole.n_pairs_images, whole.pairs]


def test_full_size_property_c4_slice():
    """A 1/8 slice of BASELINE config 4 (what one of 8 GPUs owns): 12.5 M rows, 125 000 images,
    k = 50, generated on device.  Size-independent properties: counts sum to the active rows,
    class totals equal a torch.bincount of the classes, S2 equals the sum of squares."""
    n_images, k, n_r = 125_000, 50, 100
    rows = n_images * n_r
    g = torch.Generator(device="cuda").manual_seed(0xF1E155)
    img = (torch.arange(rows, device="cuda", dtype=torch.int64) // n_r).to(torch.int32)
    cls = torch.randint(0, k, (rows,), device="cuda", generator=g, dtype=torch.int64).to(torch.uint8)
    act = (torch.rand(rows, device="cuda", generator=g) < 0.95).to(torch.uint8)
    counts, partials = engine.label_tally_device(img, cls, act, n_images, k)
    p = partials.cpu().numpy()
    engine.check_tally(p, k, rows)
    d = engine.partials_dict(p, k)
    assert d["R"] == int(act.sum()) == int(counts.sum())
    want_tot = torch.bincount(cls[act.bool()].to(torch.int64), minlength=k)
    assert np.array_equal(d["class_totals"], want_tot.cpu().numpy())
    assert d["S2"] == int((counts.to(torch.int64) ** 2).sum())
    flat = torch.bincount(img[act.bool()].to(torch.int64) * k + cls[act.bool()].to(torch.int64), minlength=n_images * k)
    assert torch.equal(flat.view(n_images, k).to(torch.int32), counts)


def test_distinct_images_per_annotator(ref_labels):
    from ics_b200.crud import classificacao_crud

    class Db:
        classificacoes = ref_labels["classificacoes"]

    got = classificacao_crud.contagem_classificacoes_todos(Db())
    for case in ref_labels["distinct_count"]:
        if case["id_con"] is not None:
            assert got[case["id_con"]] == case["total"]


def test_distinct_images_long_inactive_runs_and_host_entry():
    """Runs of one (annotator, image) pair with thousands of rows, active ones at the start, the end, or nowhere:
    a run counts once iff it holds an active row (the kernel scans each run forward once: linear)."""
    from ics_b200 import hostapi
    rng = np.random.default_rng(12)
    ann, img, act = [], [], []
    want = np.zeros(7, dtype=np.int64)
    for a in range(7):
        for i in range(int(rng.integers(1, 9))):
            length = int(rng.choice([1, 2, 33, 700, 5000]))
            flags = np.zeros(length, dtype=np.uint8)
            mode = int(rng.integers(0, 4))
            if mode == 1:
                flags[0] = 1
            elif mode == 2:
                flags[-1] = 1
            elif mode == 3:
                flags[rng.integers(0, length, size=max(1, length // 3))] = 1
            ann += [a] * length
            img += [i * 3] * length
            act.append(flags)
            want[a] += int(flags.any())
    ann, img, act = np.array(ann, np.int32), np.array(img, np.int32), np.concatenate(act)
    got = hostapi.distinct_images_host(ann, img, act, 7)
    assert got.tolist() == want.tolist()
    got_d = engine.distinct_images_per_annotator_device(torch.from_numpy(ann).cuda(), torch.from_numpy(img).cuda(),
                                                        torch.from_numpy(act).cuda(), 7)
    assert got_d.cpu().tolist() == want.tolist()


@pytest.mark.parametrize("sorted_rows", [True, False])
def test_agreement_histogram_gives_general_kappa_from_integers(sorted_rows):
    """Variable ratings per image (active w.p. 0.8, some images empty or single-rated): the histogram of the tally
    pass equals the oracle's bit for bit and the kappa computed from it matches the count-matrix formula to 1e-12;
    shards add up exactly, so the general kappa is identical for any GPU count too."""
    from oracle import agreement_hist
    from ics_b200.dist import shard_rows_by_image
    n_images, k = 30_011, 50
    rng = np.random.default_rng(9)
    per = rng.integers(0, 40, n_images)
    per[::97] = 1
    img = np.repeat(np.arange(n_images, dtype=np.int32), per)
    cls = rng.integers(0, k, img.size).astype(np.uint8)
    cls[rng.random(img.size) < 0.5] = 3
    act = (rng.random(img.size) < 0.8).astype(np.uint8)
    if not sorted_rows:
        perm = rng.permutation(img.size)
        img_in, cls_in, act_in = img[perm], cls[perm], act[perm]
    else:
        img_in, cls_in, act_in = img, cls, act
    t = labels.label_tally(img_in, cls_in, act_in, n_images, k, sorted_by_image=sorted_rows)
    counts = label_tally(img, cls, act, n_images, k)
    assert np.array_equal(t.counts, counts)
    assert np.array_equal(t.agree_hist, agreement_hist(counts))
    kg = t.kappa_general()
    assert abs(kg - fleiss_kappa_general(counts)) <= 1e-12 * abs(kg)
    if sorted_rows:
        tot = np.zeros_like(t.agree_hist)
        for r in range(3):
            lo, hi, r0, r1 = shard_rows_by_image(img, n_images, r, 3)
            tot += labels.label_tally(img[r0:r1], cls[r0:r1], act[r0:r1], hi - lo, k, image_base=lo).agree_hist
        assert np.array_equal(tot, t.agree_hist)
    # device-pointer form + the count-matrix form agree with it
    d_counts = torch.from_numpy(counts).cuda()
    hist = torch.empty(1024, dtype=torch.int64, device="cuda")
    engine.fleiss_partials_device(d_counts, agree_hist=hist)
    assert np.array_equal(hist.cpu().numpy(), t.agree_hist)


def test_agreement_histogram_overflow_bin():
    """Images with >= B2_AGREE_BINS ratings are counted in bin 0 and kappa_general refuses to answer."""
    img = np.concatenate([np.zeros(1500, np.int32), np.ones(10, np.int32)])
    cls = (np.arange(img.size) % 4).astype(np.uint8)
    t = labels.label_tally(img, cls, np.ones(img.size, np.uint8), 2, 4)
    assert int(t.agree_hist[0]) == 1 and int(t.agree_hist[10]) == 2 * 3 * 3 + 2 * 2 * 2 - 10
    with pytest.raises(ValueError):
        t.kappa_general()
