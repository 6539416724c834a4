"""GPU parity: pp_cyclone / pp_phase_scores (through the C ABI) vs reference goldens and the CPU oracle."""
import numpy as np
import pytest

from conftest import CYCLONE_CASE_NAMES
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _ffi():
    from pypairs_b200 import _ffi
    return _ffi


def default_setup(rng, n_cells, n_genes, kind="poisson", dtype=np.float32):
    from pypairs_b200 import datasets
    markers = datasets.default_cc_marker()
    genes = sorted({a for v in markers.values() for p in v for a in p})
    names = genes + ["X{}".format(i) for i in range(n_genes - len(genes))]
    rng.shuffle(names)
    if kind == "poisson":
        x = rng.poisson(2.0, (n_cells, n_genes)).astype(dtype)
    else:
        x = rng.lognormal(0.0, 1.0, (n_cells, n_genes)).astype(dtype)
    return markers, names, x


@pytest.mark.parametrize("name", CYCLONE_CASE_NAMES)
def test_golden_cases(cyclone_cases, name):
    """Scores of the reference's Numba kernels, bit for bit (MT19937 replay), incl. the NaN / 0.0 quirks."""
    from pypairs_b200.cyclone import filter_marker_pairs
    npz, meta = cyclone_cases
    m = meta[name]
    x = npz[name + "_x"]
    markers = {k: [tuple(p) for p in v] for k, v in m["markers"].items()}
    pairs, used = filter_marker_pairs(markers, m["gene_names"])
    ks = m["classes"]
    for k in ks:
        assert np.array_equal(pairs[k], npz["{}_pairs_{}".format(name, k)])
        assert np.array_equal(used[k], npz["{}_used_{}".format(name, k)])
    sc, tm, oh, ot = _ffi().cyclone(x, [np.where(used[k])[0] for k in ks], [pairs[k] for k in ks], m["iterations"],
                                    m["min_iter"], m["min_pairs"], want_observed=True)
    for i, k in enumerate(ks):
        assert np.array_equal(sc[i], npz["{}_score_{}".format(name, k)], equal_nan=True), k
        _, eh, et = orc.phase_scores(x, 1, 1, 1, pairs[k], used[k], return_observed=True)
        assert np.array_equal(oh[i], eh) and np.array_equal(ot[i], et)  # observed (hits, total) exact
        # reference-kernel-shaped single-class entry point
        one = _ffi().phase_scores(x, m["iterations"], m["min_iter"], m["min_pairs"], pairs[k], used[k])
        assert np.array_equal(one, sc[i], equal_nan=True)


def test_perm_tables(mt_golden):
    import hashlib
    f = _ffi()
    for n, info in mt_golden["tables"].items():
        if "rows" in info:
            assert f.perm_table("mt19937", 99, 0, 7, 6).tolist() == info["rows"]
        else:
            t = f.perm_table("mt19937", 2803, 0, int(n), 1000)
            assert hashlib.sha256(t.astype("<i8").tobytes()).hexdigest() == info["sha256_int64le_1000xn"]
    for seed, cls, n, it in [(2803, 0, 1195, 40), (2**40 + 5, 2, 17, 100), (1, 1, 1, 3), (7, 0, 2, 9)]:
        assert np.array_equal(f.perm_table("philox", seed, cls, n, it), orc.perm_table_philox(seed, cls, n, it))


@pytest.mark.parametrize("kind,dtype", [("poisson", np.float32), ("lognormal", np.float64)])
def test_default_markers_vs_oracle(kind, dtype):
    from pypairs_b200.cyclone import filter_marker_pairs
    rng = np.random.default_rng(42)
    markers, names, x = default_setup(rng, 150, 2500, kind, dtype)
    x[7] = x[6]
    pairs, used = filter_marker_pairs(markers, names)
    ks = list(pairs.keys())
    u = [np.where(used[k])[0] for k in ks]
    p = [pairs[k] for k in ks]
    sc, tm = _ffi().cyclone(x, u, p, 1000, 100, 50)
    for i, k in enumerate(ks):
        ref = orc.phase_scores(x, 1000, 100, 50, pairs[k], used[k])
        assert np.array_equal(sc[i], ref, equal_nan=True), k
    assert np.array_equal(sc[:, 6], sc[:, 7])
    # cells are independent: a permuted matrix gives permuted scores; a row range gives that slice
    perm = rng.permutation(150)
    sc2, _ = _ffi().cyclone(x[perm], u, p, 1000, 100, 50)
    assert np.array_equal(sc2, sc[:, perm], equal_nan=True)
    sc3, _ = _ffi().cyclone(x, u, p, 1000, 100, 50, row_begin=33, n_cells=70)
    assert np.array_equal(sc3, sc[:, 33:103], equal_nan=True)
    # validation mode: the reference's exported permutation tables give the same scores
    tabs = [orc.perm_table(2803, len(ui), 1000) for ui in u]
    sc4, _ = _ffi().cyclone(x, u, p, 1000, 100, 50, rng="table", perm_tables=tabs)
    assert np.array_equal(sc4, sc, equal_nan=True)


def test_philox_mode():
    """Bit-exact against the CPU restatement of the same Philox stream; statistically equivalent to MT19937."""
    from pypairs_b200.cyclone import filter_marker_pairs
    rng = np.random.default_rng(43)
    markers, names, x = default_setup(rng, 512, 2000, "lognormal", np.float32)
    pairs, used = filter_marker_pairs(markers, names)
    ks = list(pairs.keys())
    u = [np.where(used[k])[0] for k in ks]
    p = [pairs[k] for k in ks]
    sp, _ = _ffi().cyclone(x, u, p, 1000, 100, 50, rng="philox", seed=2803)
    sm, _ = _ffi().cyclone(x, u, p, 1000, 100, 50, rng="mt19937", seed=2803)
    for i, k in enumerate(ks):
        tab = orc.perm_table_philox(2803, i, len(u[i]), 1000)
        ref = orc.phase_scores_table(x, tab, 100, 50, pairs[k], used[k])
        assert np.array_equal(sp[i], ref, equal_nan=True)
    d = np.abs(sp - sm)
    bound = 4.0 * np.sqrt(2.0 * sm * (1.0 - sm) / 1000.0) + 1e-3 + 1.0 / 1000.0
    assert d.mean() <= 0.02                      # north-star tolerance, stated as the mean |delta score|
    assert np.mean(d <= bound) >= 0.999          # per-cell z-bound (two independent 1000-draw nulls)
    print("philox vs mt19937: mean |d| {:.4f}, max |d| {:.4f}, within 0.02: {:.3f}".format(
        d.mean(), d.max(), np.mean(d <= 0.02)))


def test_api_frame_and_anndata():
    from pandas import DataFrame
    from pypairs_b200 import pairs
    rng = np.random.default_rng(44)
    markers, names, x = default_setup(rng, 40, 1700, "poisson", np.float64)
    sn = ["c{}".format(i) for i in range(40)]
    df = pairs.cyclone(DataFrame(x, index=sn, columns=names))  # marker_pairs=None -> default_cc_marker
    assert list(df.columns) == ["G1", "S", "G2M", "max_class", "cc_prediction"] and list(df.index) == sn
    pr, us = orc.filter_marker_pairs(markers, names)
    ref = {k: orc.phase_scores(x, 1000, 100, 50, pr[k], us[k]) for k in ["G1", "S", "G2M"]}
    exp = orc.cyclone_epilogue(ref, sn)
    for k in ["G1", "S", "G2M"]:
        assert np.array_equal(df[k].values, exp[k].values, equal_nan=True)
    assert list(df["max_class"]) == list(exp["max_class"])
    assert [str(v) for v in df["cc_prediction"]] == [str(v) for v in exp["cc_prediction"]]

    class FakeAnnData:
        def __init__(self):
            self.X = x
            self.obs = DataFrame(index=sn)
            self.obs_names = sn
            self.var_names = names

    ad = FakeAnnData()
    df2 = pairs.cyclone(ad, markers, iterations=200, min_iter=10, min_pairs=1)
    assert all("pypairs_" + c in ad.obs.columns for c in df2.columns)
    assert np.array_equal(ad.obs["pypairs_G1"].values, df2["G1"].values, equal_nan=True)
    with pytest.raises(SystemExit):
        pairs.cyclone(x, {"A": [("nope", "nada")]}, names, sn)


def test_full_size_slice_properties():
    """cfg2 shape (20k genes) on 4096 device-generated cells: oracle on a row subset + invariants."""
    import torch
    from pypairs_b200.cyclone import filter_marker_pairs
    rng = np.random.default_rng(45)
    markers, names, _ = default_setup(rng, 1, 20000)
    gen = torch.Generator(device="cuda").manual_seed(20262)
    x = torch.poisson(torch.full((4096, 20000), 2.0, device="cuda"), generator=gen).to(torch.float32)
    pairs, used = filter_marker_pairs(markers, names)
    ks = list(pairs.keys())
    u = [np.where(used[k])[0] for k in ks]
    p = [pairs[k] for k in ks]
    sc, tm = _ffi().cyclone(x, u, p, 1000, 100, 50)
    assert np.all((sc >= 0) & (sc <= 1)) and np.all(sc * 1000 == np.round(sc * 1000))
    rows = np.sort(rng.choice(4096, 48, replace=False))
    xs = x[torch.from_numpy(rows).cuda()].cpu().numpy()
    for i, k in enumerate(ks):
        assert np.array_equal(sc[i, rows], orc.phase_scores(xs, 1000, 100, 50, pairs[k], used[k]))
    print("cfg2 slice 4096 cells: score kernel {:.1f} ms, prep {:.1f} ms".format(tm["main_ms"], tm["prep_ms"]))


def test_host_input_is_streamed_in_chunks():
    """Host matrices larger than one wave of tiles take the copy/compute-overlapped path: same scores."""
    import torch
    from pypairs_b200.cyclone import filter_marker_pairs
    rng = np.random.default_rng(46)
    markers, names, _ = default_setup(rng, 1, 1700)
    n = 148 * 64 * 2 + 777
    x = rng.poisson(2.0, (n, 1700)).astype(np.float32)
    pairs, used = filter_marker_pairs(markers, names)
    ks = list(pairs.keys())
    u = [np.where(used[k])[0] for k in ks]
    p = [pairs[k] for k in ks]
    s_host, tm_h, oh, ot = _ffi().cyclone(x, u, p, 40, 10, 5, want_observed=True)
    s_dev, tm_d, oh2, ot2 = _ffi().cyclone(torch.from_numpy(x).cuda(), u, p, 40, 10, 5, want_observed=True)
    assert np.array_equal(s_host, s_dev, equal_nan=True)
    assert np.array_equal(oh, oh2) and np.array_equal(ot, ot2)
    pin = torch.from_numpy(x).pin_memory()
    s_pin, _ = _ffi().cyclone(pin.numpy(), u, p, 40, 10, 5)
    assert np.array_equal(s_pin, s_dev, equal_nan=True)
    rows = np.sort(rng.choice(n, 40, replace=False))
    for i, k in enumerate(ks):
        assert np.array_equal(s_host[i, rows], orc.phase_scores(x[rows], 40, 10, 5, pairs[k], used[k]))


def _random_marker_case(rng, n_genes, n_pairs_per_class, n_classes, n_cells):
    names = ["g{}".format(i) for i in range(n_genes)]
    markers = {}
    for c in range(n_classes):
        a = rng.integers(0, n_genes, n_pairs_per_class)
        b = rng.integers(0, n_genes, n_pairs_per_class)
        markers["class{}".format(c)] = [(names[i], names[j]) for i, j in zip(a, b)]
    x = rng.poisson(3.0, (n_cells, n_genes)).astype(np.float64)
    return names, markers, x


def test_large_marker_sets_take_the_general_path():
    """More marker genes than the shared-memory tile holds (global-memory tile, integer compares) and more
    pairs per class than a 16-bit counter holds: still bit-exact against the oracle."""
    from pypairs_b200.cyclone import filter_marker_pairs
    rng = np.random.default_rng(47)
    for n_genes, n_pairs, n_cells, iters in [(2600, 9000, 70, 30), (300, 70000, 40, 12), (5000, 3000, 130, 25)]:
        names, markers, x = _random_marker_case(rng, n_genes, n_pairs, 2, n_cells)
        pairs, used = filter_marker_pairs(markers, names)
        ks = list(pairs.keys())
        sc, tm, oh, ot = _ffi().cyclone(x, [np.where(used[k])[0] for k in ks], [pairs[k] for k in ks], iters, 5, 3,
                                        want_observed=True)
        for i, k in enumerate(ks):
            ref, eh, et = orc.phase_scores(x, iters, 5, 3, pairs[k], used[k], return_observed=True)
            assert np.array_equal(oh[i], eh) and np.array_equal(ot[i], et), (n_genes, n_pairs, k)
            assert np.array_equal(sc[i], ref, equal_nan=True), (n_genes, n_pairs, k)


def test_sandbag_then_cyclone_flow():
    """configs[0] shape (leng15-like: 247 sorted cells 76/80/91, 19 084 genes; the real matrix is not in the
    checkout): markers from pairs.sandbag feed pairs.cyclone; both checked against the oracle."""
    from pypairs_b200 import pairs
    rng = np.random.default_rng(48)
    g, sizes = 19084, {"G2M": 76, "S": 80, "G1": 91}
    lab = np.concatenate([np.full(n, i) for i, n in enumerate(sizes.values())])
    base = rng.normal(0.0, 1.5, g)
    de = rng.random(g) < 0.03
    de_class = rng.integers(0, 3, g)
    lam = 3.0 * np.exp(base)[None, :] * np.where((lab[:, None] == de_class[None, :]) & de[None, :], 6.0, 1.0)
    x = rng.poisson(lam).astype(np.float32)
    gn = ["gene{}".format(i) for i in range(g)]
    sn = ["cell{}".format(i) for i in range(len(lab))]
    ann = {k: np.where(lab == i)[0].tolist() for i, k in enumerate(sizes.keys())}
    markers = pairs.sandbag(x, ann, gn, sn, fraction=0.6)
    assert set(markers.keys()) == set(sizes.keys()) and sum(len(v) for v in markers.values()) > 1000
    # oracle on a gene subset that contains marker genes
    top = sorted({a for v in markers.values() for a, _ in v[:40]} | {b for v in markers.values() for _, b in v[:40]})
    cols = np.array(sorted(int(n[4:]) for n in top))[:160]
    keep = cols[(x[:, cols] != 0).any(axis=0)]
    rr, cc, kk = orc.sandbag_arrays(x[:, keep].astype(np.float64), lab, 3, 0.6)
    names = list(sizes.keys())
    exp = {(gn[keep[r]], gn[keep[c]], names[k]) for r, c, k in zip(rr, cc, kk)}
    kept_names = {gn[i] for i in keep}
    got = {(a, b, k) for k, v in markers.items() for a, b in v if a in kept_names and b in kept_names}
    assert got == exp
    # score an "unsorted" batch with the markers just found (thousands of pairs per class)
    y = rng.poisson(3.0 * np.exp(base)[None, :] * np.ones((60, 1))).astype(np.float32)
    sub = {k: v[:3000] for k, v in markers.items()}
    df = pairs.cyclone(y, sub, gn, ["u{}".format(i) for i in range(60)], iterations=100, min_iter=10, min_pairs=5)
    pr, us = orc.filter_marker_pairs(sub, gn)
    for k in sub:
        assert np.array_equal(df[k].values, orc.phase_scores(y, 100, 10, 5, pr[k], us[k]), equal_nan=True)


def test_device_resident_flow_with_marker_table():
    """The same flow with both matrices resident on the device and the markers kept columnar (utils.MarkerTable):
    nothing is uploaded twice, no tuples, no per-pair name look-ups -- and the frame equals the host / dict flow."""
    import torch
    from pypairs_b200 import pairs, utils
    rng = np.random.default_rng(52)
    g, sizes = 6000, {"G2M": 50, "S": 61, "G1": 70}
    lab = np.concatenate([np.full(n, i) for i, n in enumerate(sizes.values())])
    base = rng.normal(0.0, 1.5, g)
    de = rng.random(g) < 0.03
    de_class = rng.integers(0, 3, g)
    lam = 3.0 * np.exp(base)[None, :] * np.where((lab[:, None] == de_class[None, :]) & de[None, :], 6.0, 1.0)
    x = rng.poisson(lam).astype(np.float32)
    y = rng.poisson(3.0 * np.exp(base)[None, :] * np.ones((77, 1))).astype(np.float32)
    gn = ["gene{}".format(i) for i in range(g)]
    sn, un = ["cell{}".format(i) for i in range(len(lab))], ["u{}".format(i) for i in range(77)]
    ann = {k: np.where(lab == i)[0].tolist() for i, k in enumerate(sizes.keys())}
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    table = pairs.sandbag(xd, ann, gn, sn, fraction=0.6, as_table=True)
    assert isinstance(table, utils.MarkerTable) and len(table) > 1000
    df_dev = pairs.cyclone(yd, table, gn, un, iterations=200, min_iter=20, min_pairs=5)
    markers = pairs.sandbag(x, ann, gn, sn, fraction=0.6)
    assert list(table.to_dict().items()) == list(markers.items())
    df_host = pairs.cyclone(y, markers, gn, un, iterations=200, min_iter=20, min_pairs=5)
    assert df_dev.equals(df_host)
    pr, us = orc.filter_marker_pairs(markers, gn)
    for k in markers:
        assert np.array_equal(df_dev[k].values[:12], orc.phase_scores(y[:12], 200, 20, 5, pr[k], us[k]), equal_nan=True)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_host_gather_path(dtype):
    """Wide host matrices (marker genes <= 1/4 of the columns): marker columns are gathered on the host into
    pinned staging buffers before the copy.  Same scores as the device-resident path and as the oracle."""
    import torch
    from pypairs_b200.cyclone import filter_marker_pairs
    rng = np.random.default_rng(49)
    markers, names, _ = default_setup(rng, 1, 8000)
    n = 148 * 64 + 500  # two chunks
    x = rng.poisson(2.0, (n, 8000)).astype(dtype)
    x[5] = rng.normal(size=8000)  # one non-count row (sort path of the rank kernel)
    pairs, used = filter_marker_pairs(markers, names)
    ks = list(pairs.keys())
    u = [np.where(used[k])[0] for k in ks]
    p = [pairs[k] for k in ks]
    s_host, tm_h = _ffi().cyclone(x, u, p, 30, 10, 5)
    s_dev, tm_d = _ffi().cyclone(torch.from_numpy(x).cuda(), u, p, 30, 10, 5)
    assert np.array_equal(s_host, s_dev, equal_nan=True)
    s_part, _ = _ffi().cyclone(x, u, p, 30, 10, 5, row_begin=1000, n_cells=777)  # single short chunk
    assert np.array_equal(s_part, s_dev[:, 1000:1777], equal_nan=True)
    rows = np.sort(rng.choice(n, 24, replace=False))
    rows[0] = 5
    for i, k in enumerate(ks):
        assert np.array_equal(s_host[i, rows], orc.phase_scores(x[rows], 30, 10, 5, pairs[k], used[k]))


@pytest.mark.parametrize("dtype,feed", [(np.float32, "auto"), (np.float32, "zerocopy"), (np.float64, "zerocopy"),
                                        (np.float32, "gather"), (np.float32, "copy"), (np.float64, "both"),
                                        (np.float32, "both")])
def test_pinned_input_zero_copy_lane(monkeypatch, dtype, feed):
    """A pinned host matrix can be read in place by the zero-copy gather kernel (mapped memory, marker columns only),
    alone or beside the host gather: every way of feeding the chunks (PP_FEED) gives the scores of the device path."""
    import torch
    from pypairs_b200.cyclone import filter_marker_pairs
    rng = np.random.default_rng(50)
    markers, names, _ = default_setup(rng, 1, 6400)
    n = 3 * 148 * 64 + 777  # several full chunks per lane
    xt = torch.from_numpy(rng.poisson(2.0, (n, 6400)).astype(dtype)).pin_memory()
    x = xt.numpy()
    pairs, used = filter_marker_pairs(markers, names)
    ks = list(pairs.keys())
    u = [np.where(used[k])[0] for k in ks]
    p = [pairs[k] for k in ks]
    s_dev, _ = _ffi().cyclone(xt.cuda(), u, p, 20, 10, 5)
    monkeypatch.setenv("PP_FEED", feed)
    s_pin, tm = _ffi().cyclone(x, u, p, 20, 10, 5)
    assert np.array_equal(s_pin, s_dev, equal_nan=True)
    s_part, _ = _ffi().cyclone(x, u, p, 20, 10, 5, row_begin=100, n_cells=n - 150)  # a row range of the pinned matrix
    assert np.array_equal(s_part, s_dev[:, 100:n - 50], equal_nan=True)
    s_page, _ = _ffi().cyclone(np.array(x), u, p, 20, 10, 5)  # pageable copy: no zero-copy lane, same answer
    assert np.array_equal(s_page, s_dev, equal_nan=True)
    assert np.array_equal(s_dev[0, :16], orc.phase_scores(x[:16], 20, 10, 5, pairs[ks[0]], used[ks[0]]), equal_nan=True)


def _score_vs_oracle(x, pairs, used, iterations, min_iter, min_pairs):
    got = _ffi().phase_scores(x, iterations, min_iter, min_pairs, pairs, used)
    exp = orc.phase_scores(x, iterations, min_iter, min_pairs, pairs, used)
    assert np.array_equal(got, exp, equal_nan=True)
    return got


def test_degenerate_cells_and_parameters():
    """Cells where every pair ties (score 0 by cyclone.py:223-229 'cur_score == 0' rule), cells below min_pairs,
    min_iter above what the permutations can reach, iterations = 1, and a single marker pair."""
    rng = np.random.default_rng(41)
    n, g = 150, 120
    x = rng.poisson(1.0, (n, g)).astype(np.float64)
    x[3] = 2.0                                        # all ties: no informative pair
    x[4, :] = 0.0
    x[4, 7] = 1.0                                     # one or two informative pairs only
    used = (rng.random(g) < 0.6).astype(np.uint8)     # pairs index the used-compacted genes (cyclone.py:303-309)
    used[7] = 1
    pairs = rng.integers(0, int(used.sum()), (300, 2))
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    for iterations, min_iter, min_pairs in [(200, 50, 10), (200, 250, 10), (1, 1, 1), (64, 1, 400), (100, 100, 0),
                                            (0, 0, 1), (0, 1, 1), (3, 0, 1)]:
        _score_vs_oracle(x, pairs, used, iterations, min_iter, min_pairs)
    u1 = np.zeros(g, dtype=np.uint8)
    u1[[5, 9]] = 1
    _score_vs_oracle(x, np.array([[0, 1]]), u1, 50, 10, 1)


def test_repeated_and_mirrored_pairs():
    """The reference keeps duplicate pairs and both orientations (cyclone.py:273-311 never dedups)."""
    rng = np.random.default_rng(43)
    n, g = 200, 60
    x = rng.poisson(2.0, (n, g)).astype(np.float32)
    used = (rng.random(g) < 0.7).astype(np.uint8)
    base = rng.integers(0, int(used.sum()), (80, 2))
    base = base[base[:, 0] != base[:, 1]]
    pairs = np.concatenate([base, base[:30], base[:40, ::-1]])
    _score_vs_oracle(x, pairs, used, 300, 100, 5)


@pytest.mark.parametrize("n_cells,n_used,n_pairs,iterations", [(5, 300, 900, 1000), (213, 700, 3000, 1000),
                                                               (70, 2500, 6000, 333), (130, 400, 500, 64)])
def test_few_cells_split_iterations_over_ctas(n_cells, n_used, n_pairs, iterations):
    """Fewer (class, tile) units than SMs: the iterations are split into slices over several CTAs (shared-memory and
    global-tile paths); scores must not depend on how the work was cut."""
    rng = np.random.default_rng(n_cells)
    g = 3000
    x = rng.poisson(1.5, (n_cells, g)).astype(np.float32)
    used = np.zeros(g, dtype=np.uint8)
    used[rng.choice(g, n_used, replace=False)] = 1
    pairs = rng.integers(0, n_used, (n_pairs, 2))
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    got = _score_vs_oracle(x, pairs, used, iterations, 10, 5)
    assert np.isfinite(got).all()
    sc, tm, oh, ot = _ffi().cyclone(x, [np.where(used)[0]], [pairs], iterations, 10, 5, want_observed=True)
    _, eh, et = orc.phase_scores(x, 1, 1, 1, pairs, used, return_observed=True)
    assert np.array_equal(sc[0], got) and np.array_equal(oh[0], eh) and np.array_equal(ot[0], et)


@pytest.mark.parametrize("n_cells,n_genes,n_used,n_pairs,host", [(300, 2500, 400, 3000, False),
                                                                (150, 3000, 2200, 5000, False),
                                                                (10500, 400, 120, 300, True)])
def test_table_passes_give_identical_scores(monkeypatch, n_cells, n_genes, n_used, n_pairs, host):
    """The permutation tables are built for k_pass iterations at a time when they would not fit the memory budget;
    forcing a 1 MB budget (many passes) must not change a single score or observed counter."""
    import torch
    rng = np.random.default_rng(n_cells)
    x = rng.poisson(1.5, (n_cells, n_genes)).astype(np.float32)
    used = np.sort(rng.choice(n_genes, n_used, replace=False))
    pairs = rng.integers(0, n_used, (n_pairs, 2))
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    xin = x if host else torch.from_numpy(x).cuda()
    ref = _ffi().cyclone(xin, [used, used[: n_used // 2]], [pairs, pairs[pairs.max(axis=1) < n_used // 2]], 500, 10, 5,
                         want_observed=True)
    monkeypatch.setenv("PP_TABLE_BUDGET_MB", "1")
    got = _ffi().cyclone(xin, [used, used[: n_used // 2]], [pairs, pairs[pairs.max(axis=1) < n_used // 2]], 500, 10, 5,
                         want_observed=True)
    assert got[1]["n_launches"] > ref[1]["n_launches"]
    assert np.array_equal(got[0], ref[0], equal_nan=True)
    assert np.array_equal(got[2], ref[2]) and np.array_equal(got[3], ref[3])
    um = np.zeros(n_genes, dtype=np.uint8)
    um[used] = 1
    sub = slice(0, 64)
    assert np.array_equal(got[0][0][sub], orc.phase_scores(x[sub], 500, 10, 5, pairs, um))


def test_seven_bit_kernel_boundary_and_mid_call_switch():
    """The four-cells-per-word kernel runs while no ranked cell has more than 128 distinct values (max dense rank 127);
    one more distinct value switches to the fp16 kernel -- also in the middle of a chunked call, where early chunks
    were scored by one kernel and later chunks by the other.  Scores must be the oracle's either way."""
    rng = np.random.default_rng(77)
    n_genes, n_used = 600, 300
    used = np.sort(rng.choice(n_genes, n_used, replace=False))
    um = np.zeros(n_genes, dtype=np.uint8)
    um[used] = 1
    pairs = rng.integers(0, n_used, (1500, 2))
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]

    def cells(n, n_distinct):
        x = rng.integers(0, n_distinct, (n, n_genes)).astype(np.float32)
        x[:, used[:n_distinct]] = np.arange(n_distinct, dtype=np.float32)  # every cell really has n_distinct values
        return x

    for n_distinct in (128, 129):
        x = cells(100, n_distinct)
        got = _ffi().phase_scores(x, 300, 10, 5, pairs, um)
        assert np.array_equal(got, orc.phase_scores(x, 300, 10, 5, pairs, um), equal_nan=True), n_distinct
    # host input longer than one wave of tiles is chunked: 9 600 low-rank cells, then 400 cells with 200 distinct values
    x = np.concatenate([cells(9600, 20), cells(400, 200)])
    got = _ffi().phase_scores(x, 200, 10, 5, pairs, um)
    for sl in (slice(0, 96), slice(9472, 9472 + 64), slice(9600, 9600 + 64), slice(9936, 10000)):
        assert np.array_equal(got[sl], orc.phase_scores(x[sl], 200, 10, 5, pairs, um), equal_nan=True)


@pytest.mark.parametrize("n_cells", [1, 127, 129, 1184, 1185, 9472, 9473, 21000])
def test_host_pipeline_chunk_boundaries(n_cells):
    """Host input goes through the gather + ramped-chunk pipeline, device input through one launch: same scores
    for cell counts around every chunk boundary (an eighth / quarter / half / full wave of 148 x 64 cells)."""
    import torch
    rng = np.random.default_rng(n_cells)
    n_genes, n_used = 700, 90
    x = rng.poisson(1.2, (n_cells, n_genes)).astype(np.float32)
    used = np.sort(rng.choice(n_genes, n_used, replace=False))
    pairs = rng.integers(0, n_used, (200, 2))
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    host = _ffi().cyclone(x, [used], [pairs], 120, 10, 3, want_observed=True)
    dev = _ffi().cyclone(torch.from_numpy(x).cuda(), [used], [pairs], 120, 10, 3, want_observed=True)
    assert np.array_equal(host[0], dev[0], equal_nan=True)
    assert np.array_equal(host[2], dev[2]) and np.array_equal(host[3], dev[3])
    um = np.zeros(n_genes, dtype=np.uint8)
    um[used] = 1
    k = min(n_cells, 40)
    assert np.array_equal(host[0][0][-k:], orc.phase_scores(x[-k:], 120, 10, 3, pairs, um), equal_nan=True)


def test_quantile_transform_pass_through():
    """quantile_transform=True (pypairs/tools/cyclone.py:138-139): sklearn on the host, then the same scoring; the
    oracle scores the sklearn-transformed matrix.  <= 10 000 cells: sklearn does not sub-sample, so the fit is the
    reference's own (unseeded) one."""
    from sklearn.preprocessing import QuantileTransformer
    from pypairs_b200 import pairs
    rng = np.random.default_rng(46)
    markers, names, x = default_setup(rng, 60, 1600, "poisson", np.float64)
    sn = ["c{}".format(i) for i in range(60)]
    df = pairs.cyclone(x, markers, names, sn, iterations=120, min_iter=20, min_pairs=5, quantile_transform=True)
    xt = QuantileTransformer().fit_transform(x)
    assert not np.array_equal(xt, x)
    pr, us = orc.filter_marker_pairs(markers, names)
    for k in ["G1", "S", "G2M"]:
        assert np.array_equal(df[k].values, orc.phase_scores(xt, 120, 20, 5, pr[k], us[k]), equal_nan=True), k
    import scipy.sparse as sp
    df2 = pairs.cyclone(sp.csr_matrix(x), markers, names, sn, iterations=120, min_iter=20, min_pairs=5,
                        quantile_transform=True)
    assert df2.equals(df)
    import torch
    with pytest.raises(ValueError):
        pairs.cyclone(torch.from_numpy(x).cuda(), markers, names, sn, quantile_transform=True)


def test_device_inputs_are_ordered_after_torch_work():
    """The library runs on its own non-blocking streams: a CUDA tensor that torch is still producing (large async
    conversion / arithmetic enqueued right before the call) must be complete when the kernels read it."""
    import torch
    from pypairs_b200.cyclone import filter_marker_pairs
    rng = np.random.default_rng(47)
    markers, names, x = default_setup(rng, 4000, 6000, "poisson", np.float32)
    pairs, used = filter_marker_pairs(markers, names)
    ks = list(pairs.keys())
    u = [np.where(used[k])[0] for k in ks]
    p = [pairs[k] for k in ks]
    want, _ = _ffi().cyclone(x, u, p, 40, 10, 5)
    xi = torch.from_numpy(x.astype(np.int32)).cuda()
    torch.cuda.synchronize()
    for _ in range(3):
        big = torch.empty((1 << 28,), dtype=torch.float32, device="cuda")
        big.normal_()                       # keeps torch's stream busy ...
        big.mul_(2.0).add_(1.0)
        xd = (xi.to(torch.float64) * 3.0 + 1.0).to(torch.float32)  # ... ahead of a monotone transform of the counts
        got, _ = _ffi().cyclone(xd, u, p, 40, 10, 5)               # and int32 input is converted by matrix_arg itself
        assert np.array_equal(got, want, equal_nan=True)
        got, _ = _ffi().cyclone(xi, u, p, 40, 10, 5)
        assert np.array_equal(got, want, equal_nan=True)
    # sandbag through a torch sparse-CSR tensor: col_indices() are int64 and converted right before the call
    import scipy.sparse as sp
    n, g, c = 600, 900, 3
    lab = rng.integers(0, c, n)
    xs = rng.poisson(1.0, (n, g)).astype(np.float32)
    xs[:, 0] += 1
    order = np.argsort(lab, kind="stable")
    sizes = np.bincount(lab, minlength=c)
    thr = orc.calc_thresholds(sizes, 0.6)
    k0, c0, _ = _ffi().sandbag(xs, order, sizes, np.arange(g), thr)
    m = sp.csr_matrix(xs)
    for _ in range(3):
        big = torch.empty((1 << 28,), dtype=torch.float32, device="cuda").normal_()
        t = torch.sparse_csr_tensor(torch.from_numpy(m.indptr.astype(np.int64)).cuda(),
                                    torch.from_numpy(m.indices.astype(np.int64)).cuda(),
                                    torch.from_numpy(m.data).cuda(), size=m.shape)
        k1, c1, _ = _ffi().sandbag(t, order, sizes, np.arange(g), thr)
        assert np.array_equal(k0, k1) and np.array_equal(c0, c1)
        del big


def test_malformed_device_csr_is_rejected():
    import torch
    f = _ffi()
    n, g = 8, 12
    crow = torch.tensor([0, 2, 4, 4, 4, 4, 4, 4, 4], dtype=torch.int64)
    vals = torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float32)
    bad_col = torch.tensor([1, 40, 2, 3], dtype=torch.int64)
    t = torch.sparse_csr_tensor(crow.cuda(), bad_col.cuda(), vals.cuda(), size=(n, g))
    with pytest.raises(f.PyPairsB200Error) as e:
        f.sandbag(t, np.arange(n), np.array([4, 4]), np.arange(g), np.array([3, 3]))
    assert "column index out of range" in str(e.value)
    bad_crow = torch.tensor([0, 3, 2, 4, 4, 4, 4, 4, 4], dtype=torch.int64)
    t = torch.sparse_csr_tensor(bad_crow.cuda(), torch.tensor([1, 5, 2, 3]).cuda(), vals.cuda(), size=(n, g))
    with pytest.raises(f.PyPairsB200Error) as e:
        f.sandbag(t, np.arange(n), np.array([4, 4]), np.arange(g), np.array([3, 3]))
    assert "indptr decreases" in str(e.value)
