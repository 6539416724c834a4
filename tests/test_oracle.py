"""Pins the CPU oracle (oracle/pp_oracle.c) to the reference: its own KAT, fixtures generated
by running the reference's Numba kernels (oracle/make_golden.py), and Numba MT19937 digests."""
import hashlib

import numpy as np
import pytest

from conftest import CYCLONE_CASE_NAMES, SANDBAG_CASE_NAMES
from oracle import oracle as orc


def test_minref_known_answer(sandbag_minref):
    """pypairs/tests/test_sandbag.py:26-59 -- the only in-tree KAT of the reference."""
    m = sandbag_minref
    x = np.array(m["matrix_genes_by_samples"], dtype=np.float64).T  # samples x genes
    keep_genes = ~np.all(x == 0, axis=0)
    names = list(m["annotation"].keys())
    lab = np.full(x.shape[0], -1)
    for k, idx in enumerate(m["annotation"].values()):
        lab[idx] = k
    keep = lab >= 0
    gn = np.array(m["gene_names"])[keep_genes]
    expected_set = {
        "G1": {("X0", "X2"), ("X0", "X7"), ("X10", "X3"), ("X5", "X4"), ("X8", "X4"), ("X5", "X7"), ("X8", "X7"),
               ("X10", "X7"), ("X8", "X9")},
        "S": {("X3", "X0"), ("X5", "X0"), ("X3", "X2"), ("X4", "X2"), ("X7", "X2"), ("X3", "X8"), ("X3", "X9"),
              ("X5", "X8"), ("X6", "X8")},
        "G2M": {("X0", "X10"), ("X2", "X10"), ("X4", "X10"), ("X9", "X5"), ("X6", "X10"), ("X8", "X10"),
                ("X9", "X10")}}
    for frac_s, ref in m["results"].items():
        frac = float(frac_s)
        sizes = np.bincount(lab[keep], minlength=len(names))
        res = orc.check_pairs(x[keep][:, keep_genes], lab[keep], orc.calc_thresholds(sizes, frac))
        got = orc.result_to_markers(res, gn, names)
        assert [[k, [list(p) for p in v]] for k, v in got.items()] == ref  # key order and list order
        if frac == 0.65:
            assert {k: set(map(tuple, v)) for k, v in got.items()} == expected_set


@pytest.mark.parametrize("name", SANDBAG_CASE_NAMES)
def test_check_pairs_vs_reference(sandbag_cases, name):
    c = sandbag_cases
    x, cats, thr = c[name + "_x"], c[name + "_cats"], c[name + "_thr"]
    sizes = np.bincount(cats, minlength=len(thr))
    assert np.array_equal(orc.calc_thresholds(sizes, float(c[name + "_frac"])), thr)
    res = orc.check_pairs(x, cats, thr)
    rows, cols = np.where(res != -1)
    assert np.array_equal(rows, c[name + "_rows"])
    assert np.array_equal(cols, c[name + "_cols"])
    assert np.array_equal(res[rows, cols], c[name + "_cls"])


def test_threshold_float_ceil_edge():
    assert orc.calc_thresholds([10], 0.7)[0] == 7
    assert orc.calc_thresholds([100], 0.55)[0] == 56  # 100*0.55 == 55.00000000000001 in IEEE double
    assert orc.calc_thresholds([100], 0.07)[0] == 8   # 7.000000000000001
    assert orc.calc_thresholds([91, 80, 76], 0.6).tolist() == [55, 48, 46]
    assert orc.calc_thresholds([91, 80, 76], 0.65).tolist() == [60, 52, 50]


def test_mt19937_known_answers(mt_golden):
    assert orc.mt_outputs(2803, 6).tolist() == [1626997105, 3209404029, 1774804572, 3134047717, 1668877626,
                                                 3659196522]
    for n, info in mt_golden["tables"].items():
        if "rows" in info:
            assert orc.perm_table(99, 7, 6).tolist() == info["rows"]
            continue
        t = orc.perm_table(2803, int(n), 1000)
        assert t[0][:8].tolist() == info["first_row_head"]
        assert t[-1][:8].tolist() == info["last_row_head"]
        assert hashlib.sha256(t.astype("<i8").tobytes()).hexdigest() == info["sha256_int64le_1000xn"]


@pytest.mark.parametrize("name", CYCLONE_CASE_NAMES)
def test_phase_scores_vs_reference(cyclone_cases, name):
    npz, meta = cyclone_cases
    m = meta[name]
    x = npz[name + "_x"]
    markers = {k: [tuple(p) for p in v] for k, v in m["markers"].items()}
    pairs, used = orc.filter_marker_pairs(markers, m["gene_names"])
    for k in m["classes"]:
        assert np.array_equal(pairs[k], npz["{}_pairs_{}".format(name, k)])
        assert np.array_equal(used[k], npz["{}_used_{}".format(name, k)])
        ref = npz["{}_score_{}".format(name, k)]
        got = orc.phase_scores(x, m["iterations"], m["min_iter"], m["min_pairs"], pairs[k], used[k])
        assert np.array_equal(got, ref, equal_nan=True)
        # table-replay form is the same function of the data
        tab = orc.perm_table(2803, int(used[k].sum()), m["iterations"])
        got2 = orc.phase_scores_table(x, tab, m["min_iter"], m["min_pairs"], pairs[k], used[k])
        assert np.array_equal(got2, ref, equal_nan=True)


def test_philox_known_answer():
    """Random123 kat_vectors: philox4x32-10 zero key/counter and the pi-digits vector."""
    assert orc.philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox([0xffffffff] * 4, [0xffffffff] * 2).tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6,
                                                                          0x6d5451fd]
    assert orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    t = orc.perm_table_philox(2803, 1, 50, 20)
    assert all(sorted(r.tolist()) == list(range(50)) for r in t)
    assert len({tuple(r.tolist()) for r in t}) == 20


def test_numba_reference_copy_matches_port():
    """oracle/_ref (the reference's own Numba kernels, copied by oracle/build_ref.py where /root/reference exists) and
    the C restatement agree bit for bit on both hot paths -- the two CPU arms bench.py reports side by side."""
    from oracle import ref_numba
    if not ref_numba.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    ref_numba.configure_threads(2)
    rng = np.random.default_rng(11)
    n, g, c = 90, 60, 3
    lab = rng.integers(0, c, n)
    x = rng.poisson(3.0 * np.exp(rng.normal(0, 1, g))[None, :] * np.where(lab[:, None] == (np.arange(g) % c)[None, :], 3, 1)
                    ).astype(np.float64)
    thr = orc.calc_thresholds(np.bincount(lab, minlength=c), 0.6)
    assert np.array_equal(ref_numba.check_pairs(x, lab, thr, 2), orc.check_pairs(x, lab, thr))
    pairs = rng.integers(0, 40, (70, 2))
    used = np.zeros(g, dtype=bool)
    used[rng.choice(g, 40, replace=False)] = True
    a = ref_numba.phase_scores(x, 50, 10, 2, pairs, used)
    assert np.array_equal(a, orc.phase_scores(x, 50, 10, 2, pairs, used), equal_nan=True)
