"""GPU parity: pp_sandbag (through the C ABI) vs reference goldens and the CPU oracle."""
import json

import numpy as np
import pytest

from conftest import SANDBAG_CASE_NAMES
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _ffi():
    from pypairs_b200 import _ffi
    return _ffi


def run_kernel(x, cats, n_classes, thr, **kw):
    order = np.argsort(cats, kind="stable")
    sizes = np.bincount(cats, minlength=n_classes)
    return _ffi().sandbag(x, order, sizes, np.arange(x.shape[1]), thr, **kw)


def synth(rng, n, g, c, frac_de=0.1, dtype=np.float64):
    lab = rng.integers(0, c, n)
    lab[:c] = np.arange(c)
    base = rng.normal(0.0, 1.5, g)
    de = rng.random(g) < frac_de
    de_class = rng.integers(0, c, g)
    sign = rng.choice([-1.0, 1.0], g)
    lam = np.tile(3.0 * np.exp(base), (n, 1))
    eff = np.where(de, 2.0 ** (2.0 * sign), 1.0)
    lam = np.where(lab[:, None] == de_class[None, :], lam * eff[None, :], lam)
    return rng.poisson(lam).astype(dtype), lab.astype(np.int64)


def test_dense_rank_matches_numpy():
    rng = np.random.default_rng(1)
    for dtype in (np.float64, np.float32):
        x = rng.poisson(3.0, (37, 700)).astype(dtype)
        x[0] = rng.normal(size=700).astype(dtype)          # all distinct, negative values
        x[1] = 5.0                                          # constant
        x[2, ::2] = -0.0                                    # signed zeros tie
        x[2, 1::2] = 0.0
        x[3] = np.linspace(1.0, 1.0 + 1e-12, 700) if dtype == np.float64 else x[3]  # float32-colliding doubles
        rows = np.array([5, 0, 1, 2, 3, 36, 20], dtype=np.int32)
        cols = np.sort(rng.choice(700, 333, replace=False)).astype(np.int32)
        got, mx = _ffi().dense_rank(x, rows, cols)
        sub = x[rows][:, cols]
        exp = np.stack([np.unique(r, return_inverse=True)[1] for r in sub])
        assert np.array_equal(got, exp.astype(np.uint16))
        assert mx == exp.max()


@pytest.mark.parametrize("dtype,n_cols,n_keep", [(np.float32, 5003, 5003), (np.float32, 6001, 4500), (np.float64, 3001, 2900),
                                                 (np.float32, 40000, 40000), (np.float64, 20000, 19000)])
def test_streaming_rank_path(dtype, n_cols, n_keep):
    """Rows of >= 8 KB with ascending, dense columns take the TMA-streamed count path (rank_stream_kernel): odd row
    lengths (rows start at every 16-byte phase), dropped columns, cells that are not counts (handed to the general
    kernel through the to-do list), more cells than resident CTAs."""
    rng = np.random.default_rng(n_cols)
    n_rows = 700
    x = rng.poisson(rng.gamma(0.5, 6.0, n_cols)[None, :] + 0.05, (n_rows, n_cols)).astype(dtype)
    x[3] = rng.lognormal(0, 1, n_cols).astype(dtype)         # real values: general kernel
    x[4, 100:110] = 0.5                                      # a few non-integers
    x[5] = 70000.0 + np.arange(n_cols) % 3                   # integers beyond 16 bits
    x[6] = 0
    x[7, -1] = 65535.0
    x[8, 0] = -1.0
    x[9] = (np.arange(n_cols) * 7) % 60000                   # thousands of distinct counts beyond the look-up table
    x[10] = x[9]                                             # ... twice in a row (upper bitmap words are cleared in between)
    x[12, ::3] = -0.0                                        # signed zeros tie
    cols = np.sort(rng.choice(n_cols, n_keep, replace=False)).astype(np.int32)
    rows = rng.permutation(n_rows)[:650].astype(np.int32)
    rows[:14] = [11, 3, 4, 5, 6, 7, 8, 9, 10, 9, 12, 13, 10, 2]
    got, mx = _ffi().dense_rank(x, rows, cols)
    sub = x[rows][:, cols]
    exp = np.stack([np.unique(r, return_inverse=True)[1] for r in sub])
    assert np.array_equal(got, exp.astype(np.uint16))
    assert mx == exp.max()
    xn = x.copy()
    xn[11, cols[5]] = np.nan
    with pytest.raises(_ffi().PyPairsB200Error) as e:
        _ffi().dense_rank(xn, rows, cols)
    assert e.value.code == _ffi().PP_ENAN


@pytest.mark.parametrize("dtype,n_genes", [(np.float32, 20000), (np.float32, 1479), (np.float64, 12000), (np.float64, 15000),
                                             (np.float32, 30000)])
def test_real_valued_cells_sort_paths(monkeypatch, dtype, n_genes):
    """Tie-free / real-valued cells: keys sorted in shared memory (rank_sort_kernel) when two key buffers fit, in the
    global scratch of rank_kernel otherwise (float64 x 15 000, float32 x 30 000) or when PP_NO_SMEM_SORT is set; ties,
    signed zeros, infinities, a count cell and a few-distinct cell in between."""
    rng = np.random.default_rng(n_genes)
    x = rng.lognormal(0.0, 2.0, (70, n_genes)).astype(dtype)
    x[1] = np.round(x[1], 1)                       # many ties, more than 1024 distinct values
    x[2, ::2] = -x[2, ::2]
    x[3, :50] = [0.0, -0.0] * 25
    x[4, 7], x[4, 8] = np.inf, -np.inf
    x[5] = rng.poisson(3.0, n_genes)               # count cell (bitmap path)
    x[6] = rng.integers(0, 300, n_genes) / 7.0     # 300 distinct real values (hash path)
    x[7] = x[7, 0]                                 # constant
    rows = np.arange(70, dtype=np.int32)[::-1].copy()
    cols = np.sort(rng.choice(n_genes, n_genes - n_genes // 10, replace=False)).astype(np.int32)
    exp = np.stack([np.unique(r, return_inverse=True)[1] for r in x[rows][:, cols]]).astype(np.uint16)
    got, mx = _ffi().dense_rank(x, rows, cols)
    assert np.array_equal(got, exp) and mx == exp.max()
    monkeypatch.setenv("PP_NO_SMEM_SORT", "1")
    got2, mx2 = _ffi().dense_rank(x, rows, cols)
    assert np.array_equal(got2, exp) and mx2 == mx


def test_large_pageable_input_is_staged():
    """A pageable host matrix of more than 64 MB goes through the threaded staging copy (pp_copy_h2d: chunks of 32 MB
    through two pinned buffers); a pinned one and a device tensor do not -- same markers from all three."""
    import torch
    rng = np.random.default_rng(8)
    x, lab = synth(rng, 1301, 13001, 3, frac_de=0.02, dtype=np.float32)  # 67.7 MB, odd sizes: a partial last chunk
    thr = np.ceil(np.bincount(lab, minlength=3) * 0.65).astype(np.int64)
    k_page, c_page, _ = run_kernel(x, lab, 3, thr)
    xt = torch.from_numpy(x)
    k_dev, c_dev, _ = run_kernel(xt.cuda(), lab, 3, thr)
    k_pin, c_pin, _ = run_kernel(xt.pin_memory().numpy(), lab, 3, thr)
    assert len(k_dev) > 100
    assert np.array_equal(k_page, k_dev) and np.array_equal(c_page, c_dev)
    assert np.array_equal(k_pin, k_dev) and np.array_equal(c_pin, c_dev)
    sub = np.sort(rng.choice(13001, 300, replace=False))
    rr, cc, kk = orc.sandbag_arrays(x[:, sub].astype(np.float64), lab, 3, 0.65)
    pos = np.full(13001, -1)
    pos[sub] = np.arange(300)
    rows, cols = (k_page // np.uint64(13001)).astype(np.int64), (k_page % np.uint64(13001)).astype(np.int64)
    m = (pos[rows] >= 0) & (pos[cols] >= 0)
    assert np.array_equal(pos[rows[m]], rr) and np.array_equal(pos[cols[m]], cc) and np.array_equal(c_page[m], kk)


@pytest.mark.parametrize("name", SANDBAG_CASE_NAMES)
def test_golden_cases(sandbag_cases, name):
    c = sandbag_cases
    x, cats, thr = c[name + "_x"], c[name + "_cats"], c[name + "_thr"]
    g = x.shape[1]
    keys, cls, tm = run_kernel(x, cats, len(thr), thr)
    assert np.array_equal(keys, c[name + "_rows"].astype(np.uint64) * np.uint64(g) + c[name + "_cols"].astype(np.uint64))
    assert np.array_equal(cls, c[name + "_cls"])
    # float32 copy of integer counts must give the same answer
    keys32, cls32, _ = run_kernel(x.astype(np.float32) if name != "c3_f65" else x, cats, len(thr), thr)
    assert np.array_equal(keys32, keys) and np.array_equal(cls32, cls)


def test_minref_api(sandbag_minref):
    """The reference's own KAT (tests/test_sandbag.py:26-113) through pairs.sandbag: lists AND key order."""
    from pypairs_b200 import pairs, utils
    m = sandbag_minref
    x = np.array(m["matrix_genes_by_samples"]).T
    for frac_s, ref in m["results"].items():
        got = pairs.sandbag(x, m["annotation"], m["gene_names"], m["sample_names"], fraction=float(frac_s))
        assert [[str(k), [[str(a), str(b)] for a, b in v]] for k, v in got.items()] == ref
    got = pairs.sandbag(x, m["annotation"], m["gene_names"], m["sample_names"])
    ref65 = {k: [tuple(p) for p in v] for k, v in m["results"]["0.65"]}
    assert utils.same_marker(got, ref65)


def test_api_inputs(sandbag_api):
    """DataFrame / ndarray, annotation as int / str / bool, filters, class order, emptied class."""
    from pandas import DataFrame
    from pypairs_b200 import pairs
    a = sandbag_api
    x = np.array(a["x"])
    sn, gn, lab = a["sample_names"], a["gene_names"], np.array(a["labels"])
    names3 = ["G1", "S", "G2M"]
    ann = {"int": {k: np.where(lab == i)[0].tolist() for i, k in enumerate(names3)},
           "str": {k: [sn[j] for j in np.where(lab == i)[0]] for i, k in enumerate(names3)},
           "bool": {k: (lab == i).tolist() for i, k in enumerate(names3)}}
    ann["int_rev"] = {k: ann["int"][k] for k in ["G2M", "G1", "S"]}
    ann["int_empty"] = dict(ann["int"], EMPTY=[58])
    for run in a["runs"]:
        kw = dict(run["kwargs"])
        annotation = ann[kw.pop("annotation")]
        if kw.pop("dataframe", False):
            got = pairs.sandbag(DataFrame(x, index=sn, columns=gn), annotation, **kw)
        else:
            got = pairs.sandbag(x, annotation, gn, sn, **kw)
        assert [[str(k), [[str(p), str(q)] for p, q in v]] for k, v in got.items()] == run["result"], run["tag"]


def test_anndata_duck_type(sandbag_api):
    """AnnData path: classes from obs['category'] are np.unique-sorted (utils.py:219-226)."""
    from pandas import DataFrame
    from pypairs_b200 import pairs
    a = sandbag_api
    x = np.array(a["x"])
    lab = np.array(a["labels"])
    keep = lab >= 0
    names = np.array(["G1", "S", "G2M"])

    class FakeAnnData:
        def __init__(self):
            self.X = x[keep]
            self.obs = DataFrame({"category": names[lab[keep]]}, index=np.array(a["sample_names"])[keep])
            self.obs_names = list(self.obs.index)
            self.var_names = a["gene_names"]

        def obs_keys(self):
            return list(self.obs.columns)

    got = pairs.sandbag(FakeAnnData(), fraction=0.65)
    ann_sorted = {k: np.where(names[lab[keep]] == k)[0].tolist() for k in sorted(names)}
    exp = pairs.sandbag(x[keep], ann_sorted, a["gene_names"], list(np.array(a["sample_names"])[keep]), fraction=0.65)
    assert list(got.items()) == list(exp.items())


@pytest.mark.parametrize("n,g,c,frac,dtype", [
    (500, 600, 3, 0.65, np.float64), (333, 450, 10, 0.65, np.float32), (1000, 300, 2, 0.6, np.float64),
    (64, 129, 1, 0.7, np.float64), (97, 257, 4, 0.5, np.float32), (2100, 200, 3, 0.65, np.float64)])
def test_random_vs_oracle(n, g, c, frac, dtype):
    rng = np.random.default_rng(n * 7 + g)
    x, lab = synth(rng, n, g, c, dtype=dtype)
    sizes = np.bincount(lab, minlength=c)
    thr = orc.calc_thresholds(sizes, frac)
    probes = np.sort(rng.integers(0, g, (200, 2)), axis=1)
    probes = probes[probes[:, 0] < probes[:, 1]]
    keys, cls, tm, ppos, pneg = run_kernel(x, lab, c, thr, probe_pairs=probes)
    rr, cc, kk = orc.sandbag_arrays(x, lab, c, frac)
    assert np.array_equal(keys, (rr * g + cc).astype(np.uint64))
    assert np.array_equal(cls, kk.astype(np.int32))
    epos, eneg = orc.pair_counts(x, lab, c, probes)
    assert np.array_equal(ppos, epos) and np.array_equal(pneg, eneg)  # per-class num_pos / num_neg bit-exact


def test_continuous_values_need_many_planes():
    """Tie-free doubles: every gene of a cell has its own rank -> 12 planes at 4000 genes."""
    rng = np.random.default_rng(5)
    n, g, c = 150, 4000, 3
    lab = rng.integers(0, c, n)
    x = rng.lognormal(0, 1, (n, g)) * np.where(rng.random((1, g)) < 0.05, 1.0 + lab[:, None], 1.0)
    sizes = np.bincount(lab, minlength=c)
    thr = orc.calc_thresholds(sizes, 0.65)
    keys, cls, tm = run_kernel(x, lab, c, thr)
    assert tm["n_planes"] == 12
    sub = 400  # oracle on the first 400 genes; restricting genes never changes a pair's decision
    rr, cc, kk = orc.sandbag_arrays(x[:, :sub], lab, c, 0.65)
    rows, cols = (keys // np.uint64(g)).astype(np.int64), (keys % np.uint64(g)).astype(np.int64)
    m = (rows < sub) & (cols < sub)
    assert np.array_equal(rows[m], rr) and np.array_equal(cols[m], cc) and np.array_equal(cls[m], kk.astype(np.int32))


def test_cell_order_and_shards_do_not_matter():
    rng = np.random.default_rng(11)
    x, lab = synth(rng, 400, 500, 3)
    thr = orc.calc_thresholds(np.bincount(lab, minlength=3), 0.65)
    keys, cls, _ = run_kernel(x, lab, 3, thr)
    perm = rng.permutation(400)
    keys2, cls2, _ = run_kernel(x[perm], lab[perm], 3, thr)
    assert np.array_equal(keys, keys2) and np.array_equal(cls, cls2)
    parts = [run_kernel(x, lab, 3, thr, shard=s, n_shards=3) for s in range(3)]
    allk = np.concatenate([p[0] for p in parts])
    allc = np.concatenate([p[1] for p in parts])
    o = np.argsort(allk, kind="stable")
    assert np.array_equal(allk[o], keys) and np.array_equal(allc[o], cls)
    assert sum(p[2]["n_units"] for p in parts) == 500 * 499 // 2


def test_errors():
    f = _ffi()
    x = np.ones((4, 5))
    with pytest.raises(f.PyPairsB200Error):
        f.sandbag(x, [0, 1, 9], [2, 1], np.arange(5), [1, 1])       # row out of range
    with pytest.raises(f.PyPairsB200Error):
        f.sandbag(x, [0, 1, 2], [2, 2], np.arange(5), [1, 1])       # sizes do not sum
    xn = x.copy()
    xn[1, 2] = np.nan
    with pytest.raises(f.PyPairsB200Error) as e:
        f.sandbag(xn, [0, 1, 2], [2, 1], np.arange(5), [1, 1])
    assert e.value.code == f.PP_ENAN
    keys, cls, _ = f.sandbag(x, [0, 1, 2], [2, 1], np.arange(1), [1, 1])  # single gene: no pairs
    assert len(keys) == 0


def test_device_tensor_input():
    import torch
    rng = np.random.default_rng(3)
    x, lab = synth(rng, 200, 300, 3, dtype=np.float32)
    thr = orc.calc_thresholds(np.bincount(lab, minlength=3), 0.65)
    keys, cls, _ = run_kernel(x, lab, 3, thr)
    keys2, cls2, tm = run_kernel(torch.from_numpy(x).cuda(), lab, 3, thr)
    assert np.array_equal(keys, keys2) and np.array_equal(cls, cls2)
    assert tm["h2d_ms"] < 1.0


def test_full_size_properties():
    """cfg3 shape (20k genes x 10k cells, 3 classes): size-independent checks + oracle on a gene subset."""
    import torch
    g, n, c = 20000, 10000, 3
    gen = torch.Generator(device="cuda").manual_seed(20263)
    lab = torch.randint(0, c, (n,), generator=gen, device="cuda")
    base = torch.randn(g, generator=gen, device="cuda") * 1.5
    de = torch.rand(g, generator=gen, device="cuda") < 0.02
    de_class = torch.randint(0, c, (g,), generator=gen, device="cuda")
    sign = torch.where(torch.rand(g, generator=gen, device="cuda") < 0.5, -1.0, 1.0)
    eff = torch.where(de, torch.pow(2.0, 2.0 * sign), torch.ones_like(sign))
    lam = 3.0 * torch.exp(base)[None, :] * torch.where(lab[:, None] == de_class[None, :], eff[None, :], 1.0)
    x = torch.poisson(lam, generator=gen).to(torch.float32)
    x[:, 0] += 1.0
    lab_h = lab.cpu().numpy().astype(np.int64)
    thr = orc.calc_thresholds(np.bincount(lab_h, minlength=c), 0.65)
    keys, cls, tm = run_kernel(x, lab_h, c, thr)
    assert np.all(keys[1:] > keys[:-1])                       # sorted, unique == np.where order
    rows, cols = (keys // np.uint64(g)).astype(np.int64), (keys % np.uint64(g)).astype(np.int64)
    assert np.all(rows != cols) and cls.min() >= 0 and cls.max() < c
    assert tm["n_units"] == g * (g - 1) // 2
    sub = np.sort(np.random.default_rng(0).choice(g, 250, replace=False))
    xs = x[:, torch.from_numpy(sub).cuda()].cpu().numpy().astype(np.float64)
    rr, cc, kk = orc.sandbag_arrays(xs, lab_h, c, 0.65)
    pos = {int(v): i for i, v in enumerate(sub)}
    m = np.isin(rows, sub) & np.isin(cols, sub)
    got = sorted((pos[int(r)], pos[int(q)], int(k)) for r, q, k in zip(rows[m], cols[m], cls[m]))
    assert got == sorted(zip(rr.tolist(), cc.tolist(), kk.tolist()))
    print("cfg3 full size: {} markers, {} planes, kernel {:.1f} ms, prep {:.1f} ms".format(
        len(keys), tm["n_planes"], tm["main_ms"], tm["prep_ms"]))


def test_wide_counters_and_many_classes():
    """Classes above 65 535 cells (16-bit counters folded into 32-bit ones) and more than 254 classes."""
    rng = np.random.default_rng(21)
    # (a) one class of 70 001 cells next to a small one
    n, g = 70001 + 4000, 130
    lab = np.concatenate([np.zeros(70001, dtype=np.int64), np.ones(4000, dtype=np.int64)])
    lam = 2.0 * np.exp(rng.normal(0, 1.0, g))[None, :] * np.where((lab[:, None] == 1) & (np.arange(g) % 5 == 0)[None, :], 5.0, 1.0)
    x = rng.poisson(lam).astype(np.float32)
    perm = rng.permutation(n)
    x, lab = x[perm], lab[perm]
    thr = orc.calc_thresholds(np.bincount(lab, minlength=2), 0.65)
    probes = np.array([[0, 5], [3, 100], [7, 129], [10, 11]])
    keys, cls, tm, ppos, pneg = run_kernel(x, lab, 2, thr, probe_pairs=probes)
    rr, cc, kk = orc.sandbag_arrays(x, lab, 2, 0.65)
    assert np.array_equal(keys, (rr * g + cc).astype(np.uint64)) and np.array_equal(cls, kk.astype(np.int32))
    assert len(keys) > 0
    epos, eneg = orc.pair_counts(x, lab, 2, probes)
    assert np.array_equal(ppos, epos) and np.array_equal(pneg, eneg) and epos.max() > 65535
    # (b) 300 classes
    n, g, c = 3000, 90, 300
    lab = np.arange(n) % c
    x = rng.poisson(3.0, (n, g)).astype(np.float64) + (lab[:, None] == (np.arange(g) * 7 % c)[None, :]) * 20.0
    for frac in (0.65, 0.3):
        thr = orc.calc_thresholds(np.bincount(lab, minlength=c), frac)
        keys, cls, tm = run_kernel(x, lab, c, thr)
        rr, cc, kk = orc.sandbag_arrays(x, lab, c, frac)
        assert np.array_equal(keys, (rr * g + cc).astype(np.uint64)) and np.array_equal(cls, kk.astype(np.int32))


def test_atlas_shape_properties():
    """configs[3] shape (30k genes x 50k cells, 10 classes): invariants at full size + oracle on a gene subset."""
    import torch
    g, n, c = 30000, 50000, 10
    gen = torch.Generator(device="cuda").manual_seed(20264)
    lab = torch.randint(0, c, (n,), generator=gen, device="cuda")
    base = torch.randn(g, generator=gen, device="cuda") * 1.5
    de = torch.rand(g, generator=gen, device="cuda") < 0.02
    de_class = torch.randint(0, c, (g,), generator=gen, device="cuda")
    sign = torch.where(torch.rand(g, generator=gen, device="cuda") < 0.5, -1.0, 1.0)
    eff = torch.where(de, torch.pow(2.0, 2.0 * sign), torch.ones_like(sign))
    x = torch.empty((n, g), dtype=torch.float32, device="cuda")
    for i in range(0, n, 5000):
        lam = 3.0 * torch.exp(base)[None, :] * torch.where(lab[i:i + 5000, None] == de_class[None, :], eff[None, :], 1.0)
        x[i:i + 5000] = torch.poisson(lam, generator=gen)
    x[:, 0] += 1.0
    lab_h = lab.cpu().numpy().astype(np.int64)
    thr = orc.calc_thresholds(np.bincount(lab_h, minlength=c), 0.65)
    keys, cls, tm = run_kernel(x, lab_h, c, thr)
    assert np.all(keys[1:] > keys[:-1]) and tm["n_units"] == g * (g - 1) // 2
    rows, cols = (keys // np.uint64(g)).astype(np.int64), (keys % np.uint64(g)).astype(np.int64)
    sub = np.sort(np.random.default_rng(1).choice(g, 120, replace=False))
    xs = x[:, torch.from_numpy(sub).cuda()].cpu().numpy().astype(np.float64)
    rr, cc, kk = orc.sandbag_arrays(xs, lab_h, c, 0.65)
    pos = {int(v): i for i, v in enumerate(sub)}
    m = np.isin(rows, sub) & np.isin(cols, sub)
    got = sorted((pos[int(r)], pos[int(q)], int(k)) for r, q, k in zip(rows[m], cols[m], cls[m]))
    assert got == sorted(zip(rr.tolist(), cc.tolist(), kk.tolist()))
    print("cfg4 full size: {} markers, {} planes, kernel {:.1f} ms, prep {:.1f} ms -> {:.3g} pair*cell/s".format(
        len(keys), tm["n_planes"], tm["main_ms"], tm["prep_ms"], g * (g - 1) / 2 * n / (tm["main_ms"] * 1e-3)))


def test_marker_table_roundtrip_and_cyclone():
    """as_table=True gives the same markers without tuples; pairs.cyclone takes the table directly."""
    from pypairs_b200 import pairs, utils
    rng = np.random.default_rng(51)
    x, lab = synth(rng, 300, 400, 3, frac_de=0.3, dtype=np.float32)
    gn = ["g{}".format(i) for i in range(400)]
    sn = ["c{}".format(i) for i in range(300)]
    ann = {k: np.where(lab == i)[0].tolist() for i, k in enumerate(["G1", "S", "G2M"])}
    d = pairs.sandbag(x, ann, gn, sn, fraction=0.6)
    t = pairs.sandbag(x, ann, gn, sn, fraction=0.6, as_table=True)
    assert isinstance(t, utils.MarkerTable) and list(t.to_dict().items()) == list(d.items())
    y = rng.poisson(3.0, (50, 400)).astype(np.float32)
    yn = ["u{}".format(i) for i in range(50)]
    a = pairs.cyclone(y, d, gn, yn, iterations=60, min_iter=10, min_pairs=3)
    b = pairs.cyclone(y, t, gn, yn, iterations=60, min_iter=10, min_pairs=3)
    assert a.equals(b) and list(a.columns)[:len(d)] == list(d.keys())


def _check_against_oracle(x, lab, thr):
    g = x.shape[1]
    keys, cls, tm = run_kernel(x, lab, len(thr), thr)
    res = orc.check_pairs(x, lab, thr)
    rr, cc = np.where(res != -1)
    assert np.array_equal(keys, (rr * g + cc).astype(np.uint64))
    assert np.array_equal(cls, res[rr, cc].astype(np.int32))
    return len(keys), tm


@pytest.mark.parametrize("thr", [[0, 0, 0], [-3, 5, 0], [1, 1, 1], [10 ** 6, 2, 2], [40, 0, 10 ** 6]])
def test_degenerate_thresholds(thr):
    """Thresholds the reference never clamps (fraction <= 0 gives thr <= 0, sandbag.py:212-217): zero / negative
    thresholds make every class 'up' and 'down' at once, thresholds above the class size can never be met."""
    rng = np.random.default_rng(23)
    x, lab = synth(rng, 130, 310, 3, frac_de=0.3)
    _check_against_oracle(x, lab, np.array(thr, dtype=np.int64))


def test_second_sweep_runs_many_rounds():
    """Strong two-class signal: most pairs of a 128 x 64 tile are candidates, so the second sweep needs several
    rounds of its 512-entry candidate list, and the hit buffer is regrown."""
    rng = np.random.default_rng(29)
    n, g = 260, 700
    lab = (np.arange(n) % 2).astype(np.int64)
    level = rng.permutation(g).astype(np.float64)
    x = np.where(lab[:, None] == 0, level[None, :], level[::-1][None, :]) + rng.integers(0, 2, (n, g))
    n_hits, tm = _check_against_oracle(x, lab, orc.calc_thresholds(np.bincount(lab), 0.9))
    assert n_hits > 100_000


def test_values_that_stress_the_rank_transform():
    """Negative values, signed zeros, huge magnitudes, denormals and all-equal cells."""
    rng = np.random.default_rng(31)
    n, g = 90, 260
    lab = rng.integers(0, 3, n)
    lab[:3] = np.arange(3)
    x = rng.normal(0, 3, (n, g)).round(0)           # many ties, half of them negative
    x[x == 0] = np.where(rng.random((x == 0).sum()) < 0.5, -0.0, 0.0)
    x[5] = 7.0                                       # a cell where all genes tie
    x[6] = rng.normal(0, 1, g) * 1e300
    x[7] = rng.normal(0, 1, g) * 5e-324 * 1000       # denormals
    x[:, 17] += np.where(lab == 1, 50.0, 0.0)
    x[:, 40] -= np.where(lab == 2, 50.0, 0.0)
    n_hits, _ = _check_against_oracle(x, lab, orc.calc_thresholds(np.bincount(lab, minlength=3), 0.6))
    assert n_hits > 0
    xc = np.full((n, g), 3.0)                        # a constant matrix: no pair differs anywhere
    assert _check_against_oracle(xc, lab, np.array([1, 1, 1]))[0] == 0


def test_single_cell_classes_and_tiny_inputs():
    rng = np.random.default_rng(37)
    x = rng.poisson(2.0, (3, 70)).astype(np.float64)  # one cell per class
    lab = np.arange(3)
    _check_against_oracle(x, lab, np.array([1, 1, 1]))
    x2 = rng.poisson(2.0, (5, 2)).astype(np.float64)  # two genes: one pair
    _check_against_oracle(x2, np.array([0, 0, 1, 1, 1]), np.array([2, 2]))


@pytest.mark.parametrize("c,frac", [(2, 0.3), (2, 0.5), (3, 0.34), (5, 0.2)])
def test_mixed_rank_widths_and_gene_order(c, frac):
    """Genes of very different rank widths (constant, binary, low counts, high counts, tie-free) interleaved, so the
    width-sorted positions invert the gene order of many pairs, and thresholds low enough that a class can be 'up' and
    'down' at once: with two classes both branches of sandbag.py:192-197 then hold and the if/elif order decides,
    which depends on which gene the reference calls g1."""
    rng = np.random.default_rng(100 + c)
    n, g = 240, 520
    lab = rng.integers(0, c, n)
    lab[:c] = np.arange(c)
    kind = rng.integers(0, 5, g)
    x = np.empty((n, g))
    shift = rng.normal(0, 1.0, (c, g))
    for j in range(g):
        s = shift[lab, j]
        if kind[j] == 0:
            x[:, j] = 3.0                                         # constant
        elif kind[j] == 1:
            x[:, j] = (rng.random(n) < 0.3 + 0.2 * np.tanh(s)).astype(float)   # binary
        elif kind[j] == 2:
            x[:, j] = rng.poisson(np.exp(0.5 * s))                # low counts
        elif kind[j] == 3:
            x[:, j] = rng.poisson(200.0 * np.exp(0.5 * s))        # high counts
        else:
            x[:, j] = rng.lognormal(0.3 * s, 1.0)                 # continuous, tie-free
    thr = orc.calc_thresholds(np.bincount(lab, minlength=c), frac)
    n_hits, tm = _check_against_oracle(x, lab, thr)
    assert n_hits > 0 and tm["avg_planes"] < tm["n_planes"]
    probes = np.sort(rng.integers(0, g, (300, 2)), axis=1)
    probes = probes[probes[:, 0] < probes[:, 1]]
    _, _, _, ppos, pneg = run_kernel(x, lab, c, thr, probe_pairs=probes)
    epos, eneg = orc.pair_counts(x, lab, c, probes)
    assert np.array_equal(ppos, epos) and np.array_equal(pneg, eneg)
