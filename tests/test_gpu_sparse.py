"""GPU parity for CSR input: the reference densifies sparse matrices on the host (pypairs/utils.py:278-301);
here they are densified chunk by chunk on the device.  Results must equal the dense-input results."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def sparse_counts(rng, n, g, density, dtype):
    m = sp.random(n, g, density=density, random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="csr",
                  data_rvs=lambda k: rng.poisson(3.0, k) + 1.0).astype(dtype)
    return m


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
def test_sandbag_sparse_equals_dense(dtype):
    from pypairs_b200 import pairs
    rng = np.random.default_rng(31)
    n, g = 400, 700
    m = sparse_counts(rng, n, g, 0.15, dtype).tolil()
    lab = rng.integers(0, 2, n)
    for k in range(2):  # class-specific genes so that markers exist
        m[np.where(lab == k)[0][:, None], np.arange(k * 20, k * 20 + 20)[None, :]] = 9
    m = m.tocsr()
    m[:, 650] = 0          # gene with only explicit zeros: must be dropped like an all-zero dense column
    m.data[::7] *= 1       # keep explicit structure
    m = m.tocsr()
    assert (m[:, 650].toarray() == 0).all()
    dense = m.toarray()
    gn = ["g{}".format(i) for i in range(g)]
    sn = ["c{}".format(i) for i in range(n)]
    ann = {name: np.where(lab == k)[0].tolist() for k, name in enumerate(["A", "B"])}
    exp = pairs.sandbag(dense, ann, gn, sn, fraction=0.6)
    assert sum(len(v) for v in exp.values()) > 50
    for form in (m, m.tocsc(), m.tocoo()):
        got = pairs.sandbag(form, ann, gn, sn, fraction=0.6)
        assert list(got.items()) == list(exp.items())
    # unsorted column indices inside the rows
    u = m.copy()
    for r in range(n):
        a, b = u.indptr[r], u.indptr[r + 1]
        p = rng.permutation(b - a)
        u.indices[a:b] = u.indices[a:b][p]
        u.data[a:b] = u.data[a:b][p]
    u.has_sorted_indices = False
    got = pairs.sandbag(u, ann, gn, sn, fraction=0.6)
    assert list(got.items()) == list(exp.items())


def test_sandbag_sparse_many_chunks_vs_oracle():
    """20 000 genes as float64: the densify buffer holds ~1 600 cells, so 2 500 cells take two runs."""
    from pypairs_b200 import _ffi
    rng = np.random.default_rng(32)
    n, g, c = 2500, 20000, 3
    m = sparse_counts(rng, n, g, 0.03, np.float64)
    lab = rng.integers(0, c, n)
    sizes = np.bincount(lab, minlength=c)
    thr = orc.calc_thresholds(sizes, 0.3)
    order = np.argsort(lab, kind="stable")
    keys, cls, tm = _ffi.sandbag(m, order, sizes, np.arange(g), thr)
    keys_d, cls_d, _ = _ffi.sandbag(m.toarray(), order, sizes, np.arange(g), thr)
    assert np.array_equal(keys, keys_d) and np.array_equal(cls, cls_d)
    assert tm["h2d_bytes"] < 0.1 * n * g * 8
    sub = 300
    rr, cc, kk = orc.sandbag_arrays(m[:, :sub].toarray(), lab, c, 0.3)
    rows, cols = (keys // np.uint64(g)).astype(np.int64), (keys % np.uint64(g)).astype(np.int64)
    sel = (rows < sub) & (cols < sub)
    assert np.array_equal(rows[sel], rr) and np.array_equal(cols[sel], cc) and np.array_equal(cls[sel], kk.astype(np.int32))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cyclone_sparse_equals_dense(dtype):
    import torch
    from pypairs_b200 import _ffi, datasets, pairs
    from pypairs_b200.cyclone import filter_marker_pairs
    rng = np.random.default_rng(33)
    markers = datasets.default_cc_marker()
    genes = sorted({a for v in markers.values() for p in v for a in p})
    names = genes + ["X{}".format(i) for i in range(6000 - len(genes))]
    rng.shuffle(names)
    n = 148 * 64 + 321  # two chunks of cells
    m = sparse_counts(rng, n, 6000, 0.3, dtype)
    dense = m.toarray()
    pr, used = filter_marker_pairs(markers, names)
    ks = list(pr.keys())
    u = [np.where(used[k])[0] for k in ks]
    p = [pr[k] for k in ks]
    s_dense, _ = _ffi.cyclone(dense, u, p, 25, 10, 5)
    s_sparse, tm, oh, ot = _ffi.cyclone(m, u, p, 25, 10, 5, want_observed=True)
    assert np.array_equal(s_sparse, s_dense, equal_nan=True)
    assert 0 < tm["h2d_bytes"] < 0.7 * dense.nbytes
    s_part, _ = _ffi.cyclone(m, u, p, 25, 10, 5, row_begin=4000, n_cells=1234)       # a cell shard
    assert np.array_equal(s_part, s_dense[:, 4000:5234], equal_nan=True)
    t = torch.sparse_csr_tensor(torch.from_numpy(m.indptr.astype(np.int64)), torch.from_numpy(m.indices.astype(np.int64)),
                                torch.from_numpy(m.data), size=m.shape).cuda()
    s_dev, tm_d = _ffi.cyclone(t, u, p, 25, 10, 5, row_begin=100, n_cells=3000)      # device-resident CSR
    assert np.array_equal(s_dev, s_dense[:, 100:3100], equal_nan=True) and tm_d["h2d_bytes"] == 0
    rows = np.sort(rng.choice(n, 16, replace=False))
    for i, k in enumerate(ks):
        ref, eh, et = orc.phase_scores(dense[rows], 25, 10, 5, pr[k], used[k], return_observed=True)
        assert np.array_equal(s_sparse[i, rows], ref) and np.array_equal(oh[i, rows], eh) and np.array_equal(ot[i, rows], et)
    # public API, CSC input
    sn = ["c{}".format(i) for i in range(300)]
    df_s = pairs.cyclone(m[:300].tocsc(), markers, names, sn, iterations=25, min_iter=10, min_pairs=5)
    df_d = pairs.cyclone(dense[:300], markers, names, sn, iterations=25, min_iter=10, min_pairs=5)
    assert df_s.equals(df_d)
