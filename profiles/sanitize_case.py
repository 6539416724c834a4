"""Tiny end-to-end calls of both paths for compute-sanitizer (memcheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scipy.sparse as sp  # noqa: E402

from pypairs_b200 import _ffi, datasets  # noqa: E402
from pypairs_b200.cyclone import filter_marker_pairs  # noqa: E402

rng = np.random.default_rng(0)
# sandbag: 3 classes, 333 genes, 257 cells (ragged words / blocks), counts + one continuous row, f32 / f64 / CSR
n, g, c = 257, 333, 3
lab = rng.integers(0, c, n)
x = rng.poisson(3.0 * np.exp(rng.normal(0, 1, g))[None, :] * np.where(lab[:, None] == (np.arange(g) % c)[None, :], 4, 1)).astype(np.float64)
x[7] = rng.normal(size=g)
sizes = np.bincount(lab, minlength=c)
thr = np.ceil(sizes * 0.6).astype(np.int64)
order = np.argsort(lab, kind="stable")
probes = np.array([[0, 5], [3, 100], [7, 332]])
for m in (x, x.astype(np.float32), sp.csr_matrix(x)):
    out = _ffi.sandbag(m, order, sizes, np.arange(g), thr, probe_pairs=probes, drop_unexpressed=True)
    print("sandbag", type(m).__name__, len(out[0]), "markers")
# cyclone: default markers on 200 cells x 1700 genes, 20 iterations, dense / CSR / general path (2600 genes)
markers = datasets.default_cc_marker()
genes = sorted({a for v in markers.values() for p in v for a in p})
names = genes + ["X{}".format(i) for i in range(1700 - len(genes))]
y = rng.poisson(2.0, (200, 1700)).astype(np.float32)
pr, used = filter_marker_pairs(markers, names)
ks = list(pr.keys())
for m in (y, sp.csr_matrix(y)):
    sc, tm = _ffi.cyclone(m, [np.where(used[k])[0] for k in ks], [pr[k] for k in ks], 20, 10, 5)
    print("cyclone", type(m).__name__, float(np.nansum(sc)))
big = ["g{}".format(i) for i in range(2600)]
mk = {"a": [(big[i], big[(i * 7 + 3) % 2600]) for i in range(2600)], "b": [(big[(i * 5) % 2600], big[(i * 11 + 1) % 2600]) for i in range(1500)]}
z = rng.poisson(3.0, (70, 2600)).astype(np.float64)
pr, used = filter_marker_pairs(mk, big)
sc, tm = _ffi.cyclone(z, [np.where(used[k])[0] for k in mk], [pr[k] for k in mk], 12, 5, 3)
print("cyclone general path", float(np.nansum(sc)))
print(_ffi.perm_table("philox", 5, 1, 100, 30).sum())
# round 2 kernels: streaming rank (rows >= 8 KB, odd row length, dropped columns), shared-memory sort of real-valued cells,
# look-up table + upper bitmap words, pinned input read by the zero-copy gather kernel
import torch  # noqa: E402

n2, g2 = 96, 2601
lab2 = rng.integers(0, 3, n2)
w = rng.poisson(rng.gamma(0.5, 6.0, g2)[None, :] + 0.05, (n2, g2)).astype(np.float32)
w[3] = rng.lognormal(0, 1, g2)
w[4] = (np.arange(g2) * 7) % 60000
w[:, 11] = 0
sizes2 = np.bincount(lab2, minlength=3)
out = _ffi.sandbag(w, np.argsort(lab2, kind="stable"), sizes2, np.arange(g2), np.ceil(sizes2 * 0.6).astype(np.int64), drop_unexpressed=True)
print("sandbag streaming rank", len(out[0]), "markers")
rk, mx = _ffi.dense_rank(w.astype(np.float64), np.arange(n2, dtype=np.int32), np.arange(5, g2 - 3, dtype=np.int32))
print("dense_rank f64", int(mx))
yp = torch.from_numpy(rng.poisson(2.0, (700, 6400)).astype(np.float32)).pin_memory()
names2 = genes + ["X{}".format(i) for i in range(6400 - len(genes))]
pr2, used2 = filter_marker_pairs(markers, names2)
for feed in ("zerocopy", "both", "gather"):
    os.environ["PP_FEED"] = feed
    sc, tm = _ffi.cyclone(yp.numpy(), [np.where(used2[k])[0] for k in ks], [pr2[k] for k in ks], 10, 5, 5)
    print("cyclone pinned", feed, float(np.nansum(sc)))
