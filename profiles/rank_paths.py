"""Rank-kernel paths on the cfg3 shape (20k genes x 10k cells): raw counts (bitmap path), normalised counts
(few-distinct hash path), tie-free values (radix sort path).   python profiles/rank_paths.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from pypairs_b200 import _ffi  # noqa: E402

x, lab = bench.sandbag_matrix_device()
n, g = x.shape
rows = np.arange(n, dtype=np.int32)
cols = np.arange(g, dtype=np.int32)
size = x.sum(dim=1, keepdim=True)
cases = {"raw counts": x, "log1p(counts / size factor)": torch.log1p(x / size * 1e4),
         "tie-free": x + torch.rand_like(x)}
sizes = np.bincount(lab, minlength=3)
thr = np.ceil(sizes * 0.65).astype(np.int64)
order = np.argsort(lab, kind="stable")
for name, m in cases.items():
    m = m.contiguous()
    for _ in range(2):
        keys, cls, tm = _ffi.sandbag(m, order, sizes, cols, thr)
    r, mx = _ffi.dense_rank(m[:64].contiguous(), rows[:64], cols)
    ref = np.unique(m[7].cpu().numpy(), return_inverse=True)[1]
    assert np.array_equal(r[7], ref.astype(np.uint16)), name
    print("{:30s} max rank {:6d}: rank + bit-slice {:.2f} ms, pair kernel {:.1f} ms ({:.2f} planes per pair), {} markers".format(
        name, tm["n_planes"] and mx, tm["prep_ms"], tm["main_ms"], tm["avg_planes"], len(keys)))
