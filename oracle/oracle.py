"""Python face of the CPU oracle (ctypes over oracle/libpp_oracle.so + small numpy restatements).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(pypairs_b200/) never imports this module.

Parity status: pinned -- see the header of pp_oracle.c and tests/test_oracle.py.
"""
import ctypes
import math
import os
import subprocess
from collections import defaultdict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile libpp_oracle.so with the system gcc (a few hundred ms)."""
    so = os.path.join(_HERE, "libpp_oracle.so")
    src = os.path.join(_HERE, "pp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call(
            [cc, "-O2", "-fPIC", "-fopenmp", "-fvisibility=hidden", "-std=c11", "-shared", "-o", so, src, "-lm"]
        )
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        c = ctypes
        vp, i64, i32, u32, u64, dbl = c.c_void_p, c.c_int64, c.c_int, c.c_uint32, c.c_uint64, c.c_double
        _LIB.orc_max_threads.restype = i32
        _LIB.orc_check_pairs.argtypes = [vp, vp, vp, i64, i64, i64, vp, i32]
        _LIB.orc_pair_counts.argtypes = [vp, vp, i64, i64, i64, vp, i64, vp, vp]
        _LIB.orc_mt_outputs.argtypes = [u32, i64, vp]
        _LIB.orc_perm_table.argtypes = [u32, i64, i64, vp]
        _LIB.orc_phase_scores.argtypes = [vp, i64, i64, vp, vp, i64, i64, i64, i64, i32, u32, vp, vp, vp, i32]
        _LIB.orc_philox.argtypes = [u32, u32, u32, u32, u32, u32, vp]
        _LIB.orc_perm_table_philox.argtypes = [u64, u32, i64, i64, vp]
        _LIB.orc_phase_scores_table.argtypes = [vp, i64, i64, vp, vp, i64, vp, i64, i64, i64, vp, i32]
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def max_threads():
    return int(lib().orc_max_threads())


# ----------------------------------------------------------------------------- sandbag
def calc_thresholds(class_sizes, fraction):
    """ceil(n_k * fraction) as a Python float product (pypairs/tools/sandbag.py:212-217)."""
    return np.array([math.ceil(int(n) * fraction) for n in class_sizes], dtype=np.int64)


def check_pairs(x, cats, thr, n_threads=0):
    """Result matrix int64[G,G] of pypairs/tools/sandbag.py:160-199 (x is [cells, genes])."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    cats = np.ascontiguousarray(cats, dtype=np.int64)
    thr = np.ascontiguousarray(thr, dtype=np.int64)
    n, g = x.shape
    out = np.empty((g, g), dtype=np.int64)
    lib().orc_check_pairs(_p(x), _p(cats), _p(thr), n, g, len(thr), _p(out), n_threads)
    return out


def pair_counts(x, cats, n_classes, pairs):
    x = np.ascontiguousarray(x, dtype=np.float64)
    cats = np.ascontiguousarray(cats, dtype=np.int64)
    pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
    n, g = x.shape
    pos = np.empty((len(pairs), n_classes), dtype=np.int64)
    neg = np.empty_like(pos)
    lib().orc_pair_counts(_p(x), _p(cats), n, g, n_classes, _p(pairs), len(pairs), _p(pos), _p(neg))
    return pos, neg


def result_to_markers(result, gene_names, category_names):
    """Row-major scan of the full G x G matrix into the marker dict (sandbag.py:129-140,157)."""
    rows, cols = np.where(result != -1)
    out = defaultdict(list)
    for r, c in zip(rows, cols):
        out[category_names[result[r, c]]].append((gene_names[r], gene_names[c]))
    return dict(out)


def sandbag_arrays(x, cats, n_classes, fraction, n_threads=0):
    """(rows, cols, cls) of all markers in the reference's scan order, from filtered arrays."""
    sizes = np.bincount(np.asarray(cats, dtype=np.int64), minlength=n_classes)
    thr = calc_thresholds(sizes, fraction)
    res = check_pairs(x, cats, thr, n_threads)
    rows, cols = np.where(res != -1)
    return rows.astype(np.int64), cols.astype(np.int64), res[rows, cols].astype(np.int64)


# ----------------------------------------------------------------------------- cyclone
def filter_marker_pairs(marker_pairs, gene_names):
    """Restatement of pypairs/tools/cyclone.py:273-311 (dict order, duplicates kept, last name wins)."""
    name_to_idx = {g: i for i, g in enumerate(gene_names)}
    out_pairs, out_used = {}, {}
    for cat, pairs in marker_pairs.items():
        used = np.zeros(len(gene_names), dtype=bool)
        kept = []
        for a, b in pairs:
            if a in name_to_idx and b in name_to_idx:
                ia, ib = name_to_idx[a], name_to_idx[b]
                used[ia] = True
                used[ib] = True
                kept.append((ia, ib))
        new_idx = {u: i for i, u in enumerate(np.where(used)[0].tolist())}
        out_pairs[cat] = np.array([[new_idx[a], new_idx[b]] for a, b in kept], dtype=np.int64).reshape(-1, 2)
        out_used[cat] = used
    return out_pairs, out_used


def phase_scores(matrix, iterations, min_iter, min_pairs, pairs, used, fix_seed=True, seed=2803,
                 n_threads=0, return_observed=False):
    """get_phase_scores (pypairs/tools/cyclone.py:223-257): per-cell reseeded MT19937 shuffles."""
    m = np.ascontiguousarray(matrix, dtype=np.float64)
    used8 = np.ascontiguousarray(used, dtype=np.uint8)
    pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
    n, g = m.shape
    scores = np.empty(n, dtype=np.float64)
    oh = np.empty(n, dtype=np.int64)
    ot = np.empty(n, dtype=np.int64)
    lib().orc_phase_scores(_p(m), n, g, _p(used8), _p(pairs), len(pairs), iterations, min_iter, min_pairs,
                           1 if fix_seed else 0, seed, _p(scores), _p(oh), _p(ot), n_threads)
    if return_observed:
        return scores, oh, ot
    return scores


def phase_scores_table(matrix, table, min_iter, min_pairs, pairs, used, n_threads=0):
    m = np.ascontiguousarray(matrix, dtype=np.float64)
    used8 = np.ascontiguousarray(used, dtype=np.uint8)
    pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
    table = np.ascontiguousarray(table, dtype=np.int32)
    n, g = m.shape
    scores = np.empty(n, dtype=np.float64)
    lib().orc_phase_scores_table(_p(m), n, g, _p(used8), _p(pairs), len(pairs), _p(table), table.shape[0],
                                 min_iter, min_pairs, _p(scores), n_threads)
    return scores


def mt_outputs(seed, n):
    out = np.empty(n, dtype=np.uint32)
    lib().orc_mt_outputs(seed, n, _p(out))
    return out


def perm_table(seed, n, iters):
    """int32[iters, n]: cumulative np.random.shuffle states of arange(n) after seeding once."""
    out = np.empty((iters, n), dtype=np.int32)
    lib().orc_perm_table(seed, n, iters, _p(out))
    return out


def philox(c, k):
    out = np.empty(4, dtype=np.uint32)
    lib().orc_philox(c[0], c[1], c[2], c[3], k[0], k[1], _p(out))
    return out


def perm_table_philox(seed, class_index, n, iters):
    out = np.empty((iters, n), dtype=np.int32)
    lib().orc_perm_table_philox(seed, class_index, n, iters, _p(out))
    return out


def cyclone_epilogue(scores, sample_names):
    """DataFrame epilogue of pypairs/tools/cyclone.py:169-181 restated with positional access
    (the reference's `sum_row[i]` / `scores_cc[i]` are positional lookups that pandas >= 3 rejects)."""
    from pandas import DataFrame

    keys = list(scores.keys())
    df = DataFrame(scores, columns=keys)
    df.index = sample_names
    vals = df[keys].to_numpy(dtype=np.float64)
    sum_row = np.nansum(vals, axis=1)  # DataFrame.sum skips NaN
    max_class = []
    for i in range(vals.shape[0]):
        row = vals[i]
        if np.all(np.isnan(row)):
            mc = np.nan
        else:
            mc = keys[int(np.nanargmax(row))]
        max_class.append(mc if sum_row[i] > 0 else "N/A")
    df["max_class"] = max_class
    if len(keys) == 3 and all(e in keys for e in ["G1", "S", "G2M"]):
        sub = df.loc[:, ["G1", "G2M"]].to_numpy(dtype=np.float64)
        pred = []
        for i in range(sub.shape[0]):
            row = sub[i]
            mx = np.nan if np.all(np.isnan(row)) else np.nanmax(row)
            if mx < 0.5:
                pred.append("S")
            else:
                pred.append(np.nan if np.all(np.isnan(row)) else ["G1", "G2M"][int(np.nanargmax(row))])
        df["cc_prediction"] = pred
    return df


def default_cc_marker():
    """The reference's bundled leng15 marker set (pypairs/datasets/__init__.py:100-120), read from the compact
    fixture tests/golden/default_cc_marker.json that make_golden.py derived from the reference's JSON."""
    import json
    d = json.load(open(os.path.join(os.path.dirname(_HERE), "tests", "golden", "default_cc_marker.json")))
    genes = d["genes"]
    return {k: [[genes[a], genes[b]] for a, b in zip(*v)] for k, v in d["classes"].items()}
