"""Side measurements that are not bench lines: the general (global-tile) cyclone path, tie-free input,
small-problem latency (cfg1 shapes) and cfg5 (1M cells) on one GPU.

    python profiles/paths_probe.py [general] [small] [million]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from pypairs_b200 import _ffi, pairs, settings  # noqa: E402

settings.verbosity = 0


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best, out = 1e30, None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, out


def general():
    rng = np.random.default_rng(5)
    res = []
    markers, names, keys, used_idx, pair_idx, _, _ = bench.cyclone_setup()
    cells = 148 * 64 * 2
    gen = torch.Generator(device="cuda").manual_seed(1)
    xp = torch.poisson(torch.full((cells, bench.CYC_GENES), 2.0, device="cuda"), generator=gen)
    xl = torch.exp(torch.randn((cells, bench.CYC_GENES), device="cuda", generator=gen))
    n_pairs = sum(len(p) for p in pair_idx)
    for tag, x in (("default markers, poisson", xp), ("default markers, lognormal (tie-free)", xl)):
        ms, (sc, tm) = timed(lambda: _ffi.cyclone(x, used_idx, pair_idx, 1000, 100, 50))
        res.append({"case": tag, "cells": cells, "genes_used": int(len(np.unique(np.concatenate(used_idx)))),
                    "pairs": n_pairs, "score_ms": tm["main_ms"], "prep_ms": tm["prep_ms"],
                    "pair_evals_per_s": cells * 1001.0 * n_pairs / (tm["main_ms"] * 1e-3)})
    for n_used, n_p in ((1600, 10000), (1700, 10000), (3000, 10000), (6000, 30000), (20000, 100000)):
        u = np.sort(rng.choice(bench.CYC_GENES, n_used, replace=False))
        p = rng.integers(0, n_used, (n_p, 2))
        p = p[p[:, 0] != p[:, 1]]
        c = cells if n_p <= 30000 else cells // 2
        ms, (sc, tm) = timed(lambda: _ffi.cyclone(xp[:c], [u], [p], 1000, 100, 50), reps=2)
        res.append({"case": "one class, random pairs", "cells": c, "genes_used": n_used, "pairs": int(len(p)),
                    "score_ms": tm["main_ms"], "prep_ms": tm["prep_ms"],
                    "pair_evals_per_s": c * 1001.0 * len(p) / (tm["main_ms"] * 1e-3)})
    return res


def small():
    """cfg1 shapes: sandbag on 247 cells x 19 084 genes (3 classes, fraction 0.6), cyclone on 213 cells."""
    rng = np.random.default_rng(1)
    g = 19084
    lab = np.repeat([0, 1, 2], [76, 80, 91])
    x, _ = None, None
    base = rng.normal(0, 1.5, g)
    de = rng.random(g) < 0.1
    dc = rng.integers(0, 3, g)
    lam = np.tile(3.0 * np.exp(base), (247, 1))
    lam = np.where((lab[:, None] == dc[None, :]) & de[None, :], lam * 4.0, lam)
    x = rng.poisson(lam).astype(np.float64)
    gn = ["g{}".format(i) for i in range(g)]
    sn = ["c{}".format(i) for i in range(247)]
    ann = {k: np.where(lab == i)[0] for i, k in enumerate(["G2M", "S", "G1"])}
    ms_sb, m = timed(lambda: pairs.sandbag(x, ann, gn, sn, fraction=0.6), reps=2)
    ms_tab, tab = timed(lambda: pairs.sandbag(x, ann, gn, sn, fraction=0.6, as_table=True), reps=2)
    tm_sb = dict(pairs.sandbag.last_timings)
    x2 = rng.poisson(np.tile(3.0 * np.exp(base), (213, 1))).astype(np.float64)
    sn2 = ["d{}".format(i) for i in range(213)]
    ms_cy, df = timed(lambda: pairs.cyclone(x2, tab, gn, sn2, iterations=1000, min_iter=100, min_pairs=50), reps=2)
    tm_cy = dict(pairs.cyclone.last_timings)
    return {"sandbag_dict_ms": ms_sb, "sandbag_table_ms": ms_tab, "markers": len(tab), "sandbag_device": tm_sb,
            "cyclone_with_those_markers_ms": ms_cy, "cyclone_device": tm_cy}


def million():
    markers, names, keys, used_idx, pair_idx, _, _ = bench.cyclone_setup()
    cells = 1000000
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = torch.empty((cells, bench.CYC_GENES), dtype=torch.float32, device="cuda")
    for i in range(0, cells, 20000):
        x[i:i + 20000] = torch.poisson(torch.full((20000, bench.CYC_GENES), 2.0, device="cuda"), generator=gen)
    ms, (sc, tm) = timed(lambda: _ffi.cyclone(x, used_idx, pair_idx, 1000, 100, 50), reps=2)
    return {"cells": cells, "wall_ms": ms, "cells_per_s": cells / (ms * 1e-3), "timings": tm,
            "nan_scores": int(np.isnan(sc).sum()), "mean_score": [float(s.mean()) for s in sc]}


if __name__ == "__main__":
    what = sys.argv[1:] or ["general", "small"]
    out = {}
    for w in what:
        out[w] = {"general": general, "small": small, "million": million}[w]()
        print(w, json.dumps(out[w], default=float), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "paths_probe.json"), "w"), indent=1, default=float)
