"""Small, fixed workloads for ncu (launch list and --set full captures).

    python profiles/prof_case.py cyclone [cells]     default 9472 cells = one wave of 64-cell tiles on 148 SMs
    python profiles/prof_case.py sandbag [genes] [cells] [workload]   default 6000 genes x 10000 cells of bench.py's
                                                                      "sandbag" workload (or sandbag_lognormal / sandbag_atlas)
Each runs the hot path twice (first call = warm-up) with the matrix generated on the device.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from pypairs_b200 import _ffi  # noqa: E402


def main():
    path = sys.argv[1]
    if path == "cyclone":
        cells = int(sys.argv[2]) if len(sys.argv) > 2 else 9472
        markers, names, keys, used_idx, pair_idx, _, _ = bench.cyclone_setup()
        gen = torch.Generator(device="cuda").manual_seed(1)
        kind = sys.argv[3] if len(sys.argv) > 3 else "poisson"
        if kind == "poisson":
            x = torch.poisson(torch.full((cells, bench.CYC_GENES), 2.0, device="cuda"), generator=gen)
        else:
            x = torch.exp(torch.randn((cells, bench.CYC_GENES), device="cuda", generator=gen))
        for _ in range(2):
            sc, tm = _ffi.cyclone(x, used_idx, pair_idx, 1000, 100, 50)
        print("cyclone", cells, "cells:", tm)
    else:
        g = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
        n = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
        wl = sys.argv[4] if len(sys.argv) > 4 else "sandbag"
        G, N, c, kind, seed = bench.SB_SHAPES[wl]
        x, lab = bench.sandbag_matrix_device(G, N, c, kind, seed)
        x = x[:n, :g].contiguous()
        lab = lab[:n]
        sizes = np.bincount(lab, minlength=c)
        thr = np.ceil(sizes * 0.65).astype(np.int64)
        order = np.argsort(lab, kind="stable")
        for _ in range(3):
            keys, cls, tm = _ffi.sandbag(x, order, sizes, np.arange(g), thr)
        print("sandbag", g, "genes x", n, "cells:", len(keys), "markers", tm)


if __name__ == "__main__":
    main()
