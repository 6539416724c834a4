"""Where does the end-to-end time of pairs.sandbag / pairs.cyclone go?  (cProfile on the GPU box)"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from pypairs_b200 import pairs, settings  # noqa: E402

settings.verbosity = 0
which = sys.argv[1]
if which == "sandbag":
    x, lab = bench.sandbag_matrix_device()
    xh_t = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
    xh_t.copy_(x)
    xh = xh_t.numpy()
    g, n = bench.SB_GENES, bench.SB_CELLS
    gn = np.array(["g{}".format(i) for i in range(g)])
    sn = ["c{}".format(i) for i in range(n)]
    ann = {k: np.where(lab == i)[0] for i, k in enumerate(["G1", "S", "G2M"])}
    as_table = len(sys.argv) > 2 and sys.argv[2] == "table"
    call = lambda: pairs.sandbag(xh, ann, gn, sn, fraction=0.65, as_table=as_table)  # noqa: E731
else:
    markers, names, keys, used_idx, pair_idx, _, _ = bench.cyclone_setup()
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.poisson(torch.full((bench.CYC_CELLS, bench.CYC_GENES), 2.0, device="cuda"), generator=gen)
    xh_t = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
    xh_t.copy_(x)
    xh = xh_t.numpy()
    sn = ["cell{}".format(i) for i in range(bench.CYC_CELLS)]
    call = lambda: pairs.cyclone(xh, markers, names, sn)  # noqa: E731
for _ in range(3):
    call()
t = time.perf_counter()
call()
print("wall ms", (time.perf_counter() - t) * 1e3, "lib timings", (pairs.sandbag if which == "sandbag" else pairs.cyclone).last_timings)
cProfile.run("call()", "/tmp/e2e.prof")
pstats.Stats("/tmp/e2e.prof").sort_stats("cumtime").print_stats(14)
