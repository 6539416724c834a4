import os, sys, time
sys.path.insert(0, '/root/repo' if os.path.exists('/root/repo/bench.py') else '.')
import numpy as np, torch
import bench
from pypairs_b200 import _ffi, pairs, settings
settings.verbosity = 0
markers, names, keys, used_idx, pair_idx, _, _ = bench.cyclone_setup()
x = torch.empty((100000, 20000), dtype=torch.float32, device="cuda")
bench.fill_cells(x, 1, "poisson")
xh = np.array(x.cpu().numpy())
sn = ["cell{}".format(i) for i in range(100000)]
for it in range(4):
    t0 = time.perf_counter()
    sc, tm = _ffi.cyclone(xh, used_idx, pair_idx, 1000, 100, 50)
    t1 = time.perf_counter()
    print("ffi call %.1f ms  tm %s" % ((t1 - t0) * 1e3, {k: round(v, 2) for k, v in tm.items() if k.endswith("_ms")}), flush=True)
for it in range(3):
    t0 = time.perf_counter()
    df = pairs.cyclone(xh, markers, names, sn)
    print("pairs.cyclone %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
