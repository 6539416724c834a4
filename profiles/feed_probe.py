"""How should a host matrix reach the GPU in pp_cyclone?  cfg2 (100k x 20k f32, default markers), pinned and pageable,
every feed mode of the library (PP_FEED): auto / gather (host threads + DMA) / zerocopy (gather kernel over mapped
memory) / copy (full rows by DMA).  Prints the library time of the call (no DataFrame) per mode.

    PP_TRACE=1 python profiles/feed_probe.py [cells] [threads]
    python -m torch.distributed.run --nproc-per-node 8 ... profiles/feed_probe.py 100000 4    # all ranks at once, one GPU each:
                                                           what the ways cost when 8 processes share the host (gloo barriers)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from pypairs_b200 import _ffi  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    barrier = lambda: None  # noqa: E731
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("gloo")
        barrier = dist.barrier
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    if len(sys.argv) > 2:
        os.environ["PP_HOST_THREADS"] = sys.argv[2]
    markers, names, keys, used_idx, pair_idx, _, _ = bench.cyclone_setup()
    x = torch.empty((cells, bench.CYC_GENES), dtype=torch.float32, device="cuda")
    bench.fill_cells(x, 1, "poisson")
    ref, _ = _ffi.cyclone(x, used_idx, pair_idx, 1000, 100, 50)
    xp = bench.pinned_copy(x)
    del x
    for label, arr in (("pinned", xp.numpy()), ("pageable", np.array(xp.numpy()))):
        for feed in ("auto", "gather", "zerocopy", "both", "copy"):
            if label == "pageable" and feed in ("zerocopy", "both"):
                continue
            if label == "pageable" and feed == "copy" and world > 1:
                continue  # 0.7 s per call on an idle host: known to be hopeless
            os.environ["PP_FEED"] = feed
            ts = []
            for _ in range(4):
                barrier()
                t0 = time.perf_counter()
                sc, tm = _ffi.cyclone(arr, used_idx, pair_idx, 1000, 100, 50)
                ts.append((time.perf_counter() - t0) * 1e3)
            assert np.array_equal(sc, ref, equal_nan=True)
            print("rank {}/{} {:9s} PP_FEED={:9s} {:7.1f} ms (best of 3 after warm-up; h2d_bytes {:.2f} GB, threads {})".format(
                rank, world, label, feed, min(ts[1:]), tm["h2d_bytes"] / 1e9, os.environ.get("PP_HOST_THREADS", "all")), flush=True)


if __name__ == "__main__":
    main()
