"""GPU parity of the multi-GPU entry points (pp_sandbag_dist / pp_cyclone_dist, the library's NCCL collectives).

* one rank: a communicator of world size 1 on the test GPU -- the whole _dist code path (regions, all-reduce,
  all-gather, gather_to) against the single-GPU entry points;
* two ranks (when >= 2 GPUs are visible): real NCCL ranks, real kernels (tests/_nccl_worker.py), results equal to
  the 1-rank answer and to the CPU oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from test_gpu_sandbag import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_one_rank_communicator():
    from pypairs_b200 import _ffi
    from test_gpu_cyclone import default_setup
    from pypairs_b200.cyclone import filter_marker_pairs
    if _ffi.comm_info()[1] == 0:
        _ffi.comm_init(_ffi.comm_unique_id(), 0, 1)
    assert _ffi.comm_info() == (0, 1)
    rng = np.random.default_rng(3)
    x, lab = synth(rng, 500, 400, 3, dtype=np.float32)
    order, sizes = np.argsort(lab, kind="stable"), np.bincount(lab, minlength=3)
    thr = np.ceil(sizes * 0.6).astype(np.int64)
    k1, c1, _ = _ffi.sandbag(x, order, sizes, np.arange(400), thr)
    kd, cd, _ = _ffi.sandbag(x, order, sizes, np.arange(400), thr, dist=True)
    assert len(k1) > 0 and np.array_equal(k1, kd) and np.array_equal(c1, cd)
    markers, names, xc = default_setup(rng, 200, 2500)
    pr, used = filter_marker_pairs(markers, names)
    ui, pi = [np.where(used[k])[0] for k in pr], [pr[k] for k in pr]
    s1, _ = _ffi.cyclone(xc, ui, pi, 100, 10, 5)
    sd, _ = _ffi.cyclone(xc, ui, pi, 100, 10, 5, shard_cells=[200])
    assert np.array_equal(s1, sd, equal_nan=True)
    _ffi.comm_destroy()
    assert _ffi.comm_info()[1] == 0
    with pytest.raises(_ffi.PyPairsB200Error) as e:
        _ffi.sandbag(x, order, sizes, np.arange(400), thr, dist=True)
    assert e.value.code == _ffi.PP_ENCCL


def _spawn(ws, timeout=420):
    """ws worker processes; the first failing rank ends the test at once (its peers would wait in a collective)."""
    import tempfile
    import time
    port = 29500 + os.getpid() % 2000
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(ws), PYTHONFAULTHANDLER="1")
    logs = [tempfile.TemporaryFile() for _ in range(ws)]
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_nccl_worker.py")],
                              env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=logs[r], stderr=subprocess.STDOUT, cwd=ROOT)
             for r in range(ws)]
    t0 = time.time()
    while any(p.poll() is None for p in procs):
        if any(p.poll() not in (None, 0) for p in procs) or time.time() - t0 > timeout:
            time.sleep(2)
            for p in procs:
                if p.poll() is None:
                    p.kill()
            break
        time.sleep(0.2)
    outs = []
    for p, f in zip(procs, logs):
        p.wait()
        f.seek(0)
        outs.append(f.read().decode(errors="replace"))
    msg = "\n".join("---- rank {} (rc {}) ----\n{}".format(r, p.returncode, o[-4000:]) for r, (p, o) in enumerate(zip(procs, outs)))
    assert all(p.returncode == 0 for p in procs), msg
    assert all("NCCL_OK" in o for o in outs), msg


def test_two_ranks_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _spawn(2)


def test_all_visible_ranks_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 3:
        pytest.skip("needs more than two GPUs")
    _spawn(n)
