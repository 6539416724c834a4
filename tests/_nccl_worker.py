"""Worker of tests/test_gpu_dist.py: one of WORLD_SIZE NCCL ranks, one GPU each, real kernels and real collectives.

Every rank checks pairs.sandbag / pairs.cyclone under torch.distributed (pp_sandbag_dist / pp_cyclone_dist: the
library's own NCCL collectives) against the single-GPU entry points on the same device and against the CPU oracle.

PP_WORKER_EMULATE=1 (tests/test_host.py, no GPU): the same script over gloo with the CUDA entry points replaced by the
oracle-backed stand-ins of tests/_gloo_worker.py -- it proves that every rank issues the same sequence of collectives
for every scenario below (a mismatch is a deadlock on the GPU box).  A watchdog dumps the stacks and exits after
PP_WORKER_WATCHDOG seconds."""
import faulthandler
import os
import sys

import numpy as np

EMULATE = os.environ.get("PP_WORKER_EMULATE") == "1"
faulthandler.dump_traceback_later(int(os.environ.get("PP_WORKER_WATCHDOG", "200")), exit=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import oracle as orc  # noqa: E402
from pypairs_b200 import _dist, _ffi, pairs, settings  # noqa: E402
from test_gpu_sandbag import synth  # noqa: E402

settings.verbosity = 0
FORMS = ("host", "csr") if EMULATE else ("host", "device", "csr")


def to_form(x, form):
    import scipy.sparse as sp
    return x if form == "host" else torch.from_numpy(x).cuda() if form == "device" else sp.csr_matrix(x)


def sandbag_checks(rank, ws):
    import scipy.sparse as sp
    rng = np.random.default_rng(11)
    for n, g, c, frac, dtype in [(997, 700, 3, 0.6, np.float32), (250, 333, 10, 0.5, np.float64), (40, 90, 2, 0.65, np.float64),
                                 (5, 64, 3, 0.6, np.float32)]:
        x, lab = synth(rng, n, g, c, dtype=dtype)
        x[:, 7] = 0                                      # an unexpressed gene (dropped on the device, flags max-reduced)
        if n > 100:
            x[: n // 3] = rng.lognormal(0, 1, (n // 3, g)).astype(dtype)  # wide ranks on the first ranks' rows only
            x[: n // 3, 7] = 0
        lab[:c] = np.arange(c)
        perm = rng.permutation(n)                        # classes interleaved over the row blocks of the ranks
        x, lab = x[perm], lab[perm]
        sn = ["s{}".format(i) for i in range(n)]
        gn = ["g{}".format(i) for i in range(g)]
        ann = {"k{}".format(k): np.where(lab == k)[0].tolist() for k in range(c)}
        ann["k0"] = ann["k0"][:-1] if len(ann["k0"]) > 1 else ann["k0"]  # one uncategorised sample
        # single-GPU library call on this rank's device = the 1-rank answer
        p = __import__("pypairs_b200.sandbag", fromlist=["prepare"]).prepare(x, ann, gn, sn, frac)
        k1, c1, _, kept1 = _ffi.sandbag(p["raw"], p["cell_rows"], p["class_sizes"], p["gene_cols"], p["thresholds"],
                                        drop_unexpressed=True)
        for form in FORMS:
            xin = to_form(x, form)
            kd, cd, _, keptd = _ffi.sandbag(xin, p["cell_rows"], p["class_sizes"], p["gene_cols"], p["thresholds"],
                                            drop_unexpressed=True, dist=True, gather_to=-1)
            assert np.array_equal(kd, k1) and np.array_equal(cd, c1) and np.array_equal(keptd, kept1), (n, g, c, form)
            kd, cd, _, _ = _ffi.sandbag(xin, p["cell_rows"], p["class_sizes"], p["gene_cols"], p["thresholds"],
                                        drop_unexpressed=True, dist=True, gather_to=ws - 1)
            assert (len(kd) == len(k1) and np.array_equal(kd, k1)) if rank == ws - 1 else len(kd) == 0
        # probes: per-class counters from the all-gathered planes
        pp = np.array([[0, 1], [2, g - 2], [5, 6]])
        pp = pp[pp[:, 1] < len(kept1)]
        r1 = _ffi.sandbag(x, p["cell_rows"], p["class_sizes"], kept1, p["thresholds"], probe_pairs=pp)
        rd = _ffi.sandbag(x, p["cell_rows"], p["class_sizes"], kept1, p["thresholds"], probe_pairs=pp, dist=True)
        assert np.array_equal(r1[3], rd[3]) and np.array_equal(r1[4], rd[4])
        # public API, every rank the full dict in the reference's order == oracle
        d = pairs.sandbag(x, ann, gn, sn, fraction=frac)
        d1 = pairs.sandbag(x, ann, gn, sn, fraction=frac, gather_to=0)
        assert (d1 is None) == (rank != 0)
        cats = np.full(n, -1)
        for k in range(c):
            cats[np.array(ann["k{}".format(k)], dtype=int)] = k
        keep = cats >= 0
        expressed = (x != 0).any(axis=0)
        order = np.argsort(cats[keep], kind="stable")
        xs = x[keep][order][:, expressed].astype(np.float64)
        sizes = np.bincount(cats[keep], minlength=c)
        thr = np.array([int(np.ceil(int(s) * frac)) for s in sizes])
        res = orc.check_pairs(xs, np.sort(cats[keep]), thr)
        rr, cc = np.where(res != -1)
        gk = np.array(gn)[expressed]
        exp = {}
        for a, b, k in zip(rr, cc, res[rr, cc]):
            exp.setdefault("k{}".format(k), []).append((gk[a], gk[b]))
        assert list(d.keys()) == list(exp.keys()), (list(d.keys()), list(exp.keys()))
        assert all([tuple(map(str, t)) for t in d[k]] == [tuple(map(str, t)) for t in exp[k]] for k in exp)
        if rank == 0:
            assert list(d1.items()) == list(d.items())
    # NaN on one rank's rows fails the call on every rank
    x, lab = synth(rng, 300, 100, 3, dtype=np.float32)
    x[299, 3] = np.nan
    try:
        _ffi.sandbag(x, np.argsort(lab, kind="stable"), np.bincount(lab, minlength=3), np.arange(100), np.array([50, 50, 50]),
                     dist=True)
        raise AssertionError("NaN not reported")
    except _ffi.PyPairsB200Error as e:
        assert e.code == _ffi.PP_ENAN


def cyclone_checks(rank, ws):
    import scipy.sparse as sp
    from test_gpu_cyclone import default_setup
    rng = np.random.default_rng(12)
    for n, kind, dtype in [(301, "poisson", np.float32), (70, "lognormal", np.float64), (3, "poisson", np.float32)]:
        markers, names, x = default_setup(rng, n, 3000, kind, dtype)
        sn = ["c{}".format(i) for i in range(n)]
        pr, used = __import__("pypairs_b200.cyclone", fromlist=["filter_marker_pairs"]).filter_marker_pairs(markers, names)
        keys = list(pr.keys())
        ui, pi = [np.where(used[k])[0] for k in keys], [pr[k] for k in keys]
        s1, _, oh1, ot1 = _ffi.cyclone(x, ui, pi, 200, 20, 5, want_observed=True)   # 1-rank answer on this device
        if n >= 70:
            for i, k in enumerate(keys):  # and the oracle on a subset
                assert np.array_equal(orc.phase_scores(x[:16], 200, 20, 5, pr[k], used[k]), s1[i][:16], equal_nan=True)
        sizes = _dist.shard_sizes(n, ws)
        b, e = _dist.cell_shard(n, rank, ws)
        for form in FORMS:
            xin = to_form(x, form)
            sd, _, ohd, otd = _ffi.cyclone(xin, ui, pi, 200, 20, 5, row_begin=b, n_cells=e - b, shard_cells=sizes,
                                           want_observed=True)
            assert np.array_equal(sd, s1, equal_nan=True) and np.array_equal(ohd, oh1) and np.array_equal(otd, ot1), form
            sd, _ = _ffi.cyclone(xin, ui, pi, 200, 20, 5, row_begin=b, n_cells=e - b, shard_cells=sizes, gather_to=0)
            assert (sd is None) if rank else np.array_equal(sd, s1, equal_nan=True)
        full = pairs.cyclone(x, markers, names, sn, iterations=200, min_iter=20, min_pairs=5)
        assert all(np.array_equal(full[k].values, s1[i], equal_nan=True) for i, k in enumerate(keys)) and list(full.index) == sn
        # local shards of unequal size (an empty one when there are fewer cells than ranks)
        cut = [0] + [min(n, (q + 1) * (n // ws) + 5) for q in range(ws - 1)] + [n]
        lo, hi = cut[rank], cut[rank + 1]
        loc = pairs.cyclone(x[lo:hi], markers, names, sn[lo:hi], iterations=200, min_iter=20, min_pairs=5, local_shard=True)
        assert loc.equals(full)
        loc0 = pairs.cyclone(x[lo:hi], markers, names, sn, iterations=200, min_iter=20, min_pairs=5, local_shard=True,
                             gather_to=ws - 1)
        assert (loc0 is None) == (rank != ws - 1) and (loc0 is None or loc0.equals(full))
    settings.rng = "philox"
    ph = pairs.cyclone(x, markers, names, sn, iterations=200, min_iter=20, min_pairs=5)
    settings.rng = "mt19937"
    sp1, _ = _ffi.cyclone(x, ui, pi, 200, 20, 5, rng="philox", seed=settings.seed)
    assert all(np.array_equal(ph[k].values, sp1[i], equal_nan=True) for i, k in enumerate(keys))


def main():
    rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if EMULATE:
        import _gloo_worker as emu
        _ffi.sandbag, _ffi.cyclone = emu.fake_sandbag, emu.fake_cyclone
        _dist.ensure_comm = lambda device=None: _dist.world()
        dist.init_process_group("gloo", rank=rank, world_size=ws)
    else:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", torch.cuda.current_device()))
        assert _dist.ensure_comm() == (rank, ws) and _ffi.comm_info() == (rank, ws)
    sandbag_checks(rank, ws)
    print("sandbag checks done, rank", rank, flush=True)
    cyclone_checks(rank, ws)
    dist.barrier()
    if not EMULATE:
        _dist.drop_comm()
    dist.destroy_process_group()
    print("NCCL_OK rank", rank)


if __name__ == "__main__":
    main()
