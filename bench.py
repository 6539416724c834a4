#!/usr/bin/env python
"""bench.py -- benchmark of the two PyPairs hot paths on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workloads all|LIST]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).

Headline = BASELINE.json configs[1]: cyclone, default_cc_marker, synthetic 100k cells x 20k genes (float32 Poisson
counts), 1000 iterations, PER GPU (weak scaling: every rank scores its own 100k cells).  A step = one pass of the
hot path over that batch.  `value` = cells/s with the matrix resident in HBM, timed around the library call -- at
N > 1 that is pp_cyclone_dist: kernels + the NCCL all-gather of the score columns + the copy of the gathered
columns to the host, on every rank.  `e2e` = the same through pairs.cyclone() from a pinned host ndarray to the
result DataFrame.

Further workloads are reported in the same line (each with value / e2e / roofline / result sha256; cpu_baseline at
N = 1), selectable with --workloads (comma list of the names below, default all):
    sandbag            configs[2]: 20k genes x 10k cells, 3 classes, fraction 0.65, Poisson counts (strong scaling)
    sandbag_lognormal  the same shape, tie-free lognormal float32
    cyclone_lognormal  configs[1] with tie-free lognormal float32 (fp16 kernel instead of the 7-bit one)
    sandbag_atlas      configs[3]: 30k genes x 50k cells, 10 classes (strong scaling)
    cyclone_1m         configs[4]: 1M cells x 20k genes over the N ranks (strong scaling, generated on the device)
    flow               configs[0] shape (leng15-like: 247 sorted / 213 unsorted cells, 19 084 genes): sandbag(fraction 0.6)
                       then cyclone with the markers just found -- host ndarrays + dict vs device tensors + MarkerTable (N = 1)
Strong-scaling workloads use the same data at every N, so their result sha256 must not change with N.

--impl reference times the reference's CPU implementation (oracle/_ref Numba path, else the C port) on the host.
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CYC_CELLS, CYC_GENES, CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS = 100000, 20000, 1000, 100, 50
CYC_1M_CELLS, CYC_1M_BLOCKS = 1000000, 8
SB_FRACTION = 0.65
SB_SHAPES = {"sandbag": (20000, 10000, 3, "poisson", 20263), "sandbag_lognormal": (20000, 10000, 3, "lognormal", 20263),
             "sandbag_atlas": (30000, 50000, 10, "poisson", 20264)}
WORKLOADS = ["sandbag", "sandbag_lognormal", "cyclone_lognormal", "sandbag_atlas", "cyclone_1m", "flow"]
METRIC = "cyclone cells/s @1000 iters"
UNIT = "cells/s"


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz, self.power = index, False, [], set(), None, []

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # pragma: no cover
            self.reasons.add("nvml_unavailable:{}".format(type(e).__name__))

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "power_w_max": max(self.power) if self.power else None,
                "samples": len(sm)}


def ncu_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (None if absent)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]["dram_bytes"]
    except Exception:
        return None


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def max_over_ranks(v, world):
    if world == 1:
        return v
    import torch
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sync_all(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def timed(fn, steps, world):
    """K calls of fn bracketed by barrier + synchronize on both sides -> (ms per step = max over ranks, last result,
    list of the library's per-call timing dicts)."""
    sync_all(world)
    t0 = time.perf_counter()
    tms, res = [], None
    for _ in range(steps):
        res = fn()
        tms.append(res[1] if isinstance(res, tuple) and len(res) > 1 and isinstance(res[1], dict) else
                   res[2] if isinstance(res, tuple) and len(res) > 2 and isinstance(res[2], dict) else None)
    sync_all(world)
    ms = (time.perf_counter() - t0) * 1e3 / steps
    return max_over_ranks(ms, world), res, tms


def tm_mean(tms, key, world):
    vals = [t[key] for t in tms if t]
    return max_over_ranks(float(np.mean(vals)) if vals else 0.0, world)


def cyclone_setup(seed=20262):
    """Gene names = the 1 479 default-marker genes + fillers, shuffled."""
    from pypairs_b200 import datasets
    from pypairs_b200.cyclone import filter_marker_pairs
    markers = datasets.default_cc_marker()
    genes = sorted({g for v in markers.values() for p in v for g in p})
    names = genes + ["X{}".format(i) for i in range(CYC_GENES - len(genes))]
    np.random.default_rng(seed).shuffle(names)
    pairs, used = filter_marker_pairs(markers, names)
    keys = list(pairs.keys())
    return markers, names, keys, [np.where(used[k])[0] for k in keys], [pairs[k] for k in keys], pairs, used


def fill_cells(x, seed, kind):
    """Synthetic expression values into the rows of the device tensor x, 10 000 rows at a time (bounded temporaries)."""
    import torch
    gen = torch.Generator(device="cuda").manual_seed(seed)
    for i in range(0, x.shape[0], 10000):
        n = min(10000, x.shape[0] - i)
        if kind == "poisson":
            x[i:i + n] = torch.poisson(torch.full((n, x.shape[1]), 2.0, device="cuda"), generator=gen)
        else:
            x[i:i + n] = torch.exp(torch.randn((n, x.shape[1]), device="cuda", generator=gen))


def sandbag_matrix_device(g, n, c, kind, seed):
    """cfg3 / cfg4 synthetic generator: 2 % DE genes, effect 2^(+-2) in the DE class; Poisson
    counts (tie-heavy) or the same means with lognormal noise (tie-free)."""
    import torch
    gen = torch.Generator(device="cuda").manual_seed(seed)
    lab = torch.randint(0, c, (n,), generator=gen, device="cuda")
    base = torch.randn(g, generator=gen, device="cuda") * 1.5
    de = torch.rand(g, generator=gen, device="cuda") < 0.02
    de_class = torch.randint(0, c, (g,), generator=gen, device="cuda")
    sign = torch.where(torch.rand(g, generator=gen, device="cuda") < 0.5, -1.0, 1.0)
    eff = torch.where(de, torch.pow(2.0, 2.0 * sign), torch.ones_like(sign))
    x = torch.empty((n, g), dtype=torch.float32, device="cuda")
    for i in range(0, n, 5000):
        j = min(n, i + 5000)
        lam = 3.0 * torch.exp(base)[None, :] * torch.where(lab[i:j, None] == de_class[None, :], eff[None, :], 1.0)
        if kind == "poisson":
            x[i:j] = torch.poisson(lam, generator=gen)
        else:
            x[i:j] = lam * torch.exp(0.5 * torch.randn((j - i, g), device="cuda", generator=gen))
    x[:, 0] += 1.0
    return x, lab.cpu().numpy().astype(np.int64)


def host_memory_allows(extra_bytes):
    """True if `extra_bytes` more of host memory (summed over the ranks of this host) leave a comfortable margin."""
    try:
        import psutil
        return psutil.virtual_memory().available > 2 * extra_bytes + (16 << 30)
    except Exception:
        return True


def pinned_copy(x):
    import torch
    t = torch.empty(tuple(x.shape), dtype=x.dtype, pin_memory=True)
    t.copy_(x)
    torch.cuda.synchronize()
    return t


# ------------------------------------------------------------------------------------------------ CPU arms
def reference_setup(seed=20262):
    """The cyclone workload from oracle/ and tests/golden only (the reference arm must not load the product)."""
    from oracle import oracle as orc
    markers = orc.default_cc_marker()
    genes = sorted({g for v in markers.values() for p in v for g in p})
    names = genes + ["X{}".format(i) for i in range(CYC_GENES - len(genes))]
    np.random.default_rng(seed).shuffle(names)
    pairs, used = orc.filter_marker_pairs(markers, names)
    return list(pairs.keys()), pairs, used


def cpu_cyclone_runner(n_threads):
    """(step(xs) -> [scores per class], port(xs), description dict): the reference's own Numba gufunc from oracle/_ref
    when that copy is present (kind "reference"), else the C/OpenMP restatement (kind "port"); `n_threads` threads."""
    from oracle import oracle as orc, ref_numba
    keys, pairs, used = reference_setup()
    info = {"cores": n_threads}
    port = lambda xs: [orc.phase_scores(xs, CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS, pairs[k], used[k],  # noqa: E731
                                        n_threads=n_threads) for k in keys]
    if ref_numba.available():
        ref_numba.configure_threads(n_threads)
        r = ref_numba.load()
        step = lambda xs: [ref_numba.phase_scores(xs, CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS, pairs[k], used[k])  # noqa: E731
                           for k in keys]
        info.update(kind="reference", cores=int(r.threads), numba=r.numba.__version__,
                    impl="pypairs.tools.cyclone.get_phase_scores (Numba gufunc, target=parallel) from oracle/_ref")
    else:
        orc.build()
        step = port
        info.update(kind="port", impl="oracle/pp_oracle.c (C + OpenMP restatement; oracle/_ref absent)")
    return step, port, info


def cpu_cyclone_baseline(xs_all, gpu_scores, seconds=12.0):
    """cpu_baseline of a cyclone workload: the first cells of the workload's matrix on all host cores; the GPU scores
    of those cells must be bit-identical.  xs_all: host ndarray [>= 2048 cells, genes]."""
    n_threads = os.cpu_count() or 1
    step, port, info = cpu_cyclone_runner(n_threads)
    step(xs_all[:64])  # JIT / thread-pool start
    t0 = time.perf_counter()
    step(xs_all[:64])
    rate = 64 / (time.perf_counter() - t0)
    ncell = int(min(len(xs_all), max(128, rate * seconds)))
    t0 = time.perf_counter()
    ref = step(xs_all[:ncell])
    dt = time.perf_counter() - t0
    assert all(np.array_equal(ref[i], gpu_scores[i][:ncell], equal_nan=True) for i in range(len(ref))), \
        "GPU != CPU on the timed sample"
    out = dict(info, value=ncell / dt, unit=UNIT,
               sample="first {} cells of the workload (all 3 classes, 1000 iterations); cells are independent, so the rate "
                      "extrapolates linearly; GPU scores bit-identical on the sample".format(ncell))
    if info["kind"] == "reference":
        from oracle import ref_numba
        out["threading_layer"] = ref_numba.threading_layer()
        m = min(ncell, 256)
        t0 = time.perf_counter()
        chk = port(xs_all[:m])
        out["port_value"] = m / (time.perf_counter() - t0)
        assert all(np.array_equal(a[:m], b, equal_nan=True) for a, b in zip(ref, chk)), "Numba reference != C port"
    return out


def cpu_sandbag_baseline(x_dev, lab, c, keys, cls, seconds=12.0):
    """cpu_baseline of a sandbag workload: all cells x the first G' genes (work scales with G'(G'-1)/2); the GPU markers
    among those genes must be identical."""
    from oracle import oracle as orc, ref_numba
    n_threads = os.cpu_count() or 1
    n, g = x_dev.shape
    sizes = np.bincount(lab, minlength=c)
    thr = orc.calc_thresholds(sizes, SB_FRACTION)
    order = np.argsort(lab, kind="stable")
    cats = lab[order]
    use_ref = ref_numba.available()
    if use_ref:
        ref_numba.configure_threads(n_threads)
        fn = ref_numba.compiled_check_pairs(n_threads)
        run = lambda xs: fn(xs, cats, thr)  # noqa: E731
        run(np.zeros((len(cats), 8)))  # JIT outside the timed region
    else:
        run = lambda xs: orc.check_pairs(xs, cats, thr, n_threads)  # noqa: E731
    sub = lambda gs: np.ascontiguousarray(x_dev[:, :gs].cpu().numpy().astype(np.float64)[order])  # noqa: E731
    gs = 300
    xs = sub(gs)
    t0 = time.perf_counter()
    run(xs)
    rate = gs * (gs - 1) / 2.0 * n / (time.perf_counter() - t0)
    gs = int(min(4000, g, max(300, np.sqrt(2.0 * rate * seconds / n))))
    xs = sub(gs)
    t0 = time.perf_counter()
    res = run(xs)
    dt = time.perf_counter() - t0
    rr, cc = np.where(res != -1)
    kk = res[rr, cc]
    rows, cols = (keys // np.uint64(g)).astype(np.int64), (keys % np.uint64(g)).astype(np.int64)
    m = (rows < gs) & (cols < gs)
    assert np.array_equal(rows[m], rr) and np.array_equal(cols[m], cc) and np.array_equal(cls[m], kk), \
        "GPU != CPU on the timed sample"
    out = {"value": gs * (gs - 1) / 2.0 * n / dt, "unit": "pair*cell/s", "cores": n_threads,
           "kind": "reference" if use_ref else "port",
           "impl": "utils.parallel_njit(check_pairs) from oracle/_ref (Numba, JIT excluded)" if use_ref else "oracle/pp_oracle.c",
           "sample": "all {} cells x first {} genes ({} markers, identical to the GPU's among those genes); extrapolated: "
                     "work scales with G(G-1)/2".format(n, gs, len(rr))}
    if use_ref:
        t0 = time.perf_counter()
        chk = orc.check_pairs(xs, cats, thr, n_threads)
        out["port_value"] = gs * (gs - 1) / 2.0 * n / (time.perf_counter() - t0)
        assert np.array_equal(chk, res), "Numba reference != C port"
        out["threading_layer"] = ref_numba.threading_layer()
    return out


def oracle_spot_check_cyclone(xs, gpu_scores, n=48):
    """Every run is a parity run: the C port on a few cells (explicit thread count: torchrun exports OMP_NUM_THREADS=1)."""
    from oracle import oracle as orc
    keys, pairs, used = reference_setup()
    for i, k in enumerate(keys):
        ref = orc.phase_scores(xs[:n], CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS, pairs[k], used[k], n_threads=os.cpu_count() or 1)
        assert np.array_equal(ref, gpu_scores[i][:n], equal_nan=True), "GPU != oracle ({})".format(k)
    return n


def oracle_spot_check_sandbag(x_dev, lab, c, keys, cls, gs=160):
    from oracle import oracle as orc
    g = x_dev.shape[1]
    order = np.argsort(lab, kind="stable")
    xs = np.ascontiguousarray(x_dev[:, :gs].cpu().numpy().astype(np.float64)[order])
    thr = orc.calc_thresholds(np.bincount(lab, minlength=c), SB_FRACTION)
    res = orc.check_pairs(xs, lab[order], thr, os.cpu_count() or 1)
    rr, cc = np.where(res != -1)
    rows, cols = (keys // np.uint64(g)).astype(np.int64), (keys % np.uint64(g)).astype(np.int64)
    m = (rows < gs) & (cols < gs)
    assert np.array_equal(rows[m], rr) and np.array_equal(cols[m], cc) and np.array_equal(cls[m], res[rr, cc]), "GPU != oracle"
    return gs


# ------------------------------------------------------------------------------------------------ b200 arm: cyclone
def bench_cyclone(args, ctx, kind="poisson", headline=True):
    """configs[1] per GPU (weak scaling).  kind: poisson (tie-heavy counts -> 7-bit kernel) / lognormal (tie-free -> fp16)."""
    import torch
    from pypairs_b200 import _dist, _ffi, pairs
    rank, local_rank, world, peaks = ctx["rank"], ctx["local_rank"], ctx["world"], ctx["peaks"]
    markers, names, keys, used_idx, pair_idx, pairs_d, used_d = cyclone_setup()
    n_pairs = sum(len(p) for p in pair_idx)
    n_union = len(np.unique(np.concatenate(used_idx)))
    steps = args.steps if headline else min(args.steps, 3)
    x = torch.empty((CYC_CELLS, CYC_GENES), dtype=torch.float32, device="cuda")
    fill_cells(x, 20262 + rank, kind)
    shard = np.full(world, CYC_CELLS, dtype=np.int64)
    if world > 1:
        run = lambda: _ffi.cyclone(x, used_idx, pair_idx, CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS, shard_cells=shard)  # noqa: E731
    else:
        run = lambda: _ffi.cyclone(x, used_idx, pair_idx, CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS)  # noqa: E731
    warm = [run() for _ in range(args.warmup)]  # kept alive: the library's pinned result pool reaches its steady size
    del warm
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_step, (sc, _), tms = timed(run, steps, world)
    clocks = sampler.result()
    dev_step, main_step, prep_step = (tm_mean(tms, k, world) for k in ("total_ms", "main_ms", "prep_ms"))
    launches = sum(t["n_launches"] for t in tms)
    total_cells = CYC_CELLS * world
    assert sc.shape == (len(keys), total_cells)
    mine = sc[:, rank * CYC_CELLS:(rank + 1) * CYC_CELLS]
    checked = oracle_spot_check_cyclone(x[:64].cpu().numpy(), sc) if rank == 0 else 0  # rank 0's shard comes first

    # ---- end to end through the public API: host ndarray -> DataFrame
    sn = ["cell{}".format(i) for i in range(total_cells)]  # local_shard: the names of all cells, no exchange of names
    xh_t = pinned_copy(x)
    xh = xh_t.numpy()
    e2e_call = lambda a: pairs.cyclone(a, markers, names, sn, iterations=CYC_ITERS, min_iter=CYC_MIN_ITER,  # noqa: E731
                                       min_pairs=CYC_MIN_PAIRS, local_shard=(world > 1))
    warm = [e2e_call(xh), e2e_call(xh)]
    del warm
    e2e_ms, df, _ = timed(lambda: e2e_call(xh), steps, world)
    e2e_tm = pairs.cyclone.last_timings
    assert len(df) == total_cells
    assert np.array_equal(df[keys[0]].values[rank * CYC_CELLS:(rank + 1) * CYC_CELLS], mine[0], equal_nan=True)
    # the same call with the shard already resident on the device (a CUDA tensor): no feed, kernels + collectives + frame
    dev_call = lambda: pairs.cyclone(x, markers, names, sn, iterations=CYC_ITERS, min_iter=CYC_MIN_ITER,  # noqa: E731
                                     min_pairs=CYC_MIN_PAIRS, local_shard=(world > 1))
    warm = [dev_call(), dev_call()]
    del warm
    dev_ms, df_dev, _ = timed(dev_call, min(steps, 3), world)
    assert df_dev.equals(df)
    device_resident = {"value": total_cells / (dev_ms * 1e-3), "ms_per_step": dev_ms,
                       "api": "pairs.cyclone(CUDA tensor{}) -> DataFrame of all {} cells on every rank".format(
                           ", local_shard=True" if world > 1 else "", total_cells)}
    del df_dev
    pageable = None
    if headline and not args.quick and host_memory_allows(xh.nbytes * world):
        xp = np.array(xh)  # ordinary (pageable) ndarray, what a user holds
        e2e_call(xp)
        pg_ms, _, _ = timed(lambda: e2e_call(xp), min(steps, 3), world)
        pageable = {"value": total_cells / (pg_ms * 1e-3), "ms_per_step": pg_ms, "api": "the same call on a pageable ndarray"}
        del xp

    evals = float(CYC_CELLS) * (CYC_ITERS + 1) * n_pairs  # pair evaluations per launch (per GPU)
    alg_ops = 2.0 * evals                                   # `!=` and `>` per evaluation (cyclone.py:212-215)
    prep_bytes = CYC_CELLS * CYC_GENES * 4 + CYC_CELLS * n_union * 2
    out = {
        "value": total_cells / (ms_step * 1e-3), "ms_per_step": ms_step, "device_ms_per_step": dev_step,
        "kernel_ms": {"score": main_step, "prep_and_tables": prep_step},
        "gpu_launches": launches, "clocks": clocks, "steps": steps,
        "timed_call": "pp_cyclone_dist (kernels + NCCL all-gather of the score columns + D2H of [3, {}] on every rank)".format(
            total_cells) if world > 1 else "pp_cyclone (kernels + D2H of the scores)",
        "result_sha256_rank0_shard": sha(sc[:, :CYC_CELLS]), "oracle_checked_cells": checked,
        "e2e": {"value": total_cells / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(e2e_tm["h2d_bytes"]), "d2h_bytes_per_step": len(keys) * total_cells * 8,
                "host_matrix_bytes": CYC_CELLS * CYC_GENES * 4, "pageable": pageable, "device_resident": device_resident,
                "api": "pypairs_b200.pairs.cyclone(pinned ndarray{}) -> DataFrame of all {} cells on every rank; the "
                       "library moves only the {} marker-gene columns of each cell over PCIe".format(
                           ", local_shard=True" if world > 1 else "", total_cells, n_union)},
        "roofline": {"bound": "int_alu", "kernel": "score_kernel",
                     "achieved": alg_ops / (main_step * 1e-3) / 1e9, "peak": peaks["lop3_lane_ops_per_s"] / 1e9,
                     "unit": "Gop/s", "frac": alg_ops / (main_step * 1e-3) / peaks["lop3_lane_ops_per_s"],
                     "peak_source": "pp_alu_peak micro-benchmark (independent LOP3 stream), measured in this run",
                     "algorithmic_ops_per_unit": 2.0 * (CYC_ITERS + 1) * n_pairs, "units_per_launch": CYC_CELLS,
                     "note": "count data: four 7-bit ranks per 32-bit lane, so the algorithmic rate may exceed the "
                             "one-op-per-lane peak; ncu: issue slots 88 %, ALU pipe 76 %" if kind == "poisson" else
                             "tie-free data: two integer-valued fp16 ranks per lane (HSET2 / HADD2)",
                     "traffic": ncu_traffic("score_kernel") if kind == "poisson" else None},
        "roofline_prep": {"bound": "hbm", "kernel": "rank_kernel+tables", "achieved": prep_bytes / (prep_step * 1e-3) / 1e9,
                          "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": prep_bytes / (prep_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
                          "peak_source": peaks["hbm_source"]},
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_cyclone_baseline(x[:2048].cpu().numpy(), sc, 12.0 if headline else 5.0)
    del x, xh_t, xh
    torch.cuda.empty_cache()
    return out


def bench_cyclone_1m(args, ctx):
    """configs[4]: 1M cells x 20k genes sharded over the ranks (strong scaling); the matrix is generated on the device in
    8 blocks of 125k cells with fixed per-block seeds, so the merged result is the same at every N."""
    import torch
    from pypairs_b200 import _ffi, pairs
    rank, local_rank, world, peaks = ctx["rank"], ctx["local_rank"], ctx["world"], ctx["peaks"]
    if CYC_1M_BLOCKS % world:
        return {"skipped": "world size {} does not divide the {} data blocks".format(world, CYC_1M_BLOCKS)}
    markers, names, keys, used_idx, pair_idx, pairs_d, used_d = cyclone_setup()
    per_block = CYC_1M_CELLS // CYC_1M_BLOCKS
    my_blocks = range(rank * CYC_1M_BLOCKS // world, (rank + 1) * CYC_1M_BLOCKS // world)
    n_local = per_block * len(my_blocks)
    free = torch.cuda.mem_get_info()[0]
    if free < n_local * CYC_GENES * 4 * 1.15 + (8 << 30):
        return {"skipped": "needs {:.0f} GB of device memory per rank, {:.0f} GB free".format(n_local * CYC_GENES * 4 / 1e9, free / 1e9)}
    x = torch.empty((n_local, CYC_GENES), dtype=torch.float32, device="cuda")
    for i, b in enumerate(my_blocks):
        fill_cells(x[i * per_block:(i + 1) * per_block], 5000 + b, "poisson")
    steps = min(args.steps, 3)
    shard = np.full(world, n_local, dtype=np.int64)
    if world > 1:
        run = lambda: _ffi.cyclone(x, used_idx, pair_idx, CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS, shard_cells=shard)  # noqa: E731
    else:
        run = lambda: _ffi.cyclone(x, used_idx, pair_idx, CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS)  # noqa: E731
    warm = [run() for _ in range(min(args.warmup, 3))]
    del warm
    ms_step, (sc, _), tms = timed(run, steps, world)
    main_step = tm_mean(tms, "main_ms", world)
    assert sc.shape == (len(keys), CYC_1M_CELLS)
    checked = oracle_spot_check_cyclone(x[:32].cpu().numpy(), sc, 32) if rank == 0 else 0
    n_pairs = sum(len(p) for p in pair_idx)
    alg_ops = 2.0 * float(n_local) * (CYC_ITERS + 1) * n_pairs
    out = {"metric": METRIC, "unit": UNIT, "scaling": "strong", "steps": steps,
           "config": {"workload": "configs[4]: cyclone 1M cells x 20k genes, 1000 iterations, cells sharded over {} rank(s), "
                                  "NCCL all-gather of the score columns".format(world), "cells_per_gpu": n_local},
           "value": CYC_1M_CELLS / (ms_step * 1e-3), "ms_per_step": ms_step, "device_ms_per_step": tm_mean(tms, "total_ms", world),
           "kernel_ms": {"score": main_step, "prep_and_tables": tm_mean(tms, "prep_ms", world)},
           "gpu_launches": sum(t["n_launches"] for t in tms), "result_sha256": sha(sc), "oracle_checked_cells": checked,
           "roofline": {"bound": "int_alu", "kernel": "score_kernel", "achieved": alg_ops / (main_step * 1e-3) / 1e9,
                        "peak": peaks["lop3_lane_ops_per_s"] / 1e9, "unit": "Gop/s",
                        "frac": alg_ops / (main_step * 1e-3) / peaks["lop3_lane_ops_per_s"]}}
    if n_local <= 250000:  # end to end from the host only where the shard fits comfortably in pinned host memory
        sn = ["cell{}".format(i) for i in range(CYC_1M_CELLS)]
        xh_t = pinned_copy(x)
        xh = xh_t.numpy()
        e2e_call = lambda: pairs.cyclone(xh, markers, names, sn, iterations=CYC_ITERS, min_iter=CYC_MIN_ITER,  # noqa: E731
                                         min_pairs=CYC_MIN_PAIRS, local_shard=True, gather_to=0)
        e2e_call()
        e2e_ms, df, _ = timed(e2e_call, steps, world)
        assert (df is None) == (rank != 0) and (df is None or len(df) == CYC_1M_CELLS)
        out["e2e"] = {"value": CYC_1M_CELLS / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                      "h2d_bytes_per_step": int(pairs.cyclone.last_timings["h2d_bytes"]),
                      "d2h_bytes_per_step": len(keys) * CYC_1M_CELLS * 8,
                      "api": "pairs.cyclone(pinned shard, local_shard=True, gather_to=0) -> one DataFrame of 1M cells on rank 0"}
        del xh_t, xh
    else:
        out["e2e"] = None
        out["e2e_note"] = "the {:.0f} GB shard of a rank is generated on the device (SURVEY.md 8d); no host copy at this N".format(
            n_local * CYC_GENES * 4 / 1e9)
    del x
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ b200 arm: sandbag
def bench_sandbag(args, ctx, name):
    import torch
    from pypairs_b200 import _ffi, pairs
    rank, local_rank, world, peaks = ctx["rank"], ctx["local_rank"], ctx["world"], ctx["peaks"]
    g, n, c, kind, seed = SB_SHAPES[name]
    x, lab = sandbag_matrix_device(g, n, c, kind, seed)
    sizes = np.bincount(lab, minlength=c)
    thr = np.array([int(np.ceil(int(s) * SB_FRACTION)) for s in sizes], dtype=np.int64)
    order = np.argsort(lab, kind="stable")
    steps = args.steps if name == "sandbag" else min(args.steps, 3)
    run = lambda: _ffi.sandbag(x, order, sizes, np.arange(g), thr, dist=(world > 1))  # noqa: E731
    warm = [run() for _ in range(min(args.warmup, 3))]  # kept alive: the library's pinned result pool reaches its steady size
    del warm
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_step, (keys, cls, tm), tms = timed(run, steps, world)
    clocks = sampler.result()
    main_step, prep_step = tm_mean(tms, "main_ms", world), tm_mean(tms, "prep_ms", world)
    units = g * (g - 1) / 2.0 * n  # gene-pair x cell evaluations of the whole job
    checked = oracle_spot_check_sandbag(x, lab, c, keys, cls) if rank == 0 else 0

    # ---- end to end through pairs.sandbag: host ndarray -> marker dict (built on rank 0 when there are several ranks)
    class_names = ["G1", "S", "G2M"] if c == 3 else ["type{}".format(i) for i in range(c)]
    xh_t = pinned_copy(x)
    xh = xh_t.numpy()
    gn = np.array(["g{}".format(i) for i in range(g)])
    sn = ["c{}".format(i) for i in range(n)]
    ann = {k: np.where(lab == i)[0] for i, k in enumerate(class_names)}
    gt = 0 if world > 1 else None
    # the columnar MarkerTable first, so that the cyclic GC is not busy with millions of tuples of the dict results
    tab_call = lambda a=xh: pairs.sandbag(a, ann, gn, sn, fraction=SB_FRACTION, as_table=True, gather_to=gt)  # noqa: E731
    warm = [tab_call(), tab_call()]
    del warm
    e2e_tab_ms, tab, _ = timed(tab_call, steps, world)
    e2e_call = lambda a=xh: pairs.sandbag(a, ann, gn, sn, fraction=SB_FRACTION, gather_to=gt)  # noqa: E731
    e2e_call()
    e2e_ms, d, _ = timed(e2e_call, steps, world)
    e2e_tm = pairs.sandbag.last_timings
    n_markers = len(keys)
    if rank == 0:
        assert len(tab) == n_markers and sum(len(v) for v in d.values()) == n_markers
    pageable = None
    if name == "sandbag" and not args.quick:
        xp = np.array(xh)
        tab_call(xp)
        pg_ms, _, _ = timed(lambda: tab_call(xp), min(steps, 3), world)
        pageable = {"value": units / (pg_ms * 1e-3), "ms_per_step": pg_ms, "api": "as_table call on a pageable ndarray"}
        del xp
    my_pairs = tm["n_units"]
    alg_ops = 2.0 * my_pairs * n  # one `>` and one `<` per gene pair x cell (sandbag.py:179-182), this rank's launch
    rank_bytes = float(n) * g * 6 / world  # f32 in, u16 out, this rank's cells
    cfg = {"sandbag": "configs[2]", "sandbag_lognormal": "configs[2] shape, tie-free values", "sandbag_atlas": "configs[3]"}[name]
    out = {
        "metric": "sandbag gene-pair x cell compares/s", "unit": "pair*cell/s", "scaling": "strong", "steps": steps,
        "config": {"workload": "{}: sandbag synthetic {} genes x {} cells, {} classes, fraction={}".format(cfg, g, n, c, SB_FRACTION),
                   "dtype_in": "float32 " + ("Poisson counts" if kind == "poisson" else "lognormal values (tie-free)"),
                   "markers_found": n_markers, "bit_planes": tm["n_planes"], "bit_planes_compared_per_pair": tm["avg_planes"]},
        "value": units / (ms_step * 1e-3), "ms_per_step": ms_step, "device_ms_per_step": tm_mean(tms, "total_ms", world),
        "kernel_ms": {"pair": main_step, "rank_bitslice_and_plane_allgather": prep_step},
        "gpu_launches": sum(t["n_launches"] for t in tms), "clocks": clocks,
        "timed_call": "pp_sandbag_dist (prep of 1/N of the cells, NCCL all-gather of the bit planes, 1/N of the gene-pair tiles, "
                      "NCCL all-gather + sort of the marker keys, D2H on every rank)" if world > 1 else "pp_sandbag",
        "result_sha256": sha(keys, cls), "oracle_checked_genes": checked,
        "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "pair*cell/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(e2e_tm["h2d_bytes"]), "d2h_bytes_per_step": int(n_markers * 12),
                "api": "pypairs_b200.pairs.sandbag(pinned ndarray{}) -> dict".format(", gather_to=0" if world > 1 else ""),
                "as_table": {"value": units / (e2e_tab_ms * 1e-3), "ms_per_step": e2e_tab_ms,
                             "api": "pairs.sandbag(..., as_table=True) -> utils.MarkerTable (index arrays, no tuples)"},
                "pageable": pageable},
        "roofline": {"bound": "int_alu", "kernel": "pair_kernel", "achieved": alg_ops / (main_step * 1e-3) / 1e9,
                     "peak": peaks["lop3_lane_ops_per_s"] / 1e9, "unit": "Gop/s",
                     "frac": alg_ops / (main_step * 1e-3) / peaks["lop3_lane_ops_per_s"],
                     "machine_frac": tm["avg_planes"] / 32.0 * my_pairs * n / (main_step * 1e-3) / peaks["lop3_lane_ops_per_s"],
                     "peak_source": "pp_alu_peak micro-benchmark (independent LOP3 stream), measured in this run",
                     "algorithmic_ops_per_unit": 2, "units_per_launch": my_pairs * n,
                     "machine_lop3_per_unit": tm["avg_planes"] / 32.0,
                     "traffic": ncu_traffic("pair_kernel") if (world == 1 and name == "sandbag") else None},
        "roofline_prep": {"bound": "hbm", "kernel": "rank_kernel+col_or+bitslice(+plane all-gather)",
                          "achieved": rank_bytes / (prep_step * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": rank_bytes / (prep_step * 1e-3) / 1e9 / peaks["hbm_gbs"]},
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_sandbag_baseline(x, lab, c, keys, cls, 12.0 if name == "sandbag" else 5.0)
    del x, xh_t, xh
    torch.cuda.empty_cache()
    return out


def bench_flow(args, ctx):
    """configs[0]-shaped train -> predict flow (README.rst:55-72 of the reference; the leng15 matrix itself is not in
    the checkout): pairs.sandbag on 247 class-sorted cells, pairs.cyclone on 213 other cells with the markers found.
    (a) the reference's own way: host ndarrays, marker dict of tuples; (b) both matrices resident on the device,
    markers as a columnar MarkerTable -- no upload, no tuples, no name dict of 4 M pairs."""
    import torch
    from pypairs_b200 import pairs
    if ctx["world"] > 1:
        return {"skipped": "single-GPU flow"}
    g, sizes = 19084, {"G2M": 76, "S": 80, "G1": 91}
    gen = torch.Generator(device="cuda").manual_seed(20261)
    lab = np.concatenate([np.full(n, i) for i, n in enumerate(sizes.values())])
    base = torch.randn(g, generator=gen, device="cuda") * 1.5
    de = torch.rand(g, generator=gen, device="cuda") < 0.03
    de_class = torch.randint(0, 3, (g,), generator=gen, device="cuda")
    lam0 = 3.0 * torch.exp(base)
    lab_d = torch.from_numpy(lab).cuda()
    lam = lam0[None, :] * torch.where((lab_d[:, None] == de_class[None, :]) & de[None, :], 6.0, 1.0)
    xd_train = torch.poisson(lam, generator=gen).to(torch.float32)
    lab_t = torch.randint(0, 3, (213,), generator=gen, device="cuda")
    lam_t = lam0[None, :] * torch.where((lab_t[:, None] == de_class[None, :]) & de[None, :], 6.0, 1.0)
    xd_test = torch.poisson(lam_t, generator=gen).to(torch.float32)
    gn = ["gene{}".format(i) for i in range(g)]
    sn_train = ["cell{}".format(i) for i in range(len(lab))]
    sn_test = ["u{}".format(i) for i in range(213)]
    ann = {k: np.where(lab == i)[0] for i, k in enumerate(sizes.keys())}
    xh_train, xh_test = xd_train.cpu().numpy(), xd_test.cpu().numpy()

    def host_flow():
        m = pairs.sandbag(xh_train, ann, gn, sn_train, fraction=0.6)
        return m, pairs.cyclone(xh_test, m, gn, sn_test)

    def device_flow():
        t = pairs.sandbag(xd_train, ann, gn, sn_train, fraction=0.6, as_table=True)
        return t, pairs.cyclone(xd_test, t, gn, sn_test)

    out = {"config": {"workload": "configs[0] shape: sandbag(247 cells x 19084 genes, fraction 0.6) -> cyclone(213 cells, "
                                  "markers just found, 1000 iterations); synthetic counts, 3 % DE genes"},
           "metric": "train -> predict flow, seconds", "unit": "s", "higher_is_better": False}
    res = {}
    for name, fn in (("device_resident_table", device_flow), ("host_dict", host_flow)):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            m, df = fn()
        torch.cuda.synchronize()
        res[name] = (m, df)
        out[name + "_s"] = (time.perf_counter() - t0) / 2
    t, df_d = res["device_resident_table"]
    m, df_h = res["host_dict"]
    assert df_d.equals(df_h) and list(t.to_dict().items()) == list(m.items()), "device-resident flow != host flow"
    out["markers_found"] = len(t)
    out["marker_genes"] = int(len(np.unique(np.concatenate([t.rows, t.cols]))))
    out["saved_s"] = out["host_dict_s"] - out["device_resident_table_s"]
    out["prediction_accuracy_on_synthetic_labels"] = float(np.mean(
        np.array(list(sizes.keys()))[lab_t.cpu().numpy()] == df_d["max_class"].values.astype(str)))
    return out


def run_b200(args):
    # libraries (NCCL's version banner, build output) write to fd 1: keep stdout for the one JSON line
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    rank, local_rank, world = env_world()
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    sync_all(world)
    from pypairs_b200 import _dist, _ffi, settings
    settings.verbosity = 0
    if world > 1:
        _dist.ensure_comm()
    lop3, n_sm = _ffi.alu_peak(5)
    peaks = {"lop3_lane_ops_per_s": lop3, "n_sm": n_sm, "hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)"}
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        peaks["hbm_gbs"] = json.load(open(mp))["hbm_gbs"]
        peaks["hbm_source"] = "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    ctx = {"rank": rank, "local_rank": local_rank, "world": world, "peaks": peaks}
    line = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16 dense ranks (from f32), compared as packed 7-bit integers (count data) or fp16 pairs",
            "data": "synthetic",
            "config": {"workload": "configs[1]: cyclone default_cc_marker, synthetic {} cells x {} genes per GPU, "
                                   "{} iterations".format(CYC_CELLS, CYC_GENES, CYC_ITERS),
                       "cells_per_gpu": CYC_CELLS, "genes": CYC_GENES, "iterations": CYC_ITERS, "rng": "mt19937 replay",
                       "parallelism": "cells sharded, {} rank(s)".format(world),
                       "l2": "inputs (8 GB matrix, 33 MB pair tables re-read per cell tile) exceed the 126 MB L2"},
            "peaks": peaks}
    line.update(bench_cyclone(args, ctx, "poisson", headline=True))
    line["steps"] = args.steps
    extra = {}
    for name in args.workloads:
        t0 = time.perf_counter()
        if name in SB_SHAPES:
            extra[name] = bench_sandbag(args, ctx, name)
        elif name == "cyclone_lognormal":
            extra[name] = bench_cyclone(args, ctx, "lognormal", headline=False)
            extra[name]["config"] = {"workload": "configs[1] with tie-free lognormal float32 values, per GPU"}
            extra[name]["metric"], extra[name]["unit"], extra[name]["scaling"] = METRIC, UNIT, "weak"
        elif name == "cyclone_1m":
            extra[name] = bench_cyclone_1m(args, ctx)
        elif name == "flow":
            extra[name] = bench_flow(args, ctx)
        extra[name]["bench_wall_s"] = time.perf_counter() - t0
        line["gpu_launches"] += extra[name].get("gpu_launches", 0)
    if "sandbag" in extra:
        line["sandbag"] = extra.pop("sandbag")
    line["workloads"] = extra
    if world > 1:
        import torch.distributed as dist
        _dist.drop_comm()
        dist.destroy_process_group()
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's CPU implementation of the path on all host cores, on bounded samples of the same workload:
    per step 512 cells of the cyclone matrix.  Loads nothing of the product (oracle/ and tests/golden only); the
    thread count is explicit because torch.distributed.run exports OMP_NUM_THREADS=1."""
    rank, _, world = env_world()
    if rank != 0:
        return
    n_threads = os.cpu_count() or 1
    step, port, info = cpu_cyclone_runner(n_threads)
    cells = 512
    rng = np.random.default_rng(20262)
    for _ in range(max(args.warmup, 1)):  # the first call also pays Numba's lazy thread-pool start
        step(rng.poisson(2.0, (64, CYC_GENES)).astype(np.float32))
    xs_all = [rng.poisson(2.0, (cells, CYC_GENES)).astype(np.float32) for _ in range(args.steps)]
    t0 = time.perf_counter()
    for xs in xs_all:
        out = step(xs)
    dt = time.perf_counter() - t0
    v = cells * args.steps / dt
    if info["kind"] == "reference":
        from oracle import ref_numba
        info["threading_layer"] = ref_numba.threading_layer()
        t1 = time.perf_counter()  # the C port beside it, one step, and the two must agree bit for bit
        chk = port(xs_all[-1])
        info["port_value"] = cells / (time.perf_counter() - t1)
        assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(out, chk)), "Numba reference != C port"
    sample = "{} cells per step of the 100000 x 20000 matrix, all 3 classes, 1000 iterations".format(cells)
    base = dict(info, value=v, unit=UNIT, sample=sample)
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "value": v, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: cyclone default_cc_marker, synthetic 100000 cells x 20000 genes, "
                                   "1000 iterations (bounded sample: {})".format(sample)},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workloads", default="all", help="comma list of {} | all | none".format(",".join(WORKLOADS)))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--quick", action="store_true", help="skip the pageable-input end-to-end legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    args.workloads = WORKLOADS if args.workloads == "all" else [] if args.workloads == "none" else \
        [w for w in args.workloads.split(",") if w]
    bad = [w for w in args.workloads if w not in WORKLOADS]
    if bad:
        ap.error("unknown workloads: {}".format(bad))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
