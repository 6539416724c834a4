#!/usr/bin/env python
"""bench.py -- headline benchmark of the two PyPairs hot paths on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--path both|cyclone|sandbag]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  Headline = BASELINE.json configs[1]: cyclone, default_cc_marker,
synthetic 100k cells x 20k genes (float32 Poisson counts), 1000 iterations, per GPU (weak scaling: every rank
scores its own 100k cells, as configs[4] shards 1M cells over 8 GPUs).  A step = one pass of the hot path over
that batch.  `value` = cells/s with the matrix resident in HBM; `e2e` = the same through pairs.cyclone() from a
pinned host ndarray to the result DataFrame.  The sandbag path (configs[2]: 20k genes x 10k cells, 3 classes,
fraction 0.65, gene-pair tiles sharded over the ranks = strong scaling) is reported in the same line under
"sandbag".  --impl reference times the CPU port of the reference (oracle/) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CYC_CELLS, CYC_GENES, CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS = 100000, 20000, 1000, 100, 50
SB_GENES, SB_CELLS, SB_CLASSES, SB_FRACTION = 20000, 10000, 3, 0.65
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
            names = {nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8: "hw_slowdown",
                     0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
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


def sandbag_matrix_device(seed=20263):
    """cfg3 synthetic generator: Poisson counts, 2 % DE genes, 3 classes."""
    import torch
    g, n, c = SB_GENES, SB_CELLS, SB_CLASSES
    gen = torch.Generator(device="cuda").manual_seed(seed)
    lab = torch.randint(0, c, (n,), generator=gen, device="cuda")
    base = torch.randn(g, generator=gen, device="cuda") * 1.5
    de = torch.rand(g, generator=gen, device="cuda") < 0.02
    de_class = torch.randint(0, c, (g,), generator=gen, device="cuda")
    sign = torch.where(torch.rand(g, generator=gen, device="cuda") < 0.5, -1.0, 1.0)
    eff = torch.where(de, torch.pow(2.0, 2.0 * sign), torch.ones_like(sign))
    lam = 3.0 * torch.exp(base)[None, :] * torch.where(lab[:, None] == de_class[None, :], eff[None, :], 1.0)
    x = torch.poisson(lam, generator=gen).to(torch.float32)
    x[:, 0] += 1.0
    return x, lab.cpu().numpy().astype(np.int64)


# ------------------------------------------------------------------------------------------------ b200 arm
def bench_cyclone(args, rank, local_rank, world, peaks):
    import torch
    from oracle import oracle as orc  # cpu_baseline leg only
    from pypairs_b200 import _ffi, pairs, settings
    settings.verbosity = 0
    markers, names, keys, used_idx, pair_idx, pairs_d, used_d = cyclone_setup()
    n_pairs = sum(len(p) for p in pair_idx)
    n_union = len(np.unique(np.concatenate(used_idx)))
    gen = torch.Generator(device="cuda").manual_seed(20262 + rank)
    x = torch.empty((CYC_CELLS, CYC_GENES), dtype=torch.float32, device="cuda")
    for i in range(0, CYC_CELLS, 10000):  # bounded temporaries
        x[i:i + 10000] = torch.poisson(torch.full((10000, CYC_GENES), 2.0, device="cuda"), generator=gen)
    run = lambda: _ffi.cyclone(x, used_idx, pair_idx, CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS)  # noqa: E731
    for _ in range(args.warmup):
        sc, tm = run()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sync_all(world)
    t0 = time.perf_counter()
    dev_ms, main_ms, prep_ms, launches = 0.0, 0.0, 0.0, 0
    for _ in range(args.steps):
        sc, tm = run()
        dev_ms += tm["total_ms"]
        main_ms += tm["main_ms"]
        prep_ms += tm["prep_ms"]
        launches += tm["n_launches"]
    sync_all(world)
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.result()
    ms_step = max_over_ranks(wall_ms / args.steps, world)
    dev_step = max_over_ranks(dev_ms / args.steps, world)
    main_step = max_over_ranks(main_ms / args.steps, world)
    prep_step = max_over_ranks(prep_ms / args.steps, world)

    # ---- end to end through the public API: pinned host ndarray -> DataFrame
    sn = ["cell{}".format(i) for i in range(CYC_CELLS)]
    xh_t = torch.empty((CYC_CELLS, CYC_GENES), dtype=torch.float32, pin_memory=True)
    xh_t.copy_(x)
    xh = xh_t.numpy()
    e2e_call = lambda: pairs.cyclone(xh, markers, names, sn, iterations=CYC_ITERS, min_iter=CYC_MIN_ITER,  # noqa: E731
                                     min_pairs=CYC_MIN_PAIRS, local_shard=(world > 1))
    e2e_call()
    sync_all(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        df = e2e_call()
    sync_all(world)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps, world)
    e2e_tm = pairs.cyclone.last_timings
    assert np.array_equal(df[keys[0]].values[rank * CYC_CELLS:(rank + 1) * CYC_CELLS], sc[0])  # same answer both ways

    total_cells = CYC_CELLS * world
    evals = float(CYC_CELLS) * (CYC_ITERS + 1) * n_pairs  # pair evaluations per launch (per GPU)
    alg_ops = 2.0 * evals                                   # `!=` and `>` per evaluation (cyclone.py:212-215)
    out = {
        "value": total_cells / (ms_step * 1e-3), "ms_per_step": ms_step, "device_ms_per_step": dev_step,
        "kernel_ms": {"score": main_step, "prep_and_tables": prep_step},
        "gpu_launches": launches, "clocks": clocks,
        "e2e": {"value": total_cells / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(e2e_tm["h2d_bytes"]), "d2h_bytes_per_step": len(keys) * CYC_CELLS * 8,
                "host_matrix_bytes": CYC_CELLS * CYC_GENES * 4,
                "api": "pypairs_b200.pairs.cyclone(ndarray pinned) -> DataFrame; per chunk of cells the library gathers "
                       "the {} marker-gene columns on the host (threads, pinned staging), copies them and runs rank + "
                       "score, all overlapped; the gather (16 host threads) and the kernels take about the same time".format(n_union)},
        "roofline": {"bound": "int_alu", "kernel": "score_kernel",
                     "achieved": alg_ops / (main_step * 1e-3) / 1e9, "peak": peaks["lop3_lane_ops_per_s"] / 1e9,
                     "unit": "Gop/s", "frac": alg_ops / (main_step * 1e-3) / peaks["lop3_lane_ops_per_s"],
                     "peak_source": "pp_alu_peak micro-benchmark (independent LOP3 stream), measured in this run",
                     "algorithmic_ops_per_unit": 2.0 * (CYC_ITERS + 1) * n_pairs, "units_per_launch": CYC_CELLS,
                     "note": "count data: four 7-bit ranks per 32-bit lane, ~8.8 instructions per comparison slot of "
                             "128 cells, so the algorithmic rate may exceed the one-op-per-lane peak; ncu: issue slots "
                             "88 %, ALU pipe 76 %",
                     "traffic": ncu_traffic("score_kernel")},
        "roofline_prep": {"bound": "hbm", "kernel": "rank_kernel+tables",
                          "achieved": (CYC_CELLS * CYC_GENES * 4 + CYC_CELLS * n_union * 2) / (prep_step * 1e-3) / 1e9,
                          "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": (CYC_CELLS * CYC_GENES * 4 + CYC_CELLS * n_union * 2) / (prep_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
                          "peak_source": peaks["hbm_source"]},
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        ncell = 64
        xs = x[:2048].cpu().numpy()
        t0 = time.perf_counter()
        for k in keys:
            orc.phase_scores(xs[:ncell], CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS, pairs_d[k], used_d[k])
        rate = ncell / (time.perf_counter() - t0)
        ncell = int(min(2048, max(128, rate * 15)))
        t0 = time.perf_counter()
        ref = [orc.phase_scores(xs[:ncell], CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS, pairs_d[k], used_d[k]) for k in keys]
        dt = time.perf_counter() - t0
        assert all(np.array_equal(ref[i], sc[i][:ncell]) for i in range(len(keys))), "GPU != CPU on the timed sample"
        out["cpu_baseline"] = {"value": ncell / dt, "unit": UNIT, "cores": orc.max_threads(), "kind": "port",
                               "sample": "first {} of the 100000 cells, all 3 classes, 1000 iterations; GPU scores "
                                         "bit-identical on the sample".format(ncell)}
    del x, xh_t
    torch.cuda.empty_cache()
    return out


def bench_sandbag(args, rank, local_rank, world, peaks):
    import torch
    from oracle import oracle as orc  # cpu_baseline leg only
    from pypairs_b200 import _ffi, pairs, settings
    settings.verbosity = 0
    x, lab = sandbag_matrix_device()
    g, n, c = SB_GENES, SB_CELLS, SB_CLASSES
    sizes = np.bincount(lab, minlength=c)
    thr = np.array([int(np.ceil(int(s) * SB_FRACTION)) for s in sizes], dtype=np.int64)
    order = np.argsort(lab, kind="stable")
    run = lambda: _ffi.sandbag(x, order, sizes, np.arange(g), thr, shard=rank, n_shards=world)  # noqa: E731
    for _ in range(args.warmup):
        keys, cls, tm = run()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sync_all(world)
    t0 = time.perf_counter()
    dev_ms = main_ms = prep_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        keys, cls, tm = run()
        dev_ms += tm["total_ms"]
        main_ms += tm["main_ms"]
        prep_ms += tm["prep_ms"]
        launches += tm["n_launches"]
    sync_all(world)
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.result()
    ms_step = max_over_ranks(wall_ms / args.steps, world)
    main_step = max_over_ranks(main_ms / args.steps, world)
    prep_step = max_over_ranks(prep_ms / args.steps, world)
    units = g * (g - 1) / 2.0 * n  # gene-pair x cell evaluations of the whole job
    # ---- end to end through pairs.sandbag: pinned host ndarray -> marker dict
    xh_t = torch.empty((n, g), dtype=torch.float32, pin_memory=True)
    xh_t.copy_(x)
    xh = xh_t.numpy()
    gn = np.array(["g{}".format(i) for i in range(g)])
    sn = ["c{}".format(i) for i in range(n)]
    ann = {k: np.where(lab == i)[0] for i, k in enumerate(["G1", "S", "G2M"])}
    # supplementary: the same call returning the columnar MarkerTable (no Python tuples); measured first so that the
    # cyclic GC is not busy with the millions of tuples of the dict results
    pairs.sandbag(xh, ann, gn, sn, fraction=SB_FRACTION, as_table=True)
    sync_all(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tab = pairs.sandbag(xh, ann, gn, sn, fraction=SB_FRACTION, as_table=True)
    sync_all(world)
    e2e_tab_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps, world)
    e2e_call = lambda: pairs.sandbag(xh, ann, gn, sn, fraction=SB_FRACTION)  # noqa: E731
    e2e_call()
    sync_all(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d = e2e_call()
    sync_all(world)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps, world)
    e2e_tm = pairs.sandbag.last_timings
    n_markers = sum(len(v) for v in d.values())
    assert len(tab) == n_markers
    my_pairs = tm["n_units"]
    alg_ops = 2.0 * my_pairs * n  # one `>` and one `<` per gene pair x cell (sandbag.py:179-182), this rank's launch
    out = {
        "metric": "sandbag gene-pair x cell compares/s", "unit": "pair*cell/s", "scaling": "strong",
        "config": {"workload": "configs[2]: sandbag synthetic 20k genes x 10k cells, 3 classes, fraction=0.65",
                   "dtype_in": "float32 Poisson counts", "markers_found": n_markers, "bit_planes": tm["n_planes"],
                   "bit_planes_compared_per_pair": tm["avg_planes"]},
        "value": units / (ms_step * 1e-3), "ms_per_step": ms_step, "device_ms_per_step": max_over_ranks(dev_ms / args.steps, world),
        "kernel_ms": {"pair": main_step, "rank_and_bitslice": prep_step}, "gpu_launches": launches, "clocks": clocks,
        "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "pair*cell/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(e2e_tm["h2d_bytes"]), "d2h_bytes_per_step": int(n_markers * 12 / max(world, 1)),
                "api": "pypairs_b200.pairs.sandbag(ndarray pinned) -> dict",
                "as_table": {"value": units / (e2e_tab_ms * 1e-3), "ms_per_step": e2e_tab_ms,
                             "api": "pairs.sandbag(..., as_table=True) -> utils.MarkerTable (index arrays, no tuples)"}},
        "roofline": {"bound": "int_alu", "kernel": "pair_kernel", "achieved": alg_ops / (main_step * 1e-3) / 1e9,
                     "peak": peaks["lop3_lane_ops_per_s"] / 1e9, "unit": "Gop/s",
                     "frac": alg_ops / (main_step * 1e-3) / peaks["lop3_lane_ops_per_s"],
                     "peak_source": "pp_alu_peak micro-benchmark (independent LOP3 stream), measured in this run",
                     "algorithmic_ops_per_unit": 2, "units_per_launch": my_pairs * n,
                     "machine_lop3_per_unit": tm["avg_planes"] / 32.0,
                     "traffic": ncu_traffic("pair_kernel") if world == 1 else None},
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        gs = 600
        xs = x[:, :gs].cpu().numpy().astype(np.float64)
        t0 = time.perf_counter()
        rr, cc, kk = orc.sandbag_arrays(xs, lab, c, SB_FRACTION)
        rate = gs * (gs - 1) / 2.0 * n / (time.perf_counter() - t0)
        gs = int(min(4000, max(600, np.sqrt(2.0 * rate * 15 / n))))
        xs = x[:, :gs].cpu().numpy().astype(np.float64)
        t0 = time.perf_counter()
        rr, cc, kk = orc.sandbag_arrays(xs, lab, c, SB_FRACTION)
        dt = time.perf_counter() - t0
        rows, cols = (keys // np.uint64(g)).astype(np.int64), (keys % np.uint64(g)).astype(np.int64)
        m = (rows < gs) & (cols < gs)
        assert np.array_equal(rows[m], rr) and np.array_equal(cols[m], cc) and np.array_equal(cls[m], kk), \
            "GPU != CPU on the timed sample"
        out["cpu_baseline"] = {"value": gs * (gs - 1) / 2.0 * n / dt, "unit": "pair*cell/s", "cores": orc.max_threads(),
                               "kind": "port", "sample": "all 10000 cells x first {} genes ({} markers, GPU markers "
                                                         "identical on the sample)".format(gs, len(rr))}
    del x, xh_t
    torch.cuda.empty_cache()
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
    from pypairs_b200 import _ffi
    lop3, n_sm = _ffi.alu_peak(5)
    peaks = {"lop3_lane_ops_per_s": lop3, "n_sm": n_sm, "hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)"}
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        peaks["hbm_gbs"] = json.load(open(mp))["hbm_gbs"]
        peaks["hbm_source"] = "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    line = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16 dense ranks (from f32), compared as packed 7-bit integers (count data) or fp16 pairs",
            "data": "synthetic",
            "config": {"workload": "configs[1]: cyclone default_cc_marker, synthetic {} cells x {} genes per GPU, "
                                   "{} iterations".format(CYC_CELLS, CYC_GENES, CYC_ITERS),
                       "cells_per_gpu": CYC_CELLS, "genes": CYC_GENES, "iterations": CYC_ITERS, "rng": "mt19937 replay",
                       "parallelism": "cells sharded, {} rank(s)".format(world),
                       "l2": "inputs (8 GB matrix, 33 MB pair tables re-read per cell tile) exceed the 126 MB L2"},
            "peaks": peaks}
    if args.path in ("both", "cyclone"):
        line.update(bench_cyclone(args, rank, local_rank, world, peaks))
    if args.path in ("both", "sandbag"):
        sb = bench_sandbag(args, rank, local_rank, world, peaks)
        if args.path == "sandbag":
            line.update(sb)
        else:
            line["sandbag"] = sb
            line["gpu_launches"] += sb["gpu_launches"]
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)


# ------------------------------------------------------------------------------------------------ reference arm
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
    """(step(xs) -> [scores per class], description dict): the reference's own Numba gufunc from oracle/_ref when that
    copy is present (kind "reference"), else the C/OpenMP restatement (kind "port"); both with `n_threads` threads."""
    from oracle import oracle as orc, ref_numba
    keys, pairs, used = reference_setup()
    info = {"cores": n_threads}
    if ref_numba.available():
        ref_numba.configure_threads(n_threads)
        r = ref_numba.load()
        step = lambda xs: [ref_numba.phase_scores(xs, CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS, pairs[k], used[k])  # noqa: E731
                           for k in keys]
        info.update(kind="reference", cores=int(r.threads), numba=r.numba.__version__,
                    impl="pypairs.tools.cyclone.get_phase_scores (Numba gufunc, target=parallel) from oracle/_ref")
    else:
        orc.build()
        step = lambda xs: [orc.phase_scores(xs, CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS, pairs[k], used[k],  # noqa: E731
                                            n_threads=n_threads) for k in keys]
        info.update(kind="port", impl="oracle/pp_oracle.c (C + OpenMP restatement; oracle/_ref absent)")
    port = lambda xs: [orc.phase_scores(xs, CYC_ITERS, CYC_MIN_ITER, CYC_MIN_PAIRS, pairs[k], used[k],  # noqa: E731
                                        n_threads=n_threads) for k in keys]
    return step, port, info


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
    ap.add_argument("--path", default="both", choices=["both", "cyclone", "sandbag"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
