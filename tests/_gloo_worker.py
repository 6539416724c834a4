"""Worker of tests/test_host.py::test_two_rank_gloo_merge (one of two gloo ranks on CPU).

The CUDA entry points are replaced by the CPU oracle HERE, in the test, to exercise the host-side sharding logic
of pairs.sandbag / pairs.cyclone without a GPU.  The collectives the library issues with NCCL inside
pp_sandbag_dist / pp_cyclone_dist are emulated with gloo object gathers (same contract: gather_to, shard_cells)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist  # noqa: E402

from oracle import oracle as orc  # noqa: E402
from pypairs_b200 import _dist, _ffi, pairs, settings  # noqa: E402

settings.verbosity = 0
calls = {"sandbag": [], "cyclone": []}


def fake_sandbag(x, cell_rows, class_sizes, gene_cols, thresholds, shard=0, n_shards=1, probe_pairs=None, **kw):
    is_dist, gather_to = bool(kw.get("dist")), kw.get("gather_to", -1)
    if is_dist:
        shard, n_shards = dist.get_rank(), dist.get_world_size()
    calls["sandbag"].append((shard, n_shards, is_dist, gather_to))
    x = np.asarray(x.toarray() if hasattr(x, "toarray") else x, dtype=np.float64)
    if np.isnan(x).any():
        raise _ffi.PyPairsB200Error(_ffi.PP_ENAN, "NaN in the expression matrix")
    gene_cols = np.asarray(gene_cols)
    if kw.get("drop_unexpressed"):
        gene_cols = gene_cols[(x != 0).any(axis=0)[gene_cols]]
    cats = np.concatenate([np.full(n, k) for k, n in enumerate(class_sizes)])
    xs = x[np.asarray(cell_rows)][:, gene_cols]
    res = orc.check_pairs(xs, cats, np.asarray(thresholds))
    r, c = np.where(res != -1)
    keys = (r * len(gene_cols) + c).astype(np.uint64)
    sel = ((r // 7 + c // 5) % n_shards) == shard  # some tile-like split of the pair space
    keys, cls = keys[sel], res[r, c].astype(np.int32)[sel]
    if is_dist:  # what pp_sandbag_dist does on the device: gather the compacted lists, sort
        parts = [None] * n_shards
        dist.all_gather_object(parts, (keys, cls))
        keys, cls = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
        order = np.argsort(keys, kind="stable")
        keys, cls = keys[order], cls[order]
        if gather_to >= 0 and gather_to != shard:
            keys, cls = keys[:0], cls[:0]
    out = (keys, cls, {})
    if probe_pairs is not None:
        pos, neg = orc.pair_counts(xs, cats, len(class_sizes), np.asarray(probe_pairs))
        out = out + (pos.astype(np.int32), neg.astype(np.int32))
    if kw.get("drop_unexpressed"):
        out = out + (gene_cols.astype(np.int32),)
    return out


def fake_cyclone(x, used_idx_list, pairs_list, iterations, min_iter, min_pairs, rng="mt19937", seed=2803,
                 perm_tables=None, row_begin=0, n_cells=None, shard_cells=None, gather_to=-1, want_observed=False, **kw):
    calls["cyclone"].append((row_begin, n_cells, None if shard_cells is None else list(shard_cells), gather_to))
    x = np.asarray(x.toarray() if hasattr(x, "toarray") else x, dtype=np.float64)
    if n_cells is None:
        n_cells = x.shape[0] - row_begin
    x = x[row_begin:row_begin + n_cells]
    out, oh, ot = [], [], []
    for u, p in zip(used_idx_list, pairs_list):
        used = np.zeros(x.shape[1], dtype=bool)
        used[u] = True
        if rng == "philox":  # the emulation has no Philox stream: any deterministic per-cell function will do
            sc, h, t = orc.phase_scores(x, iterations, min_iter, min_pairs, p, used, seed=int(seed) % 1000 + 1, return_observed=True)
        else:
            sc, h, t = orc.phase_scores(x, iterations, min_iter, min_pairs, p, used, seed=seed, return_observed=True)
        out.append(sc)
        oh.append(h)
        ot.append(t)
    res = [np.stack(out).reshape(len(out), n_cells), np.stack(oh).reshape(len(out), n_cells).astype(np.int32),
           np.stack(ot).reshape(len(out), n_cells).astype(np.int32)]
    if shard_cells is not None:  # what pp_cyclone_dist does on the device: all-gather of the score columns
        assert shard_cells[dist.get_rank()] == n_cells
        for i in range(3):
            parts = [None] * dist.get_world_size()
            dist.all_gather_object(parts, res[i])
            res[i] = np.concatenate(parts, axis=1) if gather_to < 0 or gather_to == dist.get_rank() else None
    if want_observed:
        return res[0], {}, res[1], res[2]
    return res[0], {}


def main():
    rng = np.random.default_rng(5)
    n, g = 61, 40
    lab = rng.integers(0, 3, n)
    x = rng.poisson(3.0 * np.exp(rng.normal(0, 1, g))[None, :] * np.where(lab[:, None] == (np.arange(g) % 3)[None, :], 3, 1)
                    ).astype(np.float64)
    x[:, 0] += 1
    sn = ["s{}".format(i) for i in range(n)]
    gn = ["g{}".format(i) for i in range(g)]
    ann = {k: np.where(lab == i)[0].tolist() for i, k in enumerate(["G1", "S", "G2M"])}
    mk = {"G1": [(gn[i], gn[(i * 3 + 1) % g]) for i in range(30)], "S": [(gn[(i * 5) % g], gn[(i + 7) % g]) for i in range(25)],
          "G2M": [(gn[(i * 7) % g], gn[(i * 2 + 3) % g]) for i in range(28)]}

    _ffi.sandbag, _ffi.cyclone = fake_sandbag, fake_cyclone
    _dist.ensure_comm = lambda device=None: _dist.world()  # no GPU here: the NCCL communicator is not created
    single_m = pairs.sandbag(x, ann, gn, sn, fraction=0.6)           # before init: world size 1
    single_s = pairs.cyclone(x, mk, gn, sn, iterations=60, min_iter=10, min_pairs=2)
    assert calls["sandbag"][-1] == (0, 1, False, -1) and calls["cyclone"][-1] == (0, None, None, -1)

    dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank = dist.get_rank()
    multi_m = pairs.sandbag(x, ann, gn, sn, fraction=0.6)
    multi_s = pairs.cyclone(x, mk, gn, sn, iterations=60, min_iter=10, min_pairs=2)
    assert calls["sandbag"][-1] == (rank, 2, True, -1)
    assert calls["cyclone"][-1] == ((0, 31, [31, 30], -1) if rank == 0 else (31, 30, [31, 30], -1))
    assert list(multi_m.items()) == list(single_m.items())           # same lists, same order, same key order
    assert sum(len(v) for v in multi_m.values()) > 0
    assert multi_s.equals(single_s)

    # gather_to: one rank builds the result, the others return None
    one_m = pairs.sandbag(x, ann, gn, sn, fraction=0.6, gather_to=1)
    one_s = pairs.cyclone(x, mk, gn, sn, iterations=60, min_iter=10, min_pairs=2, gather_to=0)
    assert (one_m is None) == (rank != 1) and (one_s is None) == (rank != 0)
    if rank == 1:
        assert list(one_m.items()) == list(single_m.items())
    if rank == 0:
        assert one_s.equals(single_s)

    # local_shard: every rank passes only its own cells; names are exchanged as bytes, or given for all cells
    lo, hi = (0, 40) if rank == 0 else (40, n)                       # deliberately unequal shards
    loc = pairs.cyclone(x[lo:hi], mk, gn, sn[lo:hi], iterations=60, min_iter=10, min_pairs=2, local_shard=True)
    assert calls["cyclone"][-1] == (0, hi - lo, [40, n - 40], -1)
    assert loc.equals(single_s) and list(loc.index) == sn
    loc2 = pairs.cyclone(x[lo:hi], mk, gn, sn, iterations=60, min_iter=10, min_pairs=2, local_shard=True)
    assert loc2.equals(single_s)
    # one rank holds every cell, the other none: "names of my cells" and "names of all cells" look the same to rank 0,
    # the ranks agree on the meaning through the exchanged counts (no rank-dependent collective)
    lo, hi = (0, n) if rank == 0 else (n, n)
    for names_arg in (sn, sn[lo:hi]):
        one = pairs.cyclone(x[lo:hi], mk, gn, names_arg, iterations=60, min_iter=10, min_pairs=2, local_shard=True)
        assert one.equals(single_s)
    try:
        pairs.cyclone(x[lo:hi], mk, gn, sn[:5], iterations=60, min_iter=10, min_pairs=2, local_shard=True)
        raise AssertionError("inconsistent sample names accepted")
    except ValueError:
        pass
    lo, hi = (0, 40) if rank == 0 else (40, n)
    uni = ["z\u00e4{}".format(i) * (1 + i % 3) for i in range(n)]     # non-ASCII, ragged widths
    got = _dist.allgather_names(uni[lo:hi])
    assert list(got) == uni
    assert list(_dist.allgather_counts(hi - lo)) == [40, n - 40]
    dist.barrier()
    dist.destroy_process_group()
    print("GLOO_OK rank", rank)


if __name__ == "__main__":
    main()
