"""pairs.cyclone -- permutation-null scoring; host side of pypairs/tools/cyclone.py:17-194.

The host translates marker gene names to indices (filter_marker_pairs, cyclone.py:273-311) and
builds the result DataFrame (cyclone.py:169-191); scoring of every cell against every class
(get_phase_scores, cyclone.py:223-257) runs in the CUDA library behind pp_cyclone().
"""
import numpy as np
from pandas import DataFrame

from . import _dist, _ffi, datasets, log as logg, settings, utils


def filter_marker_pairs(marker_pairs, gene_names):
    """Per class (dict order): ascending used gene columns and the pairs re-indexed into that
    compacted vector, keeping input order and duplicates, dropping pairs with an unknown gene;
    with duplicated gene names the last column wins (cyclone.py:273-311).

    Returns (pairs: dict class -> int32[P, 2], used: dict class -> bool[G]).
    """
    name_to_idx = {g: i for i, g in enumerate(gene_names)}
    n = len(gene_names)
    out_pairs, out_used = {}, {}
    removed = 0
    if isinstance(marker_pairs, utils.MarkerTable):  # columnar markers: map each distinct gene once
        t = marker_pairs
        g2i = np.fromiter((name_to_idx.get(g, -1) for g in t.gene_names.tolist()), dtype=np.int64,
                          count=len(t.gene_names))
        per_class = [(t.class_names[k], g2i[t.rows[t.cls == k]], g2i[t.cols[t.cls == k]]) for k in t.class_order()]
    else:
        per_class = []
        for cat, pairs in marker_pairs.items():
            per_class.append((cat,) + utils.pair_indices(pairs, name_to_idx))
    for cat, ia, ib in per_class:
        ok = (ia >= 0) & (ib >= 0)
        removed += int((~ok).sum())
        ia, ib = ia[ok], ib[ok]
        used = np.zeros(n, dtype=bool)
        used[ia] = True
        used[ib] = True
        new_idx = np.cumsum(used) - 1
        out_pairs[cat] = np.stack([new_idx[ia], new_idx[ib]], axis=1).astype(np.int32).reshape(-1, 2)
        out_used[cat] = used
    logg.hint("translated marker pairs, {} removed".format(removed))
    return out_pairs, out_used


def _label_column(codes, labels):
    """Column of labels[codes] (code -1 = missing) in the dtype pandas infers for a list of str.  With an Arrow-backed
    default string dtype (pandas >= 3) the column is decoded from a dictionary array in C++ instead of converting
    one Python str per row."""
    import pandas as pd
    dtype = pd.Series(["a"]).dtype
    if isinstance(dtype, pd.StringDtype) and getattr(dtype, "storage", None) == "pyarrow":
        try:
            import pyarrow as pa
            idx = pa.array(codes.astype(np.int32), mask=(codes < 0))
            arr = pa.DictionaryArray.from_arrays(idx, pa.array(list(labels), type=pa.large_string())).dictionary_decode()
            return pd.array(arr, dtype=dtype)
        except Exception:  # pragma: no cover - fall back to the generic conversion
            pass
    col = np.array(list(labels) + [np.nan], dtype=object)[codes]
    return col.tolist()


def _names_index(sample_names):
    """Index of the result frame.  A list of str goes through an Arrow array when pandas' inferred string dtype is
    Arrow-backed: same Index as `pd.Index(list)`, built in a third of the time."""
    import pandas as pd
    if isinstance(sample_names, pd.Index):
        return sample_names
    if isinstance(sample_names, (list, tuple)) and len(sample_names) > 1000 and all(
            type(sample_names[i]) is str for i in (0, len(sample_names) // 2, -1)):
        dtype = pd.Series(["a"]).dtype
        if isinstance(dtype, pd.StringDtype) and getattr(dtype, "storage", None) == "pyarrow":
            try:
                import pyarrow as pa
                return pd.Index(pd.array(pa.array(sample_names, type=pa.large_string()), dtype=dtype))
            except Exception:  # a non-str element somewhere: the generic conversion decides
                pass
    return pd.Index(sample_names)


def scores_to_frame(scores, keys, sample_names):
    """DataFrame epilogue of cyclone.py:169-181 with positional semantics.

    max_class = first class with the largest score, 'N/A' unless the row sum is > 0;
    cc_prediction (only for exactly {G1, S, G2M}) = 'S' if max(G1, G2M) < 0.5 else argmax(G1, G2M).
    The label codes come from one pass in the library (pp_score_labels).
    """
    keys = list(keys)
    vals = np.ascontiguousarray(scores, dtype=np.float64).reshape(len(keys), len(sample_names))  # [class, cell]
    df = DataFrame({k: vals[i] for i, k in enumerate(keys)}, columns=keys)
    df.index = _names_index(sample_names)
    cc3 = len(keys) == 3 and all(e in keys for e in ["G1", "S", "G2M"])
    if len(keys):
        mc, cc = _ffi.score_labels(vals, keys.index("G1") if cc3 else -1, keys.index("G2M") if cc3 else -1)
        df["max_class"] = _label_column(mc, keys + ["N/A"])
    else:
        df["max_class"] = []
    if cc3:
        df["cc_prediction"] = _label_column(cc, ["G1", "G2M", "S"])
    return df


def _quantile_transform(raw, local_shard):
    """sklearn QuantileTransformer().fit_transform on the host, as cyclone.py:138-139 (a pass-through, not part of
    the CUDA path).  The transform is fitted on ALL cells, so it needs the whole matrix on the host of every rank:
    it is refused for `local_shard` (a shard-wise fit would make the scores depend on the sharding) and for CUDA
    tensors.  Matrices of more than 10 000 cells are sub-sampled by sklearn; the sub-sample is seeded with
    `settings.seed` so that all ranks (and repeated calls) fit the same quantiles -- the reference's is unseeded."""
    if local_shard:
        raise ValueError("quantile_transform=True needs the whole matrix: not available with local_shard=True")
    if utils.is_cuda_tensor(raw):
        raise ValueError("quantile_transform=True runs sklearn on the host: pass a host matrix, not a CUDA tensor")
    from sklearn.preprocessing import QuantileTransformer
    dense = raw.toarray() if utils.issparse(raw) else raw
    return QuantileTransformer(random_state=int(settings.seed) % (2 ** 32)).fit_transform(
        np.asarray(dense, dtype=np.float64))


def cyclone(data, marker_pairs=None, gene_names=None, sample_names=None, iterations=1000, min_iter=100,
            min_pairs=50, quantile_transform=False, *, local_shard=False, gather_to=None):
    """Score each sample for each category of marker pairs.

    Same signature, defaults and return value as pypairs.pairs.cyclone
    (pypairs/tools/cyclone.py:17-26): a DataFrame indexed by sample with one score column per
    category plus 'max_class' and, for the categories {G1, S, G2M}, 'cc_prediction'; for AnnData
    input the columns are also written to ``data.obs['pypairs_<column>']``.

    `local_shard=True` (keyword-only extension, torch.distributed only): `data` holds just this
    rank's cells -- for matrices that do not fit one host -- and the frame covers all ranks' cells
    in rank order (`sample_names` may name this rank's cells or, to skip the exchange of names, all cells).
    `gather_to=r` (keyword-only extension): only rank r receives the scores and builds the frame, the
    other ranks return None.

    `settings.rng` selects the permutation stream ('mt19937' = bit-exact with the reference,
    'philox' = device-generated).  Under torch.distributed the cells are sharded over the ranks, the score
    columns are all-gathered on the device by the library (pp_cyclone_dist) and every rank returns the
    complete frame.
    """
    logg.info("predicting category scores with cyclone", r=True)
    if marker_pairs is None:
        logg.hint("no marker pairs passed, using default cell cycle prediction marker")
        marker_pairs = datasets.default_cc_marker()
    given_names = None
    if local_shard and sample_names is not None and hasattr(data, "shape"):
        # the names of this rank's cells or of ALL ranks' cells: decided below, by all ranks together
        given_names, sample_names = sample_names, range(data.shape[0])
    raw, gene_names, sample_names = utils.parse_data(data, gene_names, sample_names)
    pairs, used = filter_marker_pairs(marker_pairs, gene_names)
    total = sum(len(p) for p in pairs.values())
    for k, p in pairs.items():
        if len(p) == 0:
            logg.error("no genes present for marker pairs for category {}".format(k))
    if total == 0:  # the reference logs and calls exit(1) here (cyclone.py:122-124)
        logg.error("no genes present in marker pairs. aborting.")
        raise SystemExit(1)
    empty = [k for k, p in pairs.items() if len(p) == 0]
    if empty:  # the reference's gufunc rejects the empty pair array of such a class
        raise ValueError("no genes present for marker pairs for category {}".format(empty[0]))
    logg.hint("received {} marker pairs".format(total))
    if quantile_transform:
        raw = _quantile_transform(raw, local_shard)

    keys = list(pairs.keys())
    rank, ws = _dist.world()
    rng = settings.rng if settings.fix_seed else "philox"
    seed = settings.seed if settings.fix_seed else int(np.random.SeedSequence().generate_state(1, np.uint64)[0])
    used_idx, pair_idx = [np.where(used[k])[0] for k in keys], [pairs[k] for k in keys]
    if ws > 1:
        _dist.ensure_comm()
        if not settings.fix_seed:  # all shards must share one stream
            seed = int(_dist.broadcast_object(seed, 0))
        names_global = False
        if local_shard:
            begin, end = 0, raw.shape[0]
            info = _dist.allgather_ints([raw.shape[0], len(given_names) if given_names is not None else -1])
            shard_cells = np.ascontiguousarray(info[:, 0])
            if given_names is not None:  # one decision for all ranks (a rank holding every cell cannot tell on its own)
                names_global = bool(np.all(info[:, 1] == shard_cells.sum()))
                if not names_global and not np.all(info[:, 1] == info[:, 0]):
                    raise ValueError("local_shard: pass the sample names of this rank's cells, or of all cells, on every rank "
                                     "(cells {}, names {})".format(info[:, 0].tolist(), info[:, 1].tolist()))
        else:
            begin, end = _dist.cell_shard(raw.shape[0], rank, ws)
            shard_cells = _dist.shard_sizes(raw.shape[0], ws)
        scores, tm = _ffi.cyclone(raw, used_idx, pair_idx, iterations, min_iter, min_pairs, rng=rng, seed=seed,
                                  row_begin=begin, n_cells=end - begin, shard_cells=shard_cells,
                                  gather_to=-1 if gather_to is None else int(gather_to))
        if given_names is not None:
            sample_names = given_names if names_global else _dist.allgather_names(given_names)
        elif local_shard:
            sample_names = _dist.allgather_names(sample_names)
        if scores is None:  # not the receiving rank
            logg.info("finished", time_=True)
            cyclone.last_timings = tm
            return None
    else:
        if given_names is not None:
            if len(given_names) != raw.shape[0]:
                raise ValueError("{} sample names for {} cells".format(len(given_names), raw.shape[0]))
            sample_names = given_names
        scores, tm = _ffi.cyclone(raw, used_idx, pair_idx, iterations, min_iter, min_pairs, rng=rng, seed=seed)
    df = scores_to_frame(scores, keys, sample_names)
    if utils.is_anndata(data):
        logg.hint('adding scores with key "pypairs_{category}" to `data.obs`"')
        for name in df.columns:
            data.obs["pypairs_{}".format(name)] = df[name].values
    logg.info("finished", time_=True)
    cyclone.last_timings = tm
    return df


cyclone.last_timings = None
