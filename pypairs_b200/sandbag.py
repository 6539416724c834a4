"""pairs.sandbag -- marker-pair discovery; host side of pypairs/tools/sandbag.py:15-157.

The host only does mask / index arithmetic (gene and sample filtering, class sorting of cells,
thresholds); the all-gene-pairs evaluation (check_pairs, sandbag.py:160-199) and the scan of the
result matrix (:129) run in the CUDA library behind pp_sandbag().
"""
from math import ceil

import numpy as np

from . import _dist, _ffi, log as logg, settings, utils


def calc_thresholds(class_sizes, fraction):
    """ceil(n_k * fraction) with the reference's float arithmetic (sandbag.py:212-217)."""
    return np.array([ceil(int(n) * fraction) for n in class_sizes], dtype=np.int64)


def prepare(data, annotation=None, gene_names=None, sample_names=None, fraction=0.65, filter_genes=None,
            filter_samples=None):
    """Everything sandbag.py:98-122 derives before calling the kernel, as index arrays.

    Returns a dict with: raw (matrix as given), all gene names, class names (non-empty, in order),
    gene_cols (candidate gene columns after `filter_genes`, ascending -- the all-zero-gene test of
    utils.py:368 needs a pass over the matrix and is done on the device), cell_rows (kept samples,
    class-sorted), class_sizes, thresholds.
    """
    raw, gene_names, sample_names, class_names, cats = utils.parse_data_and_annotation(
        data, annotation, gene_names, sample_names)
    gene_mask = utils.to_boolean_mask(filter_genes, gene_names)
    sample_mask = np.logical_and(np.any(cats, axis=0), utils.to_boolean_mask(filter_samples, sample_names))
    kept_gene_names = np.array(gene_names)
    cats = cats[:, sample_mask]
    rows_kept = np.where(sample_mask)[0]
    # remove_empty_categories (sandbag.py:202-209)
    nonempty = cats.sum(axis=1) > 0
    cats = cats[nonempty]
    class_names = np.array(class_names)[nonempty]
    if cats.shape[0] == 0:
        raise ValueError("no samples left in any category after filtering")
    # the reference builds `cats` with np.where(categories.T)[1] (sandbag.py:120); a sample that sits
    # in two classes silently misaligns that vector with the data rows -- refuse instead
    if np.any(cats.sum(axis=0) != 1):
        raise ValueError("every kept sample must belong to exactly one category")
    class_sizes = cats.sum(axis=1).astype(np.int64)
    thresholds = calc_thresholds(class_sizes, fraction)
    cell_rows = np.concatenate([rows_kept[cats[k]] for k in range(cats.shape[0])]).astype(np.int32)
    return {"raw": raw, "gene_names": kept_gene_names, "class_names": class_names,
            "gene_cols": np.where(gene_mask)[0].astype(np.int32), "cell_rows": cell_rows,
            "class_sizes": class_sizes, "thresholds": thresholds}


def markers_to_dict(keys, cls, n_genes, gene_names, class_names):
    """Ascending keys == row-major np.where order (sandbag.py:129-140): dict key order is the order
    of first appearance, tuples are (gene_names[row], gene_names[col]) as numpy strings."""
    if len(keys) == 0:
        return {}
    k64 = np.ascontiguousarray(keys).view(np.int64)  # keys < 2^32: the signed view is exact and divides faster
    rows, cols = k64 // n_genes, k64 % n_genes
    names = list(gene_names)  # the np.str_ scalars `gene_names[g]` yields, made once
    # classes in order of first appearance in the scan: smallest position of every class
    first = utils.first_positions(cls, len(class_names))
    per_class = utils.pair_lists(names, rows, cols, cls, len(class_names))
    out = {}
    for k in np.argsort(first, kind="stable"):
        if first[k] == len(cls):
            break
        out[class_names[k]] = per_class[k]
    return out


def sandbag(data, annotation=None, gene_names=None, sample_names=None, fraction=0.65, filter_genes=None,
            filter_samples=None, *, as_table=False, gather_to=None):
    """Calculate the pairs of genes serving as marker pairs for each category.

    Same signature, defaults and return value as pypairs.pairs.sandbag
    (pypairs/tools/sandbag.py:15-23): a dict mapping each category that has markers to a list of
    (gene_high, gene_low) tuples in the reference's order.  `data` may be an AnnData (classes from
    ``obs['category']`` unless `annotation` is given), a DataFrame, an ndarray / scipy-sparse matrix
    with `gene_names` and `sample_names`, or a CUDA torch tensor.

    `as_table=True` (keyword-only extension) returns a `utils.MarkerTable` -- the same markers as index arrays,
    without building Python tuples (`.to_dict()` gives the dict); pairs.cyclone accepts it directly.

    Under torch.distributed (one process per GPU) every rank passes the same arguments; the library splits the
    prep over the cells and the gene-pair tiles over the ranks and merges the marker lists with NCCL
    (pp_sandbag_dist).  Every rank returns the complete dict, or -- `gather_to=r` (keyword-only extension) -- only
    rank r does and the others return None without building the Python result.
    """
    logg.info("identifying marker pairs with sandbag", r=True)
    logg.hint("sandbag running with fraction of {}".format(fraction))
    p = prepare(data, annotation, gene_names, sample_names, fraction, filter_genes, filter_samples)
    rank, ws = _dist.world()
    if ws > 1:
        _dist.ensure_comm()
        keys, cls, tm, kept = _ffi.sandbag(p["raw"], p["cell_rows"], p["class_sizes"], p["gene_cols"], p["thresholds"],
                                           drop_unexpressed=True, dist=True,
                                           gather_to=-1 if gather_to is None else int(gather_to))
    else:
        keys, cls, tm, kept = _ffi.sandbag(p["raw"], p["cell_rows"], p["class_sizes"], p["gene_cols"], p["thresholds"],
                                           drop_unexpressed=True)
    sandbag.last_timings = tm
    if ws > 1 and gather_to is not None and rank != int(gather_to):
        logg.info("finished", time_=True)
        return None
    if as_table:
        k64 = np.ascontiguousarray(keys).view(np.int64)
        n = max(len(kept), 1)
        out = utils.MarkerTable(k64 // n, k64 % n, cls, p["gene_names"][kept], p["class_names"])
        logg.info("finished", time_=True)
        if settings.verbosity >= 4:
            logg.hint("found {} marker pairs ({})".format(len(out), out.counts()))
        return out
    out = markers_to_dict(keys, cls, len(kept), p["gene_names"][kept], p["class_names"])
    logg.info("finished", time_=True)
    logg.hint("found {} marker pairs ({})".format(len(keys), ", ".join(
        "{}: {}".format(k, len(v)) for k, v in out.items())))
    return out


sandbag.last_timings = None
