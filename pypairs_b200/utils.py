"""Host-side input normalisation with the semantics of pypairs/utils.py:207-385.

Everything here is cheap index / mask logic; the expression matrix itself is never copied
or filtered on the host -- the kernels gather the kept rows / columns on the device.
"""
import json
import os

import numpy as np
from pandas import DataFrame
from scipy.sparse import issparse

try:  # anndata is optional; AnnData inputs are also recognised structurally
    from anndata import AnnData as _AnnData
except ImportError:  # pragma: no cover
    _AnnData = None


try:  # optional C helper built by __graft_entry__.build(); pure-Python equivalents below
    from ._pytuples import pair_list as _c_pair_list, pair_lists as _c_pair_lists
except ImportError:  # pragma: no cover
    _c_pair_list = _c_pair_lists = None
try:
    from ._pytuples import pair_indices as _c_pair_indices
except ImportError:  # pragma: no cover
    _c_pair_indices = None


def pair_indices(pairs, name_to_idx):
    """(int64[n], int64[n]) = name_to_idx.get(p[0], -1), name_to_idx.get(p[1], -1) for every pair p."""
    if _c_pair_indices is not None:
        a, b = _c_pair_indices(pairs, name_to_idx)
        return np.frombuffer(a, dtype=np.int64), np.frombuffer(b, dtype=np.int64)
    return (np.fromiter((name_to_idx.get(p[0], -1) for p in pairs), dtype=np.int64, count=len(pairs)),
            np.fromiter((name_to_idx.get(p[1], -1) for p in pairs), dtype=np.int64, count=len(pairs)))


def first_positions(cls, n_classes):
    """Position of the first entry of every class in `cls` (len(cls) for a class that does not occur)."""
    cls = np.asarray(cls)
    first = np.full(n_classes, len(cls), dtype=np.int64)
    if n_classes <= 16:  # a handful of vectorised scans
        for k in range(n_classes):
            hit = cls == k
            i = int(hit.argmax()) if len(cls) else 0
            if len(cls) and hit[i]:
                first[k] = i
    else:
        np.minimum.at(first, cls, np.arange(len(cls), dtype=np.int64))
    return first


def pair_lists(names, rows, cols, cls, n_classes):
    """Per class k the list of (names[rows[i]], names[cols[i]]) over the entries with cls[i] == k, input order kept."""
    names = names if isinstance(names, list) else list(names)
    if _c_pair_lists is not None:
        return _c_pair_lists(names, np.ascontiguousarray(rows, dtype=np.int64), np.ascontiguousarray(cols, dtype=np.int64),
                             np.ascontiguousarray(cls, dtype=np.int32), int(n_classes))
    obj = np.empty(len(names), dtype=object)
    obj[:] = names
    cls = np.asarray(cls)
    return [list(zip(obj[rows[cls == k]].tolist(), obj[cols[cls == k]].tolist())) for k in range(n_classes)]


def pair_list(names, rows, cols):
    """[(names[rows[i]], names[cols[i]]) ...] -- `names` is an object array / list of the gene-name scalars."""
    if _c_pair_list is not None:
        return _c_pair_list(names if isinstance(names, list) else names.tolist(),
                            np.ascontiguousarray(rows, dtype=np.int64), np.ascontiguousarray(cols, dtype=np.int64))
    names = np.asarray(names, dtype=object)
    return list(zip(names[rows].tolist(), names[cols].tolist()))


def is_anndata(data):
    if _AnnData is not None and isinstance(data, _AnnData):
        return True
    return all(hasattr(data, a) for a in ("X", "obs", "obs_names", "var_names"))


def is_cuda_tensor(data):
    return type(data).__module__.startswith("torch") and getattr(data, "is_cuda", False)


def _matrix(x):
    """Sparse input stays sparse: the reference calls .toarray() here (pypairs/utils.py:278-281, 290-292,
    299-301); the CUDA library takes CSR and densifies chunk by chunk on the device instead."""
    return x.tocsr() if issparse(x) else x


def parse_data(data, gene_names=None, sample_names=None):
    """-> (raw_data [samples, genes], gene_names, sample_names); pypairs/utils.py:264-308.

    Besides the reference's AnnData / DataFrame / ndarray / scipy-sparse inputs a CUDA torch
    tensor is accepted (names required), so that matrices larger than host memory can be scored.
    """
    if is_anndata(data):
        if sample_names is None:
            sample_names = list(data.obs_names)
        if gene_names is None:
            gene_names = list(data.var_names)
        raw = _matrix(data.X)
    elif isinstance(data, DataFrame):
        if gene_names is None:
            gene_names = list(data.columns)
        if sample_names is None:
            sample_names = list(data.index)
        raw = _matrix(data.values)
    elif isinstance(data, np.ndarray) or issparse(data) or is_cuda_tensor(data):
        if gene_names is None or sample_names is None:
            raise ValueError("Provide gene names and sample names in ``gene_names`` and ``sample_names``")
        raw = _matrix(data)
    else:
        raise ValueError("data can only be of type AnnData, DataFrame or ndarray")
    if len(raw.shape) != 2:
        raise ValueError("data must be 2-dimensional (samples x genes)")
    if raw.shape[0] != len(sample_names) or raw.shape[1] != len(gene_names):
        raise ValueError("data of shape {} x {} does not match {} sample names x {} gene names".format(
            raw.shape[0], raw.shape[1], len(sample_names), len(gene_names)))
    return raw, gene_names, sample_names


def to_boolean_mask(selected, names):
    """Index / name / boolean selection -> boolean mask (pypairs/utils.py:311-339)."""
    n = len(names)
    if selected is None:
        return np.ones(n, dtype=bool)
    selected = np.array(selected)
    if selected.size == 0:
        return np.ones(n, dtype=bool)
    if selected.dtype == bool:
        if selected.shape != (n,):
            raise ValueError("boolean selection has length {} but {} entries were expected".format(len(selected), n))
        return selected
    mask = np.zeros(n, dtype=bool)
    if np.issubdtype(selected.dtype, np.integer):
        mask[selected] = True
    elif selected.dtype.type is np.str_:
        mask = np.isin(np.asarray(names, dtype=str), selected)
    else:
        raise ValueError("Categories must be array-like of type bool, int or str")
    return mask


def parse_annotation(annotation, sample_names):
    """{'class': [samples...]} -> (class names in dict order, bool[C, N]); pypairs/utils.py:243-261."""
    names = np.array(list(annotation.keys()))
    cats = np.zeros((len(names), len(sample_names)), dtype=bool)
    for i, k in enumerate(annotation.keys()):
        cats[i] = to_boolean_mask(np.array(annotation[k]), sample_names)
    return names, cats


def parse_data_and_annotation(data, annotation=None, gene_names=None, sample_names=None):
    """pypairs/utils.py:207-240: classes from `annotation` (dict order) or, for AnnData without
    annotation, from obs['category'] (np.unique -> sorted)."""
    raw, gene_names, sample_names = parse_data(data, gene_names, sample_names)
    if annotation:
        names, cats = parse_annotation(annotation, sample_names)
    elif is_anndata(data):
        keys = data.obs_keys() if hasattr(data, "obs_keys") else list(data.obs.keys())
        if "category" not in keys:
            raise ValueError("Provide categories as data.var['category'] or in ``annotation``")
        col = np.asarray(data.obs["category"])
        names = np.unique(col)
        cats = np.zeros((len(names), len(sample_names)), dtype=bool)
        for i, name in enumerate(names):
            cats[i] = np.isin(col, name)
    else:
        raise ValueError("Provide categories in ``annotation``")
    return raw, gene_names, sample_names, names, cats


def expressed_gene_mask(raw):
    """~all(data == 0, axis=0) over ALL samples, before sample filtering (pypairs/utils.py:368)."""
    if is_cuda_tensor(raw):
        return (raw != 0).any(dim=0).cpu().numpy()
    if issparse(raw):
        m = raw.tocsr()
        out = np.zeros(m.shape[1], dtype=bool)
        out[m.indices[m.data != 0]] = True
        return out
    raw = np.asarray(raw)
    out = np.zeros(raw.shape[1], dtype=bool)
    step = max(1, (1 << 24) // max(1, raw.shape[1]))  # bounded temporaries on large matrices
    for i in range(0, raw.shape[0], step):
        out |= (raw[i:i + step] != 0).any(axis=0)
    return out


def get_filter_masks(raw, gene_names, sample_names, categories, filter_genes, filter_samples):
    """(gene_mask, sample_mask) of pypairs/utils.py:364-385."""
    gene_mask = np.logical_and(expressed_gene_mask(raw), to_boolean_mask(filter_genes, gene_names))
    categorized = np.any(categories, axis=0)
    sample_mask = np.logical_and(categorized, to_boolean_mask(filter_samples, sample_names))
    return gene_mask, sample_mask


def same_marker(a, b):
    """Set equality per class, the comparator of the reference's tests (pypairs/utils.py:409-423)."""
    if set(a.keys()) != set(b.keys()):
        return False
    for k in a:
        if {tuple(p) for p in a[k]} != {tuple(p) for p in b[k]}:
            return False
    return True


def export_marker(marker, fname, defaultpath=True):
    """Marker pairs -> JSON file under `settings.writedir` (or at `fname` itself with defaultpath=False).
    I/O errors are logged, not raised (pypairs/utils.py:97-130)."""
    from . import log as logg, settings
    fpath = settings.writedir + fname if defaultpath else fname
    try:
        if defaultpath and settings.writedir and not os.path.isdir(settings.writedir):
            os.makedirs(settings.writedir, exist_ok=True)
        with open(fpath, "w") as f:
            json.dump({str(k): [[str(a), str(b)] for a, b in v] for k, v in dict(marker).items()}, f)
        logg.hint("marker pairs written to: " + str(fpath))
    except IOError as e:
        logg.error("could not write to {}. Please verify that the path exists and is writable. "
                   "Or change `writedir` via `pypairs_b200.settings.writedir`".format(fpath))
        logg.error(str(e))


def load_marker(fname, defaultpath=True):
    """JSON file -> marker dict; None (and a logged error) if it cannot be read (pypairs/utils.py:133-166)."""
    from . import log as logg, settings
    fpath = settings.writedir + fname if defaultpath else fname
    try:
        with open(fpath) as f:
            marker = json.load(f)
    except IOError:
        logg.error("could not read from {}.\n Please verify that the path exists and is writable.".format(fpath))
        return None
    logg.hint("loaded {} marker pairs".format(sum(len(p) for p in marker.values())))
    return marker


def evaluate_prediction(prediction, reference):
    """F1 score, recall and precision of a cyclone prediction (pypairs/utils.py:23-94).

    Returns a DataFrame indexed by the sorted union of labels plus "average", with columns "f1", "recall",
    "precision", "average" (the row-wise mean of the three).  Same numbers as the reference's sklearn calls
    (per-label scores with `labels=np.unique(ref + pred)`, macro average, 0 where a denominator is 0).
    """
    ref = np.array(reference)
    pred = np.array(prediction)
    labels = np.unique(list(ref) + list(pred))
    tp = np.array([np.sum((ref == k) & (pred == k)) for k in labels], dtype=np.float64)
    n_ref = np.array([np.sum(ref == k) for k in labels], dtype=np.float64)
    n_pred = np.array([np.sum(pred == k) for k in labels], dtype=np.float64)

    def ratio(a, b):
        return np.divide(a, b, out=np.zeros_like(a), where=b != 0)

    recall = ratio(tp, n_ref)
    precision = ratio(tp, n_pred)
    f1 = ratio(2.0 * tp, n_ref + n_pred)
    rows = {"f1": np.append(f1, f1.mean()), "recall": np.append(recall, recall.mean()),
            "precision": np.append(precision, precision.mean())}
    df = DataFrame(rows, index=np.append(labels, "average"), columns=["f1", "recall", "precision"])
    df["average"] = df[["f1", "recall", "precision"]].mean(axis=1)
    return df


class MarkerTable(object):
    """Columnar form of a sandbag result: index arrays instead of hundreds of thousands of Python tuples.

    `rows[i]`, `cols[i]` index `gene_names`, `cls[i]` indexes `class_names`; entries are in the reference's
    scan order (row-major over the kept-gene matrix).  `to_dict()` gives exactly what pairs.sandbag returns;
    pairs.cyclone accepts a MarkerTable directly and maps the genes by name without building tuples.
    """

    def __init__(self, rows, cols, cls, gene_names, class_names):
        self.rows = np.asarray(rows, dtype=np.int64)
        self.cols = np.asarray(cols, dtype=np.int64)
        self.cls = np.asarray(cls, dtype=np.int64)
        self.gene_names = np.asarray(gene_names)
        self.class_names = np.asarray(class_names)

    def __len__(self):
        return len(self.rows)

    def class_order(self):
        """Class indices in order of first appearance (the reference's dict key order)."""
        first = first_positions(self.cls, len(self.class_names))
        return [int(k) for k in np.argsort(first, kind="stable") if first[k] < len(self.cls)]

    def keys(self):
        return [self.class_names[k] for k in self.class_order()]

    def counts(self):
        return {self.class_names[k]: int(np.sum(self.cls == k)) for k in self.class_order()}

    def to_dict(self):
        per_class = pair_lists(list(self.gene_names), self.rows, self.cols, self.cls, len(self.class_names))
        return {self.class_names[k]: per_class[k] for k in self.class_order()}
