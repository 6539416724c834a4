"""ctypes binding of libpypairs_b200.so (include/pypairs_b200.h).

There is no CPU fallback: if the CUDA library is missing or no sm_100a device is
present, every compute entry point raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpypairs_b200.so")

PP_F32, PP_F64 = 0, 1
PP_RNG_MT19937, PP_RNG_PHILOX, PP_RNG_TABLE = 0, 1, 2
PP_ENAN = 4
PP_ENCCL = 6
PP_COMM_ID_BYTES = 128

SYMBOLS = ["pp_version", "pp_device_count", "pp_ctx_create", "pp_ctx_destroy", "pp_last_error", "pp_free",
           "pp_get_timings", "pp_sandbag", "pp_sandbag_csr", "pp_cyclone", "pp_cyclone_csr", "pp_phase_scores",
           "pp_perm_table", "pp_dense_rank", "pp_alu_peak", "pp_nccl_load", "pp_nccl_version", "pp_comm_unique_id",
           "pp_comm_init", "pp_comm_destroy", "pp_comm_info", "pp_sandbag_dist", "pp_sandbag_csr_dist",
           "pp_cyclone_dist", "pp_cyclone_csr_dist", "pp_score_labels", "pp_alloc_result"]


class Csr(ctypes.Structure):
    """struct pp_csr of include/pypairs_b200.h"""
    _fields_ = [("indptr", ctypes.c_void_p), ("indptr_is_64", ctypes.c_int32), ("indices", ctypes.c_void_p),
                ("data", ctypes.c_void_p), ("data_dtype", ctypes.c_int32), ("is_device", ctypes.c_int32)]


class Timings(ctypes.Structure):
    _fields_ = [("h2d_ms", ctypes.c_float), ("prep_ms", ctypes.c_float), ("main_ms", ctypes.c_float),
                ("post_ms", ctypes.c_float), ("total_ms", ctypes.c_float), ("n_planes", ctypes.c_int32),
                ("n_launches", ctypes.c_int32), ("n_units", ctypes.c_int64), ("h2d_bytes", ctypes.c_int64),
                ("avg_planes", ctypes.c_float), ("reserved", ctypes.c_int32)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class PyPairsB200Error(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "pypairs_b200 error {}: {}".format(code, msg))
        self.code = code


_lib = None
_ctxs = {}


def load():
    """dlopen the library and declare every prototype of include/pypairs_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "pypairs_b200: {} is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.".format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i64, i32, u64 = c.c_void_p, c.c_int64, c.c_int32, c.c_uint64
    lib.pp_version.restype = c.c_int
    lib.pp_device_count.restype = c.c_int
    lib.pp_ctx_create.argtypes = [c.c_int, c.POINTER(vp)]
    lib.pp_ctx_destroy.argtypes = [vp]
    lib.pp_ctx_destroy.restype = None
    lib.pp_last_error.argtypes = [vp]
    lib.pp_last_error.restype = c.c_char_p
    lib.pp_free.argtypes = [vp]
    lib.pp_free.restype = None
    lib.pp_alloc_result.argtypes = [c.c_size_t]
    lib.pp_alloc_result.restype = vp
    lib.pp_get_timings.argtypes = [vp, c.POINTER(Timings)]
    lib.pp_sandbag.argtypes = [vp, vp, c.c_int, c.c_int, i64, i64, i64, vp, i64, vp, i32, vp, i64, vp, i32, i32,
                               c.POINTER(vp), c.POINTER(vp), c.POINTER(i64), vp, i64, vp, vp, i32, c.POINTER(vp),
                               c.POINTER(i64)]
    lib.pp_cyclone.argtypes = [vp, vp, c.c_int, c.c_int, i64, i64, i64, i64, i64, i32, vp, vp, vp, vp, i32, i32, i32,
                               i32, u64, vp, vp, vp, vp]
    lib.pp_sandbag_csr.argtypes = [vp, c.POINTER(Csr), i64, i64, vp, i64, vp, i32, vp, i64, vp, i32, i32,
                                   c.POINTER(vp), c.POINTER(vp), c.POINTER(i64), vp, i64, vp, vp, i32, c.POINTER(vp),
                                   c.POINTER(i64)]
    lib.pp_cyclone_csr.argtypes = [vp, c.POINTER(Csr), i64, i64, i64, i64, i32, vp, vp, vp, vp, i32, i32, i32, i32, u64,
                                   vp, vp, vp, vp]
    # multi-GPU (one process per GPU)
    lib.pp_nccl_load.argtypes = [c.c_char_p]
    lib.pp_nccl_version.restype = c.c_int
    lib.pp_comm_unique_id.argtypes = [vp]
    lib.pp_comm_init.argtypes = [vp, vp, i32, i32]
    lib.pp_comm_destroy.argtypes = [vp]
    lib.pp_comm_info.argtypes = [vp, c.POINTER(i32), c.POINTER(i32)]
    lib.pp_sandbag_dist.argtypes = [vp, vp, c.c_int, c.c_int, i64, i64, i64, vp, i64, vp, i32, vp, i64, vp, i32,
                                    c.POINTER(vp), c.POINTER(vp), c.POINTER(i64), vp, i64, vp, vp, i32, c.POINTER(vp),
                                    c.POINTER(i64)]
    lib.pp_sandbag_csr_dist.argtypes = [vp, c.POINTER(Csr), i64, i64, vp, i64, vp, i32, vp, i64, vp, i32,
                                        c.POINTER(vp), c.POINTER(vp), c.POINTER(i64), vp, i64, vp, vp, i32,
                                        c.POINTER(vp), c.POINTER(i64)]
    lib.pp_cyclone_dist.argtypes = [vp, vp, c.c_int, c.c_int, i64, i64, i64, i64, i64, i32, vp, vp, vp, vp, i32, i32,
                                    i32, i32, u64, vp, vp, i32, i32, vp, vp, vp]
    lib.pp_cyclone_csr_dist.argtypes = [vp, c.POINTER(Csr), i64, i64, i64, i64, i32, vp, vp, vp, vp, i32, i32, i32, i32,
                                        u64, vp, vp, i32, i32, vp, vp, vp]
    lib.pp_score_labels.argtypes = [vp, i32, i64, i32, i32, vp, vp]
    lib.pp_phase_scores.argtypes = [vp, vp, c.c_int, c.c_int, i64, i64, i64, i32, i32, i32, vp, i64, vp, vp]
    lib.pp_perm_table.argtypes = [vp, i32, u64, i32, i32, i32, vp]
    lib.pp_dense_rank.argtypes = [vp, vp, c.c_int, c.c_int, i64, i64, i64, vp, i64, vp, i64, vp, c.POINTER(i32)]
    lib.pp_alu_peak.argtypes = [vp, i32, c.POINTER(c.c_double), c.POINTER(i32)]
    _lib = lib
    return lib


def _current_device():
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except ImportError:
        pass
    return -1


def ctx(device=None):
    """One context per device per process (streams, events, memory pool)."""
    lib = load()
    if device is None:
        device = _current_device()
    if device not in _ctxs:
        h = ctypes.c_void_p()
        rc = lib.pp_ctx_create(device, ctypes.byref(h))
        if rc != 0:
            raise PyPairsB200Error(rc, lib.pp_last_error(None).decode())
        _ctxs[device] = h
    return _ctxs[device]


def check(rc, h):
    if rc != 0:
        raise PyPairsB200Error(rc, load().pp_last_error(h).decode())


def timings(h):
    t = Timings()
    check(load().pp_get_timings(h, ctypes.byref(t)), h)
    return t.as_dict()


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _order_after_torch(t):
    """Stream contract of the C ABI (include/pypairs_b200.h): the library reads device inputs on its own
    non-blocking streams, so everything torch has enqueued on the tensor's device -- the conversions made
    here and whatever the caller launched to produce the tensor -- must have completed before the call."""
    import torch
    torch.cuda.current_stream(t.device).synchronize()


def _pick_device(device, input_device):
    """Context device of a call: the device the input lives on (an explicit `device` must agree with it)."""
    if input_device is None:
        return device
    if device is not None and int(device) != int(input_device):
        raise ValueError("device={} but the input tensor lives on cuda:{}".format(device, input_device))
    return int(input_device)


def matrix_arg(x):
    """(keepalive, pointer, dtype code, is_device, n_rows, n_cols, ld, device index or None) for an ndarray or
    a CUDA torch tensor.

    float32 stays float32 (widening to the reference's float64 is exact and order preserving);
    every other dtype is converted to float64 exactly as `data.astype(float)` does
    (pypairs/tools/sandbag.py:122, cyclone.py:135).
    """
    if type(x).__module__.startswith("torch"):
        import torch
        if not x.is_cuda:
            x = x.numpy()
        else:
            if x.dtype not in (torch.float32, torch.float64):
                x = x.to(torch.float64)
            if x.dim() != 2:
                raise ValueError("expression matrix must be 2-D")
            if x.stride(1) != 1 or x.stride(0) < x.shape[1]:
                x = x.contiguous()
            code = PP_F32 if x.dtype == torch.float32 else PP_F64
            _order_after_torch(x)
            return x, ctypes.c_void_p(x.data_ptr()), code, 1, x.shape[0], x.shape[1], x.stride(0), x.device.index
    x = np.asarray(x)
    if x.ndim != 2:
        raise ValueError("expression matrix must be 2-D")
    if x.dtype != np.float32 and x.dtype != np.float64:
        x = x.astype(np.float64)
    item = x.dtype.itemsize
    if not (x.shape[1] == 0 or (x.strides[1] == item and x.strides[0] % item == 0 and x.strides[0] >= x.shape[1] * item)):
        x = np.ascontiguousarray(x)
    ld = x.strides[0] // item if x.shape[0] > 1 else x.shape[1]
    code = PP_F32 if x.dtype == np.float32 else PP_F64
    return x, ctypes.c_void_p(x.ctypes.data), code, 0, x.shape[0], x.shape[1], max(ld, x.shape[1]), None


def is_sparse(x):
    """scipy.sparse matrix / array, or a torch sparse-CSR tensor."""
    if type(x).__module__.startswith("torch"):
        import torch
        return x.layout == torch.sparse_csr
    try:
        from scipy.sparse import issparse
    except ImportError:  # pragma: no cover
        return False
    return issparse(x)


def csr_arg(x):
    """(keepalive, Csr struct, n_rows, n_cols, device index or None) for a scipy sparse matrix or a torch
    sparse-CSR tensor.

    float32 values stay float32, anything else becomes float64 (as `data.astype(float)` in the reference);
    duplicate entries are summed first (what `.toarray()` would do; a torch CSR tensor is rebuilt through COO
    coalescing when a row holds a repeated column), explicit zeros are kept.  Device-resident arrays are
    bounds-checked by the library (csr_validate_kernel).
    """
    if type(x).__module__.startswith("torch"):
        import torch
        col64, crow0 = x.col_indices(), x.crow_indices()
        if col64.numel() > 1:
            # repeated (row, column) entries: adjacent after sorting; rows that are already sorted (the usual case)
            # are checked with two element-wise passes, anything else through a sort of (row, column) keys
            inner = torch.ones(col64.numel() - 1, dtype=torch.bool, device=col64.device)
            starts = crow0[1:-1]
            starts = starts[(starts > 0) & (starts < col64.numel())]
            inner[starts - 1] = False  # pairs (k-1, k) that straddle a row boundary
            dup = bool(((col64[1:] == col64[:-1]) & inner).any())
            if not dup and bool(((col64[1:] < col64[:-1]) & inner).any()):
                row_of = torch.repeat_interleave(torch.arange(x.shape[0], device=col64.device), crow0[1:] - crow0[:-1])
                skey = torch.sort(row_of * x.shape[1] + col64).values
                dup = bool((skey[1:] == skey[:-1]).any())
            if dup:  # (to_sparse_coo() would mark the result coalesced without looking)
                row_of = torch.repeat_interleave(torch.arange(x.shape[0], device=col64.device), crow0[1:] - crow0[:-1])
                x = torch.sparse_coo_tensor(torch.stack([row_of, col64]), x.values(), x.shape).coalesce().to_sparse_csr()
                col64 = x.col_indices()
        vals, crow, col = x.values(), x.crow_indices(), col64.to(torch.int32)
        if vals.dtype not in (torch.float32, torch.float64):
            vals = vals.to(torch.float64)
        vals, crow, col = vals.contiguous(), crow.contiguous(), col.contiguous()
        on_dev = vals.is_cuda
        if not on_dev:
            vals, crow, col = vals.numpy(), crow.numpy(), col.numpy()
            ptr = lambda a: a.ctypes.data  # noqa: E731
            is64, f32 = crow.dtype == np.int64, vals.dtype == np.float32
        else:
            ptr = lambda a: a.data_ptr()  # noqa: E731
            is64, f32 = crow.dtype == torch.int64, vals.dtype == torch.float32
        st = Csr(ptr(crow), 1 if is64 else 0, ptr(col), ptr(vals), PP_F32 if f32 else PP_F64, 1 if on_dev else 0)
        if on_dev:
            _order_after_torch(vals)
        return (vals, crow, col), st, x.shape[0], x.shape[1], (vals.device.index if on_dev else None)
    m = x.tocsr()
    if not m.has_canonical_format:
        m = m.copy()
        m.sum_duplicates()
    data = m.data if m.data.dtype in (np.float32, np.float64) else m.data.astype(np.float64)
    data = np.ascontiguousarray(data)
    indices = np.ascontiguousarray(m.indices, dtype=np.int32)
    indptr = np.ascontiguousarray(m.indptr)
    if indptr.dtype not in (np.int32, np.int64):
        indptr = indptr.astype(np.int64)
    st = Csr(indptr.ctypes.data, 1 if indptr.dtype == np.int64 else 0, indices.ctypes.data, data.ctypes.data,
             PP_F32 if data.dtype == np.float32 else PP_F64, 0)
    return (data, indices, indptr), st, m.shape[0], m.shape[1], None


def sandbag(x, cell_rows, class_sizes, gene_cols, thresholds, shard=0, n_shards=1, probe_pairs=None, device=None,
            drop_unexpressed=False, dist=False, gather_to=-1):
    """-> (keys uint64[n], cls int32[n], timings dict[, probe_pos, probe_neg][, kept_gene_cols])

    drop_unexpressed=True: `gene_cols` are candidates, genes that are zero in every row of x are dropped on
    the device and the kept columns are returned as the last element (keys index the kept list).

    dist=True: pp_sandbag_dist over the ranks of the context's communicator (`comm_init`): every rank passes the same
    arguments, prepares its own block of cells, evaluates its share of the gene-pair tiles; the merged, sorted list
    arrives on every rank (gather_to < 0) or on rank `gather_to` only (n = 0 elsewhere)."""
    lib = load()
    sparse = is_sparse(x)
    if sparse:
        keep, csr, n_rows, n_cols, xdev = csr_arg(x)
    else:
        keep, px, code, is_dev, n_rows, n_cols, ld, xdev = matrix_arg(x)
    h = ctx(_pick_device(device, xdev))
    cell_rows = np.ascontiguousarray(cell_rows, dtype=np.int32)
    class_sizes = np.ascontiguousarray(class_sizes, dtype=np.int64)
    gene_cols = np.ascontiguousarray(gene_cols, dtype=np.int32)
    thresholds = np.ascontiguousarray(thresholds, dtype=np.int64)
    if len(thresholds) != len(class_sizes):
        raise ValueError("thresholds and class_sizes differ in length")
    ok, oc, on = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
    pp = ppos = pneg = None
    n_probe = 0
    if probe_pairs is not None:
        pp = np.ascontiguousarray(probe_pairs, dtype=np.int64).reshape(-1, 2)
        n_probe = len(pp)
        ppos = np.zeros((n_probe, len(class_sizes)), dtype=np.int32)
        pneg = np.zeros_like(ppos)
    kg, kn = ctypes.c_void_p(), ctypes.c_int64()
    split = (int(gather_to),) if dist else (shard, n_shards)
    tail = (_ptr(cell_rows), len(cell_rows), _ptr(class_sizes), len(class_sizes), _ptr(gene_cols), len(gene_cols),
            _ptr(thresholds)) + split + (ctypes.byref(ok), ctypes.byref(oc), ctypes.byref(on), _ptr(pp), n_probe,
                                         _ptr(ppos), _ptr(pneg), 1 if drop_unexpressed else 0, ctypes.byref(kg),
                                         ctypes.byref(kn))
    if sparse:
        fn = lib.pp_sandbag_csr_dist if dist else lib.pp_sandbag_csr
        rc = fn(h, ctypes.byref(csr), n_rows, n_cols, *tail)
    else:
        fn = lib.pp_sandbag_dist if dist else lib.pp_sandbag
        rc = fn(h, px, code, is_dev, n_rows, n_cols, ld, *tail)
    kept = None
    if kg.value:
        kept = np.ctypeslib.as_array(ctypes.cast(kg, ctypes.POINTER(ctypes.c_int32)), shape=(max(kn.value, 1),))[
            :kn.value].copy()
        lib.pp_free(kg)
    check(rc, h)
    n = on.value
    if n:
        keys, cls = _adopt(ok, np.uint64, n), _adopt(oc, np.int32, n)
    else:
        keys, cls = np.empty(0, dtype=np.uint64), np.empty(0, dtype=np.int32)
        lib.pp_free(ok)
        lib.pp_free(oc)
    del keep
    out = (keys, cls, timings(h))
    if probe_pairs is not None:
        out = out + (ppos, pneg)
    if drop_unexpressed:
        out = out + (kept if kept is not None else np.empty(0, dtype=np.int32),)
    return out


def _adopt(ptr, dtype, n):
    """ndarray over a result buffer the library malloc'ed for the caller, without copying it; pp_free runs when the
    last array viewing the buffer goes away."""
    import weakref
    buf = (ctypes.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(ptr.value)
    weakref.finalize(buf, load().pp_free, ctypes.c_void_p(ptr.value))
    return np.frombuffer(buf, dtype=dtype, count=n)


def _result_array(shape, dtype):
    """Uninitialised output array for the library to fill: pinned memory from the library's pool when it is large (the
    result is copied into it by the DMA engine, no staging copy and no page faults of a fresh allocation)."""
    n = int(np.prod(shape))
    nbytes = n * np.dtype(dtype).itemsize
    if nbytes >= (64 << 10):
        p = load().pp_alloc_result(nbytes)
        if p:
            return _adopt(ctypes.c_void_p(p), dtype, n).reshape(shape)
    return np.empty(shape, dtype=dtype)


def score_labels(scores, idx_g1=-1, idx_g2m=-1):
    """(max_class codes, cc codes or None) of pp_score_labels for float64[C, N] scores."""
    lib = load()
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    nc, n = scores.shape
    mc = np.empty(n, dtype=np.int32)
    cc = np.empty(n, dtype=np.int32) if idx_g1 >= 0 and idx_g2m >= 0 else None
    rc = lib.pp_score_labels(_ptr(scores), nc, n, idx_g1, idx_g2m, _ptr(mc), _ptr(cc))
    if rc != 0:
        raise PyPairsB200Error(rc, lib.pp_last_error(None).decode())
    return mc, cc


_RNG = {"mt19937": PP_RNG_MT19937, "philox": PP_RNG_PHILOX, "table": PP_RNG_TABLE}


def cyclone(x, used_idx_list, pairs_list, iterations, min_iter, min_pairs, rng="mt19937", seed=2803,
            perm_tables=None, row_begin=0, n_cells=None, want_observed=False, device=None, shard_cells=None,
            gather_to=-1):
    """-> (scores float64[C, n_cells], timings[, obs_hits, obs_total])

    shard_cells (one entry per rank of the context's communicator, `comm_init`): pp_cyclone_dist -- this call scores
    this rank's `n_cells` cells and the score columns of all ranks are all-gathered on the device; the returned arrays
    are [C, sum(shard_cells)] on every rank (gather_to < 0) or on rank `gather_to` only (None elsewhere)."""
    lib = load()
    sparse = is_sparse(x)
    if sparse:
        keep, csr, n_rows, n_cols, xdev = csr_arg(x)
    else:
        keep, px, code, is_dev, n_rows, n_cols, ld, xdev = matrix_arg(x)
    h = ctx(_pick_device(device, xdev))
    if n_cells is None:
        n_cells = n_rows - row_begin
    nc = len(used_idx_list)
    used_off = np.zeros(nc + 1, dtype=np.int64)
    pair_off = np.zeros(nc + 1, dtype=np.int64)
    for c in range(nc):
        used_off[c + 1] = used_off[c] + len(used_idx_list[c])
        pair_off[c + 1] = pair_off[c] + len(pairs_list[c])
    used = np.ascontiguousarray(np.concatenate([np.asarray(u, dtype=np.int32).ravel() for u in used_idx_list])
                                if nc else np.empty(0), dtype=np.int32)
    pairs = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.int32).reshape(-1, 2) for p in pairs_list])
                                 if nc else np.empty((0, 2)), dtype=np.int32)
    tabs = None
    if rng == "table":
        if perm_tables is None:
            raise ValueError("rng='table' needs perm_tables")
        tabs = np.ascontiguousarray(np.concatenate([np.asarray(t, dtype=np.int32).ravel() for t in perm_tables]),
                                    dtype=np.int32)
    dist = shard_cells is not None
    receive = True
    n_out = n_cells
    if dist:
        shard_cells = np.ascontiguousarray(shard_cells, dtype=np.int64)
        me, world = comm_info(h)
        if len(shard_cells) != world:
            raise ValueError("shard_cells has {} entries for {} ranks".format(len(shard_cells), world))
        receive = gather_to < 0 or gather_to == me
        n_out = int(shard_cells.sum())
    scores = _result_array((nc, n_out), np.float64) if receive else None
    oh = _result_array((nc, n_out), np.int32) if want_observed and receive else None
    ot = _result_array((nc, n_out), np.int32) if want_observed and receive else None
    head = (row_begin, n_cells, nc, _ptr(used), _ptr(used_off), _ptr(pairs), _ptr(pair_off), iterations, min_iter,
            min_pairs, _RNG[rng], seed, _ptr(tabs))
    outs = (_ptr(scores), _ptr(oh), _ptr(ot))
    if dist:
        head = head + (_ptr(shard_cells), int(gather_to), 1 if want_observed else 0)
    if sparse:
        fn = lib.pp_cyclone_csr_dist if dist else lib.pp_cyclone_csr
        rc = fn(h, ctypes.byref(csr), n_rows, n_cols, *(head + outs))
    else:
        fn = lib.pp_cyclone_dist if dist else lib.pp_cyclone
        rc = fn(h, px, code, is_dev, n_rows, n_cols, ld, *(head + outs))
    check(rc, h)
    del keep
    if want_observed:
        return scores, timings(h), oh, ot
    return scores, timings(h)


# ---- multi-GPU: one communicator per context -------------------------------------------------------------------
def nccl_library_path():
    """The libnccl.so.2 torch itself uses (nvidia/nccl/lib of the venv), or None: the library then falls back to
    the copy already mapped into the process / the loader path."""
    try:
        import nvidia.nccl as pkg
        for d in list(pkg.__path__):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
    except ImportError:
        pass
    return None


def comm_unique_id():
    """128 opaque bytes made by rank 0 (ncclGetUniqueId); every rank passes them to comm_init."""
    lib = load()
    p = nccl_library_path()
    rc = lib.pp_nccl_load(p.encode() if p else None)
    if rc != 0:
        raise PyPairsB200Error(rc, lib.pp_last_error(None).decode())
    buf = ctypes.create_string_buffer(PP_COMM_ID_BYTES)
    rc = lib.pp_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p))
    if rc != 0:
        raise PyPairsB200Error(rc, lib.pp_last_error(None).decode())
    return bytes(buf.raw)


def comm_init(unique_id, rank, world, device=None):
    """Join the communicator of `unique_id` with this process's context (collective over all `world` ranks)."""
    lib, h = load(), ctx(device)
    p = nccl_library_path()
    rc = lib.pp_nccl_load(p.encode() if p else None)
    if rc != 0:
        raise PyPairsB200Error(rc, lib.pp_last_error(None).decode())
    buf = ctypes.create_string_buffer(bytes(unique_id), PP_COMM_ID_BYTES)
    check(lib.pp_comm_init(h, ctypes.cast(buf, ctypes.c_void_p), int(rank), int(world)), h)


def comm_destroy(device=None):
    lib, h = load(), ctx(device)
    check(lib.pp_comm_destroy(h), h)


def comm_info(h_or_device=None):
    """(rank, world) of the context's communicator; world = 0 without one."""
    lib = load()
    h = h_or_device if isinstance(h_or_device, ctypes.c_void_p) else ctx(h_or_device)
    r, w = ctypes.c_int32(), ctypes.c_int32()
    check(lib.pp_comm_info(h, ctypes.byref(r), ctypes.byref(w)), h)
    return r.value, w.value


def phase_scores(matrix, iterations, min_iter, min_pairs, pairs, used, device=None):
    """Reference-kernel-shaped call: get_phase_scores(matrix, iterations, min_iter, min_pairs, pairs, used)
    (pypairs/tools/cyclone.py:253-257)."""
    lib = load()
    keep, px, code, is_dev, n_rows, n_cols, ld, xdev = matrix_arg(matrix)
    h = ctx(_pick_device(device, xdev))
    pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
    used = np.ascontiguousarray(used, dtype=np.uint8)
    if len(used) != n_cols:
        raise ValueError("used mask length differs from the number of genes")
    out = np.empty(n_rows, dtype=np.float64)
    check(lib.pp_phase_scores(h, px, code, is_dev, n_rows, n_cols, ld, iterations, min_iter, min_pairs, _ptr(pairs),
                              len(pairs), _ptr(used), _ptr(out)), h)
    del keep
    return out


def perm_table(rng, seed, class_index, n, iterations, device=None):
    lib, h = load(), ctx(device)
    out = np.empty((iterations, n), dtype=np.int32)
    check(lib.pp_perm_table(h, _RNG[rng], seed, class_index, n, iterations, _ptr(out)), h)
    return out


def dense_rank(x, cell_rows, gene_cols, device=None):
    lib = load()
    keep, px, code, is_dev, n_rows, n_cols, ld, xdev = matrix_arg(x)
    h = ctx(_pick_device(device, xdev))
    cell_rows = np.ascontiguousarray(cell_rows, dtype=np.int32)
    gene_cols = np.ascontiguousarray(gene_cols, dtype=np.int32)
    out = np.empty((len(cell_rows), len(gene_cols)), dtype=np.uint16)
    mx = ctypes.c_int32()
    check(lib.pp_dense_rank(h, px, code, is_dev, n_rows, n_cols, ld, _ptr(cell_rows), len(cell_rows), _ptr(gene_cols),
                            len(gene_cols), _ptr(out), ctypes.byref(mx)), h)
    del keep
    return out, mx.value


def alu_peak(repeats=5, device=None):
    """Measured LOP3 lane-operations per second of the device and its SM count."""
    lib, h = load(), ctx(device)
    v, n = ctypes.c_double(), ctypes.c_int32()
    check(lib.pp_alu_peak(h, repeats, ctypes.byref(v), ctypes.byref(n)), h)
    return v.value, n.value
