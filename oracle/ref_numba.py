"""Loader of the reference's own Numba path from oracle/_ref/ (see oracle/build_ref.py).

TEST / BENCH INFRASTRUCTURE ONLY -- imported by bench.py (--impl reference and the cpu_baseline leg) and by
tests/test_oracle.py; never by the product package.  Call `configure_threads(n)` BEFORE the first `load()`:
Numba fixes its thread pool at import time, and torch.distributed.run exports OMP_NUM_THREADS=1.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_mod = None


def available():
    return os.path.exists(os.path.join(REF_DIR, "pypairs", "tools", "cyclone.py"))


def configure_threads(n):
    """Give Numba's parallel backends (and the OpenMP runtime under them) `n` threads."""
    n = str(max(1, int(n)))
    os.environ["NUMBA_NUM_THREADS"] = n
    os.environ["OMP_NUM_THREADS"] = n


def load():
    """Import the unmodified reference: returns a namespace with the kernels of both hot paths."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise RuntimeError("oracle/_ref is absent: run `python oracle/build_ref.py` where /root/reference exists")
    for p in (os.path.join(HERE, "shims"), REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import numba
    import pypairs
    from pypairs import settings, utils
    from pypairs.tools import cyclone as ref_cyclone, sandbag as ref_sandbag
    if os.path.realpath(os.path.dirname(pypairs.__file__)) != os.path.realpath(os.path.join(REF_DIR, "pypairs")):
        raise RuntimeError("`pypairs` resolved to {} instead of oracle/_ref".format(pypairs.__file__))
    settings.verbosity = 0

    class Ref(object):
        pass

    r = Ref()
    r.numba, r.settings, r.utils, r.cyclone, r.sandbag = numba, settings, utils, ref_cyclone, ref_sandbag
    r.threads = numba.get_num_threads()
    try:
        r.threading_layer = numba.threading_layer()
    except Exception:  # decided by the first parallel run
        r.threading_layer = None
    _mod = r
    return r


def check_pairs(x, cats, thr, n_jobs):
    """utils.parallel_njit(check_pairs)(X f64[N,G], cats i64[N], thr i64[C]) -> i64[G,G]
    (pypairs/tools/sandbag.py:125-126, 160-199; a new dispatcher = a fresh JIT per call, as in the reference)."""
    import numpy as np
    r = load()
    r.settings.enable_jit = True
    r.settings.n_jobs = int(n_jobs)
    fn = r.utils.parallel_njit(r.sandbag.check_pairs)
    return fn(np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(cats, dtype=np.int64),
              np.ascontiguousarray(thr, dtype=np.int64))


def compiled_check_pairs(n_jobs):
    """The jitted dispatcher itself, so that a caller can warm it up once and time calls without the JIT."""
    r = load()
    r.settings.enable_jit = True
    r.settings.n_jobs = int(n_jobs)
    return r.utils.parallel_njit(r.sandbag.check_pairs)


def phase_scores(x, iterations, min_iter, min_pairs, pairs, used):
    """get_phase_scores (pypairs/tools/cyclone.py:253-257): the parallel gufunc over the cells of x."""
    import numpy as np
    r = load()
    return r.cyclone.get_phase_scores(np.ascontiguousarray(x, dtype=np.float64), iterations, min_iter, min_pairs,
                                      np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2),
                                      np.ascontiguousarray(used, dtype=bool))


def threading_layer():
    r = load()
    try:
        return r.numba.threading_layer()
    except Exception:
        return r.threading_layer
