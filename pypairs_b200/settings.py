"""Module-level settings, mirroring the fields of pypairs/settings.py that touch the hot paths.

Only `verbosity`, `fix_seed` and `seed` influence results (pypairs/settings.py:50-54,
pypairs/tools/cyclone.py:196-199).  Note on parity: in the reference the seed is frozen when
Numba compiles `__setseed` at import time, so the *effective* reference seed is always 2803;
here changing `seed` before a call does take effect (the default reproduces the reference).
"""
import os

verbosity = 1
"""0 silent ... 4 chatty (only messages, never results)."""

writedir = "./write/"
"""Root of utils.export_marker / utils.load_marker when `defaultpath=True` (pypairs/settings.py:19)."""

n_jobs = os.cpu_count()
"""Kept for source compatibility; the GPU path ignores it."""

enable_jit = True
enable_fastmath = True
"""Kept for source compatibility; ignored."""

fix_seed = True
"""True: every cell of a class sees the MT19937 stream seeded with `seed` (reference behaviour)."""

seed = 2803
"""Seed of the permutation stream (pypairs/settings.py:54)."""

rng = "mt19937"
"""'mt19937' = bit-exact replay of Numba's np.random stream; 'philox' = counter-based table
generated on the device (statistically equivalent, not bit-identical to the reference)."""
