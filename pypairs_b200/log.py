"""Tiny verbosity-gated logger (the reference's scanpy-style logger is out of scope)."""
import sys
import time

from . import settings

_t0 = time.time()


def _emit(level, tag, msg):
    if settings.verbosity >= level:
        print("{}{}".format(tag, msg), file=sys.stderr)


def error(msg):
    _emit(1, "ERROR: ", msg)


def warn(msg):
    _emit(2, "WARNING: ", msg)


def info(msg, r=False, time_=False):
    global _t0
    if r:
        _t0 = time.time()
    if time_:
        msg = "{} ({:.3f} s)".format(msg, time.time() - _t0)
    _emit(3, "", msg)


def hint(msg):
    _emit(4, "--> ", msg)
