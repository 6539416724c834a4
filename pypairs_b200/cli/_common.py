"""Shared plumbing of the two command-line front ends.

The option names and defaults are the reference's (pypairs/cli/sandbag/__main__.py, pypairs/cli/cyclone/__main__.py) so
that existing shell scripts keep working; everything else is our own."""
import click
import pandas as pd

_IN = click.Path(exists=True, file_okay=True, readable=True, resolve_path=True)
_OUT = click.Path(resolve_path=True, writable=True)

MATRIX_OPTIONS = [
    (("-d", "--data"), dict(required=True, type=_IN, help="expression matrix as CSV: one row per cell, one column per gene")),
    (("-s", "--sep"), dict(default=",", show_default=True, help="field separator of the matrix file")),
    (("-t", "--transpose"), dict(default=False, is_flag=True, help="the file is genes x cells: transpose it")),
    (("-v", "--verbosity"), dict(default=3, show_default=True, help="0 (silent) ... 4 (chatty)")),
]


def with_options(specs):
    """Decorator that applies a list of (flags, kwargs) click options in order."""
    def deco(fn):
        for flags, kw in reversed(specs):
            fn = click.option(*flags, **kw)(fn)
        return fn
    return deco


def read_matrix(path, sep, transpose):
    frame = pd.read_csv(path, index_col=0, sep=sep)
    return frame.T if transpose else frame
