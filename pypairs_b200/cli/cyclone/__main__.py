"""`python -m pypairs_b200.cli.cyclone`: matrix CSV (+ marker JSON) -> score CSV (stdout without -o).

Unlike the reference's script, which hands the marker *path* to cyclone(), the JSON file is parsed first."""
import json

import click

from ... import log as logg, settings
from ...cyclone import cyclone
from .._common import _IN, _OUT, MATRIX_OPTIONS, read_matrix, with_options

OPTIONS = MATRIX_OPTIONS + [
    (("-o", "--out"), dict(type=_OUT, help="write the score table to this CSV file")),
    (("-m", "--marker_pairs"), dict(type=_IN, help="marker pairs as JSON; default: the bundled cell-cycle markers")),
    (("-i", "--iterations"), dict(default=1000, show_default=True, help="permutations of the null distribution")),
    (("-x", "--min_iter"), dict(default=100, show_default=True, help="fewer valid permutations than this give NaN")),
    (("-y", "--min_pairs"), dict(default=50, show_default=True, help="fewer informative pairs than this give no proportion")),
    (("-q", "--quantile_transform"), dict(default=False, is_flag=True, help="quantile-normalise the matrix first")),
]


@click.command()
@with_options(OPTIONS)
def main(data, sep, transpose, verbosity, out, marker_pairs, iterations, min_iter, min_pairs, quantile_transform):
    settings.verbosity = verbosity
    markers = None
    if marker_pairs is not None:
        with open(marker_pairs) as handle:
            markers = json.load(handle)
    scores = cyclone(data=read_matrix(data, sep, transpose), marker_pairs=markers, iterations=iterations,
                     min_iter=min_iter, min_pairs=min_pairs, quantile_transform=quantile_transform)
    if out is None:
        print(scores)
        return
    try:
        scores.to_csv(out)
        logg.hint("scores written to {}".format(out))
    except IOError:
        logg.error("could not write {}".format(out))


if __name__ == "__main__":
    main()
