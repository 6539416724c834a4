"""`python -m pypairs_b200.cli.sandbag`: matrix CSV + annotation CSV -> marker-pair JSON (stdout without -o)."""
import csv
import json

import click

from ... import log as logg, settings
from ...sandbag import sandbag
from .._common import _IN, _OUT, MATRIX_OPTIONS, read_matrix, with_options

OPTIONS = MATRIX_OPTIONS + [
    (("-o", "--out"), dict(type=_OUT, help="write the marker pairs to this JSON file")),
    (("-a", "--annotation"), dict(required=True, type=_IN, help="CSV with (at least) a sample-name and a class column")),
    (("-as", "--ann_sep"), dict(default=",", help="field separator of the annotation file")),
    (("-an", "--ann_idx_name"), dict(default=0, help="0-based column of the sample names")),
    (("-ac", "--ann_idx_class"), dict(default=1, help="0-based column of the class labels")),
    (("--ann_header/--ann_no_header",), dict(default=True, show_default=True, help="first annotation line is a header")),
    (("-f", "--fraction"), dict(default=0.65, show_default=True, help="share of a class's cells that must agree")),
]


def read_annotation(path, sep, name_col, class_col, header):
    groups = {}
    with open(path) as handle:
        rows = iter(csv.reader(handle, delimiter=sep))
        if header:
            next(rows, None)
        for row in rows:
            groups.setdefault(row[class_col], []).append(row[name_col])
    return groups


@click.command()
@with_options(OPTIONS)
def main(data, sep, transpose, verbosity, out, annotation, ann_sep, ann_idx_name, ann_idx_class, ann_header, fraction):
    settings.verbosity = verbosity
    markers = sandbag(data=read_matrix(data, sep, transpose),
                      annotation=read_annotation(annotation, ann_sep, ann_idx_name, ann_idx_class, ann_header),
                      fraction=fraction)
    plain = {str(k): [[str(a), str(b)] for a, b in v] for k, v in markers.items()}
    if out is None:
        print(plain)
        return
    try:
        with open(out, "w") as handle:
            json.dump(plain, handle)
        logg.hint("marker pairs written to {}".format(out))
    except IOError:
        logg.error("could not write {}".format(out))


if __name__ == "__main__":
    main()
