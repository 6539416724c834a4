"""python -m pypairs_b200.cli.sandbag -- find marker pairs in an annotated CSV matrix (see _common.py)."""
