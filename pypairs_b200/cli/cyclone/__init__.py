"""python -m pypairs_b200.cli.cyclone -- score samples against marker pairs from CSV input (see _common.py)."""
