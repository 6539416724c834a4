"""Command line front ends: python -m pypairs_b200.cli.sandbag / python -m pypairs_b200.cli.cyclone."""
