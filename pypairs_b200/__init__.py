"""pypairs_b200 -- B200-native sandbag / cyclone (drop-in for pypairs.pairs)."""
from . import datasets, pairs, settings, utils  # noqa: F401

__version__ = "0.1.0"
