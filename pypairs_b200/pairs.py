"""Drop-in for pypairs.pairs (pypairs/pairs.py:1-2)."""
from .cyclone import cyclone  # noqa: F401
from .sandbag import sandbag  # noqa: F401
