"""Minimal stand-in so that the reference package imports in a container without anndata.

Only used by oracle/make_golden.py (and tests that import the live reference when
/root/reference exists).  Not part of the product.
"""


class AnnData:  # the reference only does isinstance() checks against this type
    pass
