"""Test-infrastructure stub: `duckdb` is imported at module scope by the reference's
dquartic/utils/data_loader.py:4 but only used on the parquet branch, which the oracle never
takes (NPY path only).  Not part of the product."""


def query(*a, **k):
    raise RuntimeError("duckdb is not installed in this image; the oracle uses the NPY path only")
