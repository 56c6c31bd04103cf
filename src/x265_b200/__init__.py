"""x265 1.9 lookahead cost estimation on B200 (sm_100a).

The product is the C-ABI shared library libx265cu.so (include/x265cu.h) plus the C++ host layer
libx265cu_host.so that mirrors x265's Lookahead / CostEstimateGroup interface.  This package only
holds their sources (csrc/, host/) and thin ctypes bindings used by the tests and bench.py.
"""
from . import abi  # noqa: F401
