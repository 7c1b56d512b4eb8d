"""The whole GPU test suite against the -DDR_BOUNDS_CHECK build (every volume load and every cell-major gradient reduction is
range-checked on the device), then the violation count of that one process.  Run on a B200:
    DIFFRENDER_LIB=differender_b200/libdiffrender_dbg.so python tools/bounds_suite.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytest
from differender_b200 import _lib

assert "dbg" in os.path.basename(_lib.LIB_PATH), "set DIFFRENDER_LIB to the debug build"
rc = pytest.main(["tests", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"] + sys.argv[1:])
n = _lib.load().dr_debug_oob_count()
print(f"pytest exit {int(rc)}; out-of-range volume loads / gradient reductions counted over the whole suite: {n}")
sys.exit(int(rc) or (1 if n else 0))
