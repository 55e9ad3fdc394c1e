"""vrod_b200 -- B200-native exact kNN scan behind vRod's SEARCH command.

The product is the CUDA shared library vrod_b200/libvrod_knn.so (C ABI: include/vrod_knn.h) and
the C++ host layer in vrod_b200/host/ that mirrors the reference's Command / CommandBuilder /
Database interface.  `vrod_b200.ffi` is the ctypes binding the tests and bench.py use.
"""
from . import ffi  # noqa: F401
