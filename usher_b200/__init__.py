"""usher_b200 — B200-native sample-placement engine (C ABI in include/usher_b200.h).

The product is libusher_b200.so (hand-written sm_100a kernels behind a C ABI) plus the C++ host code under
usher_b200/csrc/.  The Python in this package only builds the libraries in-tree and binds them for tests and
bench.py."""
from . import build, capi, dist  # noqa: F401
