"""gnss-sdr-rs_b200: B200 (sm_100a) implementation of the gnss-sdr-rs acquisition / correlator hot path.

The product is the C-ABI library libgnss_b200.so (include/gnss_b200.h, csrc/).  These Python modules
are a thin ctypes binding plus a host-side mirror of the reference crate's acquisition / tracking API,
used by tests/ and bench.py.  There is no CPU fallback: everything raises if the library or a CUDA
device is missing.
"""
__version__ = "0.1.0"
