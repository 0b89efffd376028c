"""B200-native (sm_100a) implementation of ULTRA's relational message-passing hot path:
`generalized_rspmm` forward + backward behind the reference's own operator boundary
(`torchdrug.layers.functional.generalized_rspmm`, reference ultra/layer.py:134-167, 336-369).

Public surface:
    ultra_torchdrug_b200.functional.generalized_rspmm   the operator (autograd-enabled)
    ultra_torchdrug_b200.functional.GraphIndex          cached CSR/CSC/by-relation edge orders
    ultra_torchdrug_b200.compat.install()               import shims so the unmodified reference modules import
    ultra_torchdrug_b200.build.build()                  nvcc build of libultra_rspmm.so (C ABI: include/ultra_rspmm.h)
"""
__version__ = "0.1.0"
