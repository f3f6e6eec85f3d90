"""B200-native batched sigma-point filter engine (host-side Python plumbing).

The product is the CUDA library `csrc/libslb.so` behind the C ABI of include/slb.h; this package
only loads it (`engine`), builds it (`build`) and generates synthetic workloads (`synth`).
"""
__all__ = ["engine", "build", "synth"]
