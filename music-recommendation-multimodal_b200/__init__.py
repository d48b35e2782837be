"""B200-native two-tower hot path (SASRec user tower, late-fusion item tower, InfoNCE,
catalog top-K) behind the reference's module interfaces.

The device work lives in ``libtt_b200.so`` (hand-written sm_100a CUDA, C ABI declared in
``include/tt_b200.h``); this package is the PyTorch-facing host side.
"""
from ._lib import lib, TTError  # noqa: F401

__all__ = ["lib", "TTError"]
