"""Quantizers of the drop-in package.

`Quantizer` is the abstract contract (fit / encode / decode / serialize); `NanoPQ` and `NanoOPQ`
keep nanopq's state layout so stored quantizers stay interchangeable with the reference.  On
the scoring path codes are never decoded: a quantized index is scored by the asymmetric-distance
kernels of libffx straight from the uint8 codes.  `device=` moves k-means and encoding of the
build side to the GPU.
"""

from .base import Quantizer
from .nanopq import NanoOPQ, NanoPQ

__all__ = ("NanoOPQ", "NanoPQ", "Quantizer")
