"""Quantizers (reference: src/fast_forward/quantizer/__init__.py)."""

from fast_forward.quantizer.base import Quantizer
from fast_forward.quantizer.nanopq import NanoOPQ, NanoPQ

__all__ = ["Quantizer", "NanoPQ", "NanoOPQ"]
