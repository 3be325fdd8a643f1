"""Encoders.  Only the pieces the re-ranking path needs: the `Encoder` interface,
`LambdaEncoder` (reference: src/fast_forward/encoder/__init__.py:32-44) and `TableEncoder`
for precomputed query vectors.  The HF transformer presets of the reference
(encoder/transformer.py) are out of scope: wrap any model in a `LambdaEncoder`."""

from __future__ import annotations

from collections.abc import Callable, Mapping, Sequence

import numpy as np

from fast_forward.encoder.base import Encoder

__all__ = ["Encoder", "LambdaEncoder", "TableEncoder"]


class LambdaEncoder(Encoder):
    """Adapter around a function that encodes ONE text."""

    def __init__(self, f: Callable[[str], np.ndarray]) -> None:
        super().__init__()
        self._f = f

    def _encode(self, texts: Sequence[str]) -> np.ndarray:
        return np.array([self._f(t) for t in texts])


class TableEncoder(Encoder):
    """Looks precomputed vectors up by query text (benchmarks, cached encoders)."""

    def __init__(self, table: Mapping[str, np.ndarray]) -> None:
        super().__init__()
        self._table = table

    def _encode(self, texts: Sequence[str]) -> np.ndarray:
        return np.stack([self._table[t] for t in texts])
