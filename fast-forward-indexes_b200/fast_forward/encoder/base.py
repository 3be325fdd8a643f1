"""Encoder base class (reference: src/fast_forward/encoder/base.py:10-23).

Query encoding is outside the accelerated path (the metric uses precomputed query vectors);
the classes here only keep `Index(query_encoder=...)` drop-in."""

from __future__ import annotations

import abc
from collections.abc import Sequence

import numpy as np


class Encoder(abc.ABC):
    """Maps a batch of texts to a `[len(texts), dim]` array."""

    @abc.abstractmethod
    def _encode(self, texts: Sequence[str]) -> np.ndarray:
        ...

    def __call__(self, texts: Sequence[str]) -> np.ndarray:
        """Encode a batch of texts."""
        return self._encode(texts)
