"""Encoders: the `Encoder` interface, `LambdaEncoder` (reference:
src/fast_forward/encoder/__init__.py:32-44), `TableEncoder` for precomputed query vectors, and
the Transformer presets of encoder/transformer.py (resolved on first use, so importing the
package does not import torch / transformers)."""

from __future__ import annotations

from collections.abc import Callable, Mapping, Sequence

import numpy as np

from fast_forward.encoder.base import Encoder

_TRANSFORMER_PRESETS = ("TransformerEncoder", "TCTColBERTQueryEncoder", "TCTColBERTDocumentEncoder",
                        "TASBEncoder", "ContrieverEncoder", "BGEEncoder")

__all__ = ["Encoder", "LambdaEncoder", "TableEncoder", *_TRANSFORMER_PRESETS]


def __getattr__(name: str):
    if name in _TRANSFORMER_PRESETS:
        from fast_forward.encoder import transformer

        return getattr(transformer, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


class LambdaEncoder(Encoder):
    """Adapter around a function that encodes ONE text."""

    def __init__(self, f: Callable[[str], np.ndarray]) -> None:
        super().__init__()
        self._f = f

    def _encode(self, texts: Sequence[str]) -> np.ndarray:
        return np.array([self._f(t) for t in texts])


class TableEncoder(Encoder):
    """Looks precomputed vectors up by query text (benchmarks, cached encoders): one dictionary
    probe per text and ONE gather from the stacked table per call."""

    def __init__(self, table: Mapping[str, np.ndarray]) -> None:
        super().__init__()
        self._row = {text: i for i, text in enumerate(table)}
        self._matrix = np.stack([np.asarray(v) for v in table.values()]) if len(self._row) else np.zeros((0, 0), np.float32)

    def _encode(self, texts: Sequence[str]) -> np.ndarray:
        return self._matrix[[self._row[t] for t in texts]]
