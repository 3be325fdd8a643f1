"""`NanoPQ` / `NanoOPQ` — drop-ins for src/fast_forward/quantizer/nanopq.py:9-149.

Same constructor arguments, same serialised state (attributes M, Ks, Ds, metric, verbose;
data `codewords`, and `R` for OPQ), same module path and class names — so a quantizer
stored in an index file by the reference deserialises here and vice versa.  The arithmetic
is provided by `fast_forward.quantizer._pq` instead of the `nanopq` package.
"""

from __future__ import annotations

from typing import Any

import numpy as np

from fast_forward.quantizer._pq import Codebook
from fast_forward.quantizer.base import Quantizer, QuantizerAttributes, QuantizerData


class _CodebookQuantizer(Quantizer):
    _ROTATED = False

    def __init__(self, M: int, Ks: int, metric: str = "dot", verbose: bool = False,
                 device: int | None = None) -> None:
        """:param M: number of subspaces. :param Ks: codewords per subspace.
        :param metric: "dot" or "l2". :param verbose: kept for API compatibility.
        :param device: CUDA ordinal to run `fit`'s k-means and `encode` on (not in the reference;
            None = on the host with scipy, like the reference)."""
        self._book = Codebook(M=M, Ks=Ks, metric=metric, verbose=verbose, rotated=self._ROTATED, device=device)
        super().__init__()

    def _fit(self, vectors: np.ndarray, **kwargs: Any) -> None:
        self._book.fit(vectors, **kwargs)

    def _get_dtype(self) -> np.dtype:
        return np.dtype(self._book.code_dtype)

    def _get_dims(self) -> tuple[int | None, int | None]:
        book = self._book
        return (None if book.Ds is None else book.Ds * book.M), book.M

    def _encode(self, vectors: np.ndarray) -> np.ndarray:
        return self._book.encode(vectors)

    def _decode(self, codes: np.ndarray) -> np.ndarray:
        return self._book.decode(codes)

    def adc_tables(self):
        if self._book.codewords is None:
            return None
        return self._book.codewords, self._book.R

    def _get_state(self) -> tuple[QuantizerAttributes, QuantizerData]:
        book = self._book
        attributes = {"M": book.M, "Ks": book.Ks, "Ds": book.Ds, "metric": book.metric,
                      "verbose": book.verbose}
        data = {}
        if book.codewords is not None:
            data["codewords"] = book.codewords
        if book.R is not None:
            data["R"] = book.R
        return attributes, data

    @classmethod
    def _from_state(cls, attributes: QuantizerAttributes, data: QuantizerData):
        quantizer = cls(M=int(attributes["M"]), Ks=int(attributes["Ks"]),
                        metric=str(attributes["metric"]), verbose=bool(attributes["verbose"]))
        if attributes.get("Ds") is not None:
            quantizer._book.Ds = int(attributes["Ds"])
        if "codewords" in data:
            quantizer._book.codewords = np.asarray(data["codewords"], np.float32)
        if "R" in data:
            quantizer._book.R = np.asarray(data["R"], np.float32)
        return quantizer


class NanoPQ(_CodebookQuantizer):
    """Product quantizer (nanopq.PQ semantics)."""

    _pq = property(lambda self: self._book)  # the attribute name of quantizer/nanopq.py:26


class NanoOPQ(_CodebookQuantizer):
    """Optimised product quantizer: learned rotation + PQ (nanopq.OPQ semantics)."""

    _ROTATED = True
    _opq = property(lambda self: self._book)  # the attribute name of quantizer/nanopq.py:94
