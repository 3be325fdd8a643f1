"""`Quantizer` — trained/attached state machine + (de)serialisation, drop-in for
src/fast_forward/quantizer/base.py:16-196 of the reference.

On the scoring path the index never calls `decode`: codes stay uint8 in HBM and the
asymmetric-distance kernel (csrc/ffx_adc.cuh) scores them against per-query lookup tables
built from `adc_tables()`.  `decode` remains for `Index.batch_iter` / `_get_vectors` users.
"""

from __future__ import annotations

import abc
import importlib
import logging
from collections.abc import Mapping
from typing import Any

import numpy as np

LOGGER = logging.getLogger(__name__)

QuantizerAttributes = Mapping[str, "str | bool | float"]
QuantizerData = Mapping[str, np.ndarray]


class Quantizer(abc.ABC):
    """Base class for vector quantizers."""

    _attached: bool = False
    _trained: bool = False

    # ---- identity -------------------------------------------------------------------
    def __eq__(self, o: object) -> bool:
        """Equal when the serialised forms agree (meta, attributes and every array)."""
        if not isinstance(o, Quantizer):
            return False
        mine, theirs = self.serialize(), o.serialize()
        if mine[0] != theirs[0] or mine[1] != theirs[1]:
            return False
        if mine[2].keys() != theirs[2].keys():
            return False
        return all(np.array_equal(v, theirs[2][k]) for k, v in mine[2].items())

    __hash__ = None

    # ---- lifecycle ------------------------------------------------------------------
    def set_attached(self) -> None:
        """Called by `Index.quantizer = ...`; afterwards `fit` is refused.

        :raises RuntimeError: if the quantizer has not been fit yet.
        """
        if not self._trained:
            raise RuntimeError(
                f"Call {type(self).__name__}.fit before attaching the quantizer to an index.")
        self._attached = True

    def fit(self, vectors: np.ndarray, **kwargs: Any) -> None:
        """Train on `vectors`; only allowed before the quantizer is attached to an index."""
        if self._attached:
            raise RuntimeError("Quantizers can only be fitted before they are attached to an index.")
        self._fit(vectors, **kwargs)
        self._trained = True

    def _require_trained(self) -> None:
        if not self._trained:
            raise RuntimeError(f"Call {type(self).__name__}.fit first.")

    def encode(self, vectors: np.ndarray) -> np.ndarray:
        """Vectors -> codes (RuntimeError before `fit`)."""
        self._require_trained()
        return self._encode(vectors)

    def decode(self, codes: np.ndarray) -> np.ndarray:
        """Codes -> approximate vectors (RuntimeError before `fit`)."""
        self._require_trained()
        return self._decode(codes)

    @property
    def dtype(self) -> np.dtype:
        """Data type of the codes."""
        return self._get_dtype()

    @property
    def dims(self) -> tuple[int | None, int | None]:
        """(dimension of the original vectors, dimension of the codes); None until known."""
        return self._get_dims()

    def adc_tables(self) -> tuple[np.ndarray, np.ndarray | None] | None:
        """`(codewords[M, Ks, Ds], R[D, D] or None)` when scores can be computed straight
        from the codes as `sum_m (qR)[m] . codewords[m, code_m]`; None otherwise."""
        return None

    # ---- serialisation --------------------------------------------------------------
    def serialize(self) -> tuple[QuantizerAttributes, QuantizerAttributes, QuantizerData]:
        """`(meta, attributes, data)` as stored in an index file (disk.py:123-136)."""
        attributes, data = self._get_state()
        meta = {"__module__": type(self).__module__, "__name__": type(self).__name__,
                "_trained": self._trained}
        return meta, attributes, data

    @classmethod
    def deserialize(cls, meta: QuantizerAttributes, attributes: QuantizerAttributes,
                    data: QuantizerData) -> "Quantizer":
        """Rebuild a quantizer from `serialize()` output (class looked up by module + name)."""
        LOGGER.debug("reconstructing %s.%s", meta["__module__"], meta["__name__"])
        klass = getattr(importlib.import_module(str(meta["__module__"])), str(meta["__name__"]))
        quantizer = klass._from_state(attributes, data)
        quantizer._trained = bool(meta["_trained"])
        return quantizer

    # ---- to be provided by implementations --------------------------------------------
    @abc.abstractmethod
    def _fit(self, vectors: np.ndarray, **kwargs: Any) -> None: ...

    @abc.abstractmethod
    def _get_dtype(self) -> np.dtype: ...

    @abc.abstractmethod
    def _get_dims(self) -> tuple[int | None, int | None]: ...

    @abc.abstractmethod
    def _encode(self, vectors: np.ndarray) -> np.ndarray: ...

    @abc.abstractmethod
    def _decode(self, codes: np.ndarray) -> np.ndarray: ...

    @abc.abstractmethod
    def _get_state(self) -> tuple[QuantizerAttributes, QuantizerData]: ...

    @classmethod
    @abc.abstractmethod
    def _from_state(cls, attributes: QuantizerAttributes, data: QuantizerData) -> "Quantizer": ...
