"""Transformer text encoders (drop-in for src/fast_forward/encoder/transformer.py:18-262).

Query/document encoding sits in front of the re-ranking path and is excluded from its metric;
these classes exist so that scripts written against the reference keep working.  One
`TransformerEncoder` runs tokenizer + model and hands the last hidden states to a pooling
function; the presets differ only in prompt decoration, tokenizer arguments and pooling.
torch and transformers are imported when an encoder is created, not when the package is.
"""

from __future__ import annotations

from collections.abc import Mapping, Sequence
from pathlib import Path
from typing import Any

import numpy as np

from fast_forward.encoder.base import Encoder


# ---- pooling: (last_hidden_state [B, T, H], attention_mask [B, T]) -> [B, H] -------------------
def _pool_first_token(hidden, mask):
    return hidden[:, 0]


def _pool_mean_after(skip: int, masked: bool):
    """Mean over token positions >= `skip`; padding excluded when `masked`."""

    def pool(hidden, mask):
        import torch

        states = hidden[:, skip:, :]
        if not masked:
            return torch.mean(states, dim=-2)
        weights = mask[:, skip:].unsqueeze(-1).expand(states.size()).float()
        return torch.sum(states * weights, 1) / torch.clamp(weights.sum(1), min=1e-9)

    return pool


def _pool_masked_mean(hidden, mask):
    kept = hidden.masked_fill(~mask[..., None].bool(), 0.0)
    return kept.sum(dim=1) / mask.sum(dim=1)[..., None]


class TransformerEncoder(Encoder):
    """A pre-trained Transformer; by default the first ([CLS]) token of the last layer."""

    _pool = staticmethod(_pool_first_token)

    def __init__(self, model: str | Path, device: str = "cpu", model_args: Mapping[str, Any] = {},
                 tokenizer_args: Mapping[str, Any] = {},
                 tokenizer_call_args: Mapping[str, Any] = {"padding": True, "truncation": True},
                 normalize: bool = False) -> None:
        """`model`: name or path; `device`: torch device; `normalize`: L2-normalise outputs."""
        from transformers import AutoModel, AutoTokenizer

        super().__init__()
        self._model = AutoModel.from_pretrained(model, **model_args)
        self._model.to(device)
        self._model.eval()
        self._tokenizer = AutoTokenizer.from_pretrained(model, **tokenizer_args)
        self._device = device
        self._tokenizer_call_args = dict(tokenizer_call_args)
        self._normalize = normalize

    def _get_tokenizer_inputs(self, texts: Sequence[str]) -> list[str]:
        """What is tokenized for `texts` (presets add their prompt markers here)."""
        return list(texts)

    def _aggregate_model_outputs(self, model_outputs, model_inputs):
        """[B, H] representations from the model outputs (the reference's override point)."""
        return self._pool(model_outputs.last_hidden_state, model_inputs["attention_mask"])

    def _encode(self, texts: Sequence[str]) -> np.ndarray:
        import torch

        model_inputs = self._tokenizer(self._get_tokenizer_inputs(texts), return_tensors="pt",
                                       **self._tokenizer_call_args).to(self._device)
        with torch.no_grad():
            pooled = self._aggregate_model_outputs(self._model(**model_inputs), model_inputs)
            if self._normalize:
                pooled = torch.nn.functional.normalize(pooled, p=2, dim=1)
        return pooled.cpu().detach().numpy()


class TCTColBERTQueryEncoder(TransformerEncoder):
    """TCT-ColBERT queries (https://aclanthology.org/2021.repl4nlp-1.17/): "[CLS] [Q] " + query,
    padded with [MASK] to `max_length`; mean of the states after the 4 marker tokens."""

    _pool = staticmethod(_pool_mean_after(4, masked=False))

    def __init__(self, model: str | Path = "castorini/tct_colbert-msmarco", device: str = "cpu",
                 max_length: int = 36) -> None:
        self._max_length = max_length
        super().__init__(model, device=device, tokenizer_call_args={
            "max_length": max_length, "truncation": True, "add_special_tokens": False})

    def _get_tokenizer_inputs(self, texts: Sequence[str]) -> list[str]:
        return ["[CLS] [Q] " + text + "[MASK]" * self._max_length for text in texts]


class TCTColBERTDocumentEncoder(TransformerEncoder):
    """TCT-ColBERT documents: "[CLS] [D] " + text; padding-aware mean after the 4 marker tokens."""

    _pool = staticmethod(_pool_mean_after(4, masked=True))

    def __init__(self, model: str | Path = "castorini/tct_colbert-msmarco", device: str = "cpu",
                 max_length: int = 512) -> None:
        self._max_length = max_length
        super().__init__(model, device=device, tokenizer_call_args={
            "max_length": max_length, "padding": True, "truncation": True, "add_special_tokens": False})

    def _get_tokenizer_inputs(self, texts: Sequence[str]) -> list[str]:
        return ["[CLS] [D] " + text for text in texts]


class TASBEncoder(TransformerEncoder):
    """TAS-B (https://dl.acm.org/doi/10.1145/3404835.3462891): [CLS] pooling."""

    def __init__(self, model: str | Path = "sebastian-hofstaetter/distilbert-dot-tas_b-b256-msmarco",
                 device: str = "cpu") -> None:
        super().__init__(model, device=device)


class ContrieverEncoder(TransformerEncoder):
    """Contriever (https://openreview.net/forum?id=jKN1pXi7b0): mean over non-padding tokens."""

    _pool = staticmethod(_pool_masked_mean)

    def __init__(self, model: str | Path = "facebook/contriever", device: str = "cpu") -> None:
        super().__init__(model, device=device)


class BGEEncoder(TransformerEncoder):
    """BGE (https://dl.acm.org/doi/10.1145/3626772.3657878): [CLS] pooling, L2-normalised."""

    def __init__(self, model: str | Path = "BAAI/bge-base-en-v1.5", device: str = "cpu") -> None:
        super().__init__(model, device=device, normalize=True)
