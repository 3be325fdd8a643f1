"""Id -> row resolution (reference: src/fast_forward/index/util.py:12-42)."""

from __future__ import annotations

from collections.abc import Iterable

from fast_forward.index.base import Mode


def get_indices(ids: Iterable[str], mode: Mode, doc_id_to_idx: dict[str, list[int]],
                psg_id_to_idx: dict[str, int]) -> tuple[list[int], list[str]]:
    """Rows needed to score each id in `mode`, and the id owning each returned row.

    MAXP/AVEP: every row of the document, in insertion order; FIRSTP: its first row;
    PASSAGE: the passage's row.  IndexError when an id resolves to no row."""
    rows: list[int] = []
    owners: list[str] = []
    doc_mode = mode is not Mode.PASSAGE
    for id_ in ids:
        if doc_mode:
            found = doc_id_to_idx.get(id_, [])
            if mode is Mode.FIRSTP:
                found = found[:1]
        else:
            row = psg_id_to_idx.get(id_)
            found = [] if row is None else [row]
        if not found:
            raise IndexError(f"ID {id_} not found in the index.")
        rows += found
        owners += [id_] * len(found)
    return rows, owners


class ChunkIndexer:
    """Vectors by id out of a row store that is cut into chunks (reference:
    src/fast_forward/index/util.py:45-113; the first chunk may be longer than the others, which all
    have the length of the second).  Host-side utility of the reference's chunked indexes — the
    device store of this package is one contiguous array and does not need it; kept so that code
    written against the reference finds it."""

    def __init__(self, chunks, doc_id_to_idx: dict[str, list[int]], psg_id_to_idx: dict[str, int]) -> None:
        self._chunks = chunks
        self._doc_id_to_idx = doc_id_to_idx
        self._psg_id_to_idx = psg_id_to_idx

    def _get_chunk_indices(self, idx: int) -> tuple[int, int]:
        """(chunk, row inside the chunk) of a flat row number."""
        first = self._chunks[0].shape[0]
        if idx < first:
            return 0, idx
        step = self._chunks[1].shape[0]
        return (idx - first) // step + 1, (idx - first) % step

    def __call__(self, ids: Iterable[str], mode: Mode):
        """The rows `get_indices` names and their owning ids.  With several chunks the rows come
        back grouped by chunk — chunks in the order their first row is asked for, rows of a chunk
        in the order asked — which is the order the reference's per-chunk gathers produce."""
        import numpy as np

        rows, owners = get_indices(ids, mode, self._doc_id_to_idx, self._psg_id_to_idx)
        if not rows:
            return np.array([]), []
        if len(self._chunks) == 1:
            return self._chunks[0][rows], owners
        flat = np.asarray(rows, np.int64)
        first = self._chunks[0].shape[0]
        step = self._chunks[1].shape[0]
        chunk = np.where(flat < first, 0, (flat - first) // step + 1)
        inner = np.where(flat < first, flat, (flat - first) % step)
        seen, first_use = np.unique(chunk, return_index=True)
        parts, out_ids = [], []
        for c in seen[np.argsort(first_use, kind="stable")].tolist():
            pick = np.flatnonzero(chunk == c)
            parts.append(self._chunks[c][inner[pick].tolist()])
            out_ids.extend(owners[i] for i in pick.tolist())
        return np.concatenate(parts), out_ids
