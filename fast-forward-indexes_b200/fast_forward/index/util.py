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
