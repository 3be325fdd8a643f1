"""Indexes (reference: src/fast_forward/index/__init__.py)."""

from fast_forward.index.base import Index, Mode
from fast_forward.index.disk import OnDiskIndex
from fast_forward.index.memory import InMemoryIndex

__all__ = ["Index", "Mode", "OnDiskIndex", "InMemoryIndex"]
