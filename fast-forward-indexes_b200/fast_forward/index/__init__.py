"""Index back-ends of the drop-in package.

`InMemoryIndex` keeps its rows in the GPU's HBM, `OnDiskIndex` persists them in the reference's
HDF5 layout and serves them from HBM as well; both score through libffx (`Index.__call__`,
`Index.rerank`).  `Mode` selects how a document's passages are reduced.
"""

from .base import Index, Mode
from .memory import InMemoryIndex
from .disk import OnDiskIndex

__all__ = ("Mode", "Index", "InMemoryIndex", "OnDiskIndex")
