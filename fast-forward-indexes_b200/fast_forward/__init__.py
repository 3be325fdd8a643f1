"""fast_forward — B200-native drop-in for the re-ranking hot path of
mrjleo/fast-forward-indexes (`Index.__call__`, `InMemoryIndex`, `OnDiskIndex.load`, `Mode`,
`Ranking.interpolate` / `cut`, the `Quantizer` classes).  Scoring runs in hand-written
sm_100a kernels behind the C ABI of `libffx.so`; there is no CPU scoring path."""

__version__ = "0.8.0+b200.1"

from fast_forward import encoder, index, quantizer, util  # noqa: E402
from fast_forward.ranking import Ranking  # noqa: E402

__all__ = ["encoder", "index", "quantizer", "util", "Ranking"]
