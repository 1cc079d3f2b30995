"""compress_b200 -- B200 (sm_100a) block-compression engine behind klauspost/compress's
zstd / s2 / huff0 block-codec hot path.  See DESIGN.md and INTEGRATION.md."""
from . import _lib  # noqa: F401  (raises when the CUDA library is missing: there is no CPU fallback)
from . import zstd  # noqa: F401
