"""gsearch_b200 -- B200-native (sm_100a) sketch-and-search hot path of GSearch.

The product is the C-ABI shared library ``libgsearch_b200.so`` (``include/gsearch_b200.h``);
this package is the thin host-side mirror of the reference's interfaces for that path
(``SeqSketcherT`` / ``DistHamming`` / ``Hnsw``), bound with ctypes.  There is no CPU path:
every compute call raises :class:`GsbError` when no B200 is visible.
"""
from ._lib import GsbError, lib, lib_path, version, device_count  # noqa: F401
from .params import (  # noqa: F401
    ALGO_PROB3A, ALGO_SUPER, ALGO_OPTDENS, ALGO_REVOPTDENS, ALGO_SUPER2, ALGO_HLL, DATA_DNA, DATA_AA,
    SIG_U32, SIG_U64, SIG_F32, SIG_U16, SPEC_NOHASH_IDENTITY, SPEC_OPTDENS_F64_DRAW,
    SeqSketcherParams, HnswParams, sig_dtype,
)
from .sketcher import Sketcher  # noqa: F401
from .distance import DistHamming  # noqa: F401
from .index import Hnsw, Neighbour  # noqa: F401
from . import synth  # noqa: F401
from . import comm  # noqa: F401
