"""pb_starphase_b200 -- B200-native (sm_100a) scoring path for pb-StarPhase.

Only the data-parallel hot path lives here: batched read/consensus-vs-allele
infix edit distance (K1), allele-/chain-pair scoring (K2) and the CYP2D6 candidate scoring
(K3 = K1 with the roles swapped), behind the C ABI declared in include/starphase_gpu.h.
The Python layer is a thin ctypes mirror of that ABI plus host-side restatements of the
reference functions that consume the integers (HLA / CYP2D6 tails).  There is no CPU
fallback: importing works anywhere, but every compute call needs the built
libstarphase_gpu.so and a B200.
"""
from .binding import (  # noqa: F401
    Comm,
    Consensus,
    Context,
    DMatrix,
    PatternSet,
    SpError,
    TargetSet,
    lib_path,
    load_library,
    SP_INFIX,
    SP_PREFIX,
)

__version__ = "0.1.0"
