"""B200-native THCM Newton-step hot path (residual, Jacobian -> CRS, FP64 CSR SpMV, GMRES / IDR(s)).

Host-side mirror of the reference's THCM / Ocean interface over the C ABI of ``libthcm_b200.so``
(include/thcm_b200.h).  All arithmetic on the hot path runs in hand-written CUDA kernels for
sm_100a; there is no CPU fallback -- loading fails loudly when the extension is missing.
"""
from .params import PAR_INDEX, PAR_NAMES, par_index            # noqa: F401
from .masks import read_mask, write_mask, synthetic_global_mask, all_ocean_mask  # noqa: F401
from .thcm import (Settings, THCM, Ocean, lib, lib_path, load_library, KrylovResult,  # noqa: F401
                   FortranABI, ThetaOcean, last_error)
from . import paramlist                                         # noqa: F401
