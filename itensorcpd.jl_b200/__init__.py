"""itensorcpd.jl_b200 -- B200-native (sm_100a) CP-ALS engine behind ITensorCPD.jl's decomposition API.

Only what the hot path needs: `csrc/` (hand-written CUDA + the C-ABI, built into lib/libitcpd_b200.so),
`engine.py` (ctypes handle) and `host.py` (the reference's public API mirrored over the C-ABI).
The directory name contains a dot, so import it through the `itcpd` shim at the repo root:

    import itcpd                      # -> this package
    cp = itcpd.decompose(T, 50, check=itcpd.FitCheck(1e-3, 100, norm_T))
"""
from ._lib import ItcpdError, LIB_PATH, DECLARED_SYMBOLS, load  # noqa: F401
from .engine import Engine, PinnedBuffer, column_to_multi_coords, multi_coords_to_column, sparse_sign_matrix  # noqa: F401
from .host import (  # noqa: F401
    ALS, BlockLevScoreSampled, CPD, CPAngleCheck, CPDFit, CPDiffCheck, CPDOptimizer, DirectNormal, FitCheck, KRPFreeNormal, KRPNormal, KSEQRCSPivProjected,
    LevScoreSampled, MttkrpAlgorithm, NoCheck, ProjectionAlgorithm, QRPivProjected, SEQRCSPivProjected, als_optimize,
    compute_als, cp_rank, decompose, increase_cpd_rank, optimize, random_CPD, random_factors, reconstruct, start, stop,
    update_samples,
)

__all__ = [n for n in dir() if not n.startswith("_")]
