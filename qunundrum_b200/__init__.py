"""qunundrum_b200 -- B200-native slice integrators behind the reference's interface.

The package is a thin host-side mirror (same names, argument meaning and error
behaviour) of the six entry points of ekera/qunundrum's slice-integration path,
over the C ABI declared in include/qunundrum_b200.h and implemented by
hand-written sm_100a CUDA kernels in qunundrum_b200/csrc.

There is NO CPU path: importing works anywhere (so that the host logic can be
unit-tested), but every compute call needs libqunundrum_b200.so and a CUDA
device and raises otherwise.
"""
from .host import (  # noqa: F401
    CriticalError,
    Context,
    Diagonal_Distribution_Slice,
    Diagonal_Parameters,
    Distribution_Slice,
    Linear_Distribution_Slice,
    Parameters,
    DISTRIBUTION_SLICE_COMPUTE_METHOD_HEURISTIC_SIGMA,
    DISTRIBUTION_SLICE_COMPUTE_METHOD_OPTIMAL_LOCAL_SIGMA,
    DISTRIBUTION_SLICE_COMPUTE_METHOD_QUICK,
    LINEAR_DISTRIBUTION_SLICE_COMPUTE_TARGET_D,
    LINEAR_DISTRIBUTION_SLICE_COMPUTE_TARGET_R,
    SLICE_FLAGS_ERROR_BOUND_WARNING,
    SLICE_FLAGS_METHOD_RICHARDSON,
    SLICE_FLAGS_METHOD_SIMPSON,
    SLICE_FLAGS_MASK_METHOD,
    default_context,
    diagonal_distribution_slice_compute,
    diagonal_distribution_slice_compute_richardson,
    distribution_slice_compute,
    distribution_slice_compute_richardson,
    heuristic_sigma,
    host_constants,
    lib,
    lib_path,
    linear_distribution_slice_compute,
    linear_distribution_slice_compute_richardson,
    # the slice text format
    TEXT_F64,
    TEXT_X87,
    diagonal_distribution_slice_export,
    diagonal_distribution_slice_import,
    distribution_slice_export,
    distribution_slice_import,
    linear_distribution_slice_export,
    linear_distribution_slice_import,
    text_pow10,
    # sampling from stored distributions: tau estimation
    Distribution,
    Linear_Distribution,
    Sampler,
    WordStream,
    tau_estimate,
    tau_estimate_linear,
    # the server's work on a finished distribution
    Resident,
    linear_distribution_init_collapse_d,
    linear_distribution_init_collapse_r,
    # the diagonal distribution's k given (j, eta) (second half)
    DiagonalKSampler,
    int_to_limbs,
    limbs_to_int,
    sample_k_from_diagonal_j_eta_pivot,
    # the exact samplers (first half; src/sample.cpp:78-410)
    EXACT_DIAGONAL,
    EXACT_REGION_DTYPE,
    EXACT_TWO_DIMENSIONAL,
    ExactSampler,
    diagonal_sample_drawn,
    pack_regions,
    ByteStream,
    sample_alpha_from_region,
    sample_j_from_alpha_r,
    sample_j_from_diagonal_alpha_r,
    sample_j_k_from_alpha_d,
    sample_j_k_from_alpha_d_r,
)
from . import host  # noqa: F401,E402

__all__ = [n for n in dir() if not n.startswith("_")]
