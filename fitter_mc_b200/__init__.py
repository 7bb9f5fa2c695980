"""fitter_mc_b200 -- B200-native bootstrap-fitting engine (hot path of ndd21/Fitter_MC).

The product is libfmc_b200.so (hand-written CUDA for sm_100a behind the C ABI in
include/fmc_b200.h).  This package is the thin Python host layer over that ABI; its
function names follow the reference's routines (bootstrap.f90): initialise_model, model_y,
derived_properties, fit_data, monte_carlo_fit_data.  There is no CPU fallback: importing
works anywhere, computing needs a CUDA device.
"""
from .api import (  # noqa: F401
    DEFAULT_SEED,
    Context,
    FmcError,
    debug_philox4x32,
    decode_accumulator,
    derived_properties,
    finalise_moments,
    fit_data,
    generate_offsets,
    histogram_quantiles,
    histogram_range_from_pilot,
    initialise_model,
    last_combine_method,
    lib,
    lib_path,
    measure_dfma_peak,
    model_count,
    model_y,
    monte_carlo_fit_data,
    nccl_available,
    nccl_unique_id,
)
