"""mipgen_b200 -- B200-native candidate-MIP enumeration + scoring (the MIPgen hot path).

The product is libmipgen_b200.so (hand-written sm_100a CUDA behind the C-ABI in
include/mipgen_b200.h) plus the drop-in C++ headers in mipgen_b200/dropin/.  This
package is the thin ctypes view of that C-ABI used by the tests and bench.py.
"""
from .panel import Config, Region  # noqa: F401
from ._capi import (Context, Panel, MgError, load_library, tile_replay, config_grid_size, design_records, describe_candidates, universal_middle, tile_regions, tile_sizes, partition_regions, TileResult,  # noqa: F401
                    MG_WANT_LOGISTIC, MG_WANT_SVR,
                    MG_WANT_FEATURES, MG_NFEAT, MG_NLRC)
