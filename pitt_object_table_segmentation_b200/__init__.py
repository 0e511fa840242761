"""pitt_object_table_segmentation_b200 — B200-native hot path of pitt_object_table_segmentation.

The product is libpitt_b200.so (hand-written sm_100a CUDA behind the C ABI in include/pitt_b200.h);
this package is the Python host mirror used by the tests and bench. No CPU fallback exists.
"""
from . import _abi  # noqa: F401
from .api import Context, Cloud, PittError, load_library, LIB_PATH, EXPORTED_SYMBOLS, DEBUG_SYMBOLS  # noqa: F401
from .api import (default_sac_params, default_support_sac_params, default_support_params,  # noqa: F401
                  default_cluster_params, default_frame_params, default_prefilter_params, default_arm_filter_params,
                  select_primitive,
                  segment_frames_batched, segment_clouds_batched)
