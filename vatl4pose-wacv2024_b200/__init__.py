"""vatlq — B200-native (sm_100a) implementation of VATL4Pose's active-learning query pass.

The package name on disk is `vatl4pose-wacv2024_b200`; `import vatlq` (repo-root shim) aliases it.
Host code is Python over a C-ABI CUDA library (libvatlq.so, include/vatlq.h); there is no CPU path.
"""
from . import _lib, synth  # noqa: F401
from . import ops, query, dist, integration  # noqa: F401
from .active_learning import (ActiveLearning, IndexCollection, WholeBodyAE, compute_entropy, compute_hybrid, compute_thc,  # noqa: F401
                              coreset_selection, heatmap_to_coord_simple, localpeak_mean)
from .query import QueryPass, QueryResult, run_query  # noqa: F401

__version__ = "0.1.0"
