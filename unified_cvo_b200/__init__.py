"""unified_cvo_b200 — B200-native hot path of Unified CVO behind the reference's API.

The product is unified_cvo_b200/csrc/libcvo_b200.so (hand-written sm_100a kernels + a
C-ABI, include/cvo_b200.h).  This package is the thin host mirror of the reference's
CvoGPU / CvoPointCloud / CvoParams / Association classes over that C-ABI.
"""
from ._abi import AlignInfo, IterTrace, Params, load_library  # noqa: F401
from .cvo import (Association, CvoError, CvoGPU, CvoParams, CvoPointCloud,  # noqa: F401
                  default_params, read_params_yaml)
from . import synthetic  # noqa: F401
from .sequence import FrameToFrameOdometry  # noqa: F401
from .multiframe import BinaryStateGPU, CvoFrameGPU, update_edges, update_edges_sharded  # noqa: F401

__all__ = ["CvoGPU", "CvoPointCloud", "CvoParams", "Association", "CvoError", "Params",
           "IterTrace", "AlignInfo", "default_params", "read_params_yaml", "synthetic",
           "load_library", "FrameToFrameOdometry", "CvoFrameGPU", "BinaryStateGPU", "update_edges", "update_edges_sharded"]
