"""cogdl_b200 -- B200 (sm_100a) native implementation of CogDL's sparse message-passing hot path:
CSR SpMM (GE-SpMM), edge-softmax, multi-head SpMM / SDDMM and scatter_max, behind CogDL's own call
signatures.  Hand-written CUDA in cogdl_b200/csrc reached through the C ABI of
include/cogdl_b200.h; PyTorch only provides device memory, streams and torch.distributed.

There is no CPU or PyTorch fallback: importing this package loads libcogdl_b200.so and fails
loudly when it has not been built.
"""
from . import _cabi

_cabi.load()  # fail at import time, not at first use

from .structure import CSRStructure, HubPlan, structure_for  # noqa: E402,F401
from .data import Graph, Adjacency  # noqa: E402,F401
from .utils.spmm_utils import (  # noqa: E402,F401
    spmm, edge_softmax, mh_spmm, fused_gat_op, SpMM, EdgeSoftmax, MultiHeadSpMM, FusedGATOp, CONFIGS,
)
from .operators import csrspmm, csr_edge_softmax, csrmhspmm, scatter_max, fused_gat_func  # noqa: E402,F401
from .install import install  # noqa: E402,F401
from . import sampling  # noqa: E402,F401

__version__ = "0.1.0"
