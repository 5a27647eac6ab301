from .spmm_utils import (  # noqa: F401
    CONFIGS, spmm, SpMM, edge_softmax, EdgeSoftmax, mh_spmm, MultiHeadSpMM, fused_gat_op, FusedGATOp,
    check_fused_gat, initialize_spmm, initialize_edge_softmax, initialize_fused_gat,
)
