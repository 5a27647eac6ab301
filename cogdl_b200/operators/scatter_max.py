"""scatter_max -- drop-in for cogdl.operators.scatter_max.scatter_max (cogdl/operators/scatter_max.py:17-37)."""
import torch

from ..structure import structure_for, CSRStructure
from ._raw import scatter_max_fwd_raw, scatter_max_bwd_raw


class ScatterMaxFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rowptr, colind, feat):
        st = rowptr if isinstance(rowptr, CSRStructure) else structure_for(rowptr, colind, feat.shape[0])
        out, max_id = scatter_max_fwd_raw(st, feat)
        ctx.n_src = feat.shape[0]
        ctx.save_for_backward(max_id)
        return out

    @staticmethod
    def backward(ctx, grad):
        (max_id,) = ctx.saved_tensors
        return None, None, scatter_max_bwd_raw(grad.contiguous(), max_id, ctx.n_src)


def scatter_max(rowptr, colind, feat):
    return ScatterMaxFunction.apply(rowptr, colind, feat)


def scatter_max_with_argmax(rowptr, colind, feat):
    """(out, max_id) without autograd -- the pair the reference's scatter_max_fp returns
    (cogdl/operators/scatter_max/scatter_max.cc:6-22)."""
    st = rowptr if isinstance(rowptr, CSRStructure) else structure_for(rowptr, colind, feat.shape[0])
    return scatter_max_fwd_raw(st, feat)
