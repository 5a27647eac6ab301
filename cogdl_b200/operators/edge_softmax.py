"""csr_edge_softmax -- drop-in for cogdl.operators.edge_softmax (cogdl/operators/edge_softmax.py:17-38)."""
import torch

from ..structure import structure_for, CSRStructure
from ._raw import edge_softmax_fwd_raw, edge_softmax_bwd_raw


class EdgeSoftmaxFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rowptr, h, colind=None):
        if isinstance(rowptr, CSRStructure):
            st = rowptr
        else:
            # the reference passes only rowptr; a structure needs colind too, but edge softmax never
            # reads it: a zero-length stand-in keeps the (rowptr-keyed) plan cache working.
            ci = colind if colind is not None else _dummy_colind(rowptr, h.shape[0])
            st = structure_for(rowptr, ci)
        out = edge_softmax_fwd_raw(st, h)
        ctx.st = st
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (out,) = ctx.saved_tensors
        return None, edge_softmax_bwd_raw(ctx.st, out, grad_out.contiguous()), None


_DUMMY = {}


def _dummy_colind(rowptr, nnz):
    key = (str(rowptr.device), nnz, rowptr.dtype)
    t = _DUMMY.get(key)
    if t is None:
        if len(_DUMMY) > 8:
            _DUMMY.clear()
        t = torch.zeros(nnz, dtype=rowptr.dtype, device=rowptr.device)
        _DUMMY[key] = t
    return t


def csr_edge_softmax(rowptr, h):
    return EdgeSoftmaxFunction.apply(rowptr, h)
