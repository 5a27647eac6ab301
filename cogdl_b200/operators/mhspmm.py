"""csrmhspmm -- drop-in for cogdl.operators.mhspmm.csrmhspmm (cogdl/operators/mhspmm.py:34-64)."""
import torch

from ..structure import structure_for, CSRStructure
from ._raw import mhspmm_raw, mhsddmm_raw


class MHSPMMFunction(torch.autograd.Function):
    """out[i,h,:] = sum_p attention[p,h] * feat[col[p],h,:]  -> [N, H, F].
    Backward: grad_feat through the cached transpose with the edge permutation fused into the
    kernel (the reference permutes `arange(E).float()` through cuSPARSE and materialises
    attention[perm], mhspmm.py:55-61 -- inexact beyond 2^24 edges); grad_attention = mhsddmm."""

    @staticmethod
    def forward(ctx, rowptr, colind, feat, attention):
        st = rowptr if isinstance(rowptr, CSRStructure) else structure_for(rowptr, colind, feat.shape[0])
        out = mhspmm_raw(st, attention, feat)
        ctx.st = st
        ctx.save_for_backward(feat, attention)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        feat, attention = ctx.saved_tensors
        st = ctx.st
        grad_out = grad_out.contiguous()
        grad_feat = grad_att = None
        if ctx.needs_input_grad[2]:
            st_t, perm = st.csc()
            grad_feat = mhspmm_raw(st_t, attention, grad_out, perm=perm)
        if ctx.needs_input_grad[3]:
            grad_att = mhsddmm_raw(st, grad_out, feat)
        return None, None, grad_feat, grad_att


def csrmhspmm(rowptr, colind, feat, attention):
    return MHSPMMFunction.apply(rowptr, colind, feat, attention)
