"""fused_gat_func -- the fused GAT operator behind cogdl.utils.fused_gat_op.

Reference: cogdl/operators/fused_gat.py:3-41 binds dgNN's fused_gatconv, but the binding is stale at
this pin (6 args / 3 outputs vs dgNN's 7 / 4), so it is dead code there.  We keep the
call signature `fused_gat_func(attn_row, attn_col, row_ptr, col_ind, col_ptr, row_ind,
negative_slope, in_feat)` and implement the op natively.  col_ptr/row_ind are accepted for
signature compatibility; the transpose comes from the cached structure.
"""
import torch

from ..structure import structure_for, CSRStructure
from ._raw import gat_fwd_raw, mhspmm_raw, mhsddmm_raw, gat_attn_bwd_raw, edge_colsum_raw


class FusedGATFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attn_row, attn_col, row_ptr, col_ind, col_ptr, row_ind, negative_slope, in_feat):
        st = row_ptr if isinstance(row_ptr, CSRStructure) else structure_for(row_ptr, col_ind, in_feat.shape[0])
        need_grad = any(t.requires_grad for t in (attn_row, attn_col, in_feat))
        out, att = gat_fwd_raw(st, attn_row, attn_col, in_feat, negative_slope, want_att=need_grad)
        ctx.st, ctx.slope = st, float(negative_slope)
        if need_grad:
            ctx.save_for_backward(attn_row, attn_col, in_feat, att)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        """Native backward (precedent dgNN fused_gatconv_kernel.cu:636-790), five library calls, no eager
        index arithmetic:  d feat = CSC mh-SpMM (perm fused) . d att = mh-SDDMM . (d edge, g_row) = softmax
        backward * LeakyReLU' + row sums in one pass (cogdl_b200_gat_attn_bwd_f32) . g_col = column sums of
        d edge through the cached transpose (cogdl_b200_edge_colsum_f32)."""
        attn_row, attn_col, feat, att = ctx.saved_tensors
        st = ctx.st
        grad_out = grad_out.contiguous()
        st_t, perm = st.csc()
        grad_feat = mhspmm_raw(st_t, att, grad_out, perm=perm) if ctx.needs_input_grad[7] else None
        g_row = g_col = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            grad_att = mhsddmm_raw(st, grad_out, feat)                        # [E,H]
            d_edge, g_row = gat_attn_bwd_raw(st, att, grad_att, attn_row, attn_col, ctx.slope)
            g_col = edge_colsum_raw(st_t, perm, d_edge)
        return g_row, g_col, None, None, None, None, None, grad_feat


def fused_gat_func(attn_row, attn_col, row_ptr, col_ind, col_ptr, row_ind, negative_slope, in_feat):
    return FusedGATFunction.apply(attn_row, attn_col, row_ptr, col_ind, col_ptr, row_ind, negative_slope, in_feat)
