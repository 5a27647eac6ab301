"""fused_gat_func -- the fused GAT operator behind cogdl.utils.fused_gat_op.

Reference: cogdl/operators/fused_gat.py:3-41 binds dgNN's fused_gatconv, but the binding is stale at
this pin (6 args / 3 outputs vs dgNN's 7 / 4), so it is dead code there.  We keep the
call signature `fused_gat_func(attn_row, attn_col, row_ptr, col_ind, col_ptr, row_ind,
negative_slope, in_feat)` and implement the op natively.  col_ptr/row_ind are accepted for
signature compatibility; the transpose comes from the cached structure.
"""
import torch

from ..structure import structure_for, CSRStructure
from ._raw import gat_fwd_raw, mhspmm_raw, mhsddmm_raw, edge_softmax_bwd_raw


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
        attn_row, attn_col, feat, att = ctx.saved_tensors
        st = ctx.st
        grad_out = grad_out.contiguous()
        st_t, perm = st.csc()
        H = att.shape[1]
        grad_feat = mhspmm_raw(st_t, att, grad_out, perm=perm)            # [N,H,F]
        grad_att = mhsddmm_raw(st, grad_out, feat)                        # [E,H]
        grad_e = edge_softmax_bwd_raw(st, att, grad_att)                  # d/d leakyrelu output
        # leakyrelu'(z) with z = attn_row[row] + attn_col[col]: recompute the sign from the inputs
        rows = torch.repeat_interleave(torch.arange(st.n_rows, device=att.device),
                                       (st.rowptr[1:] - st.rowptr[:-1]).long())
        z = attn_row[rows] + attn_col[st.colind.long()]
        grad_z = torch.where(z > 0, grad_e, grad_e * ctx.slope).contiguous()
        ones = torch.ones((st.n_cols, H, 1), dtype=torch.float32, device=att.device)
        g_row = mhspmm_raw(st, grad_z, ones[: st.n_cols]).view(st.n_rows, H)       # sum over a row's edges
        g_col = mhspmm_raw(st_t, grad_z, torch.ones((st.n_rows, H, 1), dtype=torch.float32, device=att.device),
                           perm=perm).view(st.n_cols, H)                          # sum over a column's edges
        return g_row, g_col, None, None, None, None, None, grad_feat


def fused_gat_func(attn_row, attn_col, row_ptr, col_ind, col_ptr, row_ind, negative_slope, in_feat):
    return FusedGATFunction.apply(attn_row, attn_col, row_ptr, col_ind, col_ptr, row_ind, negative_slope, in_feat)
