"""Thin tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw pointers and
the current stream.  No autograd here (see the sibling modules)."""
import torch

from .. import _cabi
from ..structure import CSRStructure, _ptr, _stream, require_cuda


def _f32c(t, name):
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def spmm_raw(st: CSRStructure, val, x):
    """Y = A @ X with A = (st, val).  x: [n_src, F] fp32 or fp16; val: [nnz] or None."""
    dev = require_cuda(x, val)
    if x.dim() != 2:
        raise ValueError("x must be [n_src, F]")
    x = x.contiguous()
    F = x.shape[1]
    if val is not None:
        if val.numel() != st.nnz:
            raise ValueError(f"edge values have {val.numel()} entries, CSR has {st.nnz}")
        val = val.contiguous().view(-1)
    with torch.cuda.device(dev):
        y = torch.empty((st.n_rows, F), dtype=x.dtype, device=dev)
        plan, keep = st.plan_struct(st.plan.n_chunks * F * 4 if st.chunk_edges > 0 else 0)
        if x.dtype == torch.float32:
            if val is not None and val.dtype != torch.float32:
                val = val.float()
            _cabi.call("cogdl_b200_spmm_csr_f32", _ptr(st.rowptr), _ptr(st.colind), _ptr(val), _ptr(x), _ptr(y),
                       st.n_rows, F, plan, _stream(dev))
        elif x.dtype == torch.float16:
            if val is not None and val.dtype != torch.float16:
                val = val.half()
            _cabi.call("cogdl_b200_spmm_csr_f16", _ptr(st.rowptr), _ptr(st.colind), _ptr(val), _ptr(x), _ptr(y),
                       st.n_rows, F, plan, _stream(dev))
        else:
            raise TypeError(f"spmm supports float32 / float16 features, got {x.dtype}")
        del keep
    return y


def spmm_2src_raw(st: CSRStructure, val, x0, x1):
    """Partitioned SpMM: columns < x0.shape[0] read x0, the others read x1 (halo rows)."""
    dev = require_cuda(x0, x1, val)
    x0, x1 = _f32c(x0, "x0"), _f32c(x1, "x1")
    F = x0.shape[1]
    if x1.shape[1] != F:
        raise ValueError("x0 / x1 feature widths differ")
    val = None if val is None else _f32c(val, "val").view(-1)
    with torch.cuda.device(dev):
        y = torch.empty((st.n_rows, F), dtype=torch.float32, device=dev)
        plan, keep = st.plan_struct(st.plan.n_chunks * F * 4 if st.chunk_edges > 0 else 0)
        x1p = x1 if x1.numel() > 0 else x0
        _cabi.call("cogdl_b200_spmm_csr_f32_2src", _ptr(st.rowptr), _ptr(st.colind), _ptr(val), _ptr(x0),
                   x0.shape[0], _ptr(x1p), _ptr(y), st.n_rows, F, plan, _stream(dev))
        del keep
    return y


def sddmm_raw(st: CSRStructure, d1, d2):
    dev = require_cuda(d1, d2)
    d1, d2 = _f32c(d1, "d1"), _f32c(d2, "d2")
    F = d1.shape[1]
    with torch.cuda.device(dev):
        out = torch.empty(st.nnz, dtype=torch.float32, device=dev)
        plan, keep = st.plan_struct(0)
        _cabi.call("cogdl_b200_sddmm_csr_f32", _ptr(st.rowptr), _ptr(st.colind), _ptr(d1), _ptr(d2), _ptr(out),
                   st.n_rows, F, plan, _stream(dev))
        del keep
    return out


def gather_rows_raw(perm, x):
    dev = require_cuda(perm, x)
    x = _f32c(x, "x")
    H = 1 if x.dim() == 1 else x.shape[1]
    with torch.cuda.device(dev):
        out = torch.empty((perm.numel(),) + tuple(x.shape[1:]), dtype=torch.float32, device=dev)
        _cabi.call("cogdl_b200_gather_rows_f32", _ptr(perm), _ptr(x), _ptr(out), perm.numel(), H, _stream(dev))
    return out


def _es_scratch(st, H):
    """Bytes of plan scratch the split-row softmax needs (per hub chunk and head: (max, sum) + a partial)."""
    return st.plan.n_chunks * H * 12 if st.chunk_edges > 0 else 0


def edge_softmax_fwd_raw(st: CSRStructure, e):
    dev = require_cuda(e)
    e = _f32c(e, "edge values")
    if e.dim() != 2 or e.shape[0] != st.nnz:
        raise ValueError(f"edge values must be [nnz={st.nnz}, H], got {tuple(e.shape)}")
    with torch.cuda.device(dev):
        out = torch.empty_like(e)
        plan, keep = st.plan_struct(_es_scratch(st, e.shape[1]))
        _cabi.call("cogdl_b200_edge_softmax_fwd_f32", _ptr(st.rowptr), _ptr(e), _ptr(out), st.n_rows, e.shape[1],
                   plan, _stream(dev))
        del keep
    return out


def edge_softmax_bwd_raw(st: CSRStructure, y, g):
    dev = require_cuda(y, g)
    y, g = _f32c(y, "y"), _f32c(g, "g")
    with torch.cuda.device(dev):
        out = torch.empty_like(y)
        plan, keep = st.plan_struct(_es_scratch(st, y.shape[1]))
        _cabi.call("cogdl_b200_edge_softmax_bwd_f32", _ptr(st.rowptr), _ptr(y), _ptr(g), _ptr(out), st.n_rows,
                   y.shape[1], plan, _stream(dev))
        del keep
    return out


def mhspmm_raw(st: CSRStructure, att, feat, perm=None):
    """out[i,h,:] = sum_p att[P(p),h] * feat[col[p],h,:];  feat [n_src,H,F], att [nnz,H]."""
    dev = require_cuda(att, feat, perm)
    att, feat = _f32c(att, "attention"), _f32c(feat, "feat")
    if feat.dim() != 3:
        raise ValueError("feat must be [N, H, F]")
    H, F = feat.shape[1], feat.shape[2]
    if att.shape != (st.nnz, H):
        raise ValueError(f"attention must be [nnz={st.nnz}, H={H}], got {tuple(att.shape)}")
    with torch.cuda.device(dev):
        out = torch.empty((st.n_rows, H, F), dtype=torch.float32, device=dev)
        plan, keep = st.plan_struct(st.plan.n_chunks * H * F * 4 if st.chunk_edges > 0 else 0)
        _cabi.call("cogdl_b200_mhspmm_f32", _ptr(st.rowptr), _ptr(st.colind), _ptr(perm), _ptr(att), _ptr(feat),
                   _ptr(out), st.n_rows, H, F, plan, _stream(dev))
        del keep
    return out


def mhsddmm_raw(st: CSRStructure, grad, feat):
    dev = require_cuda(grad, feat)
    grad, feat = _f32c(grad, "grad"), _f32c(feat, "feat")
    H, F = feat.shape[1], feat.shape[2]
    with torch.cuda.device(dev):
        out = torch.empty((st.nnz, H), dtype=torch.float32, device=dev)
        plan, keep = st.plan_struct(0)
        _cabi.call("cogdl_b200_mhsddmm_f32", _ptr(st.rowptr), _ptr(st.colind), _ptr(grad), _ptr(feat), _ptr(out),
                   st.n_rows, H, F, plan, _stream(dev))
        del keep
    return out


def scatter_max_fwd_raw(st: CSRStructure, x):
    dev = require_cuda(x)
    x = _f32c(x, "feat")
    F = x.shape[1]
    with torch.cuda.device(dev):
        out = torch.empty((st.n_rows, F), dtype=torch.float32, device=dev)
        arg = torch.empty((st.n_rows, F), dtype=torch.int32, device=dev)
        plan, keep = st.plan_struct(st.plan.n_chunks * F * 8 if st.chunk_edges > 0 else 0)
        _cabi.call("cogdl_b200_scatter_max_fwd_f32", _ptr(st.rowptr), _ptr(st.colind), _ptr(x), _ptr(out),
                   _ptr(arg), st.n_rows, F, plan, _stream(dev))
        del keep
    return out, arg


def scatter_max_bwd_raw(grad, argmax, n_src):
    dev = require_cuda(grad, argmax)
    grad = _f32c(grad, "grad")
    n, F = grad.shape
    with torch.cuda.device(dev):
        gx = torch.empty((n_src, F), dtype=torch.float32, device=dev)
        _cabi.call("cogdl_b200_scatter_max_bwd_f32", _ptr(grad), _ptr(argmax), _ptr(gx), n, n_src, F, _stream(dev))
    return gx


def gat_fwd_raw(st: CSRStructure, h_l, h_r, feat, slope, want_att):
    dev = require_cuda(h_l, h_r, feat)
    h_l, h_r, feat = _f32c(h_l, "attn_row"), _f32c(h_r, "attn_col"), _f32c(feat, "in_feat")
    H, F = feat.shape[1], feat.shape[2]
    with torch.cuda.device(dev):
        out = torch.empty((st.n_rows, H, F), dtype=torch.float32, device=dev)
        att = torch.empty((st.nnz, H), dtype=torch.float32, device=dev)   # output for training, scratch otherwise
        plan, keep = st.plan_struct(max(st.plan.n_chunks * H * F * 4, _es_scratch(st, H)) if st.chunk_edges > 0 else 0)
        _cabi.call("cogdl_b200_gat_fwd_f32", _ptr(st.rowptr), _ptr(st.colind), _ptr(h_l), _ptr(h_r), _ptr(feat),
                   float(slope), _ptr(out), _ptr(att), st.n_rows, H, F, plan, _stream(dev))
        del keep
    return out, att


def gat_attn_bwd_raw(st: CSRStructure, att, d_att, h_l, h_r, slope):
    """(d_edge [nnz,H], g_row [n_rows,H]): softmax backward * LeakyReLU' and its row sums in one pass."""
    dev = require_cuda(att, d_att, h_l, h_r)
    att, d_att, h_l, h_r = _f32c(att, "att"), _f32c(d_att, "d_att"), _f32c(h_l, "attn_row"), _f32c(h_r, "attn_col")
    H = att.shape[1]
    with torch.cuda.device(dev):
        d_edge = torch.empty_like(att)
        g_row = torch.empty((st.n_rows, H), dtype=torch.float32, device=dev)
        plan, keep = st.plan_struct(_es_scratch(st, H))
        _cabi.call("cogdl_b200_gat_attn_bwd_f32", _ptr(st.rowptr), _ptr(st.colind), _ptr(att), _ptr(d_att), _ptr(h_l),
                   _ptr(h_r), float(slope), _ptr(d_edge), _ptr(g_row), st.n_rows, H, plan, _stream(dev))
        del keep
    return d_edge, g_row


def edge_colsum_raw(st_t: CSRStructure, perm, e):
    """out[j,h] = sum of e[p,h] over the edges p whose column is j (st_t = cached transpose, perm its permutation)."""
    dev = require_cuda(perm, e)
    e = _f32c(e, "edge values")
    H = e.shape[1]
    with torch.cuda.device(dev):
        out = torch.empty((st_t.n_rows, H), dtype=torch.float32, device=dev)
        plan, keep = st_t.plan_struct(st_t.plan.n_chunks * H * 4 if st_t.chunk_edges > 0 else 0)
        _cabi.call("cogdl_b200_edge_colsum_f32", _ptr(st_t.rowptr), _ptr(perm), _ptr(e), _ptr(out), st_t.n_rows, H, plan,
                   _stream(dev))
        del keep
    return out
