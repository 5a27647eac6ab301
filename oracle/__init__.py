"""CPU oracle for the CogDL sparse message-passing hot path -- TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
legs.  The product package (cogdl_b200/) never imports this module; a product path that routed
through it would void every parity claim.

Two layers:
  * `liboracle.so`  (oracle.c)     -- our C restatement, numpy in / numpy out, via ctypes.
  * `oracle/_ref/`  (build_ref.py) -- the reference's own sources compiled in place:
        ref_module("spmm_cpu", "asis"|"o3")  -> the reference CPU SpMM (torch extension)
        ref_module("sampler",  "asis"|"o3")  -> coo2csr_cpu_index
        ref_module("<cuda op>", "cuda")      -> the reference CUDA kernels built for sm_100a
"""
import ctypes
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f32p = ctypes.POINTER(ctypes.c_float)
_i64 = ctypes.c_int64


def build(force=False):
    """Compile oracle.c -> liboracle.so (gcc, a second or two)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _a(x, dtype):
    x = np.ascontiguousarray(x, dtype=dtype)
    return x


def _p(x, ptype):
    return None if x is None else x.ctypes.data_as(ptype)


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


# ----------------------------------------------------------------------------- ops


def spmm_csr(rowptr, colind, val, X):
    rowptr, colind, X = _a(rowptr, np.int32), _a(colind, np.int32), _a(X, np.float32)
    val = None if val is None else _a(val, np.float32)
    n, F = rowptr.shape[0] - 1, X.shape[1]
    Y = np.empty((n, F), np.float32)
    lib().oracle_spmm_csr_f32(_p(rowptr, _i32p), _p(colind, _i32p), _p(val, _f32p), _p(X, _f32p),
                              _p(Y, _f32p), _i64(n), _i64(F))
    return Y


def sddmm_csr(rowptr, colind, D1, D2):
    rowptr, colind = _a(rowptr, np.int32), _a(colind, np.int32)
    D1, D2 = _a(D1, np.float32), _a(D2, np.float32)
    n, F = rowptr.shape[0] - 1, D1.shape[1]
    out = np.empty(colind.shape[0], np.float32)
    lib().oracle_sddmm_csr_f32(_p(rowptr, _i32p), _p(colind, _i32p), _p(D1, _f32p), _p(D2, _f32p),
                               _p(out, _f32p), _i64(n), _i64(F))
    return out


def edge_softmax_fwd(rowptr, e):
    rowptr, e = _a(rowptr, np.int32), _a(e, np.float32)
    squeeze = e.ndim == 1
    e2 = e.reshape(e.shape[0], 1 if squeeze else e.shape[1])
    out = np.zeros_like(e2)
    lib().oracle_edge_softmax_fwd_f32(_p(rowptr, _i32p), _p(e2, _f32p), _p(out, _f32p),
                                      _i64(rowptr.shape[0] - 1), _i64(e2.shape[1]))
    return out.reshape(-1) if squeeze else out


def edge_softmax_bwd(rowptr, y, g):
    rowptr, y, g = _a(rowptr, np.int32), _a(y, np.float32), _a(g, np.float32)
    hh = 1 if y.ndim == 1 else y.shape[1]
    y2, g2 = y.reshape(y.shape[0], hh), g.reshape(g.shape[0], hh)
    out = np.zeros_like(y2)
    lib().oracle_edge_softmax_bwd_f32(_p(rowptr, _i32p), _p(y2, _f32p), _p(g2, _f32p), _p(out, _f32p),
                                      _i64(rowptr.shape[0] - 1), _i64(y2.shape[1]))
    return out.reshape(y.shape)


def mhspmm(rowptr, colind, att, feat, perm=None):
    rowptr, colind = _a(rowptr, np.int32), _a(colind, np.int32)
    att, feat = _a(att, np.float32), _a(feat, np.float32)
    perm = None if perm is None else _a(perm, np.int32)
    n, H, F = rowptr.shape[0] - 1, feat.shape[1], feat.shape[2]
    out = np.empty((n, H, F), np.float32)
    lib().oracle_mhspmm_f32(_p(rowptr, _i32p), _p(colind, _i32p), _p(perm, _i32p), _p(att, _f32p),
                            _p(feat, _f32p), _p(out, _f32p), _i64(n), _i64(H), _i64(F))
    return out


def mhsddmm(rowptr, colind, grad, feat):
    rowptr, colind = _a(rowptr, np.int32), _a(colind, np.int32)
    grad, feat = _a(grad, np.float32), _a(feat, np.float32)
    n, H, F = rowptr.shape[0] - 1, feat.shape[1], feat.shape[2]
    out = np.empty((colind.shape[0], H), np.float32)
    lib().oracle_mhsddmm_f32(_p(rowptr, _i32p), _p(colind, _i32p), _p(grad, _f32p), _p(feat, _f32p),
                             _p(out, _f32p), _i64(n), _i64(H), _i64(F))
    return out


def gather_rows(perm, x):
    perm, x = _a(perm, np.int32), _a(x, np.float32)
    x2 = x.reshape(x.shape[0], int(np.prod(x.shape[1:])) if x.ndim > 1 else 1)
    out = np.empty((perm.shape[0], x2.shape[1]), np.float32)
    lib().oracle_gather_rows_f32(_p(perm, _i32p), _p(x2, _f32p), _p(out, _f32p),
                                 _i64(perm.shape[0]), _i64(x2.shape[1]))
    return out.reshape((perm.shape[0],) + x.shape[1:])


def scatter_max_fwd(rowptr, colind, X, reference_semantics=False):
    rowptr, colind, X = _a(rowptr, np.int32), _a(colind, np.int32), _a(X, np.float32)
    n, F = rowptr.shape[0] - 1, X.shape[1]
    out = np.empty((n, F), np.float32)
    arg = np.empty((n, F), np.int32)
    lib().oracle_scatter_max_fwd_f32(_p(rowptr, _i32p), _p(colind, _i32p), _p(X, _f32p), _p(out, _f32p),
                                     _p(arg, _i32p), _i64(n), _i64(F), ctypes.c_int(int(reference_semantics)))
    return out, arg


def scatter_max_bwd(grad, argmax, n_src=None):
    grad, argmax = _a(grad, np.float32), _a(argmax, np.int32)
    n, F = grad.shape
    n_src = n if n_src is None else n_src
    gx = np.empty((n_src, F), np.float32)
    lib().oracle_scatter_max_bwd_f32(_p(grad, _f32p), _p(argmax, _i32p), _p(gx, _f32p),
                                     _i64(n), _i64(n_src), _i64(F))
    return gx


def csr2csc(rowptr, colind, n_cols=None):
    rowptr, colind = _a(rowptr, np.int32), _a(colind, np.int32)
    n = rowptr.shape[0] - 1
    n_cols = n if n_cols is None else n_cols
    colptr = np.empty(n_cols + 1, np.int32)
    rowind = np.empty(colind.shape[0], np.int32)
    perm = np.empty(colind.shape[0], np.int32)
    lib().oracle_csr2csc(_p(rowptr, _i32p), _p(colind, _i32p), _p(colptr, _i32p), _p(rowind, _i32p),
                         _p(perm, _i32p), _i64(n), _i64(n_cols))
    return colptr, rowind, perm


def coo2csr_index(row, num_nodes):
    row = _a(row, np.int64)
    row_ptr = np.empty(num_nodes + 1, np.int64)
    reindex = np.empty(row.shape[0], np.int64)
    lib().oracle_coo2csr_index(_p(row, _i64p), _i64(row.shape[0]), _i64(num_nodes), _p(row_ptr, _i64p),
                               _p(reindex, _i64p))
    return row_ptr, reindex


def gat_fwd(rowptr, colind, h_l, h_r, feat, slope=0.2, return_att=False):
    rowptr, colind = _a(rowptr, np.int32), _a(colind, np.int32)
    h_l, h_r, feat = _a(h_l, np.float32), _a(h_r, np.float32), _a(feat, np.float32)
    n, H, F = rowptr.shape[0] - 1, feat.shape[1], feat.shape[2]
    out = np.empty((n, H, F), np.float32)
    att = np.zeros((colind.shape[0], H), np.float32) if return_att else None
    lib().oracle_gat_fwd_f32(_p(rowptr, _i32p), _p(colind, _i32p), _p(h_l, _f32p), _p(h_r, _f32p),
                             _p(feat, _f32p), ctypes.c_float(slope), _p(out, _f32p), _p(att, _f32p),
                             _i64(n), _i64(H), _i64(F))
    return (out, att) if return_att else out


# ----------------------------------------------------------------------------- reference builds


def ref_path(name, variant):
    return os.path.join(_HERE, "_ref", variant, name + ".so")


def ref_available(name, variant):
    return os.path.exists(ref_path(name, variant))


def ref_module(name, variant="asis"):
    """Import a module compiled by build_ref.py from the reference's own sources
    (a torch pybind extension; `import torch` first so libtorch is resolvable)."""
    import torch  # noqa: F401

    path = ref_path(name, variant)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path}: run `python oracle/build_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ----------------------------------------------------------------------------- sampler
_u64 = ctypes.c_uint64


def sample_draw(seed, slot, k):
    f = lib().oracle_sample_draw
    f.restype = _u64
    return int(f(_u64(seed), _i64(slot), _i64(k)))


def sample_adj(indptr, indices, node_idx, num_neighbors=-1, replace=True, seed=0, floyd_variant=0):
    """(out_indptr, out_indices, out_nodes, out_edges) -- cogdl/operators/sample/sample.cpp:6-146 with the
    counter-based generator shared with the CUDA path (see oracle.c)."""
    indptr, indices, node_idx = _a(indptr, np.int64), _a(indices, np.int64), _a(node_idx, np.int64)
    n, nb = indptr.shape[0] - 1, node_idx.shape[0]
    deg = indptr[node_idx + 1] - indptr[node_idx]
    cap = int(deg.sum()) if num_neighbors < 0 else int(nb * num_neighbors)
    out_indptr = np.zeros(nb + 1, np.int64)
    out_indices = np.empty(max(cap, 1), np.int64)
    out_edges = np.empty(max(cap, 1), np.int64)
    out_nodes = np.empty(nb + max(cap, 1), np.int64)
    f = lib().oracle_sample_adj
    f.restype = _i64
    n_nodes = f(_p(indptr, _i64p), _p(indices, _i64p), _p(node_idx, _i64p), _i64(nb), _i64(n), _i64(num_neighbors),
                ctypes.c_int(int(replace)), _u64(seed), ctypes.c_int(floyd_variant), _p(out_indptr, _i64p),
                _p(out_indices, _i64p), _p(out_edges, _i64p), _p(out_nodes, _i64p), _i64(cap))
    assert n_nodes >= 0
    ne = int(out_indptr[-1])
    return out_indptr, out_indices[:ne].copy(), out_nodes[:n_nodes].copy(), out_edges[:ne].copy()


def subgraph(indptr, indices, node_idx):
    """(out_indptr, out_indices, out_edges) -- cogdl/operators/sample/sample.cpp:148-188."""
    indptr, indices, node_idx = _a(indptr, np.int64), _a(indices, np.int64), _a(node_idx, np.int64)
    n, ns = indptr.shape[0] - 1, node_idx.shape[0]
    cap = max(int((indptr[node_idx + 1] - indptr[node_idx]).sum()), 1)
    out_indptr = np.zeros(ns + 1, np.int64)
    out_indices, out_edges = np.empty(cap, np.int64), np.empty(cap, np.int64)
    f = lib().oracle_subgraph
    f.restype = _i64
    ne = f(_p(indptr, _i64p), _p(indices, _i64p), _p(node_idx, _i64p), _i64(ns), _i64(n), _p(out_indptr, _i64p),
           _p(out_indices, _i64p), _p(out_edges, _i64p))
    return out_indptr, out_indices[:ne].copy(), out_edges[:ne].copy()


def gat_attn_bwd(rowptr, colind, att, d_att, h_l, h_r, slope=0.2, n_cols=None):
    """(d_edge [nnz,H], g_row [n,H], g_col [n_cols,H]) -- autograd of gat_layer.py:73-74, see oracle.c."""
    rowptr, colind = _a(rowptr, np.int32), _a(colind, np.int32)
    att, d_att, h_l, h_r = _a(att, np.float32), _a(d_att, np.float32), _a(h_l, np.float32), _a(h_r, np.float32)
    n, H = rowptr.shape[0] - 1, att.shape[1]
    n_cols = h_r.shape[0] if n_cols is None else n_cols
    d_edge = np.zeros_like(att)
    g_row = np.zeros((n, H), np.float32)
    g_col = np.zeros((n_cols, H), np.float32)
    lib().oracle_gat_attn_bwd_f32(_p(rowptr, _i32p), _p(colind, _i32p), _p(att, _f32p), _p(d_att, _f32p), _p(h_l, _f32p),
                                  _p(h_r, _f32p), ctypes.c_float(slope), _p(d_edge, _f32p), _p(g_row, _f32p),
                                  _p(g_col, _f32p), _i64(n), _i64(n_cols), _i64(H))
    return d_edge, g_row, g_col
