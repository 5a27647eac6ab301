"""Device-resident CSR structure in the kernels' format (int32), with everything that is derived
from the structure alone cached next to it: the hub plan and the transpose (CSC + permutation).

Why this exists: the reference re-casts `row_ptr.int(), col_indices.int()` on EVERY spmm call
(cogdl/utils/spmm_utils.py:106) and re-runs cuSPARSE csr2csc on EVERY backward of a non-symmetric
graph (cogdl/operators/spmm.py:67).  Here both are computed once per structure.

All device work goes through the C ABI (cogdl_b200._cabi); torch only owns the memory.
"""
import ctypes
from collections import OrderedDict

import torch

from . import _cabi

import os

# rows with more edges than this are cut into chunks of this many edges (hub plan)
DEFAULT_CHUNK_EDGES = int(os.environ.get("COGDL_B200_CHUNK_EDGES", "64"))
# rows + edges streamed by one warp of the row-stream kernels (0 disables the stream form)
DEFAULT_SEG_COST = int(os.environ.get("COGDL_B200_SEG_COST", "128"))


_EMPTY = {}


def _ptr(t):
    """Raw device pointer of a tensor (None -> NULL).  A zero-element tensor has data_ptr() == 0;
    the C ABI treats NULL as "argument missing", so hand it a valid dummy address instead."""
    if t is None:
        return None
    if t.numel() == 0 and t.is_cuda:
        d = _EMPTY.get(t.device)
        if d is None:
            d = _EMPTY[t.device] = torch.zeros(16, dtype=torch.uint8, device=t.device)
        return ctypes.c_void_p(d.data_ptr())
    return ctypes.c_void_p(t.data_ptr())


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    """No CPU fallback: fail loudly when handed a CPU tensor (the reference silently takes a
    slow torch path, cogdl/operators/spmm.py:11-40 `except Exception: csrspmm = None`)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "cogdl_b200 operators run on CUDA (sm_100a) only and have no CPU fallback; "
                f"got a tensor on {t.device}"
            )
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


class HubPlan:
    """Which rows are cut into edge chunks (see cogdl_b200_hub_plan_t in include/cogdl_b200.h)."""

    def __init__(self, rowptr32, chunk_edges=DEFAULT_CHUNK_EDGES, nnz=None, seg_cost=None):
        dev = rowptr32.device
        n_rows = rowptr32.numel() - 1
        self.chunk_edges = int(chunk_edges)
        self.seg_cost = int(DEFAULT_SEG_COST if seg_cost is None else seg_cost)
        if nnz is None or n_rows == 0:
            self.seg_cost = 0
        self.device = dev
        self.segs = self.edge_row = None
        with torch.cuda.device(dev):
            counts = torch.empty(4, dtype=torch.int32, device=dev)
            _cabi.call("cogdl_b200_hub_plan_count", _ptr(rowptr32), n_rows, self.chunk_edges, self.seg_cost,
                       _ptr(counts), _stream(dev))
            n_hub, n_chunks, n_empty, n_segs = (int(v) for v in counts.tolist())  # one sync per structure
            self.n_hub_rows, self.n_chunks, self.n_empty_rows, self.n_segs = n_hub, n_chunks, n_empty, n_segs
            self.hub_rows = torch.empty(max(n_hub, 1), dtype=torch.int32, device=dev)
            self.chunks = torch.empty(max(2 * n_chunks, 2), dtype=torch.int32, device=dev)
            self.counters = torch.zeros(max(n_chunks, 1), dtype=torch.int32, device=dev)
            if self.seg_cost > 0 and n_segs > 0:
                self.segs = torch.empty(2 * n_segs, dtype=torch.int32, device=dev)
                self.edge_row = torch.empty(int(nnz), dtype=torch.int32, device=dev)
                _cabi.call("cogdl_b200_edge_rows", _ptr(rowptr32), n_rows, int(nnz), _ptr(self.edge_row), _stream(dev))
            else:
                self.n_segs = 0
            if n_chunks > 0 or self.n_segs > 0:
                _cabi.call("cogdl_b200_hub_plan_fill", _ptr(rowptr32), n_rows, self.chunk_edges,
                           self.seg_cost if self.segs is not None else 0, _ptr(counts), _ptr(self.hub_rows),
                           _ptr(self.chunks), _ptr(self.segs), _stream(dev))
            # hub rows by descending degree + a host copy of the degrees (row-tiered ops size their
            # block / cluster launches from it); the chunk table refers to rows, not to list positions
            self.hub_degrees_host = None
            if n_hub > 0:
                hr = self.hub_rows[:n_hub].long()
                deg = (rowptr32[hr + 1] - rowptr32[hr]).to(torch.int32)
                order = torch.argsort(deg, descending=True, stable=True)
                self.hub_rows[:n_hub] = self.hub_rows[:n_hub][order]
                self._hub_deg_np = deg[order].cpu().numpy().astype("int32")
                self.hub_degrees_host = self._hub_deg_np.ctypes.data

    def struct(self, partial_bytes=0):
        """ctypes struct for one call; partial scratch comes from torch's caching allocator.
        Returns (struct, keepalive)."""
        s = _cabi.HubPlanStruct()
        s.chunk_edges = self.chunk_edges
        s.n_hub_rows = self.n_hub_rows
        s.n_chunks = self.n_chunks
        s.hub_rows = self.hub_rows.data_ptr()
        s.chunks = self.chunks.data_ptr()
        s.counters = self.counters.data_ptr()
        s.n_empty_rows = self.n_empty_rows
        s.hub_degrees_host = self.hub_degrees_host
        if self.segs is not None:
            s.seg_cost, s.n_segs = self.seg_cost, self.n_segs
            s.segs, s.edge_row = self.segs.data_ptr(), self.edge_row.data_ptr()
        scratch = None
        if self.n_chunks > 0 and partial_bytes > 0:
            scratch = torch.empty((partial_bytes + 15) // 16 * 4, dtype=torch.float32, device=self.device)
            s.partials = scratch.data_ptr()
            s.partials_bytes = scratch.numel() * 4
        return s, scratch


class CSRStructure:
    """int32 CSR on the device + cached hub plan + cached transpose."""

    def __init__(self, rowptr32, colind32, n_cols=None, chunk_edges=DEFAULT_CHUNK_EDGES, seg_cost=None):
        dev = require_cuda(rowptr32, colind32)
        self.seg_cost = seg_cost
        if rowptr32.dtype != torch.int32 or colind32.dtype != torch.int32:
            raise ValueError("rowptr / colind must be int32 (use CSRStructure.from_int64 for Graph tensors)")
        if not (rowptr32.is_contiguous() and colind32.is_contiguous()):
            raise ValueError("rowptr / colind must be contiguous")
        self.rowptr, self.colind = rowptr32, colind32
        self.device = dev
        self.n_rows = rowptr32.numel() - 1
        self.nnz = colind32.numel()
        self.n_cols = self.n_rows if n_cols is None else int(n_cols)
        self.chunk_edges = int(chunk_edges)
        self._plan = None
        self._csc = None

    @staticmethod
    def from_int64(row_ptr, col, n_cols=None, chunk_edges=DEFAULT_CHUNK_EDGES):
        """Narrow Graph-style int64 (row_ptr, col) once, on the device."""
        dev = require_cuda(row_ptr, col)
        if row_ptr.dtype == torch.int32 and col.dtype == torch.int32:
            return CSRStructure(row_ptr.contiguous(), col.contiguous(), n_cols, chunk_edges)
        if row_ptr.dtype != torch.int64 or col.dtype != torch.int64:
            raise ValueError("row_ptr / col must both be int64 or both be int32")
        if col.numel() >= 2**31 or row_ptr.numel() >= 2**31:
            raise ValueError("graph too large for int32 indices on one device")
        row_ptr, col = row_ptr.contiguous(), col.contiguous()
        rp = torch.empty(row_ptr.numel(), dtype=torch.int32, device=dev)
        ci = torch.empty(col.numel(), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _cabi.call("cogdl_b200_narrow_i64_i32", _ptr(row_ptr), _ptr(rp), row_ptr.numel(), _stream(dev))
            _cabi.call("cogdl_b200_narrow_i64_i32", _ptr(col), _ptr(ci), col.numel(), _stream(dev))
        return CSRStructure(rp, ci, n_cols, chunk_edges)

    @property
    def plan(self):
        if self._plan is None:
            self._plan = HubPlan(self.rowptr, self.chunk_edges, nnz=self.nnz, seg_cost=self.seg_cost)
        return self._plan

    def plan_struct(self, partial_bytes=0):
        """(ctypes pointer or None, keepalive) for one kernel call."""
        if self.chunk_edges <= 0:
            return None, None
        s, scratch = self.plan.struct(partial_bytes)
        return ctypes.byref(s), (s, scratch)

    def csc(self):
        """(CSRStructure of A^T, perm) with perm[q] = CSR position of CSC entry q.  Cached."""
        if self._csc is None:
            dev = self.device
            with torch.cuda.device(dev):
                colptr = torch.empty(self.n_cols + 1, dtype=torch.int32, device=dev)
                rowind = torch.empty(self.nnz, dtype=torch.int32, device=dev)
                perm = torch.empty(self.nnz, dtype=torch.int32, device=dev)
                wbytes = int(_cabi.load().cogdl_b200_csr2csc_workspace_bytes(self.nnz, self.n_cols))
                ws = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=dev)
                _cabi.call("cogdl_b200_csr2csc", _ptr(self.rowptr), _ptr(self.colind), self.n_rows, self.n_cols,
                           self.nnz, _ptr(colptr), _ptr(rowind), _ptr(perm), _ptr(ws), ws.numel(), _stream(dev))
            t = CSRStructure(colptr, rowind, n_cols=self.n_rows, chunk_edges=self.chunk_edges, seg_cost=self.seg_cost)
            self._csc = (t, perm)
        return self._csc


# ---------------------------------------------------------------------------------------------
# Structure cache for the operator-level API, where the caller hands us bare (rowptr, colind)
# tensors (reference signature csrspmm(rowptr, colind, x, csr_data, sym), operators/spmm.py:24).
#
# Fast path: keyed on storage identity; the key tensors are kept alive by the entry so an address
# cannot be recycled under a live key, and the in-place version counter catches mutation.
#
# Slow path: the reference's own dispatch passes `graph.row_indptr.int(), graph.col_indices.int()`
# -- FRESH tensors on every call (cogdl/utils/spmm_utils.py:106) -- so identity never matches for a
# module that captured the reference `spmm` before install().  Those calls are matched by CONTENT
# against the most recent entries of the same shape (two device compares + one sync, ~tens of us)
# instead of rebuilding the int32 copy, the hub plan (three kernels + a host sync + an argsort) and
# possibly the transpose, and no entry is created for the short-lived tensors (no dead CSR copies
# pinned by the cache).  The cache is bounded by entries AND by bytes.
# ---------------------------------------------------------------------------------------------
_CACHE = OrderedDict()
_CACHE_CAP = 8
_CACHE_MAX_BYTES = int(os.environ.get("COGDL_B200_STRUCT_CACHE_BYTES", str(8 << 30)))
_CONTENT_SCAN = 4
cache_stats = {"hit": 0, "content_hit": 0, "miss": 0}


def _entry_bytes(st, rp, ci):
    b = rp.numel() * rp.element_size() + ci.numel() * ci.element_size()
    if st.rowptr.data_ptr() != rp.data_ptr():
        b += st.rowptr.numel() * 4 + st.colind.numel() * 4
    return b + 4 * st.nnz  # + edge_row of the plan, roughly


def structure_for(rowptr, colind, n_cols=None):
    key = (rowptr.data_ptr(), colind.data_ptr(), rowptr.numel(), colind.numel(), str(rowptr.device),
           rowptr.dtype, n_cols)
    ent = _CACHE.get(key)
    if ent is not None:
        st, rp, ci, ver = ent
        if ver == (rp._version, ci._version):
            _CACHE.move_to_end(key)
            cache_stats["hit"] += 1
            return st
        del _CACHE[key]
    # content match against recent entries of the same shape (the `.int()`-per-call caller)
    scanned = 0
    for k in reversed(_CACHE):
        if scanned >= _CONTENT_SCAN:
            break
        if k[2:5] != key[2:5] or k[6] != n_cols:
            continue
        scanned += 1
        st, rp, ci, ver = _CACHE[k]
        if ver != (rp._version, ci._version):
            continue
        a_rp = rp if rp.dtype == rowptr.dtype else (st.rowptr if rowptr.dtype == torch.int32 else None)
        a_ci = ci if ci.dtype == colind.dtype else (st.colind if colind.dtype == torch.int32 else None)
        if a_rp is None or a_ci is None:
            continue
        if torch.equal(a_rp, rowptr) and torch.equal(a_ci, colind):
            _CACHE.move_to_end(k)
            cache_stats["content_hit"] += 1
            return st
    cache_stats["miss"] += 1
    st = CSRStructure.from_int64(rowptr, colind, n_cols)
    _CACHE[key] = (st, rowptr, colind, (rowptr._version, colind._version))
    total = sum(_entry_bytes(*e[:3]) for e in _CACHE.values())
    while len(_CACHE) > 1 and (len(_CACHE) > _CACHE_CAP or total > _CACHE_MAX_BYTES):
        _, old = _CACHE.popitem(last=False)
        total -= _entry_bytes(*old[:3])
    return st


def clear_structure_cache():
    _CACHE.clear()
