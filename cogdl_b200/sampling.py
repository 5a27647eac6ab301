"""Device neighbour sampling -- drop-ins for cogdl.operators.sample.{sample_adj_c, subgraph_c}
(cogdl/operators/sample.py:8-13 -> sample.cpp:6-188), which Graph.sample_adj / Graph.csr_subgraph call
per mini-batch (cogdl/data/data.py:792-832, 850-874).

Same call signatures and return tuples as the reference's pybind functions, CUDA tensors in and out
(int64, exactly as Graph stores row_ptr / col).  No CPU path: a CPU tensor raises.  Semantics and the
counter-based generator are documented in include/cogdl_b200.h and csrc/sampler.cu; `seed` is taken from
`set_seed()` and advanced by one per sampled batch, so an epoch is reproducible end to end.
"""
import ctypes

import torch

from . import _cabi
from .structure import _ptr, _stream, require_cuda

UNSEEN = 0x7FFFFFFF
_state = {"seed": 0x5EED, "calls": 0}
_assoc = {}     # (device, num_nodes) -> int32 scratch, all UNSEEN between calls


def set_seed(seed):
    _state["seed"], _state["calls"] = int(seed) & (2**64 - 1), 0


def _scratch(dev, num_nodes):
    key = (str(dev), int(num_nodes))
    a = _assoc.get(key)
    if a is None:
        if len(_assoc) >= 4:
            _assoc.pop(next(iter(_assoc)))
        a = _assoc[key] = torch.full((max(int(num_nodes), 1),), UNSEEN, dtype=torch.int32, device=dev)
    return a


def _i64c(t, name):
    if t.dtype != torch.int64:
        raise TypeError(f"{name} must be int64 (as cogdl.data.Graph stores it), got {t.dtype}")
    return t.contiguous()


def sample_adj(indptr, indices, node_idx, num_neighbors=-1, replace=True, seed=None):
    """-> (out_indptr [B+1], out_indices [E'], out_nodes [n'], out_edges [E']), all int64 on the device."""
    if not torch.is_tensor(node_idx):
        node_idx = torch.as_tensor(node_idx, dtype=torch.int64, device=indptr.device)
    dev = require_cuda(indptr, indices, node_idx)
    indptr, indices, node_idx = _i64c(indptr, "indptr"), _i64c(indices, "indices"), _i64c(node_idx.view(-1), "node_idx")
    n, nb = indptr.numel() - 1, node_idx.numel()
    if seed is None:
        seed = (_state["seed"] + _state["calls"]) & (2**64 - 1)
        _state["calls"] += 1
    lib = _cabi.load()
    with torch.cuda.device(dev):
        out_indptr = torch.empty(nb + 1, dtype=torch.int64, device=dev)
        ws = torch.empty(int(lib.cogdl_b200_sample_workspace_bytes(nb, 0)), dtype=torch.uint8, device=dev)
        _cabi.call("cogdl_b200_sample_adj_count", _ptr(indptr), _ptr(node_idx), nb, int(num_neighbors), int(bool(replace)),
                   _ptr(out_indptr), _ptr(ws), ws.numel(), _stream(dev))
        ne = int(out_indptr[-1])                       # data-dependent size: one 8-byte D2H
        out_indices = torch.empty(ne, dtype=torch.int64, device=dev)
        out_edges = torch.empty(ne, dtype=torch.int64, device=dev)
        out_nodes = torch.empty(nb + ne, dtype=torch.int64, device=dev)
        n_out = torch.empty(1, dtype=torch.int64, device=dev)
        ws = torch.empty(int(lib.cogdl_b200_sample_workspace_bytes(nb, ne)), dtype=torch.uint8, device=dev)
        _cabi.call("cogdl_b200_sample_adj_fill", _ptr(indptr), _ptr(indices), _ptr(node_idx), nb, n, int(num_neighbors),
                   int(bool(replace)), ctypes.c_uint64(seed), _ptr(out_indptr), ne, _ptr(_scratch(dev, n)),
                   _ptr(out_indices), _ptr(out_edges), _ptr(out_nodes), _ptr(n_out), _ptr(ws), ws.numel(), _stream(dev))
        out_nodes = out_nodes[: int(n_out)]
    return out_indptr, out_indices, out_nodes, out_edges


def subgraph(indptr, indices, node_idx):
    """-> (out_indptr [n_sub+1], out_indices [E'], arange(n_sub), out_edges [E'])  (sample.cpp:148-188)."""
    dev = require_cuda(indptr, indices, node_idx)
    indptr, indices, node_idx = _i64c(indptr, "indptr"), _i64c(indices, "indices"), _i64c(node_idx.view(-1), "node_idx")
    n, ns = indptr.numel() - 1, node_idx.numel()
    lib = _cabi.load()
    with torch.cuda.device(dev):
        out_indptr = torch.empty(ns + 1, dtype=torch.int64, device=dev)
        ws = torch.empty(int(lib.cogdl_b200_sample_workspace_bytes(ns, 0)), dtype=torch.uint8, device=dev)
        assoc = _scratch(dev, n)
        _cabi.call("cogdl_b200_subgraph_count", _ptr(indptr), _ptr(indices), _ptr(node_idx), ns, _ptr(assoc), _ptr(out_indptr),
                   _ptr(ws), ws.numel(), _stream(dev))
        ne = int(out_indptr[-1])
        out_indices = torch.empty(ne, dtype=torch.int64, device=dev)
        out_edges = torch.empty(ne, dtype=torch.int64, device=dev)
        _cabi.call("cogdl_b200_subgraph_fill", _ptr(indptr), _ptr(indices), _ptr(node_idx), ns, _ptr(assoc), _ptr(out_indptr),
                   _ptr(out_indices) if ne else None, _ptr(out_edges) if ne else None, _stream(dev))
    return out_indptr, out_indices, torch.arange(ns, device=dev), out_edges


# names the reference binds (cogdl/operators/sample.py:9-10)
sample_adj_c = sample_adj
subgraph_c = subgraph
