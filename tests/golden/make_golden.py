#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/*.npz by IMPORTING THE REFERENCE PACKAGE (cogdl from
/root/reference) in this container and running its own public API on CPU.  Run where
/root/reference exists:   python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed as small fixtures together
with this script.  What runs here is the reference's Python code unmodified; the only shims are
  * stub modules for optional third-party imports that are not installed (optuna, matplotlib, grave),
  * `cogdl.operators.spmm` / `cogdl.operators.sample` are pre-seeded with the reference's OWN C++
    sources compiled by oracle/build_ref.py (as-shipped flags) instead of letting the package JIT-
    compile them (its JIT would also try to nvcc-build the CUDA kernels, minutes per module).
Golden sets:
  spmm_cora.npz        cogdl.utils.spmm(graph, x) on a Cora-shaped graph (2708 nodes / 10556 edges,
                       add_remaining_self_loops + sym_norm), F = 16 and 7  [BASELINE configs[0]]
                       -> CPU path = spmm_cpu.cpp (spmm_utils.py:110-119)
  spmm_rownorm.npz     same with row_norm() on a CSR-only graph (in_norm applied around the kernel)
  edge_softmax.npz     cogdl.utils.edge_softmax CPU fallback (spmm_utils.py:149-169,185-188), logits <= 10
  mh_spmm.npz          cogdl.utils.mh_spmm CPU fallback (spmm_utils.py:216-225)
  gcn_layer.npz        cogdl.layers.GCNLayer forward with fixed weights
  gat_layer.npz        cogdl.layers.GATLayer forward (attn_drop=0) with fixed weights
  sage_mean_layer.npz  cogdl.layers.SAGELayer(aggr="mean") forward
  coo2csr.npz          coo2csr_cpu_index (sample.cpp:234-270) on a random COO list
  graph_semantics.npz  Graph: add_remaining_self_loops / sym_norm / row_norm / edge_weight / flags
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference():
    import oracle

    class _Stub(types.ModuleType):
        def __getattr__(self, name):  # any attribute of a missing optional dependency
            if name.startswith("__"):
                raise AttributeError(name)
            return lambda *a, **k: None

    for name in ["optuna", "matplotlib", "matplotlib.cm", "matplotlib.pyplot", "grave"]:
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    spmm_stub = types.ModuleType("cogdl.operators.spmm")
    spmm_stub.csrspmm = None
    spmm_stub.spmm_cpu = oracle.ref_module("spmm_cpu", "asis").csr_spmm_cpu
    sys.modules["cogdl.operators.spmm"] = spmm_stub
    smp = oracle.ref_module("sampler", "asis")
    sample_stub = types.ModuleType("cogdl.operators.sample")
    sample_stub.subgraph_c, sample_stub.sample_adj_c = smp.subgraph, smp.sample_adj
    sample_stub.coo2csr_cpu, sample_stub.coo2csr_cpu_index = smp.coo2csr_cpu, smp.coo2csr_cpu_index
    sys.modules["cogdl.operators.sample"] = sample_stub
    sys.path.insert(0, REF)
    import cogdl  # noqa: F401
    from cogdl.data import Graph
    from cogdl.utils import spmm, edge_softmax, mh_spmm
    from cogdl.layers import GCNLayer, GATLayer, SAGELayer

    return Graph, spmm, edge_softmax, mh_spmm, GCNLayer, GATLayer, SAGELayer, smp


def random_graph(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    row = torch.randint(0, n, (e,), generator=g)
    col = torch.randint(0, n, (e,), generator=g)
    return row, col


def main():
    Graph, spmm, edge_softmax, mh_spmm, GCNLayer, GATLayer, SAGELayer, smp = import_reference()
    torch.manual_seed(0)
    save = lambda name, **kw: np.savez_compressed(os.path.join(HERE, name), **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in kw.items()})

    # ---- spmm on a Cora-shaped graph (configs[0])
    n, e = 2708, 10556
    row, col = random_graph(n, e, 0)
    x16 = torch.randn(n, 16, generator=torch.Generator().manual_seed(1))
    x7 = torch.randn(n, 7, generator=torch.Generator().manual_seed(2))
    g = Graph(x=x16, edge_index=(row, col))
    g.add_remaining_self_loops()
    g.sym_norm()
    with torch.no_grad():
        y16, y7 = spmm(g, x16), spmm(g, x7)
    save("spmm_cora.npz", row=row, col=col, row_ptr=g.row_indptr, col_indices=g.col_indices, weight=g.edge_weight,
         x16=x16, y16=y16, x7=x7, y7=y7, symmetric=np.array(g.is_symmetric()))

    # ---- row_norm on a CSR-only graph: in_norm is applied outside the kernel
    g2 = Graph(x=x16, row_ptr=g.row_indptr.clone(), col=g.col_indices.clone(), num_nodes=n)
    g2._adj.row = None
    g2.row_norm()
    with torch.no_grad():
        y_rn = spmm(g2, x16)
    save("spmm_rownorm.npz", row_ptr=g2.row_indptr, col_indices=g2.col_indices, x=x16, y=y_rn,
         in_norm=g2.in_norm, has_out_norm=np.array(g2.out_norm is not None))

    # ---- edge_softmax / mh_spmm CPU fallbacks on a small graph
    n, e, H, F = 300, 2400, 8, 16
    row, col = random_graph(n, e, 3)
    gs = Graph(x=torch.zeros(n, 1), edge_index=(row, col))
    gs.add_remaining_self_loops()
    E = gs.col_indices.shape[0]
    logits = (torch.randn(E, H, generator=torch.Generator().manual_seed(4)) * 3).clamp(-10, 10)
    with torch.no_grad():
        att = edge_softmax(gs, logits.clone())   # fallback mutates its input when max > 10: never here
        h = torch.randn(n, H, F, generator=torch.Generator().manual_seed(5))
        out = mh_spmm(gs, att, h)
    save("edge_softmax.npz", row_ptr=gs.row_indptr, col_indices=gs.col_indices, logits=logits, att=att)
    save("mh_spmm.npz", row_ptr=gs.row_indptr, col_indices=gs.col_indices, att=att, h=h, out=out)

    # ---- layers with fixed parameters
    n, e = 500, 4000
    row, col = random_graph(n, e, 6)
    x = torch.randn(n, 32, generator=torch.Generator().manual_seed(7))
    gl = Graph(x=x, edge_index=(row, col))
    gl.add_remaining_self_loops()
    gl.sym_norm()
    torch.manual_seed(8)
    gcn = GCNLayer(32, 16, activation="relu")
    gcn.eval()
    with torch.no_grad():
        y = gcn(gl, x)
    save("gcn_layer.npz", row_ptr=gl.row_indptr, col_indices=gl.col_indices, weight=gl.edge_weight, x=x, y=y,
         W=gcn.linear.weight, b=gcn.linear.bias)

    gg = Graph(x=x, edge_index=(row, col))
    gg.add_remaining_self_loops()
    torch.manual_seed(9)
    gat = GATLayer(32, 8, nhead=4, attn_drop=0.0, alpha=0.2)
    gat.eval()
    with torch.no_grad():
        y = gat(gg, x)
    save("gat_layer.npz", row_ptr=gg.row_indptr, col_indices=gg.col_indices, x=x, y=y, W=gat.W, a_l=gat.a_l, a_r=gat.a_r)

    gm = Graph(x=x, edge_index=(row, col))
    gm.add_remaining_self_loops()
    torch.manual_seed(10)
    sage = SAGELayer(32, 16, aggr="mean")
    sage.eval()
    with torch.no_grad():
        y = sage(gm, x)
    save("sage_mean_layer.npz", row_ptr=gm.row_indptr, col_indices=gm.col_indices, x=x, y=y, W=sage.fc.weight, b=sage.fc.bias)

    # ---- coo2csr_cpu_index
    n, e = 1000, 20000
    row, col = random_graph(n, e, 11)
    row_ptr, reindex = smp.coo2csr_cpu_index(row, col, n)
    save("coo2csr.npz", row=row, num_nodes=np.array(n), row_ptr=row_ptr, reindex=reindex)

    # ---- Graph semantics
    n, e = 50, 300
    row, col = random_graph(n, e, 12)
    ga = Graph(x=torch.zeros(n, 1), edge_index=(row, col))
    ga.add_remaining_self_loops()
    rp, ci, w0 = ga.row_indptr.clone(), ga.col_indices.clone(), ga.edge_weight.clone()
    ga.sym_norm()
    w_sym, sym_flag = ga.edge_weight.clone(), ga.is_symmetric()
    gb = Graph(x=torch.zeros(n, 1), edge_index=(row, col))
    gb.add_remaining_self_loops()
    gb.row_norm()
    w_row, row_flag = gb.edge_weight.clone(), gb.is_symmetric()
    gc = Graph(x=torch.zeros(n, 1), edge_index=(row, col))
    gc.add_remaining_self_loops()
    gc.edge_weight = torch.arange(gc.col_indices.shape[0]).float()
    set_flag = gc.is_symmetric()
    save("graph_semantics.npz", row=row, col=col, row_ptr=rp, col_indices=ci, w0=w0, w_sym=w_sym, sym_flag=np.array(sym_flag),
         w_row=w_row, row_flag=np.array(row_flag), set_flag=np.array(set_flag))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
