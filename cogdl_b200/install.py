"""Register the sm_100a backend inside an importable `cogdl` package (the drop-in step).

Plug-in point: the module-global dict `cogdl.utils.spmm_utils.CONFIGS` and the
`cogdl.operators.*` module attributes; layer modules capture the callables at construction
(SpMM.__init__ cogdl/utils/spmm_utils.py:131, EdgeSoftmax.__init__ :195, MultiHeadSpMM.__init__
:234-235, MaxAggregator.__init__ cogdl/layers/sage_layer.py:22-25), so call install() BEFORE
building models.  Two levels are patched:

  1. operator level -- CONFIGS["fast_spmm" | "csr_edge_softmax" | "csrmhspmm" | "fused_gat_func"]
     and cogdl.operators.{spmm.csrspmm, edge_softmax.csr_edge_softmax, mhspmm.csrmhspmm,
     scatter_max.scatter_max, fused_gat.fused_gat_func}, with the *_flag latches set so the lazy
     initialisers do not overwrite them;
  2. dispatch level -- every already-imported cogdl module attribute that IS the reference's
     spmm / edge_softmax / mh_spmm / fused_gat_op function (or SpMM / EdgeSoftmax / MultiHeadSpMM
     / FusedGATOp class) is rebound to this package's version, which keeps the int32 CSR, hub
     plan and transpose cached on the graph instead of re-casting per call.
"""
import sys


def install(rebind_dispatch=True):
    import cogdl.utils.spmm_utils as ref_su  # raises ImportError if cogdl is not importable

    from . import operators as ops
    from .utils import spmm_utils as su

    ref_su.CONFIGS.update({
        "fast_spmm": ops.csrspmm, "csr_edge_softmax": ops.csr_edge_softmax, "csrmhspmm": ops.csrmhspmm,
        "fused_gat_func": ops.fused_gat_func,
        "spmm_flag": True, "mh_spmm_flag": True, "fused_gat_flag": True,
    })
    import types

    import cogdl.operators as ref_ops_pkg

    for mod_name, attrs in (
        ("cogdl.operators.spmm", {"csrspmm": ops.csrspmm}),
        ("cogdl.operators.edge_softmax", {"csr_edge_softmax": ops.csr_edge_softmax}),
        ("cogdl.operators.mhspmm", {"csrmhspmm": ops.csrmhspmm}),
        ("cogdl.operators.scatter_max", {"scatter_max": ops.scatter_max}),
        ("cogdl.operators.fused_gat", {"fused_gat_func": ops.fused_gat_func}),
    ):
        mod = sys.modules.get(mod_name)
        if mod is None:
            # Not imported yet: seed a module that already carries our operator, so a later
            # `from cogdl.operators.scatter_max import scatter_max` (sage_layer.py:23) resolves to it
            # and the reference's import-time JIT build of its own CUDA sources never runs.
            mod = types.ModuleType(mod_name)
            mod.__doc__ = "seeded by cogdl_b200.install()"
            sys.modules[mod_name] = mod
            setattr(ref_ops_pkg, mod_name.rsplit(".", 1)[1], mod)
        for k, v in attrs.items():
            setattr(mod, k, v)

    patched = []
    # Device sampler: Graph.sample_adj / csr_subgraph (cogdl/data/data.py:792-874) resolve the
    # module globals `sample_adj_c` / `subgraph_c`.  CUDA graphs go to the device sampler; CPU graphs keep
    # the reference's own host functions (they are not ours to replace: no CPU path in this package).
    import cogdl.data.data as ref_data
    from . import sampling

    ref_sample, ref_subgraph = getattr(ref_data, "sample_adj_c", None), getattr(ref_data, "subgraph_c", None)
    if not getattr(ref_sample, "_cogdl_b200", False):
        def sample_adj_c(indptr, indices, node_idx, num_neighbors=-1, replace=True):
            if not indptr.is_cuda:
                if ref_sample is None:
                    raise RuntimeError("cogdl.operators.sample failed to load and the graph is on the CPU")
                return ref_sample(indptr, indices, node_idx, num_neighbors, replace)
            import torch

            node_idx = torch.as_tensor(node_idx, dtype=torch.int64).to(indptr.device)
            rp, ci, nodes, edges = sampling.sample_adj(indptr, indices, node_idx, num_neighbors, replace)
            # data.py:818-820 pads row_ptr for the new (edge-less) nodes with a CPU tensor; hand it over padded
            pad = nodes.numel() + 1 - rp.numel()
            if pad > 0:
                rp = torch.cat([rp, rp[-1:].expand(pad)])
            return rp, ci, nodes, edges

        def subgraph_c(indptr, indices, node_idx):
            if not indptr.is_cuda:
                if ref_subgraph is None:
                    raise RuntimeError("cogdl.operators.sample failed to load and the graph is on the CPU")
                return ref_subgraph(indptr, indices, node_idx)
            return sampling.subgraph(indptr, indices, node_idx.to(indptr.device))

        sample_adj_c._cogdl_b200 = subgraph_c._cogdl_b200 = True
        ref_data.sample_adj_c, ref_data.subgraph_c = sample_adj_c, subgraph_c
        patched += ["cogdl.data.data.sample_adj_c", "cogdl.data.data.subgraph_c"]
    if rebind_dispatch:
        names = ["spmm", "edge_softmax", "mh_spmm", "fused_gat_op", "check_fused_gat", "SpMM", "EdgeSoftmax",
                 "MultiHeadSpMM", "FusedGATOp"]
        originals = {id(getattr(ref_su, n)): getattr(su, n) for n in names if hasattr(ref_su, n)}
        for mname, mod in list(sys.modules.items()):
            if mod is None or not (mname == "cogdl" or mname.startswith("cogdl.")):
                continue
            for attr, val in list(vars(mod).items()):
                new = originals.get(id(val))
                if new is not None and val is not new:
                    setattr(mod, attr, new)
                    patched.append(f"{mname}.{attr}")
    return patched
