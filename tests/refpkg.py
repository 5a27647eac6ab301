"""Import the UNMODIFIED reference package (THUDM/CogDL) inside a test process.

Where it comes from: `baseline/_ref` (pip-installed from /root/reference by __graft_entry__.build(),
git-ignored, shipped to the GPU box with the snapshot) or /root/reference itself in the build
container.  What is stubbed, and why that does not touch the path under test:
  * optuna / matplotlib / grave -- optional third-party imports of the control plane (AutoML, plots)
    that are not installed in this image;
  * cogdl.operators.{spmm, sample} -- the reference builds these with a JIT `load()` at import time;
    here the modules are pre-seeded with the SAME sources compiled ahead of time by
    oracle/build_ref.py (oracle/_ref/asis/*.so), so the reference's CPU path is its own code;
  * cogdl.operators.{edge_softmax, mhspmm, scatter_max, fused_gat} -- CUDA-only JIT builds (minutes of
    nvcc at import); seeded empty (= "failed to load", the state the reference handles by falling back
    to its Python CPU paths) until cogdl_b200.install() fills them with the sm_100a operators.
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]


def reference_dir():
    for d in CANDIDATES:
        if os.path.isdir(os.path.join(d, "cogdl")):
            return d
    return None


def import_reference():
    """Returns the imported `cogdl` module (idempotent)."""
    if "cogdl" in sys.modules and getattr(sys.modules["cogdl"], "__file__", None):
        return sys.modules["cogdl"]
    d = reference_dir()
    if d is None:
        raise ImportError("reference package not found (baseline/_ref or /root/reference)")
    import oracle

    class _Stub(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return lambda *a, **k: None

    for name in ["optuna", "matplotlib", "matplotlib.cm", "matplotlib.pyplot", "grave"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:  # noqa: BLE001
                sys.modules[name] = _Stub(name)
    spmm_mod = types.ModuleType("cogdl.operators.spmm")
    spmm_mod.csrspmm = None
    spmm_mod.spmm_cpu = oracle.ref_module("spmm_cpu", "asis").csr_spmm_cpu
    sys.modules["cogdl.operators.spmm"] = spmm_mod
    smp = oracle.ref_module("sampler", "asis")
    sample_mod = types.ModuleType("cogdl.operators.sample")
    sample_mod.subgraph_c, sample_mod.sample_adj_c = smp.subgraph, smp.sample_adj
    sample_mod.coo2csr_cpu, sample_mod.coo2csr_cpu_index = smp.coo2csr_cpu, smp.coo2csr_cpu_index
    sys.modules["cogdl.operators.sample"] = sample_mod
    for name, attrs in (("edge_softmax", {"csr_edge_softmax": None}), ("mhspmm", {"csrmhspmm": None}),
                        ("scatter_max", {"scatter_max": None}), ("fused_gat", {"fused_gat_func": None})):
        m = types.ModuleType("cogdl.operators." + name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules["cogdl.operators." + name] = m
    sys.path.insert(0, d)
    import cogdl  # noqa: F401

    return sys.modules["cogdl"]
