"""The HBM-bound configurations, a few launches each, for an `ncu --set full` capture:
products-shaped SpMM F=128 (row-stream kernel), arxiv-shaped multi-head SpMM H=8 F=128, products-shaped
scatter_max F=256."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cogdl_b200  # noqa: E402
from cogdl_b200 import synth  # noqa: E402
from cogdl_b200.operators._raw import edge_softmax_fwd_raw, mhspmm_raw, scatter_max_fwd_raw, spmm_raw  # noqa: E402

dev = torch.device("cuda")
n, e = synth.SHAPES["arxiv"]
rp, col = synth.powerlaw_csr(n, e, seed=0)
st = cogdl_b200.CSRStructure.from_int64(rp.to(dev), col.to(dev), n_cols=n)
att = edge_softmax_fwd_raw(st, torch.randn(st.nnz, 8, device=dev))
h = torch.randn(n, 8, 128, device=dev)
for _ in range(2):
    mhspmm_raw(st, att, h)
del h, att, st
n, e = synth.SHAPES["products"]
rp, col = synth.powerlaw_csr(n, e, seed=0, device=dev, self_loops=False)
st = cogdl_b200.CSRStructure.from_int64(rp, col, n_cols=n)
x = torch.rand(n, 256, device=dev) + 0.01
for _ in range(2):
    scatter_max_fwd_raw(st, x)
x = x[:, :128].contiguous()
for _ in range(2):
    spmm_raw(st, None, x)
torch.cuda.synchronize()
print("done")
