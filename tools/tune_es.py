"""Edge-softmax family timings on the arxiv shape (CUDA events, L2 flushed between runs, median of 20).
   COGDL_B200_ES_CAP=512|1024|2048 python tools/tune_es.py [H ...]"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cogdl_b200  # noqa: E402
from cogdl_b200 import _cabi, synth  # noqa: E402
from cogdl_b200.operators._raw import (edge_colsum_raw, edge_softmax_bwd_raw, edge_softmax_fwd_raw, gat_attn_bwd_raw,  # noqa: E402
                                       gat_fwd_raw, mhsddmm_raw)

dev = torch.device("cuda")
n, e = synth.SHAPES["arxiv"]
rp, col = synth.powerlaw_csr(n, e, seed=0)
st = cogdl_b200.CSRStructure.from_int64(rp.to(dev), col.to(dev), n_cols=n)
st.plan
st_t, perm = st.csc()
st_t.plan
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
peak = 6566.1


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts), min(ts)


print(f"plan: hubs {st.plan.n_hub_rows} chunks {st.plan.n_chunks} segs {st.plan.n_segs} nnz {st.nnz}  ES_CAP={os.environ.get('COGDL_B200_ES_CAP')}")
for H in [int(a) for a in sys.argv[1:]] or [8, 1, 4, 32]:
    logits = (torch.randn(st.nnz, H, device=dev) * 3).clamp_(-10, 10)
    att = edge_softmax_fwd_raw(st, logits)
    g = torch.randn(st.nnz, H, device=dev)
    hl, hr = torch.randn(n, H, device=dev), torch.randn(n, H, device=dev)
    bytes_fwd = 8 * st.nnz * H + 4 * (n + 1)
    for name, fn, nbytes in (
        ("edge_softmax_fwd", lambda: edge_softmax_fwd_raw(st, logits), bytes_fwd),
        ("edge_softmax_bwd", lambda: edge_softmax_bwd_raw(st, att, g), 12 * st.nnz * H + 4 * (n + 1)),
        ("gat_attn_bwd", lambda: gat_attn_bwd_raw(st, att, g, hl, hr, 0.2), 12 * st.nnz * H + 8 * st.nnz + 4 * (n + 1)),
        ("edge_colsum", lambda: edge_colsum_raw(st_t, perm, g), 4 * st.nnz * H + 4 * st.nnz + 4 * n * H),
    ):
        med, mn = timed(fn)
        print(f"H={H:2d} {name:18s} median {med:8.1f} us  min {mn:8.1f} us  {nbytes / med / 1e3:8.1f} GB/s  frac {nbytes / med / 1e3 / peak:5.3f}   [{_cabi.last_kernel()}]")
