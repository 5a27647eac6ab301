"""GPU tuning sweep for the SpMM kernels (run on the B200 box):
   python tools/tune_spmm.py  -> table of kernel time vs (variant, seg_cost, chunk_edges) on the
   arxiv-shaped (L2-resident) and products-shaped (HBM-bound) graphs, hidden=128."""
import itertools
import os
import statistics
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CHILD = r'''
import os, sys, statistics, torch
sys.path.insert(0, os.getcwd())
import cogdl_b200
from cogdl_b200 import synth
from cogdl_b200.operators._raw import spmm_raw
dev = torch.device("cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, steps=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), min(ts)
shape = sys.argv[1]
n, e = synth.SHAPES[shape]
if shape == "arxiv":
    rp, col = synth.powerlaw_csr(n, e, seed=0)
    w = synth.sym_norm_weights(rp, col).to(dev)
    rp, col = rp.to(dev), col.to(dev)
else:
    rp, col = synth.powerlaw_csr(n, e, seed=0, device=dev, self_loops=False)
    w = None
st = cogdl_b200.CSRStructure.from_int64(rp, col, n_cols=n)
x = torch.randn(n, 128, device=dev)
# cudaEvent timing of a single launch includes launch latency; subtract nothing, just report
med, mn = timeit(lambda: spmm_raw(st, w, x))
if shape == "arxiv" and os.environ.get("TUNE_EXTRA"):
    from cogdl_b200.operators._raw import mhspmm_raw, edge_softmax_fwd_raw
    H = 8
    logits = (torch.randn(st.nnz, H, device=dev) * 3).clamp_(-10, 10)
    m1, _ = timeit(lambda: edge_softmax_fwd_raw(st, logits), 10)
    att = edge_softmax_fwd_raw(st, logits)
    h = torch.randn(n, H, 128, device=dev)
    m2, _ = timeit(lambda: mhspmm_raw(st, att, h), 10)
    h16 = torch.randn(n, H, 16, device=dev)
    m3, _ = timeit(lambda: mhspmm_raw(st, att, h16), 10)
    from cogdl_b200.operators._raw import gat_fwd_raw, edge_softmax_bwd_raw
    hl, hr = torch.randn(n, H, device=dev), torch.randn(n, H, device=dev)
    m4, _ = timeit(lambda: gat_fwd_raw(st, hl, hr, h, 0.2, False), 10)
    m5, _ = timeit(lambda: edge_softmax_bwd_raw(st, att, logits), 10)
    print(f"RESULT extra edge_softmax_H8_us={m1*1e3:.1f} edge_softmax_bwd_us={m5*1e3:.1f} mhspmm_H8_F128_us={m2*1e3:.1f} mhspmm_H8_F16_us={m3*1e3:.1f} gat_fwd_H8_F128_us={m4*1e3:.1f}")
print(f"RESULT {shape} variant={os.environ.get('COGDL_B200_SPMM_VARIANT','0')} seg={os.environ.get('COGDL_B200_SEG_COST')} chunk={os.environ.get('COGDL_B200_CHUNK_EDGES')} median_us={med*1e3:.1f} min_us={mn*1e3:.1f} nnz={st.nnz} segs={st.plan.n_segs} chunks={st.plan.n_chunks}")
'''

def run(shape, variant, seg, chunk):
    env = dict(os.environ, COGDL_B200_SPMM_VARIANT=str(variant), COGDL_B200_SEG_COST=str(seg), COGDL_B200_CHUNK_EDGES=str(chunk))
    r = subprocess.run([sys.executable, "-c", CHILD, shape], env=env, capture_output=True, text=True)
    for line in r.stdout.splitlines():
        if line.startswith("RESULT"):
            print(line, flush=True)
    if r.returncode != 0:
        print("FAILED", shape, variant, seg, chunk, r.stderr[-500:], flush=True)

if __name__ == "__main__":
    os.environ["TUNE_EXTRA"] = "1"
    run("arxiv", 7, 128, 64)
    os.environ.pop("TUNE_EXTRA")
    run("products", 7, 128, 64)
