"""SpMM timings on the arxiv shape for several widths (CUDA events, L2 flushed, median of 30)."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cogdl_b200  # noqa: E402
from cogdl_b200 import _cabi, synth  # noqa: E402
from cogdl_b200.operators._raw import spmm_raw  # noqa: E402

dev = torch.device("cuda")
n, e = synth.SHAPES["arxiv"]
rp, col = synth.powerlaw_csr(n, e, seed=0)
w = synth.sym_norm_weights(rp, col).to(dev)
st = cogdl_b200.CSRStructure.from_int64(rp.to(dev), col.to(dev), n_cols=n)
st.plan
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for F in [int(a) for a in sys.argv[1:]] or [128, 40, 64, 16, 256]:
    x = torch.randn(n, F, device=dev)
    for wt in (w, None):
        fn = lambda: spmm_raw(st, wt, x)
        for _ in range(3):
            fn()
        ts = []
        for _ in range(30):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        algo = st.nnz * (4 * F + (8 if wt is not None else 4)) + n * (4 * F + 4)
        med = statistics.median(ts)
        print(f"F={F:4d} {'weighted  ' if wt is not None else 'unweighted'} median {med:7.1f} us min {min(ts):7.1f} us  {algo / med / 1e3:7.0f} GB/s  "
              f"frac {algo / med / 1e3 / 6566.1:5.3f}  [{_cabi.last_kernel()}]")
