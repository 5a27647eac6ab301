"""Time scatter_max (products shape, F=256) with the row-stream form on and off."""
import os
import subprocess
import sys

CHILD = r'''
import os, sys, statistics, torch
sys.path.insert(0, os.getcwd())
import cogdl_b200
from cogdl_b200 import synth
from cogdl_b200.operators._raw import scatter_max_fwd_raw
dev = torch.device("cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
n, e = synth.SHAPES["products"]
rp, col = synth.powerlaw_csr(n, e, seed=0, device=dev, self_loops=False)
st = cogdl_b200.CSRStructure.from_int64(rp, col, n_cols=n)
x = torch.rand(n, 256, device=dev) + 0.01
for _ in range(2): scatter_max_fwd_raw(st, x)
ts = []
for _ in range(6):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); scatter_max_fwd_raw(st, x); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"RESULT smax_stream={os.environ.get('COGDL_B200_SMAX_STREAM')} scatter_max_F256_products_ms={statistics.median(ts):.3f}")
'''
for v in ("1", "0"):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, COGDL_B200_SMAX_STREAM=v), capture_output=True, text=True)
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith("RESULT")) or r.stderr[-600:], flush=True)
