"""Time SDDMM / multi-head SDDMM (arxiv shape) with the row-stream form on and off (subprocess per setting)."""
import os
import subprocess
import sys

CHILD = r'''
import os, sys, statistics, torch
sys.path.insert(0, os.getcwd())
import cogdl_b200
from cogdl_b200 import synth
from cogdl_b200.operators._raw import sddmm_raw, mhsddmm_raw
dev = torch.device("cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, steps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)
n, e = synth.SHAPES["arxiv"]
rp, col = synth.powerlaw_csr(n, e, seed=0)
st = cogdl_b200.CSRStructure.from_int64(rp.to(dev), col.to(dev), n_cols=n)
x = torch.randn(n, 128, device=dev)
h = torch.randn(n, 8, 128, device=dev)
t1 = timeit(lambda: sddmm_raw(st, x, x))
t2 = timeit(lambda: mhsddmm_raw(st, h, h))
print(f"RESULT stream={os.environ.get('COGDL_B200_SDDMM_STREAM')} sddmm_F128_us={t1*1e3:.1f} mhsddmm_H8_F128_us={t2*1e3:.1f}")
'''
for v in ("1", "0"):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, COGDL_B200_SDDMM_STREAM=v), capture_output=True, text=True)
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith("RESULT")) or r.stderr[-600:], flush=True)
