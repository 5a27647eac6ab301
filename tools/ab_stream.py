"""One-process A/B of the row-stream kernels' launch shapes on a B200 (CUDA events, L2 flushed between runs):

   python tools/ab_stream.py  ->  gpurun_out/ab_stream_all.jsonl  (one JSON line per configuration)

  block     COGDL_B200_STREAM_BLOCK in {256, 128, 64, 32}: threads per block of the one-warp-per-item kernel (same SASS,
            launch parameter only; smaller blocks hand a retired warp's slot to new work sooner)
  variant   COGDL_B200_SPMM_VARIANT: unroll depth / resident blocks of the lean SpMM kernel
  chunk     edges per hub chunk of the plan
  es_warps  COGDL_B200_ES_WARPS in {8, 4, 2}: warps per block of the edge-softmax 512-float-tile kernel
  seg       segment cost of the hub plan (rows + edges per warp item)

Every configuration's output is compared BIT FOR BIT with the default configuration's (the per-row arithmetic order
does not depend on the launch shape), and the plan's arrival counters must be zero afterwards.  Knobs are
switched inside one process through cogdl_b200_reload_tuning().  Results are flushed line by line, so a run that is
cut off keeps what it measured.  (profiles/r02s_ab_launch_shapes.md is a run of this tool; at that time it also timed a
persistent ticket-counter form of the kernel, since removed.)"""
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cogdl_b200  # noqa: E402
from cogdl_b200 import _cabi, synth  # noqa: E402
from cogdl_b200.operators._raw import mhspmm_raw, spmm_raw  # noqa: E402
from cogdl_b200.structure import CSRStructure  # noqa: E402

what = "all"
dev = torch.device("cuda")
os.makedirs("gpurun_out", exist_ok=True)
out = open(os.path.join("gpurun_out", f"ab_stream_{what}.jsonl"), "a")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
t_start = time.time()


def emit(**kw):
    kw["t"] = round(time.time() - t_start, 1)
    out.write(json.dumps(kw) + "\n")
    out.flush()
    print(json.dumps(kw), flush=True)


def knobs(block=None, variant=None):
    """Set the launch-shape knobs (None = the library's default)."""
    for name, v in (("COGDL_B200_STREAM_BLOCK", block), ("COGDL_B200_SPMM_VARIANT", variant)):
        if v is None:
            os.environ.pop(name, None)
        else:
            os.environ[name] = str(v)
    _cabi.load().cogdl_b200_reload_tuning()


def timed(fn, reps):
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


def sweep(tag, st_by_seg, fn_of_st, shapes, reps):
    """shapes: list of (block, _, variant?).  Reference output = first seg, first shape."""
    ref = None
    for seg, st in st_by_seg.items():
        for shape in shapes:
            block, dyn = shape[0], shape[1]
            variant = shape[2] if len(shape) > 2 else 7
            try:
                knobs(block, variant)
                y = fn_of_st(st)
                torch.cuda.synchronize()
                if ref is None:
                    ref = y.clone()
                same = bool(torch.equal(y, ref))      # segments only group whole rows; hub chunks (64 edges) are the same
                close = float((y - ref).abs().max() / ref.abs().max())
                med, mn = timed(lambda: fn_of_st(st), reps)
                clean = int(st.plan.counters.abs().sum()) == 0
                emit(case=tag, seg=seg, block=block, variant=variant, median_us=round(med, 1), min_us=round(mn, 1),
                     bit_identical_to_default=same, max_rel_vs_default=close, counters_clean=clean,
                     kernel=_cabi.last_kernel(), segs=st.plan.n_segs, chunks=st.plan.n_chunks)
            except Exception as ex:  # noqa: BLE001
                emit(case=tag, seg=seg, block=block, error=f"{type(ex).__name__}: {ex}")
    knobs()


shapes = [(256, 0), (128, 0), (64, 0), (32, 0)]

# ---- arxiv shape (headline): weighted SpMM, F = 128 / 40 / 256; multi-head SpMM H = 8, F = 128
n, e = synth.SHAPES["arxiv"]
rp, col = synth.powerlaw_csr(n, e, seed=0)
w = synth.sym_norm_weights(rp, col).to(dev)
rp32, col32 = rp.to(dev).to(torch.int32), col.to(dev).to(torch.int32)
sts = {seg: CSRStructure(rp32, col32, n_cols=n, seg_cost=seg) for seg in (128, 96, 64)}
for st in sts.values():
    st.plan
emit(case="setup", arxiv_nnz=sts[128].nnz, gpu=torch.cuda.get_device_name(0))
for F in (128, 40, 256):
    x = torch.randn(n, F, device=dev)
    sweep(f"arxiv_spmm_F{F}", sts if F == 128 else {128: sts[128]}, lambda st: spmm_raw(st, w, x), shapes, 30 if F == 128 else 15)
    if F == 128:
        # other unroll / occupancy variants of the lean kernel at the two extreme block sizes
        sweep("arxiv_spmm_F128_variants", {128: sts[128]}, lambda st: spmm_raw(st, w, x),
              [(256, 0, 7), (256, 0, 3), (64, 0, 3), (256, 0, 1), (64, 0, 1)], 20)
        # hub chunk size (edges per chunk); the plan is per structure
        for chunk in (32, 128):
            stc = CSRStructure(rp32, col32, n_cols=n, chunk_edges=chunk, seg_cost=128)
            stc.plan
            sweep(f"arxiv_spmm_F128_chunk{chunk}", {128: stc}, lambda st: spmm_raw(st, w, x), [(256, 0), (64, 0)], 20)
            del stc
    del x
H = 8
if True:   # ---- edge softmax family
    # (arxiv, H = 8): warps per block of the staged main kernel
    from cogdl_b200.operators._raw import edge_softmax_bwd_raw, edge_softmax_fwd_raw, gat_attn_bwd_raw  # noqa: E402

    st = sts[128]
    logits = (torch.randn(st.nnz, H, device=dev) * 3).clamp_(-10, 10)
    g = torch.randn(st.nnz, H, device=dev)
    hl, hr = torch.randn(n, H, device=dev), torch.randn(n, H, device=dev)
    os.environ["COGDL_B200_ES_WARPS"] = "8"
    _cabi.load().cogdl_b200_reload_tuning()
    att0 = edge_softmax_fwd_raw(st, logits)
    for name, fn in (("es_fwd", lambda: edge_softmax_fwd_raw(st, logits)), ("es_bwd", lambda: edge_softmax_bwd_raw(st, att0, g)),
                     ("gat_attn_bwd", lambda: gat_attn_bwd_raw(st, att0, g, hl, hr, 0.2)[0])):
        ref = None
        for warps in (8, 4, 2):
            try:
                os.environ["COGDL_B200_ES_WARPS"] = str(warps)
                _cabi.load().cogdl_b200_reload_tuning()
                y = fn()
                torch.cuda.synchronize()
                ref = y.clone() if ref is None else ref
                med, mn = timed(fn, 20)
                emit(case=f"arxiv_{name}_H8", es_warps=warps, median_us=round(med, 1), min_us=round(mn, 1),
                     bit_identical_to_default=bool(torch.equal(y, ref)), counters_clean=int(st.plan.counters.abs().sum()) == 0,
                     kernel=_cabi.last_kernel())
            except Exception as ex:  # noqa: BLE001
                emit(case=f"arxiv_{name}_H8", es_warps=warps, error=f"{type(ex).__name__}: {ex}")
    os.environ["COGDL_B200_ES_WARPS"] = "8"
    _cabi.load().cogdl_b200_reload_tuning()
    del logits, g, hl, hr, att0
if True:
    att = torch.rand(sts[128].nnz, H, device=dev)
    h = torch.randn(n, H, 128, device=dev)
    sweep("arxiv_mhspmm_H8_F128", {128: sts[128]}, lambda st: mhspmm_raw(st, att, h), [s for s in shapes if s[1] == 0], 8)
    del att, h

# ---- products shape (X >> L2, HBM-bound): unweighted SpMM F = 128
n, e = synth.SHAPES["products"]
rp, col = synth.powerlaw_csr(n, e, seed=0, device=dev, self_loops=False)
st = CSRStructure.from_int64(rp, col, n_cols=n)
del rp, col
st.plan
x = torch.randn(n, 128, device=dev)
sweep("products_spmm_F128", {128: st}, lambda s_: spmm_raw(s_, None, x), shapes, 6)
emit(case="done")
