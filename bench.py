#!/usr/bin/env python
"""bench.py -- SpMM aggregated-edges/s and HBM GB/s (hidden=128) on synthetic power-law CSR graphs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extras]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one synthetic graph:
  N = 1 : BASELINE.json configs[1] -- weighted CSR SpMM, hidden=128, on the ogbn-arxiv-shaped
          graph (169 343 nodes, 1 166 243 edges + self loops, sym-normalised weights, seed 0):
          the aggregation step of GCN layer 1 (`spmm(graph, x)`).
  N > 1 : BASELINE.json configs[4] shape per GPU -- every rank owns a contiguous node range of a
          papers100M-shaped graph (1/8 of it per GPU: 13.9 M rows, 202 M edges, hidden=128) with
          locality-controlled columns; halo feature rows are exchanged, then the local two-source
          SpMM runs (cogdl_b200.dist).  Weak scaling; value = edges of all ranks / max-over-ranks time.

Timing: CUDA events on the launching (torch current) stream around each step, after W >= 3 warm-up
steps; an L2 flush (512 MiB write) runs between timed steps and is excluded from the intervals;
multi-GPU intervals are max-reduced over ranks.  `value` has the inputs resident in HBM; `e2e` is
the same step through the public API with pinned HOST feature buffers (H2D of X, kernel, D2H of Y
inside the timed interval; the CSR structure stays resident as it does across CogDL's training
steps, cogdl/trainer/trainer.py:32-45).

The `--impl reference` arm times the reference's own CPU SpMM (cogdl/operators/spmm/spmm_cpu.cpp
compiled unmodified into oracle/_ref/, -O3 build) on the same workload with all host threads.
Only that arm and the `cpu_baseline` leg may touch oracle/ (test infrastructure).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_HIDDEN = 128
FLUSH_BYTES = 512 << 20


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def spmm_bytes(n, nnz, F, weighted=True):
    """Algorithmic bytes of one SpMM launch: gather one F-float row + colind (+ val) per
    edge, write one row + read one rowptr per node; and the compulsory-traffic bound."""
    per_edge_idx = 8 if weighted else 4
    algo = nnz * (4 * F + per_edge_idx) + n * (4 * F + 4)
    minimum = 4 * (n + 1) + per_edge_idx * nnz + 8 * n * F
    return algo, minimum


def arxiv_workload():
    import torch
    from cogdl_b200 import synth

    n, e = synth.SHAPES["arxiv"]
    rp, col = synth.powerlaw_csr(n, e, seed=0, self_loops=True)
    w = synth.sym_norm_weights(rp, col)
    x = torch.randn(n, F_HIDDEN, generator=torch.Generator().manual_seed(0))
    return rp, col, w, x


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import oracle

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    rp, col, w, x = arxiv_workload()
    rp32, col32 = rp.int(), col.int()
    nnz, n = int(col.numel()), int(rp.numel() - 1)
    variant = "o3" if oracle.ref_available("spmm_cpu", "o3") else None
    if variant is not None:
        fn = oracle.ref_module("spmm_cpu", variant).csr_spmm_cpu
        kind = "reference"
        call = lambda: fn(rp32, col32, w, x)
    else:  # the reference did not compile here: the C restatement (oracle port)
        kind = "port"
        a = (rp32.numpy(), col32.numpy(), w.numpy(), x.numpy())
        call = lambda: oracle.spmm_csr(*a)
    # "all the host threads it can use": the loop's dynamic schedule stops scaling well before 128
    # threads on this box, so pick the thread count that is fastest for the reference (stated in `cores`)
    best_t, best = threads, None
    for t in sorted({threads, max(threads // 2, 1), max(threads // 4, 1), 16, 8}, reverse=True):
        if t > threads:
            continue
        oracle.set_num_threads(t)
        call()
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, best_t = dt, t
    oracle.set_num_threads(best_t)
    threads_used = best_t
    for _ in range(max(args.warmup, 1)):
        call()
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        call()
        ts.append(time.perf_counter() - t0)
    t = sum(ts) / len(ts)
    val = nnz / t
    algo, _ = spmm_bytes(n, nnz, F_HIDDEN)
    line = {
        "impl": "reference", "metric": "spmm_aggregated_edges_per_sec", "value": val, "unit": "edges/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "spmm hidden=128 on ogbn-arxiv-shaped power-law CSR (169343 nodes, 1335586 nnz incl. self loops), seed 0",
                   "kernel": "reference cogdl/operators/spmm/spmm_cpu.cpp (unmodified, -O3 -fopenmp) via oracle/_ref" if kind == "reference" else "oracle port"},
        "algorithmic_GBps": algo / t / 1e9,
        "cpu_baseline": {"value": val, "unit": "edges/s", "cores": threads_used, "host_cores": threads, "kind": kind,
                         "sample": "the full workload, every step (one SpMM over the whole graph)"},
        "e2e": {"value": val, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- our arm
def time_steps(fn, steps, warmup, flush, torch, dist_on):
    """Per-step CUDA-event intervals (ms), L2 flushed between steps; max over ranks per step."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist_on:
        import torch.distributed as dist
        dist.barrier()
    evs = []
    for _ in range(steps):
        if flush is not None:
            flush.zero_()       # evicts X / Y from the 126 MB L2; outside the timed interval
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    if dist_on:
        import torch.distributed as dist
        dist.barrier()
    ms = torch.tensor([a.elapsed_time(b) for a, b in evs], dtype=torch.float64, device="cuda")
    if dist_on:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.cpu().tolist()


def cpu_baseline_leg(rp, col, w, x, nnz):
    """Reference CPU SpMM on this box's host cores, bounded sample (rank 0, N = 1 only)."""
    import torch
    import oracle

    threads = os.cpu_count() or 1
    rp32, col32 = rp.int(), col.int()
    out = {"unit": "edges/s"}
    got = {}

    def timed(fn, budget):
        fn(rp32, col32, w, x)
        ts = []
        t_end = time.perf_counter() + budget
        while len(ts) < 30 and (time.perf_counter() < t_end or len(ts) < 3):
            t0 = time.perf_counter()
            fn(rp32, col32, w, x)
            ts.append(time.perf_counter() - t0)
        return nnz / statistics.median(ts), len(ts)

    for variant in ("o3", "asis"):
        if not oracle.ref_available("spmm_cpu", variant):
            continue
        fn = oracle.ref_module("spmm_cpu", variant).csr_spmm_cpu
        # the extension's OpenMP runtime is the system libgomp (same one liboracle.so links), so
        # oracle.set_num_threads() sets the thread count the reference loop actually runs with
        sweep = {}
        for t in sorted({threads, max(threads // 2, 1), max(threads // 4, 1), 16, 8}, reverse=True):
            if t > threads:
                continue
            oracle.set_num_threads(t)
            sweep[t] = timed(fn, 2.0 if variant == "o3" else 1.0)
        oracle.set_num_threads(threads)
        got[variant] = sweep
    if got:
        best = "o3" if "o3" in got else "asis"
        sw = got[best]
        bt = max(sw, key=lambda t: sw[t][0])
        out.update({"value": sw[bt][0], "cores": bt, "kind": "reference", "host_cores": threads,
                    "value_all_cores": sw[threads][0],
                    "threads_sweep": {str(t): v[0] for t, v in sw.items()},
                    "sample": f"full workload (one SpMM over the whole graph) per run, median of {sw[bt][1]} runs at the best "
                              f"OpenMP thread count ({bt} of {threads} host cores); reference spmm_cpu.cpp built "
                              f"{'-O3' if best == 'o3' else 'as shipped (no -O)'}"})
        if "asis" in got:
            sa = got["asis"]
            out["as_shipped_value"] = max(v[0] for v in sa.values())
    else:
        a = (rp32.numpy(), col32.numpy(), w.numpy(), x.numpy())
        oracle.spmm_csr(*a)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            oracle.spmm_csr(*a)
            ts.append(time.perf_counter() - t0)
        out.update({"value": nnz / statistics.median(ts), "kind": "port", "cores": threads,
                    "sample": "full workload, median of 5, oracle.c"})
    return out


def extras(torch, flush, steps=10):
    """Informational numbers for the other configs (not the headline): C2 layer-2 width, C3 GAT
    pieces on the arxiv shape (H=8, F=128), C4 scatter_max on a products-shaped graph."""
    import cogdl_b200
    from cogdl_b200 import synth
    from cogdl_b200.operators._raw import (edge_softmax_fwd_raw, mhspmm_raw, scatter_max_fwd_raw, spmm_raw,
                                           gat_fwd_raw, sddmm_raw)

    peak, _ = measured_peaks()
    dev = torch.device("cuda")
    res = {}

    def bench(name, fn, algo_bytes, units):
        ms = time_steps(fn, steps, 3, flush, torch, False)
        t = statistics.median(ms) / 1e3
        res[name] = {"ms": t * 1e3, "edges_per_s": units / t, "algorithmic_GBps": algo_bytes / t / 1e9,
                     "frac_of_hbm_peak": algo_bytes / t / 1e9 / peak}

    n, e = synth.SHAPES["arxiv"]
    rp, col = synth.powerlaw_csr(n, e, seed=0)
    g = cogdl_b200.Graph(row_ptr=rp, col=col, edge_weight=synth.sym_norm_weights(rp, col), num_nodes=n).to(dev)
    st = g.structure()
    nnz = st.nnz
    w = g.raw_edge_weight
    x40 = torch.randn(n, 40, device=dev)
    bench("C2_spmm_F40", lambda: spmm_raw(st, w, x40), nnz * (4 * 40 + 8) + n * (4 * 40 + 4), nnz)
    x128 = torch.randn(n, 128, device=dev)
    bench("C2_sddmm_F128", lambda: sddmm_raw(st, x128, x128), nnz * (2 * 4 * 128 + 8) + 4 * (n + 1), nnz)
    H, F = 8, 128
    logits = (torch.randn(nnz, H, device=dev) * 3).clamp_(-10, 10)
    bench("C3_edge_softmax_H8", lambda: edge_softmax_fwd_raw(st, logits), 2 * 4 * nnz * H + 4 * (n + 1), nnz)
    att = edge_softmax_fwd_raw(st, logits)
    h = torch.randn(n, H, F, device=dev)
    mh_bytes = nnz * (4 * H * F + 4 + 4 * H) + n * (4 * H * F + 4)
    bench("C3_mhspmm_H8_F128", lambda: mhspmm_raw(st, att, h), mh_bytes, nnz)
    hl, hr = torch.randn(n, H, device=dev), torch.randn(n, H, device=dev)
    bench("C3_fused_gat_H8_F128", lambda: gat_fwd_raw(st, hl, hr, h, 0.2, False), mh_bytes, nnz)
    # ---- secondary: whole training steps of the three config models on the arxiv shape (cuBLAS GEMMs +
    # our sparse kernels + autograd: forward, backward, SGD), informational
    def train_step_ms(model, graph, out_dim, reps=5):
        import torch.nn.functional as Fn

        opt = torch.optim.SGD(model.parameters(), lr=0.01)
        y = torch.randint(0, out_dim, (n,), device=dev)
        ts = []
        for i in range(reps + 2):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            opt.zero_grad(set_to_none=True)
            loss = Fn.cross_entropy(model(graph), y)
            loss.backward()
            opt.step()
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    try:
        from cogdl_b200.layers import GCN, GAT, SAGE

        g.x = torch.randn(n, 128, device=dev)
        res["C2_gcn2_train_step"] = {"ms": train_step_ms(GCN(128, 128, 40, dropout=0.0).to(dev), g, 40),
                                     "what": "2-layer GCN hidden=128, fwd+bwd+SGD, arxiv shape"}
        g2 = cogdl_b200.Graph(x=g.x, row_ptr=rp, col=col, num_nodes=n).to(dev)
        res["C3_gat2_train_step"] = {"ms": train_step_ms(GAT(128, 16, 40, nhead=8, last_nhead=1).to(dev), g2, 40),
                                     "what": "2-layer GAT 8 heads x 16, fwd+bwd+SGD, arxiv shape"}
        res["C4_sage_max2_train_step"] = {"ms": train_step_ms(SAGE(128, 128, 40, aggr="max").to(dev), g2, 40),
                                          "what": "2-layer GraphSAGE aggr=max hidden=128, fwd+bwd+SGD, arxiv shape"}
        del g2
    except Exception as ex:  # noqa: BLE001  (secondary numbers must not take the headline down)
        res["train_steps_error"] = f"{type(ex).__name__}: {ex}"
    del h, att, logits, x128, x40, g, st
    n, e = synth.SHAPES["products"]
    rp, col = synth.powerlaw_csr(n, e, seed=0, device=dev, self_loops=False)
    st = cogdl_b200.CSRStructure.from_int64(rp, col, n_cols=n)
    del rp, col
    x = torch.rand(n, 256, device=dev) + 0.01
    bench("C4_scatter_max_F256", lambda: scatter_max_fwd_raw(st, x), e * (4 * 256 + 4) + n * (8 * 256 + 4), e)
    x = x[:, :128].contiguous()
    bench("C4shape_spmm_F128_unweighted", lambda: spmm_raw(st, None, x), e * (4 * 128 + 4) + n * (4 * 128 + 4), e)
    return res


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"     # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    import cogdl_b200
    from cogdl_b200 import _cabi

    _cabi.check(_cabi.load().cogdl_b200_check_device())
    peak, peak_src = measured_peaks()
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local)

    if not dist_on:
        rp, col, w, x_host = arxiv_workload()
        n, nnz = int(rp.numel() - 1), int(col.numel())
        g = cogdl_b200.Graph(row_ptr=rp, col=col, edge_weight=w, num_nodes=n).to(dev)
        st = g.structure()
        st.plan  # build once (cached on the graph), outside the timed region
        x_dev = x_host.to(dev)
        workload = ("spmm hidden=128 (fp32) on ogbn-arxiv-shaped power-law CSR: 169343 nodes, 1166243 edges + 169343 self "
                    "loops = 1335586 nnz, sym-normalised weights, max degree %d, seed 0 [BASELINE configs[1], GCN layer-1 aggregation]"
                    % int((rp[1:] - rp[:-1]).max()))
        step = lambda: cogdl_b200.spmm(g, x_dev)
        # ---- device-resident
        l0 = _cabi.launch_count()
        sampler.start()
        ms = time_steps(step, args.steps, args.warmup, flush, torch, False)
        launches_dev = _cabi.launch_count() - l0
        # ---- end to end: pinned host X -> device, spmm through the public API, Y -> pinned host
        x_pin = x_host.pin_memory()
        y_pin = torch.empty(n, F_HIDDEN).pin_memory()
        x_in = torch.empty_like(x_dev)

        def e2e_step():
            x_in.copy_(x_pin, non_blocking=True)
            y = cogdl_b200.spmm(g, x_in)
            y_pin.copy_(y, non_blocking=True)

        l1 = _cabi.launch_count()
        ms_e2e = time_steps(e2e_step, args.steps, args.warmup, flush, torch, False)
        clocks = sampler.stop()      # sampled across both timed regions
        launches = launches_dev + (_cabi.launch_count() - l1)
        total_units = nnz
        algo, bmin = spmm_bytes(n, nnz, F_HIDDEN)
        parallelism = "single GPU"
        h2d = d2h = n * F_HIDDEN * 4
    else:
        from cogdl_b200 import dist as cdist

        part = cdist.synthetic_partition(rank, world, dev, seed=0, mode=os.environ.get("COGDL_B200_DIST_MODE"),
                                         beta=float(os.environ.get("COGDL_B200_DIST_BETA", "0.05")))
        workload = part.describe()
        x_dev = part.x_local                       # p2p mode: already inside the symmetric shard
        step = lambda: part.spmm(x_dev)
        l0 = _cabi.launch_count()
        sampler.start()
        ms = time_steps(step, args.steps, args.warmup, flush, torch, True)
        clocks = sampler.stop()
        launches_dev = _cabi.launch_count() - l0
        x_pin = x_dev.cpu().pin_memory()
        y_pin = torch.empty(part.n_local, F_HIDDEN).pin_memory()
        x_in = x_dev if part.mode == "p2p" else torch.empty_like(x_dev)   # H2D lands in the shard itself

        def e2e_step():
            x_in.copy_(x_pin, non_blocking=True)
            y = part.spmm(x_in)
            y_pin.copy_(y, non_blocking=True)

        l1 = _cabi.launch_count()
        ms_e2e = time_steps(e2e_step, max(3, args.steps // 4), 2, flush, torch, True)
        launches = launches_dev + (_cabi.launch_count() - l1)
        total_units = part.global_nnz
        n, nnz = part.n_local, part.nnz_local
        algo, bmin = spmm_bytes(n, nnz, F_HIDDEN, weighted=False)  # per rank, per launch
        parallelism = f"node-range partition x{world}; {part.exchange}; no reduce on the data path"
        h2d = d2h = part.n_local * F_HIDDEN * 4

    t = sum(ms) / len(ms) / 1e3
    t_e2e = sum(ms_e2e) / len(ms_e2e) / 1e3
    value = total_units / t
    kernel_t = t if not dist_on else part.last_kernel_seconds(step, torch)
    achieved = algo / kernel_t / 1e9
    line = {
        "metric": "spmm_aggregated_edges_per_sec", "value": value, "unit": "edges/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "ms_per_step_min": min(ms),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "hidden": F_HIDDEN, "parallelism": parallelism,
                   "l2": "flushed between timed steps (512 MiB write, excluded from the event intervals)",
                   "hub_chunk_edges": cogdl_b200.structure.DEFAULT_CHUNK_EDGES},
        "algorithmic_GBps": achieved,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": TRAFFIC_BYTES_PER_LAUNCH if not dist_on else None,
                     "peak_source": peak_src, "kernel": "cogdl_b200::stream_kernel<float4, NV=1, weighted, U=4> (row-stream SpMM)" if not dist_on else
                               "cogdl_b200::stream_kernel<float4, NV=1, unweighted, peers|two-source, U=4>",
                     "algorithmic_bytes_per_launch": algo, "compulsory_bytes_per_launch": bmin,
                     "frac_compulsory": bmin / kernel_t / 1e9 / peak,
                     "l2_resident": (not dist_on),
                     "note": "X (87 MB) fits the 126 MB L2, so algorithmic bytes/time may exceed the HBM peak; "
                             "frac_compulsory is the DRAM-side fraction" if not dist_on else
                             "per-rank SpMM kernel (max over ranks), X shard 7.1 GB >> L2; in p2p mode the same kernel "
                             "also performs the remote-row gather over NVLink"},
        "e2e": {"value": total_units / t_e2e, "unit": "edges/s", "ms_per_step": t_e2e * 1e3,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "pinned host X -> device, cogdl_b200.spmm(graph, x), Y -> pinned host; CSR resident"},
        "gpu_launches": launches, "clocks": clocks,
    }
    if not dist_on:
        line["cpu_baseline"] = cpu_baseline_leg(rp, col, w, x_host, nnz)
        if not args.no_extras:
            del g, st, x_dev, x_in
            torch.cuda.empty_cache()
            line["others"] = extras(torch, flush)
    if rank == 0:
        print(json.dumps(line))
    if dist_on:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of the SpMM kernel at the N=1 workload, from the
# committed `ncu --set full` capture (profiles/); None until a capture exists.
TRAFFIC_BYTES_PER_LAUNCH = 308029184  # 237.96 MB read + 70.07 MB write, profiles/r01f_spmm_stream_ncu_summary.txt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
