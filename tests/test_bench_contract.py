"""bench.py contract checks that run without a GPU: the `--impl reference` arm (reference CPU SpMM from
oracle/_ref, or the oracle port) prints exactly one JSON line with the required keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "spmm_aggregated_edges_per_sec" and d["unit"] == "edges/s"
    assert d["steps"] == 2 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["value"] > 0


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_multi_gpu_times_the_shard_workload_and_never_loads_the_product_library():
    """N > 1: rank 0 times a row slice of its papers100M-shaped shard (same workload string as our arm
    builds from synth.shard_description), and the process must not have dlopen-ed libcogdl_b200.so."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", COGDL_B200_BENCH_SHARD_DIV="100")
    code = (
        "import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '3'];"
        "runpy.run_path('bench.py', run_name='__main__');"
        "maps = open('/proc/self/maps').read();"
        "assert 'libcogdl_b200' not in maps, 'reference arm loaded the product library';"
        "assert 'cogdl_b200' not in sys.modules"
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["scaling"] == "weak"
    assert "papers100M-shaped" in d["config"]["workload"] and "beta=0.05" in d["config"]["workload"]
    assert d["cpu_baseline"]["cores"] >= 1 and "threads_sweep_edges_per_s" in d["cpu_baseline"]


def test_both_arms_build_the_same_workload_strings():
    sys.path.insert(0, ROOT)
    import importlib

    bench = importlib.import_module("bench")
    synth = bench.load_synth()
    from cogdl_b200 import synth as pkg_synth

    assert synth.arxiv_description(123) == pkg_synth.arxiv_description(123)
    a = synth.shard_description(10, 20, 4, 0.25, 0, 128, "strong")
    assert a == pkg_synth.shard_description(10, 20, 4, 0.25, 0, 128, "strong") and "split 4-way" in a
    assert synth.shard_sizes(2, "strong") == (111059956 // 2, 1615685872 // 2)
    assert synth.shard_sizes(2, "weak") == synth.shard_sizes(8, "weak") == (111059956 // 8, 1615685872 // 8)


def test_cpu_thread_sweep_takes_the_fastest_median():
    sys.path.insert(0, ROOT)
    import importlib
    import time

    bench = importlib.import_module("bench")
    state = {"t": 1}
    cost = {8: 0.004, 4: 0.001, 2: 0.006, 16: 0.02, 32: 0.02, 1: 0.01}

    def call():
        time.sleep(cost.get(state["t"], 0.01))

    best, res = bench.cpu_thread_sweep(call, lambda t: state.update(t=t), 8, reps=3)
    assert best == 4 and set(res) == {8, 4, 2} and state["t"] == 4


def test_shard_generator_is_deterministic_local_and_sliceable():
    import torch
    from cogdl_b200 import synth

    a = synth.shard_csr(1, 4, 5000, 60000, 0.05, seed=0)
    b = synth.shard_csr(1, 4, 5000, 60000, 0.05, seed=0)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    rp, col = a
    assert int(rp[-1]) == 60000 == col.numel() and int(col.min()) >= 0 and int(col.max()) < 20000
    remote = ((col < 5000) | (col >= 10000)).float().mean().item()
    assert 0.02 < remote < 0.06                       # beta * (P - 1) / P = 0.0375
    rp0, col0 = synth.shard_csr(2, 4, 5000, 60000, 0.0, seed=0)
    assert int(col0.min()) >= 10000 and int(col0.max()) < 15000      # beta = 0: every column inside the own range
    rps, cols = synth.shard_csr(0, 4, 5000, 60000, 0.05, seed=0, max_slice_edges=10000)
    assert 10000 <= int(rps[-1]) == cols.numel() < 12000 and rps.numel() < 5001
    full = synth.shard_csr(0, 4, 5000, 60000, 0.05, seed=0)[0]
    assert torch.equal(rps, full[: rps.numel()])       # the slice keeps the full shard's degree law


def test_elementwise_error_uses_the_row_scale():
    import importlib

    import numpy as np

    bench = importlib.import_module("bench")
    ref = np.array([[100.0, 1e-6, -50.0], [0.0, 0.0, 0.0]], np.float32)
    got = ref.copy()
    got[0, 1] += 5e-4                                   # tiny element of a large row: judged against the row scale
    assert abs(bench.elementwise_err(got, ref) - 5e-6) < 1e-7
    got[1, 2] = 1e-3                                    # an all-zero row has no scale to hide behind
    assert bench.elementwise_err(got, ref) > 1.0


def test_profiled_traffic_matches_the_instantiation_not_the_launch_shape():
    """roofline.traffic comes from a committed ncu capture and is only used when the capture is of the kernel
    instantiation that ran; a launch-shape suffix after the closing '>' (threads per block) is not part of it."""
    sys.path.insert(0, ROOT)
    import importlib

    bench = importlib.import_module("bench")
    name = "cogdl_b200::stream_kernel<float4,NV=1,weighted,SRC_ONE,U=4,MINB=5>"
    t0, src0 = bench.profiled_traffic(name)
    t1, src1 = bench.profiled_traffic(name + " block=64")
    assert t0 is not None and t0 == t1 and src0 == src1 and t0 > 100e6
    t2, why = bench.profiled_traffic("cogdl_b200::stream_kernel<float4,NV=1,weighted,SRC_ONE,U=4,MINB=4,dynamic>")
    assert t2 is None and "launched" in why
