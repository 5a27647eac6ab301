"""Two ranks on ONE GPU (`gloo` rendezvous, host-staged exchange) running the REAL kernels of the two
exchange forms of the node-range partitioned SpMM:

  * pull / halo form  : pack (gather kernel) -> all-to-all -> two-source SpMM       (PartitionedSpMM, mode "nccl")
  * push form         : boundary SpMM -> reduce-scatter of partial sums -> two-source SpMM with unit-weight
                        slots for the received partials                             (PushSpMM)

NCCL refuses two ranks on one device, so the exchange itself travels through the host here; what this
test adds over tests/test_dist_gloo.py (host logic, oracle SpMMs) is that every SpMM is the sm_100a
kernel, on renumbered shards, against the oracle on the whole graph.  The multi-GPU NVLink form is
checked inside `bench.py --gpus N` (parity block) and by tools/dist_gpu_check.py.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rowscale_err(got, ref):
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=1, keepdims=True))
    return float((np.abs(got - ref) / np.maximum(scale, 1e-30)).max())


def _worker(rank, world, port, out_dir, weighted, F):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from cogdl_b200 import dist as cdist, synth

        dev = torch.device("cuda:0")
        n, e = 6000, 90000                     # power-law: a few rows exceed the 64-edge chunk => hub chunks too
        rp, col = synth.powerlaw_csr(n, e, seed=11, locality=(world, 0.3))
        val = torch.rand(col.numel(), generator=torch.Generator().manual_seed(5)) if weighted else None
        X = torch.randn(n, F, generator=torch.Generator().manual_seed(6))
        f32 = lambda t: None if t is None else t.numpy().astype(np.float32)
        y_ref = oracle.spmm_csr(rp.numpy(), col.numpy(), f32(val), X.numpy())

        # ---- pull / halo form with the real pack + two-source kernels
        ps = cdist.partition_global_csr(rp, col, val, rank, world, dev, mode="nccl")
        lo, hi = ps.part.lo, ps.part.hi
        x_local = X[lo:hi].contiguous().to(dev)
        y_pull = ps.spmm(x_local).cpu().numpy()
        # same order inside a row as the single-process CSR loop => unsplit rows are bit-identical
        deg = np.diff(rp.numpy())[lo:hi]
        unsplit = deg <= ps.st.chunk_edges
        assert np.array_equal(y_pull[unsplit], y_ref[lo:hi][unsplit])
        assert _rowscale_err(y_pull, y_ref[lo:hi]) <= 1e-5

        # ---- push form: boundary partial sums reduce-scattered, folded into the interior SpMM
        part = cdist.PushPartition.from_global_csr(rp, col, val, rank, world, ps.part.bounds)
        push = cdist.PushSpMM(part, dev)
        assert part.n_brow > 0 and part.n_recv > 0
        y_push = push.spmm(x_local)
        y_push2 = push.spmm(x_local)
        assert torch.equal(y_push, y_push2)                       # deterministic: fixed (source rank, row) order
        assert _rowscale_err(y_push.cpu().numpy(), y_ref[lo:hi]) <= 1e-5
        # the partial sums themselves, against the oracle on the boundary block
        P = push.partials(x_local).cpu().numpy()
        P_ref = oracle.spmm_csr(part.b_rowptr.numpy(), part.b_col.numpy(), f32(part.b_val), X[lo:hi].numpy())
        assert _rowscale_err(P, P_ref) <= 1e-5
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("weighted,F", [(True, 128), (False, 40)])
def test_pull_and_push_forms_two_ranks_one_gpu(tmp_path, weighted, F):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), weighted, F), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
