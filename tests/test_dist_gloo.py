"""world_size = 2 `gloo` test (CPU) of the node-range partition host logic: row ranges, halo
renumbering, index-list exchange, row exchange.  The local SpMM itself is checked with the oracle
on the renumbered shard (the CUDA kernel needs a GPU; see test_gpu_api.py for the 2-source kernel)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from cogdl_b200 import dist as cdist, synth

        n, e, F = 3000, 30000, 8
        rp, col = synth.powerlaw_csr(n, e, seed=5, locality=(world, 0.2))
        val = torch.rand(col.numel(), generator=torch.Generator().manual_seed(1))
        X = torch.randn(n, F, generator=torch.Generator().manual_seed(2))
        ps = cdist.partition_global_csr(rp, col, val, rank, world, torch.device("cpu"))
        part = ps.part
        lo, hi = part.lo, part.hi
        # ranges are contiguous, cover everything, and are balanced by cost (not by row count)
        assert part.bounds[0] == 0 and part.bounds[-1] == n and part.bounds == sorted(part.bounds)
        costs = [int(rp[part.bounds[k + 1]] - rp[part.bounds[k]]) + part.bounds[k + 1] - part.bounds[k] for k in range(world)]
        assert max(costs) <= 1.2 * (sum(costs) / world) + 1000
        x_local = X[lo:hi].contiguous()
        halo = ps.exchange_rows(ps.pack(x_local), F)
        assert torch.equal(halo, X[part.halo])                   # every fetched row is the right row
        assert part.n_halo > 0 and int(part.col.max()) < part.n_local + part.n_halo
        # peer encoding used by the fused NVLink gather: n_local + (owner << shift | row in owner's shard)
        cp = part.col_peer
        e0, e1 = int(rp[lo]), int(rp[hi])
        glob = col[e0:e1]
        local = cp < part.n_local
        assert torch.equal(cp[local] + lo, glob[local])
        r = cp[~local] - part.n_local
        owner, idx = r >> part.peer_shift, r & ((1 << part.peer_shift) - 1)
        bt = torch.tensor(part.bounds)
        assert torch.equal(bt[owner] + idx, glob[~local]) and bool((owner != rank).all())
        assert bool((idx < (bt[owner + 1] - bt[owner])).all())
        y_local = oracle.spmm_csr(part.row_ptr.numpy(), part.col.numpy(), part.val.numpy(),
                                  torch.cat([x_local, halo]).numpy())
        y_global = oracle.spmm_csr(rp.numpy(), col.numpy(), val.numpy(), X.numpy())
        assert np.array_equal(y_local, y_global[lo:hi])          # same order inside a row => bit exact
        # sum of shards == single-process result: gather and compare on rank 0
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, y_local))
        if rank == 0:
            full = np.concatenate([g[2] for g in sorted(gathered, key=lambda t: t[0])])
            assert np.array_equal(full, y_global)
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_partitioned_spmm_host_logic_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_balanced_row_ranges_handles_hubs():
    from cogdl_b200.dist import balanced_row_ranges

    deg = torch.ones(1000, dtype=torch.int64)
    deg[0] = 100000  # one huge hub at the front
    rp = torch.zeros(1001, dtype=torch.int64)
    torch.cumsum(deg, 0, out=rp[1:])
    b = balanced_row_ranges(rp, 4)
    assert b[0] == 0 and b[-1] == 1000 and b == sorted(b)
    assert b[1] <= 2  # the hub row alone fills the first ranges
