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


# ---------------------------------------------------------------------------------------------
# Push form (reduce of boundary partial sums): host logic under gloo, the two SpMMs done by the oracle
# ---------------------------------------------------------------------------------------------
def _push_worker(rank, world, port, out_dir, weighted):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from cogdl_b200 import dist as cdist, synth

        n, e, F = 2500, 26000, 8
        rp, col = synth.powerlaw_csr(n, e, seed=7, locality=(world, 0.3))
        val = torch.rand(col.numel(), generator=torch.Generator().manual_seed(3)) if weighted else None
        X = torch.randn(n, F, generator=torch.Generator().manual_seed(4))
        bounds = cdist.balanced_row_ranges(rp, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        x_local = X[lo:hi].contiguous()
        cpu = torch.device("cpu")

        # (1) built from the replicated global CSR, (2) built from the row shard by shipping boundary edges
        p1 = cdist.PushPartition.from_global_csr(rp, col, val, rank, world, bounds)
        e0, e1 = int(rp[lo]), int(rp[hi])
        p2 = cdist.PushPartition.from_row_shard(rank, world, bounds, (rp[lo:hi + 1] - e0).clone(), col[e0:e1].clone(),
                                                None if val is None else val[e0:e1].clone())
        for a, b in [(p1.brow, p2.brow), (p1.b_rowptr, p2.b_rowptr), (p1.b_col, p2.b_col), (p1.c_rowptr, p2.c_rowptr),
                     (p1.c_col, p2.c_col), (p1.recv_row, p2.recv_row)]:
            assert torch.equal(a, b)                              # same structure either way, edge for edge
        assert p1.send_counts == p2.send_counts and p1.recv_counts == p2.recv_counts
        if weighted:
            assert torch.equal(p1.b_val, p2.b_val) and torch.equal(p1.c_val, p2.c_val)
        part = p1
        # every edge of the column slice lands in exactly one block
        col_mine = int(((col >= lo) & (col < hi)).sum())
        assert part.nnz_interior + part.nnz_boundary == col_mine
        assert part.send_counts[rank] == 0 and part.n_brow == sum(part.send_counts) and part.n_brow > 0
        bt = torch.tensor(bounds)
        owner = torch.searchsorted(bt, part.brow, right=True) - 1
        assert bool((owner != rank).all()) and bool((owner[1:] >= owner[:-1]).all())   # grouped by destination rank
        # combined block: interior columns first inside every row, then the partial slots in ascending order
        for i in torch.randint(0, part.n_local, (200,), generator=torch.Generator().manual_seed(rank)).tolist():
            c = part.c_col[int(part.c_rowptr[i]):int(part.c_rowptr[i + 1])]
            is_slot = c >= part.n_local
            k = int(is_slot.sum())
            assert not bool(is_slot[: c.numel() - k].any()) and bool(is_slot[c.numel() - k:].all())
            assert torch.equal(c[is_slot], torch.sort(c[is_slot]).values)
            assert bool((part.recv_row[c[is_slot] - part.n_local] == i).all())

        ps = cdist.PushSpMM(part, cpu)                            # host tensors: exchange plumbing only
        f32 = lambda t: None if t is None else t.numpy().astype(np.float32)
        P = oracle.spmm_csr(part.b_rowptr.numpy(), part.b_col.numpy(), f32(part.b_val), x_local.numpy())
        R = ps.reduce_scatter(torch.from_numpy(P), F)
        assert R.shape == (part.n_recv, F)
        y_local = oracle.spmm_csr(part.c_rowptr.numpy(), part.c_col.numpy(), f32(part.c_val),
                                  torch.cat([x_local, R]).numpy())
        y_global = oracle.spmm_csr(rp.numpy(), col.numpy(), f32(val), X.numpy())
        # different summation order from the single-process CSR loop => tolerance, not bit equality
        ref = y_global[lo:hi]
        scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=1, keepdims=True))
        assert float((np.abs(y_local - ref) / np.maximum(scale, 1e-30)).max()) <= 1e-5
        # exact statement: in fp64 the push form reproduces A @ X to rounding
        y64 = np.zeros((n, F))
        rows = np.repeat(np.arange(n), np.diff(rp.numpy()))
        contrib = X.numpy().astype(np.float64)[col.numpy()]
        if val is not None:
            contrib *= val.numpy().astype(np.float64)[:, None]
        np.add.at(y64, rows, contrib)
        assert float(np.abs(y_local - y64[lo:hi]).max()) <= 1e-5 * float(np.abs(y64).max())
        # determinism: the same inputs give the same bits
        R2 = ps.reduce_scatter(torch.from_numpy(P), F)
        assert torch.equal(R, R2)
        open(os.path.join(out_dir, f"push_ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,weighted", [(2, True), (3, False)])
def test_push_form_host_logic(tmp_path, world, weighted):
    mp.spawn(_push_worker, args=(world, _free_port(), str(tmp_path), weighted), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"push_ok{r}") for r in range(world))


def _push_edge_worker(rank, world, port, out_dir):
    """No boundary edges at all (every column inside the row's own range) and an empty rank: the exchange carries
    zero rows and the combined block is the interior block."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from cogdl_b200 import dist as cdist, synth

        n, e, F = 1200, 9000, 4
        rp, col = synth.powerlaw_csr(n, e, seed=9, locality=(2, 0.0), self_loops=False)   # two ranges of 600 nodes, no crossing edge
        bounds = [0, 600, 1200, 1200][: world + 1] if world == 3 else [0, 600, 1200]     # world 3: rank 2 owns nothing
        lo, hi = bounds[rank], bounds[rank + 1]
        X = torch.randn(n, F, generator=torch.Generator().manual_seed(1))
        part = cdist.PushPartition.from_global_csr(rp, col, None, rank, world, bounds)
        assert part.n_brow == 0 and part.n_recv == 0 and sum(part.send_counts) == 0
        assert part.n_local == hi - lo and part.c_rowptr.numel() == part.n_local + 1
        ps = cdist.PushSpMM(part, torch.device("cpu"))
        R = ps.reduce_scatter(torch.zeros((0, F)), F)
        assert R.shape == (0, F)
        if part.n_local:
            y = oracle.spmm_csr(part.c_rowptr.numpy(), part.c_col.numpy(), None, X[lo:hi].numpy())
            assert np.array_equal(y, oracle.spmm_csr(rp.numpy(), col.numpy(), None, X.numpy())[lo:hi])   # same order: exact
        open(os.path.join(out_dir, f"edge_ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_push_form_without_boundary_and_with_an_empty_rank(tmp_path, world):
    mp.spawn(_push_edge_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"edge_ok{r}") for r in range(world))
