"""Node-range partitioned SpMM across the GPUs of one box (new functionality -- the
reference's sparse ops are single-GPU only).

One process per GPU (torch.distributed, NCCL over NVLink 5 / NVSwitch).  Rank p owns a contiguous
range of destination rows [lo_p, hi_p) of A together with the matching rows of X and Y.  Output
rows are independent, so the only coupling is the gather of X rows owned by other ranks:

    setup (once per structure)   columns outside [lo_p, hi_p) are renumbered to n_local + k where k
                                 indexes this rank's sorted, de-duplicated halo list; the halo lists
                                 are exchanged so every owner knows which of its rows each peer needs
    step                         pack   : send_buf = X_local[send_index]          (gather kernel)
                                 move   : all-to-all of the packed rows over NVLink (NCCL)
                                 compute: Y_local = A_local @ [X_local ; X_halo]   (two-source SpMM,
                                          cogdl_b200_spmm_csr_f32_2src -- no concatenation copy)

This is the pull form of the boundary exchange; the push form ("all-reduce of boundary partial
sums", north star) moves the same order of bytes -- (#distinct remote rows) * 4F per rank -- but
needs the column-sliced matrix and an extra reduction pass, so the pull form is what is built.
There is no all-reduce on the data path: a row's sum is completed by exactly one rank.
"""
import torch
import torch.distributed as dist

from . import synth
from .structure import CSRStructure


def balanced_row_ranges(row_ptr, parts):
    """Contiguous row ranges balanced by cost = nnz + rows (power-law graphs: NOT by row count)."""
    n = row_ptr.numel() - 1
    cost = row_ptr.to(torch.int64) + torch.arange(n + 1, device=row_ptr.device, dtype=torch.int64)
    total = int(cost[-1])
    targets = torch.tensor([total * k // parts for k in range(parts + 1)], device=row_ptr.device, dtype=torch.int64)
    bounds = torch.searchsorted(cost, targets).clamp_(max=n)
    bounds[0], bounds[-1] = 0, n
    return [int(b) for b in bounds.tolist()]


class LocalPartition:
    """What one rank holds of the partitioned graph (index arithmetic only; device agnostic)."""

    def __init__(self, rank, world, bounds, row_ptr_local, col_global, val_local):
        self.rank, self.world, self.bounds = rank, world, list(bounds)
        lo, hi = bounds[rank], bounds[rank + 1]
        self.lo, self.hi, self.n_local = lo, hi, hi - lo
        col_global = col_global.to(torch.int64)
        is_local = (col_global >= lo) & (col_global < hi)
        remote = col_global[~is_local]
        self.halo = torch.unique(remote)                                # sorted global ids this rank must fetch
        self.n_halo = int(self.halo.numel())
        col_local = torch.empty_like(col_global)
        col_local[is_local] = col_global[is_local] - lo
        col_local[~is_local] = self.n_local + torch.searchsorted(self.halo, remote)
        self.row_ptr = row_ptr_local
        self.col = col_local
        self.val = val_local
        self.nnz_local = int(col_global.numel())
        bt = torch.tensor(self.bounds, device=self.halo.device, dtype=torch.int64)
        owner = torch.searchsorted(bt, self.halo, right=True) - 1       # owner rank of every halo row
        self.recv_counts = torch.bincount(owner, minlength=world).tolist()   # rows I receive from each rank
        self.halo_owner_local = self.halo - bt[owner]                   # row index inside the owner's shard
        # peer encoding (fused NVLink gather): remote column -> n_local + (owner << shift | row in owner's shard)
        max_rows = max(bounds[k + 1] - bounds[k] for k in range(world))
        self.peer_shift = max(1, int(max_rows - 1).bit_length())
        if self.n_local + (world << self.peer_shift) >= 2**31:
            self.col_peer = None      # does not fit int32: only the halo form is available
        else:
            rem_owner = torch.searchsorted(bt, remote, right=True) - 1
            col_peer = col_local.clone()
            col_peer[~is_local] = self.n_local + (rem_owner << self.peer_shift) + (remote - bt[rem_owner])
            self.col_peer = col_peer


def exchange_index_lists(part, group=None, comm_device=None):
    """Tell every owner which of its rows this rank needs.  Returns (send_index, send_counts):
    send_index = my local row ids to pack, concatenated in peer order.  comm_device: where the
    (small) index lists live during the exchange -- NCCL needs CUDA tensors even when the
    partition was computed on the host."""
    world = part.world
    dev = part.halo.device if comm_device is None else comm_device
    recv_counts = torch.tensor(part.recv_counts, dtype=torch.int64, device=dev)
    send_counts = torch.empty(world, dtype=torch.int64, device=dev)
    _all_to_all(send_counts, recv_counts, [1] * world, [1] * world, group)
    send_counts_l = [int(v) for v in send_counts.tolist()]
    send_index = torch.empty(sum(send_counts_l), dtype=torch.int64, device=dev)
    _all_to_all(send_index, part.halo_owner_local.to(dev).contiguous(), send_counts_l, part.recv_counts, group)
    return send_index.to(part.halo.device), send_counts_l


def _all_to_all(out, inp, out_splits, in_splits, group=None):
    """all_to_all_single on NCCL; point-to-point batch on backends without it (gloo in CPU tests)."""
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
        return
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    outs = list(out.split(out_splits)) if sum(out_splits) else [out[:0]] * world
    ins = list(inp.split(in_splits)) if sum(in_splits) else [inp[:0]] * world
    outs[rank].copy_(ins[rank])
    ops = []
    for peer in range(world):
        if peer == rank:
            continue
        if in_splits[peer]:
            ops.append(dist.P2POp(dist.isend, ins[peer].contiguous(), peer, group))
        if out_splits[peer]:
            ops.append(dist.P2POp(dist.irecv, outs[peer], peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class PartitionedSpMM:
    """Per-rank state + the step  y_local = A_local @ X  of the node-range partitioned SpMM.

    mode "p2p"  (default on CUDA): feature shards live in torch symmetric memory; the SpMM kernel
                reads remote rows straight from the owner's HBM over NVLink
                (cogdl_b200_spmm_csr_f32_peers) -- gather and compute are ONE kernel.
    mode "nccl": pack (gather kernel) -> all_to_all_single -> two-source SpMM on [X_local ; X_halo].
    """

    def __init__(self, part: LocalPartition, device, group=None, global_nnz=None, description="", mode=None):
        self.part, self.device, self.group = part, device, group
        self.n_local, self.n_halo, self.nnz_local = part.n_local, part.n_halo, part.nnz_local
        self.global_nnz = global_nnz
        self._desc = description
        self.world = part.world
        nccl = dist.get_backend(group) == "nccl"
        if mode is None:
            mode = "p2p" if (device.type == "cuda" and nccl and part.col_peer is not None) else "nccl"
        self.mode = mode
        self.x_local = None
        self._symm = None
        self._symm_by_F = {}
        if mode == "p2p":
            self.exchange = "fused: remote rows read from peer HBM inside the SpMM kernel (symmetric memory, NVLink P2P)"
            self.st = CSRStructure.from_int64(part.row_ptr.to(device), part.col_peer.to(device),
                                              n_cols=part.n_local)
            self.st.plan
            self.val = None if part.val is None else part.val.to(device).float().contiguous()
            rows = torch.tensor([part.n_local], device=device, dtype=torch.int64)
            dist.all_reduce(rows, op=dist.ReduceOp.MAX, group=group)
            self._symm_rows = int(rows)
            return
        comm_dev = device if nccl else None
        self.send_index, self.send_counts = exchange_index_lists(part, group, comm_dev)
        self.recv_counts = part.recv_counts
        self.exchange = "NCCL all_to_all_single of packed halo rows" if nccl else "p2p isend/irecv (gloo)"
        if device.type == "cuda":
            self.st = CSRStructure.from_int64(part.row_ptr.to(device), part.col.to(device),
                                              n_cols=part.n_local + part.n_halo)
            self.st.plan
            self.val = None if part.val is None else part.val.to(device).float().contiguous()
            self.send_index32 = self.send_index.to(device).to(torch.int32).contiguous()

    def describe(self):
        return self._desc

    # ------------------------------------------------------------------ p2p form
    def features(self, F):
        """The local feature shard [n_local, F] inside symmetric memory (write X here; peers read it).
        One symmetric buffer per feature width, kept for the life of the object (a multi-layer model
        alternates widths; re-allocating + re-rendezvousing while a slower peer's kernel still reads
        the old shard would be a use-after-free across ranks)."""
        ent = self._symm_by_F.get(F)
        if ent is None:
            import ctypes

            import torch.distributed._symmetric_memory as symm

            buf = symm.empty((self._symm_rows, F), dtype=torch.float32, device=self.device)
            hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
            ent = self._symm_by_F[F] = (buf, hdl, ptrs)
        self._symm = ent
        return ent[0][: self.n_local]

    def release(self):
        """Barrier that must precede ANY write to a shard returned by features(): a faster rank may
        only overwrite its shard once every peer's previous SpMM kernel has finished gathering from
        it (write-after-read across ranks).  spmm(x) does this itself when it copies x in."""
        if self._symm is not None:
            self._symm[1].barrier(channel=1)

    def spmm_p2p(self):
        """Y_local from the features currently in the symmetric shards (barrier, then ONE kernel)."""
        import ctypes

        from . import _cabi
        from .structure import _ptr, _stream

        buf, hdl, ptrs = self._symm
        F = buf.shape[1]
        hdl.barrier()   # every rank's shard is written before anyone gathers from it
        dev = self.device
        with torch.cuda.device(dev):
            y = torch.empty((self.n_local, F), dtype=torch.float32, device=dev)
            plan, keep = self.st.plan_struct(self.st.plan.n_chunks * F * 4)
            _cabi.call("cogdl_b200_spmm_csr_f32_peers", _ptr(self.st.rowptr), _ptr(self.st.colind), _ptr(self.val),
                       _ptr(buf), self.n_local, ctypes.cast(ptrs, ctypes.c_void_p), self.world, self.part.peer_shift,
                       _ptr(y), self.n_local, F, plan, _stream(dev))
            del keep
        return y

    # ------------------------------------------------------------------ nccl (halo) form: three stages
    def pack(self, x_local):
        if x_local.is_cuda:
            from .operators._raw import gather_rows_raw

            return gather_rows_raw(self.send_index32, x_local)
        return x_local.index_select(0, self.send_index)   # host tensors: gloo plumbing tests only

    def exchange_rows(self, send_buf, F):
        halo = torch.empty((self.n_halo, F), dtype=send_buf.dtype, device=send_buf.device)
        _all_to_all(halo.view(-1), send_buf.view(-1), [c * F for c in self.recv_counts],
                    [c * F for c in self.send_counts], self.group)
        return halo

    def local_spmm(self, x_local, x_halo):
        from .operators._raw import spmm_2src_raw

        return spmm_2src_raw(self.st, self.val, x_local, x_halo)

    # ------------------------------------------------------------------ the step
    def spmm(self, x_local):
        if self.mode == "p2p":
            shard = self.features(x_local.shape[1])
            if x_local.data_ptr() != shard.data_ptr():
                self.release()              # peers' previous kernels are done reading this shard
                shard.copy_(x_local)        # callers that keep X in features() skip barrier + copy
            return self.spmm_p2p()
        F = x_local.shape[1]
        halo = self.exchange_rows(self.pack(x_local), F)
        return self.local_spmm(x_local, halo)

    def last_kernel_seconds(self, step, torch_mod, iters=5):
        """Mean duration of the SpMM kernel alone (CUDA events, max over ranks), for the roofline."""
        if self.mode == "p2p":
            fn = self.spmm_p2p
        else:
            F = self.x_local.shape[1]
            halo = self.exchange_rows(self.pack(self.x_local), F)
            fn = lambda: self.local_spmm(self.x_local, halo)
        for _ in range(2):
            fn()
        a, b = torch_mod.cuda.Event(enable_timing=True), torch_mod.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch_mod.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / iters / 1e3], device=self.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t)


def partition_global_csr(row_ptr, col, val, rank, world, device, group=None, mode=None):
    """Slice a (replicated, host or device) global CSR into this rank's LocalPartition."""
    bounds = balanced_row_ranges(row_ptr, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    e0, e1 = int(row_ptr[lo]), int(row_ptr[hi])
    part = LocalPartition(rank, world, bounds, (row_ptr[lo:hi + 1] - e0).clone(), col[e0:e1].clone(),
                          None if val is None else val[e0:e1].clone())
    return PartitionedSpMM(part, device, group, global_nnz=int(row_ptr[-1]), mode=mode)


# papers100M-shaped graph, 1/8 of it per GPU (BASELINE configs[4]): fixed per-GPU work => weak scaling
PAPERS_ROWS_PER_GPU = synth.PAPERS_ROWS_PER_GPU
PAPERS_EDGES_PER_GPU = synth.PAPERS_EDGES_PER_GPU


def synthetic_partition(rank, world, device, seed=0, rows=PAPERS_ROWS_PER_GPU, edges=PAPERS_EDGES_PER_GPU,
                        beta=0.05, hidden=128, group=None, mode=None, scaling="weak"):
    """Generate this rank's shard directly on its GPU (synth.shard_csr: the locality-controlled
    generator of SURVEY 8d) and wrap it in a PartitionedSpMM with X resident in the feature shard."""
    row_ptr, col = synth.shard_csr(rank, world, rows, edges, beta, seed=seed, device=device)
    bounds = [k * rows for k in range(world + 1)]
    part = LocalPartition(rank, world, bounds, row_ptr, col, None)
    del col
    desc = synth.shard_description(rows, edges, world, beta, seed, hidden, scaling)
    ps = PartitionedSpMM(part, device, group, global_nnz=edges * world, description=desc, mode=mode)
    gen = torch.Generator(device=device).manual_seed(seed + rank)
    if ps.mode == "p2p":
        ps.x_local = ps.features(hidden)          # X lives in the symmetric shard: no copy per step
        ps.x_local.normal_(generator=gen)
    else:
        ps.x_local = torch.randn(rows, hidden, device=device, generator=gen)
    return ps
