"""torchrun script (N GPUs, NCCL): the node-range partitioned SpMM -- halo (NCCL all-to-all), fused NVLink
gather (p2p) and push (reduce of boundary partial sums) forms -- against the single-process oracle on a small
locality-controlled graph.  Usage:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import oracle
    from cogdl_b200 import dist as cdist, synth

    n, e, F = 20000, 300000, 128
    rp, col = synth.powerlaw_csr(n, e, seed=7, locality=(world, 0.1))
    val = torch.rand(col.numel(), generator=torch.Generator().manual_seed(1))
    X = torch.randn(n, F, generator=torch.Generator().manual_seed(2))
    full = oracle.spmm_csr(rp.numpy(), col.numpy(), val.numpy(), X.numpy())
    ok = torch.tensor([1], device=dev)
    for mode in ("nccl", "p2p"):
        ps = cdist.partition_global_csr(rp, col, val, rank, world, dev, mode=mode)
        lo, hi = ps.part.lo, ps.part.hi
        y = ps.spmm(X[lo:hi].to(dev).contiguous())
        torch.cuda.synchronize()
        ref = full[lo:hi]
        got = y.cpu().numpy()
        err = float(np.abs(got - ref).max() / np.abs(ref).max())
        deg = np.diff(rp.numpy())[lo:hi]
        exact = np.array_equal(got[deg <= ps.st.chunk_edges], ref[deg <= ps.st.chunk_edges])
        if not (err <= 1e-5 and exact):
            ok.zero_()
        print(f"[rank {rank}] mode={mode} rows [{lo},{hi}) halo {ps.n_halo} rel_err {err:.2e} unsplit_rows_bit_exact {exact}", flush=True)
        dist.barrier()
    # push form (reduce of boundary partial sums): both constructors, against the same oracle
    bounds = cdist.balanced_row_ranges(rp, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    e0, e1 = int(rp[lo]), int(rp[hi])
    x_local = X[lo:hi].to(dev).contiguous()
    for how in ("global", "row_shard"):
        if how == "global":
            part = cdist.PushPartition.from_global_csr(rp, col, val, rank, world, bounds, comm_device=dev)
        else:
            part = cdist.PushPartition.from_row_shard(rank, world, bounds, (rp[lo:hi + 1] - e0).to(dev), col[e0:e1].to(dev),
                                                      val[e0:e1].to(dev))
        push = cdist.PushSpMM(part, dev)
        y = push.spmm(x_local)
        same = torch.equal(y, push.spmm(x_local))
        torch.cuda.synchronize()
        ref = full[lo:hi]
        got = y.cpu().numpy()
        scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=1, keepdims=True))
        err = float((np.abs(got - ref) / np.maximum(scale, 1e-30)).max())
        if not (err <= 1e-5 and same):
            ok.zero_()
        print(f"[rank {rank}] mode=push({how}) rows [{lo},{hi}) boundary rows sent {part.n_brow} received {part.n_recv} "
              f"elementwise_err {err:.2e} deterministic {same}", flush=True)
        dist.barrier()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if int(ok) == 1 else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(ok) == 1 else 1)


if __name__ == "__main__":
    main()
