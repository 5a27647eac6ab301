"""Deterministic synthetic power-law CSR graphs -- the benchmark / test inputs.

Degrees follow node weights w_i ~ rank_i^(-alpha) (alpha = 0.8), randomly permuted over the node
ids; E destination endpoints are drawn from w (inverse-CDF sampling, so it scales to 10^8 nodes on
the GPU), columns are uniform -- or, with `locality`, drawn inside the row's own node-range
partition with probability 1 - beta (multi-GPU experiments).  Within-row order is left as
generated (unsorted), like a CSR built from an arbitrary COO list.
"""
import math

import torch

# (nodes, edges) of the shapes named in BASELINE.json
SHAPES = {
    "cora": (2708, 10556),
    "arxiv": (169343, 1166243),
    "products": (2449029, 61859140),
    "papers100M": (111059956, 1615685872),
}


def powerlaw_degrees(n, e, seed=0, device="cpu", alpha=0.8, chunk=1 << 26):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rank = torch.arange(1, n + 1, device=device, dtype=torch.float64)
    w = rank.pow_(-alpha)
    w = w[torch.randperm(n, generator=g, device=device)]
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    deg = torch.zeros(n, dtype=torch.int64, device=device)
    done = 0
    while done < e:
        m = min(chunk, e - done)
        u = torch.rand(m, generator=g, device=device, dtype=torch.float64)
        idx = torch.searchsorted(cdf, u).clamp_(max=n - 1)
        deg += torch.bincount(idx, minlength=n)
        done += m
    return deg, g


def powerlaw_csr(n, e, seed=0, device="cpu", self_loops=True, alpha=0.8, locality=None, index_dtype=torch.int64):
    """Returns (row_ptr [n+1], col [nnz]) of dtype `index_dtype`.

    self_loops: append one (i, i) edge at the end of every row (as add_remaining_self_loops +
                coo2csr would, cogdl/data/data.py:175-191).
    locality  : None, or (num_parts, beta): contiguous equal node ranges; a column falls inside
                the row's own range with probability 1 - beta, anywhere otherwise.
    """
    deg, g = powerlaw_degrees(n, e, seed, device, alpha)
    row_ptr_raw = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=row_ptr_raw[1:])
    col = torch.randint(0, n, (e,), generator=g, device=device, dtype=torch.int64)
    if locality is not None:
        parts, beta = locality
        psize = math.ceil(n / parts)
        rows = torch.repeat_interleave(torch.arange(n, device=device), deg)
        lo = (rows // psize) * psize
        width = torch.clamp(lo + psize, max=n) - lo
        local = lo + (torch.rand(e, generator=g, device=device, dtype=torch.float64) * width).long()
        keep_remote = torch.rand(e, generator=g, device=device) < beta
        col = torch.where(keep_remote, col, local)
        del rows, lo, width, local, keep_remote
    if self_loops:
        # interleave: row i's edges then (i, i)
        new_ptr = row_ptr_raw + torch.arange(n + 1, device=device)
        out = torch.empty(e + n, dtype=torch.int64, device=device)
        is_loop = torch.zeros(e + n, dtype=torch.bool, device=device)
        is_loop[new_ptr[1:] - 1] = True
        out[is_loop] = torch.arange(n, device=device)
        out[~is_loop] = col
        col, row_ptr_raw = out, new_ptr
    return row_ptr_raw.to(index_dtype), col.to(index_dtype)


def sym_norm_weights(row_ptr, col):
    """d_i^-1/2 * d_j^-1/2 with d = CSR row degree -- what Graph.sym_norm() bakes into the weights
    (cogdl/utils/graph_utils.py:82-89) when the graph is symmetric; used as GCN-like edge values."""
    deg = (row_ptr[1:] - row_ptr[:-1]).to(torch.float32)
    dinv = deg.pow(-0.5)
    dinv[torch.isinf(dinv)] = 0
    rows = torch.repeat_interleave(torch.arange(deg.numel(), device=row_ptr.device), (row_ptr[1:] - row_ptr[:-1]).long())
    return dinv[col.long()] * dinv[rows]


def graph_stats(row_ptr):
    deg = row_ptr[1:] - row_ptr[:-1]
    return {"n": int(deg.numel()), "nnz": int(row_ptr[-1]), "max_deg": int(deg.max()), "mean_deg": float(deg.float().mean())}


# ---------------------------------------------------------------------------------------------
# Workload definitions shared by bench.py's two arms (ours / --impl reference).  This module is
# deliberately free of any `cogdl_b200` import so the reference arm can load it by file path
# without dlopen-ing libcogdl_b200.so.
# ---------------------------------------------------------------------------------------------
PAPERS_ROWS_PER_GPU = SHAPES["papers100M"][0] // 8      # 13 882 494
PAPERS_EDGES_PER_GPU = SHAPES["papers100M"][1] // 8     # 201 960 734


def arxiv_description(max_degree):
    return ("spmm hidden=128 (fp32) on ogbn-arxiv-shaped power-law CSR: 169343 nodes, 1166243 edges + 169343 self "
            "loops = 1335586 nnz, sym-normalised weights, max degree %d, seed 0 [BASELINE configs[1], GCN layer-1 "
            "aggregation]" % int(max_degree))


def shard_description(rows, edges, world, beta, seed, hidden=128, scaling="weak"):
    n_total = rows * world
    what = ("1/8 of papers100M per GPU" if scaling == "weak" else f"papers100M split {world}-way")
    return (f"unweighted spmm hidden={hidden} (fp32), papers100M-shaped power-law CSR partitioned by node range: "
            f"{rows} rows + {edges} edges per GPU x {world} GPUs (= {n_total} nodes, {edges * world} edges), "
            f"columns remote-eligible with prob beta={beta}, seed {seed} [BASELINE configs[4] shape, {what}]")


def shard_sizes(world, scaling="weak"):
    """(rows, edges) per GPU: weak = 1/8 of papers100M per GPU whatever the GPU count;
    strong = the whole papers100M-shaped graph split `world` ways."""
    if scaling == "strong":
        n, e = SHAPES["papers100M"]
        return n // world, e // world
    return PAPERS_ROWS_PER_GPU, PAPERS_EDGES_PER_GPU


def shard_csr(rank, world, rows, edges, beta, seed=0, device="cpu", max_slice_edges=None):
    """Rank `rank`'s shard of the locality-controlled papers100M-shaped graph: power-law
    degrees inside the shard; a column is uniform over the WHOLE graph with probability beta (remote
    with prob. beta*(P-1)/P) and uniform inside the own node range otherwise.  Returns
    (row_ptr int64 [r+1], col int64 GLOBAL ids [nnz]).  max_slice_edges: keep only the leading rows
    holding at least that many edges (CPU baseline sample; the degree law is still the full shard's)."""
    n_total = rows * world
    lo = rank * rows
    deg, g = powerlaw_degrees(rows, edges, seed=seed * 1000 + rank, device=device)
    row_ptr = torch.zeros(rows + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=row_ptr[1:])
    del deg
    n_edges = edges
    if max_slice_edges is not None and max_slice_edges < edges:
        r = int(torch.searchsorted(row_ptr, torch.tensor([max_slice_edges], device=device, dtype=torch.int64))[0])
        r = max(1, min(r, rows))
        row_ptr = row_ptr[: r + 1].clone()
        n_edges = int(row_ptr[-1])
    col = torch.empty(n_edges, dtype=torch.int64, device=device)
    chunk = 1 << 26
    for s in range(0, n_edges, chunk):
        m = min(chunk, n_edges - s)
        local = lo + torch.randint(0, rows, (m,), generator=g, device=device)
        if beta > 0:
            anywhere = torch.randint(0, n_total, (m,), generator=g, device=device)
            remote = torch.rand(m, generator=g, device=device) < beta
            col[s:s + m] = torch.where(remote, anywhere, local)
            del anywhere, remote
        else:
            col[s:s + m] = local
        del local
    return row_ptr, col
