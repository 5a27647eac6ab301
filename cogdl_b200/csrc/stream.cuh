// stream.cuh -- the row-stream gather kernel shared by the SpMM and the multi-head SpMM (sm_100a).
//
// Work items, in grid order:
//   [0, n_chunks)                 hub chunks: <= chunk_edges edges of ONE hub row -> partial sum in
//                                 scratch, combined in chunk order by the last chunk to arrive;
//   [n_chunks, n_chunks+n_segs)   segments: runs of consecutive NON-hub rows (~seg_cost rows+edges).
// In multi-head mode every item exists once per 512-byte slice of the [H*F] row (adjacent warps).
//
// One warp per item.  A segment's edges are contiguous in colind/val/edge_row, so the warp walks
// them in 32-edge slabs: one coalesced load each for colind, the value (or permutation) and the
// owning row; `row != row of the next edge` gives a 32-bit "row ends here" mask with one ballot.
// Gathers are issued U at a time and are always full batches except at the segment tail, whatever
// the row lengths (a row-per-warp kernel issues only `degree` gathers behind a rowptr -> colind -> X
// dependent chain).  Per output element the accumulation is strictly in CSR order with separate
// fp32 multiply and add => bit-identical to the reference CPU loop (spmm_cpu.cpp:24-36) for every
// row that is not a hub.  At a set mask bit the accumulator is stored to its row and cleared.
#pragma once
#include "common.cuh"

namespace cogdl_b200 {

struct StreamParams {
  const int *rowptr;
  const int *colind;
  const float *val;   // MODE 1: [nnz]
  const float *att;   // MODE 2: [nnz, H]
  const int *perm;    // MODE 2, nullable: attention row of edge p is perm[p]
  const float *X0;
  const float *X1;    // second source (two-source form), rows >= n0
  int64_t n0;
  // peer form (node-range partition, fused NVLink gather): a column c >= n0 encodes
  // (owner << peer_shift | row inside the owner's shard) after subtracting n0; peers[owner] is that
  // rank's feature shard mapped into this process (symmetric memory / CUDA IPC).
  const float *peers[8];
  int n_peers;
  int peer_shift;
  float *Y;
  int ldv;            // dense row length in VecT units
  int H;              // MODE 2: heads
  int FVL;            // MODE 2: VecT units per head
  int S;              // MODE 2: 512-byte slices per row (1 otherwise)
  HubView hub;
};

enum { MODE_UNWEIGHTED = 0, MODE_WEIGHTED = 1, MODE_MULTIHEAD = 2 };
// Threads per block of every row-stream launch (COGDL_B200_STREAM_BLOCK overrides): measured on B200
// (profiles/r02s_ab_launch_shapes.md) 128 vs 256: F=256 253 -> 216 us, F=40 68.6 -> 66.6, mh-SpMM 981 -> 962,
// products F=128 4769 -> 4718, arxiv F=128 unchanged; all outputs bit-identical.
constexpr int STREAM_BLOCK_DEFAULT = 128;
enum { SRC_ONE = 0, SRC_TWO = 1, SRC_PEERS = 2 };

template <typename VecT> __device__ __forceinline__ VecT sk_zero();
template <> __device__ __forceinline__ float4 sk_zero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ float sk_zero<float>() { return 0.f; }
__device__ __forceinline__ void sk_add(float &a, const float &b) { a = __fadd_rn(a, b); }
__device__ __forceinline__ void sk_add(float4 &a, const float4 &b) { add_rn(a, b); }

// Stream the edge range [e, e_end).  ROWS: flush at row ends (segment); otherwise accumulate the
// whole range into acc (hub chunk).
template <typename VecT, int NV, int MODE, bool HAS_PERM, int SRC, int U, bool ROWS, bool PREFETCH, bool HINT>
__device__ __forceinline__ void stream_range(const StreamParams &p, int e, const int e_end, const int cv,
                                             const bool (&colok)[NV], const int head, const int lane,
                                             VecT (&acc)[NV]) {
  const VecT *X0 = reinterpret_cast<const VecT *>(p.X0) + cv;
  const VecT *X1 = reinterpret_cast<const VecT *>(p.X1) + cv;
  VecT *Y = reinterpret_cast<VecT *>(p.Y) + cv;
  const uint64_t pol = HINT ? l2_policy_evict_last() : 0;

  // slab registers: column, value / attention row, owning row, owning row of the next edge
  int c = 0, rid = -1, rnx = -1, pe = 0;
  float v = 0.f;
  auto load_slab = [&](int base, int &c_, float &v_, int &pe_, int &rid_, int &rnx_) {
    const int q = base + lane;
    c_ = 0; v_ = 0.f; pe_ = 0; rid_ = -1; rnx_ = -1;
    if (q < e_end) {
      c_ = ld_stream(p.colind + q);
      if (MODE == MODE_WEIGHTED) v_ = ld_stream(p.val + q);
      if (MODE == MODE_MULTIHEAD) pe_ = HAS_PERM ? ld_stream(p.perm + q) : q;
      if (ROWS) {
        rid_ = ld_stream(p.hub.edge_row + q);
        if (q + 1 < e_end) rnx_ = __ldg(p.hub.edge_row + q + 1);
      }
    }
  };
  load_slab(e, c, v, pe, rid, rnx);

  for (; e < e_end; e += 32) {
    const int cnt = min(32, e_end - e);
    const unsigned endmask = ROWS ? __ballot_sync(FULL, lane < cnt && rid != rnx) : 0u;
    int cn, pn, ridn, rnxn;
    float vn;
    if (PREFETCH) load_slab(e + 32, cn, vn, pn, ridn, rnxn);  // next slab rides behind this slab's gathers

#pragma unroll 1
    for (int j = 0; j < cnt; j += U) {
      const bool full = (j + U <= cnt);
      const unsigned em = endmask >> j;
      VecT x[U][NV];
      float a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int cj = __shfl_sync(FULL, c, j + u);
        const int pj = (MODE == MODE_MULTIHEAD) ? __shfl_sync(FULL, pe, j + u) : 0;
        if (full || j + u < cnt) {
          const VecT *xp;
          if (SRC == SRC_ONE || cj < p.n0) {
            xp = X0 + (int64_t)cj * p.ldv;
          } else if (SRC == SRC_TWO) {
            xp = X1 + ((int64_t)cj - p.n0) * p.ldv;
          } else {  // remote row: read it from the owner's HBM over NVLink
            const unsigned r = (unsigned)(cj - (int)p.n0);
            xp = reinterpret_cast<const VecT *>(p.peers[r >> p.peer_shift]) + cv +
                 (int64_t)(r & ((1u << p.peer_shift) - 1u)) * p.ldv;
          }
#pragma unroll
          for (int k = 0; k < NV; ++k)
            if (colok[k]) x[u][k] = HINT ? ld_gather_hint(xp + k * 32, pol) : ld_gather(xp + k * 32);
          if (MODE == MODE_MULTIHEAD && colok[0]) a[u] = __ldg(p.att + (int64_t)pj * p.H + head);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float vj = (MODE == MODE_WEIGHTED) ? __shfl_sync(FULL, v, j + u) : 0.f;
        const int rj = ROWS ? __shfl_sync(FULL, rid, j + u) : 0;
        if (full || j + u < cnt) {
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            if (colok[k]) {
              if (MODE == MODE_UNWEIGHTED) sk_add(acc[k], x[u][k]);          // 1.0f * x == x exactly
              else if (MODE == MODE_WEIGHTED) axpy_rn(acc[k], vj, x[u][k]);
              else axpy_rn(acc[k], a[u], x[u][k]);
            }
          }
          if (ROWS && ((em >> u) & 1u)) {   // last edge of row rj: store and restart (warp-uniform)
#pragma unroll
            for (int k = 0; k < NV; ++k)
              if (colok[k]) { st_stream(Y + (int64_t)rj * p.ldv + k * 32, acc[k]); acc[k] = sk_zero<VecT>(); }
          }
        }
      }
    }
    if (PREFETCH) { c = cn; v = vn; pe = pn; rid = ridn; rnx = rnxn; }
    else load_slab(e + 32, c, v, pe, rid, rnx);
  }
}

// Lean form of stream_range for the plain SpMM (NV == 1, weighted / unweighted, any source form).  The
// round-1 loop spent ~47 warp-instructions per edge (ncu: issue-bound at 62 % issue-active on the L2-resident
// arxiv shape): three SHFL broadcasts per edge, each guarded by a BRA.DIV convergence check, and two
// predicate branches with BSSY/BSYNC pairs around every gather and every accumulate.  Here the slab
// (column, value, row) is parked in shared memory once (3 STS per 32 edges) and read back with broadcast LDS
// (no convergence requirement, column + value in ONE LDS.64), full batches run without per-edge predicates, and
// the row-end test is made once per batch on the ballot mask: ~15 instructions per edge.  Same accumulation
// order and rounding as stream_range (CSR order, separate fp32 multiply and add) => still bit-identical to
// the reference CPU loop on unsplit rows.
template <typename VecT, int MODE, int SRC, int U, bool ROWS>
__device__ __forceinline__ void stream_range_lean(const StreamParams &p, int e, const int e_end, const int cv, const bool colok,
                                                  const int lane, int2 *s_cv, int *s_r, VecT &acc) {
  const VecT *X0 = reinterpret_cast<const VecT *>(p.X0) + cv;
  const VecT *X1 = reinterpret_cast<const VecT *>(p.X1) + cv;
  VecT *Y = reinterpret_cast<VecT *>(p.Y) + cv;
  auto src_ptr = [&](int cj) -> const VecT * {
    if (SRC == SRC_ONE || cj < p.n0) return X0 + (int64_t)cj * p.ldv;
    if (SRC == SRC_TWO) return X1 + ((int64_t)cj - p.n0) * p.ldv;
    const unsigned r = (unsigned)(cj - (int)p.n0);      // remote row: read it from the owner's HBM over NVLink
    return reinterpret_cast<const VecT *>(p.peers[r >> p.peer_shift]) + cv + (int64_t)(r & ((1u << p.peer_shift) - 1u)) * p.ldv;
  };
  auto accumulate = [&](float vj, const VecT &x) {
    if (MODE == MODE_UNWEIGHTED) sk_add(acc, x); else axpy_rn(acc, vj, x);
  };
  for (; e < e_end; e += 32) {
    const int cnt = min(32, e_end - e);
    const int q = e + lane;
    int c = 0, rid = -1, rnx = -1;
    float v = 0.f;
    if (q < e_end) {
      c = ld_stream(p.colind + q);
      if (MODE == MODE_WEIGHTED) v = ld_stream(p.val + q);
      if (ROWS) {
        rid = ld_stream(p.hub.edge_row + q);
        if (q + 1 < e_end) rnx = __ldg(p.hub.edge_row + q + 1);
      }
    }
    const unsigned endmask = ROWS ? __ballot_sync(FULL, lane < cnt && rid != rnx) : 0u;
    __syncwarp();                                   // everybody is done reading the previous slab
    s_cv[lane] = make_int2(c, __float_as_int(v));
    if (ROWS) s_r[lane] = rid;
    __syncwarp();
    int j = 0;
#pragma unroll 1
    for (; j + U <= cnt; j += U) {                  // full batches: no per-edge predicates
      int2 cw[U];
      VecT x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) cw[u] = s_cv[j + u];            // broadcast LDS.64
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (colok) x[u] = ld_gather(src_ptr(cw[u].x));
      const unsigned em = ROWS ? ((endmask >> j) & ((1u << U) - 1u)) : 0u;
      if (em == 0u) {                               // no row ends inside this batch
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (colok) accumulate(__int_as_float(cw[u].y), x[u]);
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (colok) accumulate(__int_as_float(cw[u].y), x[u]);
          if ((em >> u) & 1u) {                     // last edge of its row: store and restart (warp-uniform)
            const int rj = s_r[j + u];
            if (colok) { st_stream(Y + (int64_t)rj * p.ldv, acc); acc = sk_zero<VecT>(); }
          }
        }
      }
    }
#pragma unroll 1
    for (; j < cnt; ++j) {                          // tail of the last slab: one edge at a time
      const int2 cw = s_cv[j];
      if (colok) accumulate(__int_as_float(cw.y), ld_gather(src_ptr(cw.x)));
      if (ROWS && ((endmask >> j) & 1u)) {
        const int rj = s_r[j];
        if (colok) { st_stream(Y + (int64_t)rj * p.ldv, acc); acc = sk_zero<VecT>(); }
      }
    }
  }
}

// One work item (wid = item * S + slice) processed by one warp.
template <typename VecT, int NV, int MODE, bool HAS_PERM, int SRC, int U, bool PREFETCH, bool HINT>
__device__ __forceinline__ void stream_item(const StreamParams &p, const int64_t wid, const int lane, int2 *s_cv, int *s_r) {
  constexpr int TILE = 32 * NV;
  const int64_t item = (MODE == MODE_MULTIHEAD) ? wid / p.S : wid;
  const int slice = (MODE == MODE_MULTIHEAD) ? (int)(wid - item * p.S) : 0;
  VecT *Y = reinterpret_cast<VecT *>(p.Y);
  VecT *P = reinterpret_cast<VecT *>(p.hub.partials);
  constexpr bool LEAN = (NV == 1 && MODE != MODE_MULTIHEAD && !HINT);

  // column tiles: the multi-head form owns exactly one 32-vector slice; the plain form loops over
  // the row in TILE-vector passes (one pass for F <= 128 * NV)
  const int t_begin = (MODE == MODE_MULTIHEAD) ? slice * 32 : 0;
  const int t_end = (MODE == MODE_MULTIHEAD) ? t_begin + 1 : p.ldv;

  if (item < p.hub.n_chunks) {
    const WorkItem w = decode_item(item, 0, p.rowptr, p.hub);
    for (int tile0 = t_begin; tile0 < t_end; tile0 += TILE) {
      const int cv = tile0 + lane;
      bool colok[NV];
      VecT acc[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) { colok[k] = (cv + k * 32) < p.ldv; acc[k] = sk_zero<VecT>(); }
      const int head = (MODE == MODE_MULTIHEAD && colok[0]) ? cv / p.FVL : 0;
      if constexpr (LEAN) stream_range_lean<VecT, MODE, SRC, U, false>(p, w.lb, w.hb, cv, colok[0], lane, s_cv, s_r, acc[0]);
      else stream_range<VecT, NV, MODE, HAS_PERM, SRC, U, false, PREFETCH, HINT>(p, w.lb, w.hb, cv, colok, head, lane, acc);
#pragma unroll
      for (int k = 0; k < NV; ++k)
        if (colok[k]) st_cg(P + (int64_t)w.slot * p.ldv + cv + k * 32, acc[k]);
    }
    WorkItem wa = w;
    wa.n_row_chunks = w.n_row_chunks * p.S;   // every slice of every chunk arrives once
    if (hub_arrive_last<32>(wa, p.hub, lane)) {
      for (int cv = lane; cv < p.ldv; cv += 32) {
        const VecT *pp = P + (int64_t)w.first * p.ldv + cv;
        VecT s = ld_cg(pp);
        int q = 1;
        for (; q + 8 <= w.n_row_chunks; q += 8) {      // 8 partials in flight, added in chunk order (a 22 K-edge
          VecT t[8];                                   // hub has 354 of them: this loop is the kernel's tail)
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] = ld_cg(pp + (int64_t)(q + u) * p.ldv);
#pragma unroll
          for (int u = 0; u < 8; ++u) sk_add(s, t[u]);
        }
        for (; q < w.n_row_chunks; ++q) sk_add(s, ld_cg(pp + (int64_t)q * p.ldv));
        st_stream(Y + (int64_t)w.row * p.ldv + cv, s);
      }
    }
    return;
  }

  const int64_t seg = item - p.hub.n_chunks;
  if (seg >= p.hub.n_segs) return;
  const int2 rr = __ldg(p.hub.segs + seg);
  const int e_begin = __ldg(p.rowptr + rr.x), e_end = __ldg(p.rowptr + rr.y);

  for (int tile0 = t_begin; tile0 < t_end; tile0 += TILE) {
    const int cv = tile0 + lane;
    bool colok[NV];
    VecT acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) { colok[k] = (cv + k * 32) < p.ldv; acc[k] = sk_zero<VecT>(); }
    const int head = (MODE == MODE_MULTIHEAD && colok[0]) ? cv / p.FVL : 0;
    if (p.hub.n_empty_rows > 0) {
      // rows without edges never show up in the edge stream: zero-fill them here
      for (int rb = rr.x; rb < rr.y; rb += 32) {
        const int row = rb + lane;
        unsigned m = __ballot_sync(FULL, row < rr.y && __ldg(p.rowptr + row + 1) == __ldg(p.rowptr + row));
        while (m) {
          const int k0 = __ffs(m) - 1;
          m &= m - 1;
#pragma unroll
          for (int k = 0; k < NV; ++k)
            if (colok[k]) st_stream(Y + (int64_t)(rb + k0) * p.ldv + cv + k * 32, sk_zero<VecT>());
        }
      }
    }
    if constexpr (LEAN) stream_range_lean<VecT, MODE, SRC, U, true>(p, e_begin, e_end, cv, colok[0], lane, s_cv, s_r, acc[0]);
    else stream_range<VecT, NV, MODE, HAS_PERM, SRC, U, true, PREFETCH, HINT>(p, e_begin, e_end, cv, colok, head, lane, acc);
  }
}

template <typename VecT, int NV, int MODE, bool HAS_PERM, int SRC, int U, int MINB, bool PREFETCH, bool HINT>
__global__ void __launch_bounds__(256, MINB) stream_kernel(const StreamParams p) {
  const int lane = threadIdx.x & 31;
  constexpr bool LEAN = (NV == 1 && MODE != MODE_MULTIHEAD && !HINT);
  __shared__ int2 s_cv_all[LEAN ? 8 : 1][32];
  __shared__ int s_r_all[LEAN ? 8 : 1][32];
  int2 *s_cv = s_cv_all[LEAN ? (threadIdx.x >> 5) : 0];
  int *s_r = s_r_all[LEAN ? (threadIdx.x >> 5) : 0];
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  stream_item<VecT, NV, MODE, HAS_PERM, SRC, U, PREFETCH, HINT>(p, wid, lane, s_cv, s_r);
}

// Launch helper: picks the instantiation for (mode, perm, two-source); U / MINB fixed by the caller.
template <typename VecT, int NV, int U, int MINB, bool PREFETCH = true, bool HINT = false>
static int launch_stream(const StreamParams &p, int mode, cudaStream_t stream) {
  const int64_t warps = ((int64_t)p.hub.n_chunks + p.hub.n_segs) * p.S;
  // Threads per block: a launch parameter only -- the kernel indexes its per-warp slabs by warp-in-block and takes its
  // item from the global warp index, so every block size runs the same SASS and gives the same bits.  A block's warp
  // slots go to the next block only when its SLOWEST warp has retired; item durations spread (gathers that miss L2,
  // hub merges), so smaller blocks keep more warp slots busy -- most visibly for the two-vectors-per-lane
  // instantiation (F = 256).  A persistent form (grid = resident warp slots, items drawn from a ticket counter in the
  // plan) was also built and measured: slower (107 vs 101 us at arxiv F=128: the ticket loop needs 64 registers =
  // 32 warps/SM) -- removed.
  int bs = tuning("COGDL_B200_STREAM_BLOCK", STREAM_BLOCK_DEFAULT);
  if (bs != 32 && bs != 64 && bs != 128) bs = 256;
  const int64_t blocks = ceil_div(warps * 32, bs);
  if (blocks == 0) return COGDL_B200_OK;
  if (blocks > 0x7fffffffLL) return set_error(COGDL_B200_EINVAL, "stream kernel: problem too large for one launch");
  const unsigned g = (unsigned)blocks;
  const int src = p.n_peers > 0 ? SRC_PEERS : (p.n0 != INT64_MAX ? SRC_TWO : SRC_ONE);
  note_kernel("cogdl_b200::stream_kernel<%s,NV=%d,%s%s,%s,U=%d,MINB=%d%s%s>", sizeof(VecT) == 16 ? "float4" : "float", NV,
              mode == MODE_MULTIHEAD ? "multihead" : (mode == MODE_WEIGHTED ? "weighted" : "unweighted"),
              (mode == MODE_MULTIHEAD && p.perm) ? "+perm" : "",
              src == SRC_PEERS ? "SRC_PEERS" : (src == SRC_TWO ? "SRC_TWO" : "SRC_ONE"), U, MINB,
              PREFETCH ? ",prefetch" : "", HINT ? ",l2hint" : "");
  if (bs != 256) append_kernel_note(" block=%d", bs);
  if (mode == MODE_MULTIHEAD) {
    if constexpr (NV == 1) {
      if (p.perm) stream_kernel<VecT, 1, MODE_MULTIHEAD, true, SRC_ONE, U, MINB, PREFETCH, HINT><<<g, bs, 0, stream>>>(p);
      else stream_kernel<VecT, 1, MODE_MULTIHEAD, false, SRC_ONE, U, MINB, PREFETCH, HINT><<<g, bs, 0, stream>>>(p);
    } else {
      return set_error(COGDL_B200_EINVAL, "stream kernel: multi-head form needs NV == 1");
    }
  } else if (mode == MODE_WEIGHTED) {
    if (src == SRC_PEERS) stream_kernel<VecT, NV, MODE_WEIGHTED, false, SRC_PEERS, U, MINB, PREFETCH, HINT><<<g, bs, 0, stream>>>(p);
    else if (src == SRC_TWO) stream_kernel<VecT, NV, MODE_WEIGHTED, false, SRC_TWO, U, MINB, PREFETCH, HINT><<<g, bs, 0, stream>>>(p);
    else stream_kernel<VecT, NV, MODE_WEIGHTED, false, SRC_ONE, U, MINB, PREFETCH, HINT><<<g, bs, 0, stream>>>(p);
  } else {
    if (src == SRC_PEERS) stream_kernel<VecT, NV, MODE_UNWEIGHTED, false, SRC_PEERS, U, MINB, PREFETCH, HINT><<<g, bs, 0, stream>>>(p);
    else if (src == SRC_TWO) stream_kernel<VecT, NV, MODE_UNWEIGHTED, false, SRC_TWO, U, MINB, PREFETCH, HINT><<<g, bs, 0, stream>>>(p);
    else stream_kernel<VecT, NV, MODE_UNWEIGHTED, false, SRC_ONE, U, MINB, PREFETCH, HINT><<<g, bs, 0, stream>>>(p);
  }
  CB_LAUNCH_CHECK();
  return COGDL_B200_OK;
}

}  // namespace cogdl_b200
