// spmm.cu -- weighted / unweighted CSR SpMM for sm_100a.
//
//   Y[i,:] = sum_{p in row i} val[p] * X[colind[p],:]
//
// Replaces the reference's GE-SpMM kernels (cogdl/operators/spmm/spmm_kernel.cu:7-512) behind
// spmm.csr_spmm / csr_spmm_no_edge_value (spmm.cpp:22-70).  HBM/L2-bound gather: 4F+8 bytes per
// edge, no reuse => no tensor cores; what matters is coalescing, bytes in flight and balance.
//
// Design (B200):
//   * a GROUP of lanes owns one work item (a row, or one edge chunk of a hub row); each lane
//     owns NV 16-byte vectors of the feature row, so one warp-wide LDG.128 moves a full
//     512-byte row for F = 128 (the reference moves 2 x 128 B with scalar loads and re-reads
//     the indices per 64-column tile);
//   * colind/val are fetched one GROUP-wide coalesced slab at a time with L1::no_allocate and
//     broadcast by shuffle (no shared memory, no block barrier); the next slab is prefetched
//     while the current one is being gathered;
//   * UNROLL independent row gathers are issued before any is consumed (8 x 512 B per warp in
//     flight): with ~32 resident warps/SM that is ~128 KB in flight per SM, above the
//     latency x bandwidth product of HBM3e (~31 KB/SM) and of the L2 (~20 KB/SM);
//   * rows longer than the hub plan's chunk are cut into chunks that are scheduled first
//     (lowest block indices), partial sums are combined in chunk order by the last chunk to
//     arrive => deterministic; the reference serialises a hub on one warp;
//   * per output element the accumulation is in CSR order with separate fp32 mul and add, so
//     unsplit rows are bit-identical to the reference CPU SpMM (spmm_cpu.cpp:24-36);
//   * 64-bit row offsets (the reference's `int offset = colInd * k` overflows at N*F >= 2^31).
#include "common.cuh"
#include "stream.cuh"

#include <cuda_fp16.h>

#include <cstdlib>

namespace cogdl_b200 {

struct SpmmParams {
  const float *peers[8];
  int n_peers;
  int peer_shift;
  const int *rowptr;
  const int *colind;
  const float *val;
  const float *X0;
  const float *X1;
  int64_t n0;  // columns < n0 read X0, others X1[c - n0]
  float *Y;
  int64_t n_rows;
  int FV;      // feature row length in VecT units
  HubView hub;
};

template <typename VecT> __device__ __forceinline__ VecT vzero();
template <> __device__ __forceinline__ float4 vzero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ float vzero<float>() { return 0.f; }
__device__ __forceinline__ void add_rn(float &a, const float &b) { a = __fadd_rn(a, b); }

template <typename VecT, int GROUP, int NV, bool HAS_VAL>
__global__ void __launch_bounds__(256) spmm_kernel(const SpmmParams p) {
  constexpr int U0 = (NV == 1) ? 8 : (NV == 2 ? 4 : 2);
  constexpr int U = U0 < GROUP ? U0 : GROUP;
  constexpr int TILE = GROUP * NV;  // vector columns covered per pass
  const int lane = threadIdx.x & 31;
  const int gl = lane & (GROUP - 1);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t item = tid / GROUP;
  const WorkItem w = decode_item(item, p.n_rows, p.rowptr, p.hub);
  const bool warp_has_chunk = (tid - lane) / GROUP < p.hub.n_chunks;  // warp-uniform

  const VecT *X0 = reinterpret_cast<const VecT *>(p.X0);
  const VecT *X1 = reinterpret_cast<const VecT *>(p.X1);
  VecT *Y = reinterpret_cast<VecT *>(p.Y);
  VecT *P = reinterpret_cast<VecT *>(p.hub.partials);

  int maxdeg = w.hb - w.lb;
  if (GROUP < 32) maxdeg = warp_max(maxdeg);  // keep loop trip counts warp-uniform for the shuffles

  // Feature rows wider than one pass (F > 256 on the vector path) are covered by looping over
  // column tiles inside the warp (indices are re-read per tile, as the reference's grid.y tiles do).
  for (int tile0 = 0; tile0 < p.FV; tile0 += TILE) {
    const int cv = tile0 + gl;  // this lane's first vector column in this tile
    bool colok[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) colok[k] = (cv + k * GROUP) < p.FV;
    VecT acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = vzero<VecT>();

    int c = 0;
    float v = 0.f;
    {
      const int e = w.lb + gl;
      if (e < w.hb) {
        c = ld_stream(p.colind + e);
        v = HAS_VAL ? ld_stream(p.val + e) : 1.f;
      }
    }
    for (int off = 0; off < maxdeg; off += GROUP) {
      const int cnt = min(GROUP, w.hb - w.lb - off);  // <= 0 for groups whose row is already done
      int cn = 0;
      float vn = 0.f;
      {
        const int e = w.lb + off + GROUP + gl;  // prefetch the next index/value slab
        if (e < w.hb) {
          cn = ld_stream(p.colind + e);
          vn = HAS_VAL ? ld_stream(p.val + e) : 1.f;
        }
      }
#pragma unroll 1
      for (int j = 0; j < GROUP; j += U) {
        if (GROUP == 32) {
          if (j >= cnt) break;
        } else {
          if (!__any_sync(FULL, j < cnt)) break;
        }
        VecT x[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int cj = __shfl_sync(FULL, c, j + u, GROUP);
          if (j + u < cnt) {
            const VecT *xp = (cj < p.n0) ? X0 + (int64_t)cj * p.FV : X1 + ((int64_t)cj - p.n0) * p.FV;
#pragma unroll
            for (int k = 0; k < NV; ++k)
              if (colok[k]) x[u][k] = ld_gather(xp + cv + k * GROUP);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float vj = __shfl_sync(FULL, v, j + u, GROUP);
          if (j + u < cnt) {
#pragma unroll
            for (int k = 0; k < NV; ++k)
              if (colok[k]) axpy_rn(acc[k], vj, x[u][k]);
          }
        }
      }
      c = cn;
      v = vn;
    }

    if (!w.is_chunk) {
      if (w.active) {
        VecT *yp = Y + (int64_t)w.row * p.FV + cv;
#pragma unroll
        for (int k = 0; k < NV; ++k)
          if (colok[k]) st_stream(yp + k * GROUP, acc[k]);
      }
    } else {
      VecT *pp = P + (int64_t)w.slot * p.FV + cv;
#pragma unroll
      for (int k = 0; k < NV; ++k)
        if (colok[k]) st_cg(pp + k * GROUP, acc[k]);
    }
  }

  if (warp_has_chunk) {
    if (hub_arrive_last<GROUP>(w, p.hub, gl)) {
      // combine this row's partial sums in chunk order (deterministic, independent of arrival)
      for (int cv = gl; cv < p.FV; cv += GROUP) {
        const VecT *pp = P + (int64_t)w.first * p.FV + cv;
        VecT s = ld_cg(pp);
        for (int q = 1; q < w.n_row_chunks; ++q) add_rn(s, ld_cg(pp + (int64_t)q * p.FV));
        st_stream(Y + (int64_t)w.row * p.FV + cv, s);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Row-stream form (stream.cuh), used whenever the plan carries segments and a full warp owns the
// row (F >= 68, or >= 17 on the scalar path).  COGDL_B200_SPMM_VARIANT (experiments only) picks
// the unroll depth / occupancy target of the F <= 128 instantiation; the default is the variant
// measured fastest on B200 (profiles/).
// ------------------------------------------------------------------------------------------
static int spmm_variant() { return tuning("COGDL_B200_SPMM_VARIANT", 7); }

static StreamParams to_stream(const SpmmParams &p) {
  StreamParams q;
  q.rowptr = p.rowptr; q.colind = p.colind; q.val = p.val; q.att = nullptr; q.perm = nullptr;
  q.X0 = p.X0; q.X1 = p.X1; q.n0 = p.n0; q.Y = p.Y; q.ldv = p.FV; q.H = 1; q.FVL = p.FV; q.S = 1;
  q.hub = p.hub;
  q.n_peers = p.n_peers; q.peer_shift = p.peer_shift;
  for (int i = 0; i < 8; ++i) q.peers[i] = p.peers[i];
  return q;
}

template <typename VecT, int NV>
static int launch_spmm_stream(const SpmmParams &p, cudaStream_t stream) {
  const StreamParams q = to_stream(p);
  const int mode = p.val ? MODE_WEIGHTED : MODE_UNWEIGHTED;
  if constexpr (NV == 1 && sizeof(VecT) == 16) {
    // lean loop (stream_range_lean); the variants differ in gathers per batch (U) and resident blocks per SM
    switch (spmm_variant()) {
      case 1: return launch_stream<VecT, 1, 8, 4, false>(q, mode, stream);
      case 2: return launch_stream<VecT, 1, 8, 3, false>(q, mode, stream);
      case 3: return launch_stream<VecT, 1, 4, 6, false>(q, mode, stream);
      case 8: return launch_stream<VecT, 1, 4, 5, false, true>(q, mode, stream);   // round-1 loop + L2 evict_last hint on X (slower)
      // U=4, 48 registers (40 warps/SM): fastest measured (profiles/r01d_tune_stream_v2.txt, r02m_bench_n1.json)
      default: return launch_stream<VecT, 1, 4, 5, false>(q, mode, stream);
    }
  }
  constexpr int U = (NV == 1) ? 8 : (NV == 2 ? 4 : 2);
  return launch_stream<VecT, NV, U, 1>(q, mode, stream);
}

template <typename VecT, int GROUP, int NV>
static int launch_spmm(const SpmmParams &p, cudaStream_t stream) {
  const int64_t items = (int64_t)p.hub.n_chunks + p.n_rows;
  const int64_t threads = items * GROUP;
  const int64_t blocks = ceil_div(threads, 256);
  if (blocks == 0) return COGDL_B200_OK;
  if (blocks > 0x7fffffffLL)
    return set_error(COGDL_B200_EINVAL, "spmm: problem too large for one launch (blocks=%lld)", (long long)blocks);
  dim3 grid((unsigned)blocks);
  note_kernel("cogdl_b200::spmm_kernel<%s,GROUP=%d,NV=%d,%s>", sizeof(VecT) == 16 ? "float4" : "float", GROUP, NV,
              p.val ? "weighted" : "unweighted");
  if (p.val)
    spmm_kernel<VecT, GROUP, NV, true><<<grid, 256, 0, stream>>>(p);
  else
    spmm_kernel<VecT, GROUP, NV, false><<<grid, 256, 0, stream>>>(p);
  CB_LAUNCH_CHECK();
  return COGDL_B200_OK;
}

template <typename VecT>
static int dispatch_spmm(const SpmmParams &p, cudaStream_t s) {
  const int fv = p.FV;
  // peer form: only the row-stream kernel decodes (owner << shift | row) columns (SRC_PEERS); the
  // sub-warp row kernel below would read them as X1[c - n0].  NV = 1 masks lanes >= fv.
  if (p.n_peers > 0 && fv <= 32) return launch_spmm_stream<VecT, 1>(p, s);
  const bool stream = p.hub.n_segs > 0;
  // narrow rows (F = 40 -> 10 float4): the lean row-stream form with the idle lanes masked beats the sub-warp
  // row kernel once a row needs more than `stream_min_fv` vectors (measured, profiles/; tunable for experiments)
  const int stream_min_fv = tuning("COGDL_B200_SPMM_STREAM_MINFV", 9);
  if (stream && fv >= stream_min_fv && fv <= 16) return launch_spmm_stream<VecT, 1>(p, s);
  if (fv <= 1) return launch_spmm<VecT, 1, 1>(p, s);
  if (fv <= 2) return launch_spmm<VecT, 2, 1>(p, s);
  if (fv <= 4) return launch_spmm<VecT, 4, 1>(p, s);
  if (fv <= 8) return launch_spmm<VecT, 8, 1>(p, s);
  if (fv <= 16) return launch_spmm<VecT, 16, 1>(p, s);
  if (fv <= 32) return stream ? launch_spmm_stream<VecT, 1>(p, s) : launch_spmm<VecT, 32, 1>(p, s);
  if (sizeof(VecT) == 4 && fv > 64) return stream ? launch_spmm_stream<VecT, 4>(p, s) : launch_spmm<VecT, 32, 4>(p, s);
  return stream ? launch_spmm_stream<VecT, 2>(p, s) : launch_spmm<VecT, 32, 2>(p, s);
}

static int spmm_entry(const int32_t *rowptr, const int32_t *colind, const float *val, const float *X0,
                      int64_t n0, const float *X1, float *Y, int64_t n_rows, int64_t F,
                      const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream, const char *who,
                      const float *const *peers = nullptr, int n_peers = 0, int peer_shift = 0) {
  CB_REQUIRE(n_rows >= 0 && F >= 0, "%s: negative size (n_rows=%lld, F=%lld)", who, (long long)n_rows, (long long)F);
  if (n_rows == 0 || F == 0) return COGDL_B200_OK;
  CB_REQUIRE(rowptr && colind && X0 && Y, "%s: null pointer", who);
  CB_REQUIRE(n_rows < 0x7fffffffLL && F < 0x7fffffffLL, "%s: n_rows / F must fit int32", who);
  int rc = check_plan(plan, (plan ? (int64_t)plan->n_chunks : 0) * F * (int64_t)sizeof(float));
  if (rc) return rc;
  SpmmParams p;
  p.rowptr = rowptr; p.colind = colind; p.val = val; p.X0 = X0; p.X1 = X1 ? X1 : X0; p.n0 = n0;
  p.Y = Y; p.n_rows = n_rows; p.hub = hub_view(plan);
  p.n_peers = n_peers; p.peer_shift = peer_shift;
  bool peers_aligned = true;
  for (int i = 0; i < 8; ++i) {
    p.peers[i] = (i < n_peers) ? peers[i] : nullptr;
    peers_aligned = peers_aligned && aligned16(p.peers[i]);
  }
  if (n_peers > 0)
    CB_REQUIRE(p.hub.n_segs > 0 && F % 4 == 0 && F <= 512 && peers_aligned && aligned16(X0) && aligned16(Y),
               "%s: the peer form needs a plan with segments, F %% 4 == 0, F <= 512 and 16-byte aligned buffers", who);
  const bool vec = (F % 4 == 0) && aligned16(X0) && aligned16(p.X1) && aligned16(Y) &&
                   (p.hub.n_chunks == 0 || aligned16(p.hub.partials));
  if (vec) {
    p.FV = (int)(F / 4);
    return dispatch_spmm<float4>(p, (cudaStream_t)stream);
  }
  p.FV = (int)F;
  return dispatch_spmm<float>(p, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// fp16 storage, fp32 accumulation.  One lane owns 4 halves (8 bytes); a warp covers 128
// columns per pass.  (Reference: spmm_test{0,1,2}_half accumulate in half, spmm_kernel.cu:
// 222-250,311-368,442-512 -- we keep fp32 accumulators, a documented improvement.)
// ------------------------------------------------------------------------------------------
struct SpmmHalfParams {
  const int *rowptr;
  const int *colind;
  const __half *val;
  const __half *X;
  __half *Y;
  int64_t n_rows;
  int F;
  HubView hub;
};

template <bool HAS_VAL, bool VEC>
__global__ void __launch_bounds__(256) spmm_half_kernel(const SpmmHalfParams p) {
  constexpr int U = 8;
  constexpr int W = VEC ? 4 : 1;  // halves per lane per pass
  const int lane = threadIdx.x & 31;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const WorkItem w = decode_item(tid / 32, p.n_rows, p.rowptr, p.hub);
  float *P = reinterpret_cast<float *>(p.hub.partials);

  for (int tile0 = 0; tile0 < p.F; tile0 += 32 * W) {
    const int col = tile0 + lane * W;
    const bool colok = col < p.F;
    float acc[W];
#pragma unroll
    for (int t = 0; t < W; ++t) acc[t] = 0.f;

    for (int base = w.lb; base < w.hb; base += 32) {
      const int cnt = min(32, w.hb - base);
      int c = 0;
      float v = 0.f;
      if (base + lane < w.hb) {
        c = ld_stream(p.colind + base + lane);
        v = HAS_VAL ? __half2float(p.val[base + lane]) : 1.f;
      }
#pragma unroll 1
      for (int j = 0; j < cnt; j += U) {
        float x[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int cj = __shfl_sync(FULL, c, j + u);
          if (j + u < cnt && colok) {
            const __half *xp = p.X + (int64_t)cj * p.F + col;
            if constexpr (VEC) {
              const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(xp));
              const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
              const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
              x[u][0] = a.x; x[u][1] = a.y; x[u][2] = b.x; x[u][3] = b.y;
            } else {
              x[u][0] = __half2float(__ldg(xp));
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float vj = __shfl_sync(FULL, v, j + u);
          if (j + u < cnt && colok) {
#pragma unroll
            for (int t = 0; t < W; ++t) acc[t] = fmaf(vj, x[u][t], acc[t]);
          }
        }
      }
    }
    if (!w.is_chunk) {
      if (w.active && colok) {
        __half *yp = p.Y + (int64_t)w.row * p.F + col;
#pragma unroll
        for (int t = 0; t < W; ++t) yp[t] = __float2half_rn(acc[t]);
      }
    } else if (colok) {
#pragma unroll
      for (int t = 0; t < W; ++t) st_cg(P + (int64_t)w.slot * p.F + col + t, acc[t]);
    }
  }
  if ((tid - lane) / 32 < p.hub.n_chunks) {
    if (hub_arrive_last<32>(w, p.hub, lane)) {
      for (int col = lane; col < p.F; col += 32) {
        float s = 0.f;
        for (int q = 0; q < w.n_row_chunks; ++q) s += ld_cg(P + (int64_t)(w.first + q) * p.F + col);
        p.Y[(int64_t)w.row * p.F + col] = __float2half_rn(s);
      }
    }
  }
}

// Only the hub rows (degree > plan chunk) of Y = A @ X: the hub-chunk items of the row-stream kernel with
// their in-order combine, no segments.  Used by the fused GCN layer (fused_gcn.cu), which aggregates every
// other row straight into shared memory.  F % 4 == 0, 16-byte aligned X / Y.
int spmm_hub_rows_only(const int32_t *rowptr, const int32_t *colind, const float *val, const float *X, float *Y,
                       int64_t F, const cogdl_b200_hub_plan_t *plan, cudaStream_t stream) {
  SpmmParams p;
  p.rowptr = rowptr; p.colind = colind; p.val = val; p.X0 = X; p.X1 = X; p.n0 = INT64_MAX; p.Y = Y; p.n_rows = 0;
  p.hub = hub_view(plan);
  p.hub.n_segs = 0;
  p.n_peers = 0; p.peer_shift = 0;
  for (int i = 0; i < 8; ++i) p.peers[i] = nullptr;
  p.FV = (int)(F / 4);
  if (p.hub.n_chunks == 0) return COGDL_B200_OK;
  const StreamParams q = to_stream(p);
  if (p.FV <= 32) return launch_stream<float4, 1, 4, 5, false>(q, val ? MODE_WEIGHTED : MODE_UNWEIGHTED, stream);
  return launch_stream<float4, 2, 4, 1>(q, val ? MODE_WEIGHTED : MODE_UNWEIGHTED, stream);
}

}  // namespace cogdl_b200

using namespace cogdl_b200;

extern "C" int cogdl_b200_spmm_csr_f32(const int32_t *rowptr, const int32_t *colind, const float *val,
                                       const float *X, float *Y, int64_t n_rows, int64_t F,
                                       const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream) {
  return spmm_entry(rowptr, colind, val, X, INT64_MAX, nullptr, Y, n_rows, F, plan, stream,
                    "cogdl_b200_spmm_csr_f32");
}

extern "C" int cogdl_b200_spmm_csr_f32_2src(const int32_t *rowptr, const int32_t *colind, const float *val,
                                            const float *X0, int64_t n0, const float *X1, float *Y,
                                            int64_t n_rows, int64_t F, const cogdl_b200_hub_plan_t *plan,
                                            cogdl_b200_stream_t stream) {
  CB_REQUIRE(n0 >= 0, "cogdl_b200_spmm_csr_f32_2src: n0 < 0");
  CB_REQUIRE(X1 != nullptr || n0 == INT64_MAX, "cogdl_b200_spmm_csr_f32_2src: X1 is null");
  return spmm_entry(rowptr, colind, val, X0, n0, X1, Y, n_rows, F, plan, stream,
                    "cogdl_b200_spmm_csr_f32_2src");
}

extern "C" int cogdl_b200_spmm_csr_f16(const int32_t *rowptr, const int32_t *colind, const void *val,
                                       const void *X, void *Y, int64_t n_rows, int64_t F,
                                       const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream) {
  CB_REQUIRE(n_rows >= 0 && F >= 0, "cogdl_b200_spmm_csr_f16: negative size");
  if (n_rows == 0 || F == 0) return COGDL_B200_OK;
  CB_REQUIRE(rowptr && colind && X && Y, "cogdl_b200_spmm_csr_f16: null pointer");
  CB_REQUIRE(n_rows < 0x7fffffffLL && F < 0x7fffffffLL, "cogdl_b200_spmm_csr_f16: sizes must fit int32");
  int rc = check_plan(plan, (plan ? (int64_t)plan->n_chunks : 0) * F * (int64_t)sizeof(float));
  if (rc) return rc;
  SpmmHalfParams p;
  p.rowptr = rowptr; p.colind = colind; p.val = (const __half *)val; p.X = (const __half *)X;
  p.Y = (__half *)Y; p.n_rows = n_rows; p.F = (int)F; p.hub = hub_view(plan);
  const bool vec = (F % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 7u) == 0);
  const int64_t items = (int64_t)p.hub.n_chunks + n_rows;
  const int64_t blocks = ceil_div(items * 32, 256);
  CB_REQUIRE(blocks <= 0x7fffffffLL, "cogdl_b200_spmm_csr_f16: problem too large");
  dim3 grid((unsigned)blocks);
  cudaStream_t s = (cudaStream_t)stream;
  if (vec) {
    if (val) spmm_half_kernel<true, true><<<grid, 256, 0, s>>>(p);
    else spmm_half_kernel<false, true><<<grid, 256, 0, s>>>(p);
  } else {
    if (val) spmm_half_kernel<true, false><<<grid, 256, 0, s>>>(p);
    else spmm_half_kernel<false, false><<<grid, 256, 0, s>>>(p);
  }
  CB_LAUNCH_CHECK();
  return COGDL_B200_OK;
}

extern "C" int cogdl_b200_spmm_csr_f32_peers(const int32_t *rowptr, const int32_t *colind, const float *val,
                                             const float *X_local, int64_t n_local, const float *const *peer_ptrs,
                                             int32_t n_peers, int32_t owner_shift, float *Y, int64_t n_rows,
                                             int64_t F, const cogdl_b200_hub_plan_t *plan,
                                             cogdl_b200_stream_t stream) {
  CB_REQUIRE(peer_ptrs && n_peers >= 1 && n_peers <= 8, "cogdl_b200_spmm_csr_f32_peers: need 1..8 peer pointers");
  CB_REQUIRE(owner_shift >= 1 && owner_shift <= 28 && n_local >= 0, "cogdl_b200_spmm_csr_f32_peers: bad owner_shift / n_local");
  for (int i = 0; i < n_peers; ++i) CB_REQUIRE(peer_ptrs[i], "cogdl_b200_spmm_csr_f32_peers: null peer pointer %d", i);
  return spmm_entry(rowptr, colind, val, X_local, n_local, nullptr, Y, n_rows, F, plan, stream,
                    "cogdl_b200_spmm_csr_f32_peers", peer_ptrs, n_peers, owner_shift);
}
