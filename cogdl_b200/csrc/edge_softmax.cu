// edge_softmax.cu -- per-destination-row softmax over edge logits [nnz, H], forward, backward and the
// GAT attention form (logits computed on the fly), for sm_100a.
//
// Replaces edge_softmax / edge_softmax_backward (cogdl/operators/edge_softmax/edge_softmax.cu:7-98:
// one block of 32 x H threads per row, three passes over the row with lanes striding H floats apart
// -- <= 1/H sector efficiency -- and exp() evaluated twice) and, in MODE 2, the
// `leakyrelu(h_l[row] + h_r[col])` gathers + LeakyReLU + edge_softmax chain of
// cogdl/layers/gat_layer.py:73-74.
//
// Key layout fact: in CSR order the [deg, H] logits of row i are ONE contiguous block of deg*H floats
// at rowptr[i]*H, and consecutive rows are consecutive blocks.  With H a power of two <= 32, the
// element at offset t has head t % H, so a thread striding by a multiple of H keeps its head and
// per-head reductions are xor-shuffles over lane strides 16 .. H.
//
// Rows are tiered by their element count n = deg * H (power-law graphs: median row 3 edges, top hub
// 22 K edges on the arxiv shape):
//   segments   rows of degree <= plan chunk: MANY rows per warp, staged through shared memory
//              (es_seg_kernel) -- one coalesced read and write of the array, no per-row latency chain;
//   warp       hub rows, n <= 1024: one warp, 3 passes through L1 with 4 loads in flight;
//   block      hub rows, n <= 8192: one 256-thread block;
//   cluster    larger rows: a thread-block CLUSTER of 8 x 1024 threads; per-head max / sum are
//              combined across the 8 CTAs through distributed shared memory (DSMEM, cluster.sync),
//              so a 22 K-edge hub is reduced by 8 SMs instead of serialising on one.
// Without a plan every row takes the warp path.  Other head counts (not a power of two, or > 32)
// use a generic strided kernel.
//   MODE 0: forward, a = logits        MODE 1: backward, a = y, b = g  -> y * (g - sum_row y*g)
//   MODE 2: attention, a = h_l [N,H], b = h_r [N,H]: logit = leakyrelu(a[row,h] + b[col,h])
#include "common.cuh"

#include <cooperative_groups.h>
#include <math_constants.h>

namespace cg = cooperative_groups;

namespace cogdl_b200 {

struct EsParams {
  const int *rowptr;
  const float *a;
  const float *b;
  float *out;
  int64_t n_rows;         // rows (or entries of row_list) covered by a warp / block launch
  int H;
  int lgH;
  const int *row_list;    // process row_list[i] instead of row i (nullable)
  int64_t n_lo, n_hi;     // a warp / block launch handles rows with n_lo < deg*H <= n_hi
  int n_segs;             // segment kernel
  const int2 *segs;
  const int *edge_row;
  const int *colind;      // MODE 2
  float slope;            // MODE 2
};

constexpr int64_t WARP_ROW_ELEMS = 1024;    // hub rows up to this many elements: one warp
constexpr int64_t BLOCK_ROW_ELEMS = 8192;   // up to this: one 256-thread block; beyond: 8-CTA cluster
constexpr int CLUSTER_CTAS = 8;

// reduce across lanes that share lane % H  (strides 16 .. H)
__device__ __forceinline__ float head_max(float v, int H) {
  for (int s = 16; s >= H; s >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, s));
  return v;
}
__device__ __forceinline__ float head_sum(float v, int H) {
  for (int s = 16; s >= H; s >>= 1) v += __shfl_xor_sync(FULL, v, s);
  return v;
}

// Forward input of element t of the row whose block starts at edge lb (head = t % H).
template <int MODE>
__device__ __forceinline__ float es_in(const EsParams &p, int lb, int64_t t, int head, float hl) {
  if (MODE == 2) {
    const int c = __ldg(p.colind + lb + (int)(t >> p.lgH));
    const float z = hl + __ldg(p.b + (int64_t)c * p.H + head);
    return z > 0.f ? z : z * p.slope;
  }
  return __ldg(p.a + (int64_t)lb * p.H + t);
}

// One row handled by a team of NT threads striding the row's elements (tid = index inside the
// team, NT % H == 0).  reduce(v, is_max) must return the team-wide per-head reduction.
template <int MODE, int NT, typename Reduce>
__device__ __forceinline__ void es_row(const EsParams &p, int64_t row, int lb, int64_t n, int tid, Reduce reduce) {
  const int head = tid & (p.H - 1);
  float *o = p.out + (int64_t)lb * p.H;
  if (MODE == 1) {
    const float *y = p.a + (int64_t)lb * p.H, *g = p.b + (int64_t)lb * p.H;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int64_t t = tid;
    for (; t + 3 * NT < n; t += 4 * NT) {
      s0 = fmaf(__ldg(y + t), __ldg(g + t), s0); s1 = fmaf(__ldg(y + t + NT), __ldg(g + t + NT), s1);
      s2 = fmaf(__ldg(y + t + 2 * NT), __ldg(g + t + 2 * NT), s2);
      s3 = fmaf(__ldg(y + t + 3 * NT), __ldg(g + t + 3 * NT), s3);
    }
    for (; t < n; t += NT) s0 = fmaf(__ldg(y + t), __ldg(g + t), s0);
    const float s = reduce((s0 + s1) + (s2 + s3), false);
    t = tid;
#pragma unroll 4
    for (; t < n; t += NT) st_stream(o + t, __ldg(y + t) * (__ldg(g + t) - s));
    return;
  }
  const float hl = (MODE == 2) ? __ldg(p.a + row * p.H + head) : 0.f;
  float m0 = -CUDART_INF_F, m1 = m0, m2 = m0, m3 = m0;
  int64_t t = tid;
  for (; t + 3 * NT < n; t += 4 * NT) {
    const float x0 = es_in<MODE>(p, lb, t, head, hl), x1 = es_in<MODE>(p, lb, t + NT, head, hl);
    const float x2 = es_in<MODE>(p, lb, t + 2 * NT, head, hl), x3 = es_in<MODE>(p, lb, t + 3 * NT, head, hl);
    m0 = fmaxf(m0, x0); m1 = fmaxf(m1, x1); m2 = fmaxf(m2, x2); m3 = fmaxf(m3, x3);
  }
  for (; t < n; t += NT) m0 = fmaxf(m0, es_in<MODE>(p, lb, t, head, hl));
  const float m = reduce(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)), true);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  t = tid;
  for (; t + 3 * NT < n; t += 4 * NT) {
    const float x0 = es_in<MODE>(p, lb, t, head, hl), x1 = es_in<MODE>(p, lb, t + NT, head, hl);
    const float x2 = es_in<MODE>(p, lb, t + 2 * NT, head, hl), x3 = es_in<MODE>(p, lb, t + 3 * NT, head, hl);
    s0 += expf(x0 - m); s1 += expf(x1 - m); s2 += expf(x2 - m); s3 += expf(x3 - m);
  }
  for (; t < n; t += NT) s0 += expf(es_in<MODE>(p, lb, t, head, hl) - m);
  const float s = reduce((s0 + s1) + (s2 + s3), false);
  t = tid;
  for (; t + 3 * NT < n; t += 4 * NT) {
    const float x0 = es_in<MODE>(p, lb, t, head, hl), x1 = es_in<MODE>(p, lb, t + NT, head, hl);
    const float x2 = es_in<MODE>(p, lb, t + 2 * NT, head, hl), x3 = es_in<MODE>(p, lb, t + 3 * NT, head, hl);
    st_stream(o + t, expf(x0 - m) / s); st_stream(o + t + NT, expf(x1 - m) / s);
    st_stream(o + t + 2 * NT, expf(x2 - m) / s); st_stream(o + t + 3 * NT, expf(x3 - m) / s);
  }
  for (; t < n; t += NT) st_stream(o + t, expf(es_in<MODE>(p, lb, t, head, hl) - m) / s);
}

// ---------------------------------------------------------------- warp per row
template <int MODE>
__global__ void __launch_bounds__(256) es_warp_kernel(const EsParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= p.n_rows) return;  // whole warp
  const int64_t row = p.row_list ? __ldg(p.row_list + w) : w;
  const int lb = __ldg(p.rowptr + row), hb = __ldg(p.rowptr + row + 1);
  const int64_t n = (int64_t)(hb - lb) * p.H;
  if (n == 0 || n <= p.n_lo || n > p.n_hi) return;
  const int H = p.H;
  es_row<MODE, 32>(p, row, lb, n, lane, [H](float v, bool is_max) { return is_max ? head_max(v, H) : head_sum(v, H); });
}

// ---------------------------------------------------------------- block (CLUSTER == 1) or cluster per row
template <int MODE, int THREADS, int CLUSTER>
__global__ void __launch_bounds__(THREADS) es_block_kernel(const EsParams p) {
  constexpr int WARPS = THREADS / 32;
  __shared__ float red[WARPS * 32];   // per-warp per-lane partials of this CTA
  __shared__ float cta[32];           // this CTA's per-head result, read by the cluster peers
  const int row_idx = blockIdx.x / CLUSTER;
  const int crank = blockIdx.x % CLUSTER;   // == cluster rank for a 1-D cluster
  const int64_t row = p.row_list ? __ldg(p.row_list + row_idx) : row_idx;
  const int lb = __ldg(p.rowptr + row), hb = __ldg(p.rowptr + row + 1);
  const int64_t n = (int64_t)(hb - lb) * p.H;
  if (n <= p.n_lo || n > p.n_hi) return;    // uniform over the whole cluster (same row)
  const int H = p.H;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  auto reduce = [&](float v, bool is_max) {
    v = is_max ? head_max(v, H) : head_sum(v, H);
    __syncthreads();
    red[wid * 32 + lane] = v;
    __syncthreads();
    float r = red[lane];
#pragma unroll 8
    for (int q = 1; q < WARPS; ++q) r = is_max ? fmaxf(r, red[q * 32 + lane]) : r + red[q * 32 + lane];
    if (CLUSTER > 1) {
      cg::cluster_group cluster = cg::this_cluster();
      if (wid == 0) cta[lane] = r;
      cluster.sync();                         // every CTA's per-head partial is published
      float t = 0.f;
      for (int q = 0; q < CLUSTER; ++q) {     // read the peers' shared memory (DSMEM)
        const float *peer = cluster.map_shared_rank(cta, q);
        const float x = peer[lane];
        t = (q == 0) ? x : (is_max ? fmaxf(t, x) : t + x);
      }
      cluster.sync();                         // nobody overwrites `cta` while a peer still reads it
      r = t;
    }
    return r;  // every thread: result for its own head (lane % H)
  };
  es_row<MODE, THREADS * CLUSTER>(p, row, lb, n, crank * THREADS + threadIdx.x, reduce);
}

// ---------------------------------------------------------------- warp per plan segment (many short rows)
// A segment's rows are consecutive, so their [deg,H] blocks form ONE contiguous run of floats.  The
// warp stages a sub-run of whole rows (<= 32 rows, <= SEG_CAP floats) in shared memory with
// coalesced loads, every lane then owns whole (row, head) pairs -- max, sum of exp and normalise are
// private loops over the pair's column of the staged tile -- and the tile is written back coalesced.
constexpr int SEG_CAP = 2048;     // floats staged per warp (>= chunk_edges * H, checked on the host)
template <int MODE> struct SegCfg { static constexpr int WARPS = (MODE == 1) ? 2 : 4; };  // <= 48 KB static smem

template <int MODE>
__global__ void __launch_bounds__(SegCfg<MODE>::WARPS * 32) es_seg_kernel(const EsParams p) {
  constexpr int SEG_WARPS = SegCfg<MODE>::WARPS;
  __shared__ __align__(16) float tile[SEG_WARPS][SEG_CAP];
  __shared__ __align__(16) float tile2[MODE == 1 ? SEG_WARPS : 1][MODE == 1 ? SEG_CAP : 4];
  __shared__ int rps[SEG_WARPS][33];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t seg = (int64_t)blockIdx.x * SEG_WARPS + wib;
  if (seg >= p.n_segs) return;
  const int2 rr = __ldg(p.segs + seg);
  float *T = tile[wib];
  float *T2 = tile2[MODE == 1 ? wib : 0];
  int *RP = rps[wib];
  const int H = p.H;
  int rw = rr.x;
  while (rw < rr.y) {
    // window of <= 32 rows starting at rw; keep the longest prefix that fits the tile
    const int row = rw + lane;
    const int base = __ldg(p.rowptr + rw);
    int endl = 0x7fffffff;
    if (row < rr.y) endl = __ldg(p.rowptr + row + 1);
    const unsigned fit = __ballot_sync(FULL, row < rr.y && (int64_t)(endl - base) * H <= SEG_CAP);
    const int m = __popc(fit);                 // fit is a prefix mask (row ends are monotone)
    if (m == 0) {                              // cannot happen when chunk_edges * H <= SEG_CAP (host check)
      rw += 1;
      continue;
    }
    if (lane < m) RP[lane + 1] = endl - base;
    if (lane == 0) RP[0] = 0;
    const int e_end = __shfl_sync(FULL, endl, m - 1);
    const int n = (e_end - base) * H;
    __syncwarp();
    // ---- stage
    if (MODE == 2) {
#pragma unroll 4
      for (int t = lane; t < n; t += 32) {
        const int pe = base + (t >> p.lgH), h = t & (H - 1);
        const int r = __ldg(p.edge_row + pe);
        const int c = __ldg(p.colind + pe);
        const float z = __ldg(p.a + (int64_t)r * H + h) + __ldg(p.b + (int64_t)c * H + h);
        T[t] = z > 0.f ? z : z * p.slope;
      }
    } else if ((H & 3) == 0) {   // base*H and n are multiples of 4: 16-byte staging, 4 loads in flight
      const float4 *src = reinterpret_cast<const float4 *>(p.a + (int64_t)base * H);
      float4 *T4 = reinterpret_cast<float4 *>(T);
#pragma unroll 4
      for (int t = lane; t < (n >> 2); t += 32) T4[t] = __ldcs(src + t);
      if (MODE == 1) {
        const float4 *src2 = reinterpret_cast<const float4 *>(p.b + (int64_t)base * H);
        float4 *T24 = reinterpret_cast<float4 *>(T2);
#pragma unroll 4
        for (int t = lane; t < (n >> 2); t += 32) T24[t] = __ldcs(src2 + t);
      }
    } else {
      const float *src = p.a + (int64_t)base * H;
#pragma unroll 4
      for (int t = lane; t < n; t += 32) T[t] = ld_stream(src + t);
      if (MODE == 1) {
        const float *src2 = p.b + (int64_t)base * H;
#pragma unroll 4
        for (int t = lane; t < n; t += 32) T2[t] = ld_stream(src2 + t);
      }
    }
    __syncwarp();
    // ---- (row, head) pairs
    for (int q = lane; q < m * H; q += 32) {
      const int rl = q >> p.lgH, h = q & (H - 1);
      const int k0 = RP[rl], k1 = RP[rl + 1];
      if (MODE == 1) {
        float s = 0.f;
        for (int k = k0; k < k1; ++k) s = fmaf(T[k * H + h], T2[k * H + h], s);
        for (int k = k0; k < k1; ++k) T[k * H + h] = T[k * H + h] * (T2[k * H + h] - s);
      } else {
        float mx = -CUDART_INF_F;
        for (int k = k0; k < k1; ++k) mx = fmaxf(mx, T[k * H + h]);
        float s = 0.f;
        for (int k = k0; k < k1; ++k) {
          const float ex = expf(T[k * H + h] - mx);
          T[k * H + h] = ex;
          s += ex;
        }
        for (int k = k0; k < k1; ++k) T[k * H + h] = T[k * H + h] / s;
      }
    }
    __syncwarp();
    // ---- write back
    float *dst = p.out + (int64_t)base * H;
    if ((H & 3) == 0) {
      float4 *dst4 = reinterpret_cast<float4 *>(dst);
      const float4 *T4 = reinterpret_cast<const float4 *>(T);
      for (int t = lane; t < (n >> 2); t += 32) __stcs(dst4 + t, T4[t]);
    } else {
      for (int t = lane; t < n; t += 32) st_stream(dst + t, T[t]);
    }
    __syncwarp();
    rw += m;
  }
}

// ---------------------------------------------------------------- generic H: warp per row, loop over heads
template <int MODE>
__global__ void __launch_bounds__(256) es_generic_kernel(const EsParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= p.n_rows) return;
  const int lb = __ldg(p.rowptr + row), hb = __ldg(p.rowptr + row + 1);
  for (int h = 0; h < p.H; ++h) {
    if (MODE == 1) {
      float s = 0.f;
      for (int e = lb + lane; e < hb; e += 32)
        s = fmaf(__ldg(p.a + (int64_t)e * p.H + h), __ldg(p.b + (int64_t)e * p.H + h), s);
      s = head_sum(s, 1);
      for (int e = lb + lane; e < hb; e += 32) {
        const int64_t k = (int64_t)e * p.H + h;
        p.out[k] = __ldg(p.a + k) * (__ldg(p.b + k) - s);
      }
    } else {
      const float hl = (MODE == 2) ? __ldg(p.a + row * p.H + h) : 0.f;
      auto in = [&](int e) {
        if (MODE == 2) {
          const float z = hl + __ldg(p.b + (int64_t)__ldg(p.colind + e) * p.H + h);
          return z > 0.f ? z : z * p.slope;
        }
        return __ldg(p.a + (int64_t)e * p.H + h);
      };
      float m = -CUDART_INF_F;
      for (int e = lb + lane; e < hb; e += 32) m = fmaxf(m, in(e));
      m = head_max(m, 1);
      float s = 0.f;
      for (int e = lb + lane; e < hb; e += 32) s += expf(in(e) - m);
      s = head_sum(s, 1);
      for (int e = lb + lane; e < hb; e += 32) p.out[(int64_t)e * p.H + h] = expf(in(e) - m) / s;
    }
  }
}

template <int MODE, int THREADS, int CLUSTER>
static int launch_block_tier(const EsParams &q, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(q.n_rows * CLUSTER));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CB_CUDA(cudaLaunchKernelEx(&cfg, es_block_kernel<MODE, THREADS, CLUSTER>, q));
  count_launch();
  return COGDL_B200_OK;
}

// The hub tiers and the segment kernel touch disjoint rows: the hub tiers are forked onto a side
// stream (one per device, created on first use) so they overlap the segment kernel instead of
// queueing behind it -- four short kernels back to back were 120 us on the arxiv shape, the longest
// of them 50 us.  Fork / join are event record + wait, so the caller's stream order is preserved.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static int side_stream(SideStream **out) {
  static SideStream table[64];
  int dev = 0;
  CB_CUDA(cudaGetDevice(&dev));
  CB_REQUIRE(dev >= 0 && dev < 64, "edge_softmax: device index out of range");
  SideStream &ss = table[dev];
  if (!ss.stream) {
    CB_CUDA(cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking));
    CB_CUDA(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
    CB_CUDA(cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming));
  }
  *out = &ss;
  return COGDL_B200_OK;
}

// Launch plan (see the header comment).
template <int MODE>
static int es_launch(EsParams p, const cogdl_b200_hub_plan_t *plan, cudaStream_t s, const char *who) {
  const bool pow2 = p.H <= 32 && (p.H & (p.H - 1)) == 0;
  const int64_t row_blocks = ceil_div(p.n_rows * 32, 256);
  CB_REQUIRE(row_blocks <= 0x7fffffffLL, "%s: problem too large for one launch", who);
  p.lgH = 0;
  while ((1 << p.lgH) < p.H) ++p.lgH;
  p.row_list = nullptr; p.n_lo = -1; p.n_hi = INT64_MAX; p.n_segs = 0; p.segs = nullptr; p.edge_row = nullptr;
  if (!pow2) {
    es_generic_kernel<MODE><<<(unsigned)row_blocks, 256, 0, s>>>(p);
    CB_LAUNCH_CHECK();
    return COGDL_B200_OK;
  }
  const bool segs = plan && plan->chunk_edges > 0 && plan->segs && plan->edge_row && plan->n_segs > 0 &&
                    (int64_t)plan->chunk_edges * p.H <= SEG_CAP;
  if (!segs) {   // no plan: every row through the warp kernel (hubs serialise on one warp)
    es_warp_kernel<MODE><<<(unsigned)row_blocks, 256, 0, s>>>(p);
    CB_LAUNCH_CHECK();
    return COGDL_B200_OK;
  }
  cudaStream_t main_stream = s;
  SideStream *ss = nullptr;
  if (plan->n_hub_rows > 0) {   // hub tiers on the side stream, concurrently with the segment kernel
    int rcs = side_stream(&ss);
    if (rcs) return rcs;
    CB_CUDA(cudaEventRecord(ss->fork, main_stream));
    CB_CUDA(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
    s = ss->stream;
    EsParams q = p;
    const int nh = plan->n_hub_rows;
    if (plan->hub_degrees_host) {
      // hub_rows is sorted by descending degree: each tier is a contiguous slice of the list
      const int32_t *deg = plan->hub_degrees_host;
      int n_cluster = 0, n_block = 0;
      while (n_cluster < nh && (int64_t)deg[n_cluster] * p.H > BLOCK_ROW_ELEMS) ++n_cluster;
      n_block = n_cluster;
      while (n_block < nh && (int64_t)deg[n_block] * p.H > WARP_ROW_ELEMS) ++n_block;
      q.n_lo = -1; q.n_hi = INT64_MAX;
      if (n_cluster > 0) {
        q.row_list = plan->hub_rows; q.n_rows = n_cluster;
        int rc = launch_block_tier<MODE, 1024, CLUSTER_CTAS>(q, s);
        if (rc) return rc;
      }
      if (n_block > n_cluster) {
        q.row_list = plan->hub_rows + n_cluster; q.n_rows = n_block - n_cluster;
        int rc = launch_block_tier<MODE, 256, 1>(q, s);
        if (rc) return rc;
      }
      if (nh > n_block) {
        q.row_list = plan->hub_rows + n_block; q.n_rows = nh - n_block;
        es_warp_kernel<MODE><<<(unsigned)ceil_div((int64_t)q.n_rows * 32, 256), 256, 0, s>>>(q);
        CB_LAUNCH_CHECK();
      }
    } else {   // unsorted list: one warp per hub row, whatever its length
      q.row_list = plan->hub_rows; q.n_rows = nh; q.n_lo = -1; q.n_hi = INT64_MAX;
      es_warp_kernel<MODE><<<(unsigned)ceil_div((int64_t)q.n_rows * 32, 256), 256, 0, s>>>(q);
      CB_LAUNCH_CHECK();
    }
    CB_CUDA(cudaEventRecord(ss->join, ss->stream));
    s = main_stream;
  }
  p.n_segs = plan->n_segs; p.segs = reinterpret_cast<const int2 *>(plan->segs); p.edge_row = plan->edge_row;
  constexpr int SW = SegCfg<MODE>::WARPS;
  es_seg_kernel<MODE><<<(unsigned)ceil_div(p.n_segs, SW), SW * 32, 0, s>>>(p);
  CB_LAUNCH_CHECK();
  if (ss) CB_CUDA(cudaStreamWaitEvent(main_stream, ss->join, 0));
  return COGDL_B200_OK;
}

template <int MODE>
static int es_entry(const int32_t *rowptr, const float *a, const float *b, float *out, int64_t n_rows,
                    int64_t H, const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream,
                    const char *who) {
  CB_REQUIRE(n_rows >= 0 && H >= 0, "%s: negative size", who);
  if (n_rows == 0 || H == 0) return COGDL_B200_OK;
  CB_REQUIRE(rowptr && a && out && (MODE != 1 || b), "%s: null pointer", who);
  CB_REQUIRE(n_rows < 0x7fffffffLL && H < 0x7fffffffLL, "%s: sizes must fit int32", who);
  int rc = check_plan(plan, 0);
  if (rc) return rc;
  EsParams p;
  p.rowptr = rowptr; p.a = a; p.b = b; p.out = out; p.n_rows = n_rows; p.H = (int)H;
  p.colind = nullptr; p.slope = 0.f;
  return es_launch<MODE>(p, plan, (cudaStream_t)stream, who);
}

// GAT attention (used by gat_fused.cu): att = softmax_row(leakyrelu(h_l[row] + h_r[col])), any H.
int gat_attention(const int32_t *rowptr, const int32_t *colind, const float *h_l, const float *h_r, float slope,
                  float *att, int64_t n_rows, int64_t H, const cogdl_b200_hub_plan_t *plan, cudaStream_t s) {
  EsParams p;
  p.rowptr = rowptr; p.a = h_l; p.b = h_r; p.out = att; p.n_rows = n_rows; p.H = (int)H;
  p.colind = colind; p.slope = slope;
  return es_launch<2>(p, plan, s, "cogdl_b200_gat_fwd_f32");
}

}  // namespace cogdl_b200

using namespace cogdl_b200;

extern "C" int cogdl_b200_edge_softmax_fwd_f32(const int32_t *rowptr, const float *in, float *out,
                                               int64_t n_rows, int64_t H,
                                               const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream) {
  return es_entry<0>(rowptr, in, nullptr, out, n_rows, H, plan, stream, "cogdl_b200_edge_softmax_fwd_f32");
}

extern "C" int cogdl_b200_edge_softmax_bwd_f32(const int32_t *rowptr, const float *y, const float *g,
                                               float *gin, int64_t n_rows, int64_t H,
                                               const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream) {
  return es_entry<1>(rowptr, y, g, gin, n_rows, H, plan, stream, "cogdl_b200_edge_softmax_bwd_f32");
}
