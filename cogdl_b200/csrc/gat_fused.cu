// gat_fused.cu -- GAT forward without the [nnz,H] index/elementwise round trips (sm_100a; "next" row,
// SURVEY 8f-1).
//
//   e[p,h]   = leakyrelu(h_l[i,h] + h_r[colind[p],h])         p in row i
//   a[p,:]   = softmax over the row's edges, per head
//   out[i,h] = sum_p a[p,h] * feat[colind[p],h,:]
//
// Replaces the unfused chain of cogdl/layers/gat_layer.py:73-77 (h_l[row] + h_r[col] gathers ->
// LeakyReLU -> edge_softmax -> mh_spmm: two [nnz,H] gathers, an add, an activation and a softmax,
// each a kernel with its own [nnz,H] temporaries) and the stale dgNN binding
// (cogdl/operators/fused_gat.py:17-19; precedent third_party/dgNN/.../fused_gatconv_kernel.cu:26-133).
//
// Two launches:
//   1. attention kernel: logits are formed in registers from h_l / h_r and soft-maxed per row in
//      the same pass; only the normalised attention [nnz,H] is written (it is also what the
//      backward needs).  Same row tiers as edge_softmax.cu (registers / 3-pass warp / block).
//   2. the row-stream multi-head SpMM (stream.cuh) consumes it.
// A single-kernel fusion would save the attention write+read: 8*H bytes per edge against 4*H*F + ...
// bytes of feature gather (1.6 % at H=8, F=128) -- not worth a second copy of the stream kernel.
// (The first version of this file was a per-(row, slice) fused kernel that recomputed the softmax
// statistics in every slice: 4.6 ms vs 1.1 ms for the two-launch form on the arxiv shape.)
#include "common.cuh"

#include <math_constants.h>

namespace cogdl_b200 {

int mhspmm_run(const int32_t *rowptr, const int32_t *colind, const int32_t *perm, const float *att,
               const float *feat, float *out, int64_t n_rows, int64_t H, int64_t F,
               const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream, bool rows_too);

struct GatParams {
  const int *rowptr;
  const int *colind;
  const float *h_l;
  const float *h_r;
  float slope;
  float *att;
  int64_t n_rows;
  int H;
  int lgH;          // log2(H) when H is a power of two
  const int *hub_rows;
  int n_hub_rows;
  int hub_T;
};

constexpr int64_t GAT_BLOCK_ROW_ELEMS = 4096;
constexpr int GAT_HUB_THREADS = 1024;

__device__ __forceinline__ float lrelu(float z, float slope) { return z > 0.f ? z : z * slope; }

__device__ __forceinline__ float hmax(float v, int H) {
  for (int s = 16; s >= H; s >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, s));
  return v;
}
__device__ __forceinline__ float hsum(float v, int H) {
  for (int s = 16; s >= H; s >>= 1) v += __shfl_xor_sync(FULL, v, s);
  return v;
}

// logit of element t of the row's [deg, H] block (head = t % H is the caller's `head`)
__device__ __forceinline__ float gat_logit(const GatParams &p, int lb, int64_t t, int head, float hl) {
  const int c = __ldg(p.colind + lb + (int)(t >> p.lgH));
  return lrelu(hl + __ldg(p.h_r + (int64_t)c * p.H + head), p.slope);
}

// warp per row, H = 2^k <= 32: lane t owns elements t, t+32, ... of the row's contiguous [deg,H] block
__global__ void __launch_bounds__(256) gat_att_warp_kernel(const GatParams p) {
  constexpr int K = 4;
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= p.n_rows) return;
  const int lb = __ldg(p.rowptr + row), hb = __ldg(p.rowptr + row + 1);
  const int deg = hb - lb;
  const int64_t n = (int64_t)deg * p.H;
  if (deg == 0 || (p.hub_T > 0 && deg > p.hub_T && n > GAT_BLOCK_ROW_ELEMS)) return;
  const int head = lane & (p.H - 1);
  const float hl = __ldg(p.h_l + row * p.H + head);
  float *o = p.att + (int64_t)lb * p.H;
  if (n <= 32 * K) {
    float z[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int t = lane + 32 * k;
      z[k] = (t < n) ? gat_logit(p, lb, t, head, hl) : -CUDART_INF_F;
    }
    float m = z[0];
#pragma unroll
    for (int k = 1; k < K; ++k) m = fmaxf(m, z[k]);
    m = hmax(m, p.H);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      z[k] = (lane + 32 * k < n) ? expf(z[k] - m) : 0.f;
      s += z[k];
    }
    s = hsum(s, p.H);
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (lane + 32 * k < n) st_cg(o + lane + 32 * k, z[k] / s);
    return;
  }
  float m = -CUDART_INF_F;
#pragma unroll 4
  for (int64_t t = lane; t < n; t += 32) m = fmaxf(m, gat_logit(p, lb, t, head, hl));
  m = hmax(m, p.H);
  float s = 0.f;
#pragma unroll 4
  for (int64_t t = lane; t < n; t += 32) s += expf(gat_logit(p, lb, t, head, hl) - m);
  s = hsum(s, p.H);
#pragma unroll 4
  for (int64_t t = lane; t < n; t += 32) st_cg(o + t, expf(gat_logit(p, lb, t, head, hl) - m) / s);
}

// block per listed hub row with more than GAT_BLOCK_ROW_ELEMS elements
__global__ void __launch_bounds__(GAT_HUB_THREADS) gat_att_hub_kernel(const GatParams p) {
  __shared__ float smem[GAT_HUB_THREADS];
  const int row = __ldg(p.hub_rows + blockIdx.x);
  const int lb = __ldg(p.rowptr + row), hb = __ldg(p.rowptr + row + 1);
  const int64_t n = (int64_t)(hb - lb) * p.H;
  if (n <= GAT_BLOCK_ROW_ELEMS) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int head = tid & (p.H - 1);
  const float hl = __ldg(p.h_l + (int64_t)row * p.H + head);
  float *o = p.att + (int64_t)lb * p.H;
  float m = -CUDART_INF_F;
#pragma unroll 4
  for (int64_t t = tid; t < n; t += GAT_HUB_THREADS) m = fmaxf(m, gat_logit(p, lb, t, head, hl));
  m = hmax(m, p.H);
  smem[wid * 32 + lane] = m;
  __syncthreads();
  m = smem[lane];
  for (int q = 1; q < GAT_HUB_THREADS / 32; ++q) m = fmaxf(m, smem[q * 32 + lane]);
  float s = 0.f;
#pragma unroll 4
  for (int64_t t = tid; t < n; t += GAT_HUB_THREADS) s += expf(gat_logit(p, lb, t, head, hl) - m);
  s = hsum(s, p.H);
  __syncthreads();
  smem[wid * 32 + lane] = s;
  __syncthreads();
  s = smem[lane];
  for (int q = 1; q < GAT_HUB_THREADS / 32; ++q) s += smem[q * 32 + lane];
#pragma unroll 4
  for (int64_t t = tid; t < n; t += GAT_HUB_THREADS) st_cg(o + t, expf(gat_logit(p, lb, t, head, hl) - m) / s);
}

// any H: warp per row, heads one after another, lanes over edges
__global__ void __launch_bounds__(256) gat_att_generic_kernel(const GatParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= p.n_rows) return;
  const int lb = __ldg(p.rowptr + row), hb = __ldg(p.rowptr + row + 1);
  for (int h = 0; h < p.H; ++h) {
    const float hl = __ldg(p.h_l + row * p.H + h);
    float m = -CUDART_INF_F;
    for (int e = lb + lane; e < hb; e += 32)
      m = fmaxf(m, lrelu(hl + __ldg(p.h_r + (int64_t)__ldg(p.colind + e) * p.H + h), p.slope));
    m = hmax(m, 1);
    float s = 0.f;
    for (int e = lb + lane; e < hb; e += 32)
      s += expf(lrelu(hl + __ldg(p.h_r + (int64_t)__ldg(p.colind + e) * p.H + h), p.slope) - m);
    s = hsum(s, 1);
    for (int e = lb + lane; e < hb; e += 32)
      p.att[(int64_t)e * p.H + h] = expf(lrelu(hl + __ldg(p.h_r + (int64_t)__ldg(p.colind + e) * p.H + h), p.slope) - m) / s;
  }
}

}  // namespace cogdl_b200

using namespace cogdl_b200;

extern "C" int cogdl_b200_gat_fwd_f32(const int32_t *rowptr, const int32_t *colind, const float *h_l,
                                      const float *h_r, const float *feat, float negative_slope, float *out,
                                      float *att_out, int64_t n_rows, int64_t H, int64_t F,
                                      const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream) {
  CB_REQUIRE(n_rows >= 0 && H >= 0 && F >= 0, "cogdl_b200_gat_fwd_f32: negative size");
  if (n_rows == 0 || H == 0 || F == 0) return COGDL_B200_OK;
  CB_REQUIRE(rowptr && colind && h_l && h_r && feat && out, "cogdl_b200_gat_fwd_f32: null pointer");
  CB_REQUIRE(att_out, "cogdl_b200_gat_fwd_f32: att_out ([nnz,H] attention, also scratch) is required");
  CB_REQUIRE(n_rows < 0x7fffffffLL && H * F < 0x7fffffffLL, "cogdl_b200_gat_fwd_f32: sizes must fit int32");
  int rc = check_plan(plan, (plan ? (int64_t)plan->n_chunks : 0) * H * F * (int64_t)sizeof(float));
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  GatParams p;
  p.rowptr = rowptr; p.colind = colind; p.h_l = h_l; p.h_r = h_r; p.slope = negative_slope; p.att = att_out;
  p.n_rows = n_rows; p.H = (int)H; p.lgH = 0; p.hub_rows = nullptr; p.n_hub_rows = 0; p.hub_T = 0;
  const int64_t blocks = ceil_div(n_rows * 32, 256);
  CB_REQUIRE(blocks <= 0x7fffffffLL, "cogdl_b200_gat_fwd_f32: problem too large for one launch");
  const bool pow2 = H <= 32 && (H & (H - 1)) == 0;
  if (!pow2) {
    gat_att_generic_kernel<<<(unsigned)blocks, 256, 0, s>>>(p);
    CB_LAUNCH_CHECK();
  } else {
    while ((1 << p.lgH) < (int)H) ++p.lgH;
    if (plan && plan->chunk_edges > 0 && plan->n_hub_rows > 0) {
      p.hub_T = plan->chunk_edges; p.hub_rows = plan->hub_rows; p.n_hub_rows = plan->n_hub_rows;
      gat_att_hub_kernel<<<(unsigned)p.n_hub_rows, GAT_HUB_THREADS, 0, s>>>(p);
      CB_LAUNCH_CHECK();
    }
    gat_att_warp_kernel<<<(unsigned)blocks, 256, 0, s>>>(p);
    CB_LAUNCH_CHECK();
  }
  return mhspmm_run(rowptr, colind, nullptr, att_out, feat, out, n_rows, H, F, plan, stream, true);
}
