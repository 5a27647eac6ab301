// edge_softmax.cu -- per-destination-row softmax over edge logits [nnz, H], fwd + bwd (sm_100a).
//
// Replaces edge_softmax / edge_softmax_backward (cogdl/operators/edge_softmax/edge_softmax.cu:
// 7-98): one block (32 x H threads) per row, three passes over the row's logits with lanes
// striding H floats apart (<= 1/H sector efficiency) and exp() evaluated twice.
//
// Key layout fact: in CSR order the logits of row i are ONE contiguous block of deg*H floats
// starting at in + rowptr[i]*H.  So a warp reads its row with fully coalesced 128-byte loads,
// lane t owning elements t, t+32, ...; because 32 % H == 0 (H a power of two <= 32) the lane's
// head is constant (t % H) and the per-head reductions are xor-shuffles over strides 16 .. H.
// Rows with deg*H <= 128 live entirely in registers (one read, one exp, one write: the
// streaming minimum of 8*H bytes per edge); longer rows take the 3-pass form through L1/L2;
// hub rows (degree > plan chunk) get a whole 256-thread block each.
// Other head counts (H not a power of two, or > 32) use a generic strided kernel.
#include "common.cuh"

#include <math_constants.h>

namespace cogdl_b200 {

struct EsParams {
  const int *rowptr;
  const float *a;   // fwd: logits        bwd: y (softmax output)
  const float *b;   // fwd: unused        bwd: g (upstream gradient)
  float *out;
  int64_t n_rows;
  int H;
  int hub_T;              // plan chunk size: rows above it are listed in hub_rows (0: no plan)
  const int *hub_rows;
  int n_hub_rows;
};

// reduce across lanes that share lane % H  (strides 16 .. H)
__device__ __forceinline__ float head_max(float v, int H) {
  for (int s = 16; s >= H; s >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, s));
  return v;
}
__device__ __forceinline__ float head_sum(float v, int H) {
  for (int s = 16; s >= H; s >>= 1) v += __shfl_xor_sync(FULL, v, s);
  return v;
}

// Rows with more elements (deg * H) than this go to the block-per-row kernel when a plan lists them.
constexpr int64_t BLOCK_ROW_ELEMS = 4096;

// ---------------------------------------------------------------- warp per row, H = 2^k <= 32
template <bool BWD>
__global__ void __launch_bounds__(256) es_warp_kernel(const EsParams p) {
  constexpr int K = 4;
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= p.n_rows) return;  // whole warp
  const int lb = __ldg(p.rowptr + row), hb = __ldg(p.rowptr + row + 1);
  const int deg = hb - lb;
  const int64_t n = (int64_t)deg * p.H;
  // rows with more than BLOCK_ROW_ELEMS elements are listed hub rows: a whole block takes them
  if (deg == 0 || (p.hub_T > 0 && deg > p.hub_T && n > BLOCK_ROW_ELEMS)) return;
  const float *a = p.a + (int64_t)lb * p.H;
  const float *b = BWD ? p.b + (int64_t)lb * p.H : nullptr;
  float *o = p.out + (int64_t)lb * p.H;

  if (n <= 32 * K) {
    float va[K], vb[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int t = lane + 32 * k;
      va[k] = (t < n) ? ld_stream(a + t) : (BWD ? 0.f : -CUDART_INF_F);
      if (BWD) vb[k] = (t < n) ? ld_stream(b + t) : 0.f;
    }
    if (!BWD) {
      float m = va[0];
#pragma unroll
      for (int k = 1; k < K; ++k) m = fmaxf(m, va[k]);
      m = head_max(m, p.H);
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        va[k] = (lane + 32 * k < n) ? expf(va[k] - m) : 0.f;
        s += va[k];
      }
      s = head_sum(s, p.H);
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (lane + 32 * k < n) st_stream(o + lane + 32 * k, va[k] / s);
    } else {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) s = fmaf(va[k], vb[k], s);
      s = head_sum(s, p.H);
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (lane + 32 * k < n) st_stream(o + lane + 32 * k, va[k] * (vb[k] - s));
    }
    return;
  }
  // long row: 3 passes (2 for bwd), 4 independent loads in flight per lane; re-reads hit L1/L2
  if (!BWD) {
    float m0 = -CUDART_INF_F, m1 = m0, m2 = m0, m3 = m0;
    int64_t t = lane;
    for (; t + 96 < n; t += 128) {
      const float x0 = __ldg(a + t), x1 = __ldg(a + t + 32), x2 = __ldg(a + t + 64), x3 = __ldg(a + t + 96);
      m0 = fmaxf(m0, x0); m1 = fmaxf(m1, x1); m2 = fmaxf(m2, x2); m3 = fmaxf(m3, x3);
    }
    for (; t < n; t += 32) m0 = fmaxf(m0, __ldg(a + t));
    const float m = head_max(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)), p.H);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    t = lane;
    for (; t + 96 < n; t += 128) {
      const float x0 = __ldg(a + t), x1 = __ldg(a + t + 32), x2 = __ldg(a + t + 64), x3 = __ldg(a + t + 96);
      s0 += expf(x0 - m); s1 += expf(x1 - m); s2 += expf(x2 - m); s3 += expf(x3 - m);
    }
    for (; t < n; t += 32) s0 += expf(__ldg(a + t) - m);
    const float s = head_sum((s0 + s1) + (s2 + s3), p.H);
    t = lane;
    for (; t + 96 < n; t += 128) {
      const float x0 = __ldg(a + t), x1 = __ldg(a + t + 32), x2 = __ldg(a + t + 64), x3 = __ldg(a + t + 96);
      st_stream(o + t, expf(x0 - m) / s); st_stream(o + t + 32, expf(x1 - m) / s);
      st_stream(o + t + 64, expf(x2 - m) / s); st_stream(o + t + 96, expf(x3 - m) / s);
    }
    for (; t < n; t += 32) st_stream(o + t, expf(__ldg(a + t) - m) / s);
  } else {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int64_t t = lane;
    for (; t + 96 < n; t += 128) {
      s0 = fmaf(__ldg(a + t), __ldg(b + t), s0); s1 = fmaf(__ldg(a + t + 32), __ldg(b + t + 32), s1);
      s2 = fmaf(__ldg(a + t + 64), __ldg(b + t + 64), s2); s3 = fmaf(__ldg(a + t + 96), __ldg(b + t + 96), s3);
    }
    for (; t < n; t += 32) s0 = fmaf(__ldg(a + t), __ldg(b + t), s0);
    const float s = head_sum((s0 + s1) + (s2 + s3), p.H);
    t = lane;
    for (; t + 96 < n; t += 128) {
#pragma unroll
      for (int q = 0; q < 4; ++q) st_stream(o + t + 32 * q, __ldg(a + t + 32 * q) * (__ldg(b + t + 32 * q) - s));
    }
    for (; t < n; t += 32) st_stream(o + t, __ldg(a + t) * (__ldg(b + t) - s));
  }
}

// ---------------------------------------------------------------- block per hub row, H = 2^k <= 32
// 1024 threads per hub row; every pass keeps 4 independent loads in flight per thread (the hub of an
// arxiv-shaped graph is 22 K edges x 8 heads = 724 KB: latency-bound unless the loads are batched).
constexpr int HUB_THREADS = 1024;
constexpr int HUB_WARPS = HUB_THREADS / 32;

template <bool MAX>
__device__ __forceinline__ float block_head_reduce(float v, int H, float *smem /*[HUB_WARPS][32]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = MAX ? head_max(v, H) : head_sum(v, H);
  __syncthreads();  // smem reuse between successive reductions
  smem[wid * 32 + lane] = v;
  __syncthreads();
  float r = smem[lane];
#pragma unroll 8
  for (int q = 1; q < HUB_WARPS; ++q) r = MAX ? fmaxf(r, smem[q * 32 + lane]) : r + smem[q * 32 + lane];
  return r;  // every thread: result for its own head (lane % H)
}

template <bool BWD>
__global__ void __launch_bounds__(HUB_THREADS) es_hub_kernel(const EsParams p) {
  __shared__ float smem[HUB_WARPS * 32];
  const int row = __ldg(p.hub_rows + blockIdx.x);
  const int lb = __ldg(p.rowptr + row), hb = __ldg(p.rowptr + row + 1);
  const int64_t n = (int64_t)(hb - lb) * p.H;
  if (n <= BLOCK_ROW_ELEMS) return;   // small listed rows stay with the warp kernel (block-uniform)
  const float *a = p.a + (int64_t)lb * p.H;
  const float *b = BWD ? p.b + (int64_t)lb * p.H : nullptr;
  float *o = p.out + (int64_t)lb * p.H;
  const int tid = threadIdx.x;  // HUB_THREADS % H == 0 => head = tid % H is loop-invariant
  constexpr int ST = HUB_THREADS;
  if (!BWD) {
    float m0 = -CUDART_INF_F, m1 = m0, m2 = m0, m3 = m0;
    int64_t t = tid;
    for (; t + 3 * ST < n; t += 4 * ST) {
      const float x0 = __ldg(a + t), x1 = __ldg(a + t + ST), x2 = __ldg(a + t + 2 * ST), x3 = __ldg(a + t + 3 * ST);
      m0 = fmaxf(m0, x0); m1 = fmaxf(m1, x1); m2 = fmaxf(m2, x2); m3 = fmaxf(m3, x3);
    }
    for (; t < n; t += ST) m0 = fmaxf(m0, __ldg(a + t));
    const float m = block_head_reduce<true>(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)), p.H, smem);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    t = tid;
    for (; t + 3 * ST < n; t += 4 * ST) {
      const float x0 = __ldg(a + t), x1 = __ldg(a + t + ST), x2 = __ldg(a + t + 2 * ST), x3 = __ldg(a + t + 3 * ST);
      s0 += expf(x0 - m); s1 += expf(x1 - m); s2 += expf(x2 - m); s3 += expf(x3 - m);
    }
    for (; t < n; t += ST) s0 += expf(__ldg(a + t) - m);
    const float s = block_head_reduce<false>((s0 + s1) + (s2 + s3), p.H, smem);
    t = tid;
    for (; t + 3 * ST < n; t += 4 * ST) {
      const float x0 = __ldg(a + t), x1 = __ldg(a + t + ST), x2 = __ldg(a + t + 2 * ST), x3 = __ldg(a + t + 3 * ST);
      st_stream(o + t, expf(x0 - m) / s); st_stream(o + t + ST, expf(x1 - m) / s);
      st_stream(o + t + 2 * ST, expf(x2 - m) / s); st_stream(o + t + 3 * ST, expf(x3 - m) / s);
    }
    for (; t < n; t += ST) st_stream(o + t, expf(__ldg(a + t) - m) / s);
  } else {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int64_t t = tid;
    for (; t + 3 * ST < n; t += 4 * ST) {
      s0 = fmaf(__ldg(a + t), __ldg(b + t), s0); s1 = fmaf(__ldg(a + t + ST), __ldg(b + t + ST), s1);
      s2 = fmaf(__ldg(a + t + 2 * ST), __ldg(b + t + 2 * ST), s2); s3 = fmaf(__ldg(a + t + 3 * ST), __ldg(b + t + 3 * ST), s3);
    }
    for (; t < n; t += ST) s0 = fmaf(__ldg(a + t), __ldg(b + t), s0);
    const float s = block_head_reduce<false>((s0 + s1) + (s2 + s3), p.H, smem);
    t = tid;
    for (; t + 3 * ST < n; t += 4 * ST) {
#pragma unroll
      for (int q = 0; q < 4; ++q) st_stream(o + t + q * ST, __ldg(a + t + q * ST) * (__ldg(b + t + q * ST) - s));
    }
    for (; t < n; t += ST) st_stream(o + t, __ldg(a + t) * (__ldg(b + t) - s));
  }
}

// ---------------------------------------------------------------- generic H: warp per row, loop over heads
template <bool BWD>
__global__ void __launch_bounds__(256) es_generic_kernel(const EsParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= p.n_rows) return;
  const int lb = __ldg(p.rowptr + row), hb = __ldg(p.rowptr + row + 1);
  for (int h = 0; h < p.H; ++h) {
    if (!BWD) {
      float m = -CUDART_INF_F;
      for (int e = lb + lane; e < hb; e += 32) m = fmaxf(m, __ldg(p.a + (int64_t)e * p.H + h));
      m = head_max(m, 1);
      float s = 0.f;
      for (int e = lb + lane; e < hb; e += 32) s += expf(__ldg(p.a + (int64_t)e * p.H + h) - m);
      s = head_sum(s, 1);
      for (int e = lb + lane; e < hb; e += 32)
        p.out[(int64_t)e * p.H + h] = expf(__ldg(p.a + (int64_t)e * p.H + h) - m) / s;
    } else {
      float s = 0.f;
      for (int e = lb + lane; e < hb; e += 32)
        s = fmaf(__ldg(p.a + (int64_t)e * p.H + h), __ldg(p.b + (int64_t)e * p.H + h), s);
      s = head_sum(s, 1);
      for (int e = lb + lane; e < hb; e += 32) {
        const int64_t k = (int64_t)e * p.H + h;
        p.out[k] = __ldg(p.a + k) * (__ldg(p.b + k) - s);
      }
    }
  }
}

template <bool BWD>
static int es_entry(const int32_t *rowptr, const float *a, const float *b, float *out, int64_t n_rows,
                    int64_t H, const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream,
                    const char *who) {
  CB_REQUIRE(n_rows >= 0 && H >= 0, "%s: negative size", who);
  if (n_rows == 0 || H == 0) return COGDL_B200_OK;
  CB_REQUIRE(rowptr && a && out && (!BWD || b), "%s: null pointer", who);
  CB_REQUIRE(n_rows < 0x7fffffffLL && H < 0x7fffffffLL, "%s: sizes must fit int32", who);
  int rc = check_plan(plan, 0);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  EsParams p;
  p.rowptr = rowptr; p.a = a; p.b = b; p.out = out; p.n_rows = n_rows; p.H = (int)H;
  p.hub_T = 0; p.hub_rows = nullptr; p.n_hub_rows = 0;
  const int64_t blocks = ceil_div(n_rows * 32, 256);
  CB_REQUIRE(blocks <= 0x7fffffffLL, "%s: problem too large for one launch", who);
  const bool pow2 = H <= 32 && (H & (H - 1)) == 0;
  if (!pow2) {
    es_generic_kernel<BWD><<<(unsigned)blocks, 256, 0, s>>>(p);
    CB_LAUNCH_CHECK();
    return COGDL_B200_OK;
  }
  if (plan && plan->chunk_edges > 0 && plan->n_hub_rows > 0) {
    p.hub_T = plan->chunk_edges; p.hub_rows = plan->hub_rows; p.n_hub_rows = plan->n_hub_rows;
    es_hub_kernel<BWD><<<(unsigned)p.n_hub_rows, HUB_THREADS, 0, s>>>(p);  // long rows first
    CB_LAUNCH_CHECK();
  }
  es_warp_kernel<BWD><<<(unsigned)blocks, 256, 0, s>>>(p);
  CB_LAUNCH_CHECK();
  return COGDL_B200_OK;
}

}  // namespace cogdl_b200

using namespace cogdl_b200;

extern "C" int cogdl_b200_edge_softmax_fwd_f32(const int32_t *rowptr, const float *in, float *out,
                                               int64_t n_rows, int64_t H,
                                               const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream) {
  return es_entry<false>(rowptr, in, nullptr, out, n_rows, H, plan, stream, "cogdl_b200_edge_softmax_fwd_f32");
}

extern "C" int cogdl_b200_edge_softmax_bwd_f32(const int32_t *rowptr, const float *y, const float *g,
                                               float *gin, int64_t n_rows, int64_t H,
                                               const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream) {
  return es_entry<true>(rowptr, y, g, gin, n_rows, H, plan, stream, "cogdl_b200_edge_softmax_bwd_f32");
}
