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
  int hub_T;              // rows with degree > hub_T are left to the hub kernel (0: none)
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

// ---------------------------------------------------------------- warp per row, H = 2^k <= 32
template <bool BWD>
__global__ void __launch_bounds__(256) es_warp_kernel(const EsParams p) {
  constexpr int K = 4;
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= p.n_rows) return;  // whole warp
  const int lb = __ldg(p.rowptr + row), hb = __ldg(p.rowptr + row + 1);
  const int deg = hb - lb;
  if (deg == 0 || (p.hub_T > 0 && deg > p.hub_T)) return;
  const int64_t n = (int64_t)deg * p.H;
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
  // long row: 3 passes (2 for bwd); re-reads hit L1/L2
  if (!BWD) {
    float m = -CUDART_INF_F;
    for (int64_t t = lane; t < n; t += 32) m = fmaxf(m, __ldg(a + t));
    m = head_max(m, p.H);
    float s = 0.f;
    for (int64_t t = lane; t < n; t += 32) s += expf(__ldg(a + t) - m);
    s = head_sum(s, p.H);
    for (int64_t t = lane; t < n; t += 32) st_stream(o + t, expf(__ldg(a + t) - m) / s);
  } else {
    float s = 0.f;
    for (int64_t t = lane; t < n; t += 32) s = fmaf(__ldg(a + t), __ldg(b + t), s);
    s = head_sum(s, p.H);
    for (int64_t t = lane; t < n; t += 32) st_stream(o + t, __ldg(a + t) * (__ldg(b + t) - s));
  }
}

// ---------------------------------------------------------------- block per hub row, H = 2^k <= 32
template <bool MAX>
__device__ __forceinline__ float block_head_reduce(float v, int H, float *smem /*[8][32]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = MAX ? head_max(v, H) : head_sum(v, H);
  __syncthreads();  // smem reuse between successive reductions
  smem[wid * 32 + lane] = v;
  __syncthreads();
  float r = smem[lane];
#pragma unroll
  for (int q = 1; q < 8; ++q) r = MAX ? fmaxf(r, smem[q * 32 + lane]) : r + smem[q * 32 + lane];
  return r;  // every thread: result for its own head (lane % H)
}

template <bool BWD>
__global__ void __launch_bounds__(256) es_hub_kernel(const EsParams p) {
  __shared__ float smem[8 * 32];
  const int row = __ldg(p.hub_rows + blockIdx.x);
  const int lb = __ldg(p.rowptr + row), hb = __ldg(p.rowptr + row + 1);
  const int64_t n = (int64_t)(hb - lb) * p.H;
  const float *a = p.a + (int64_t)lb * p.H;
  const float *b = BWD ? p.b + (int64_t)lb * p.H : nullptr;
  float *o = p.out + (int64_t)lb * p.H;
  const int tid = threadIdx.x;  // 256 % H == 0 => head = tid % H is loop-invariant
  if (!BWD) {
    float m = -CUDART_INF_F;
    for (int64_t t = tid; t < n; t += 256) m = fmaxf(m, __ldg(a + t));
    m = block_head_reduce<true>(m, p.H, smem);
    float s = 0.f;
    for (int64_t t = tid; t < n; t += 256) s += expf(__ldg(a + t) - m);
    s = block_head_reduce<false>(s, p.H, smem);
    for (int64_t t = tid; t < n; t += 256) st_stream(o + t, expf(__ldg(a + t) - m) / s);
  } else {
    float s = 0.f;
    for (int64_t t = tid; t < n; t += 256) s = fmaf(__ldg(a + t), __ldg(b + t), s);
    s = block_head_reduce<false>(s, p.H, smem);
    for (int64_t t = tid; t < n; t += 256) st_stream(o + t, __ldg(a + t) * (__ldg(b + t) - s));
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
    es_hub_kernel<BWD><<<(unsigned)p.n_hub_rows, 256, 0, s>>>(p);  // long rows first
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
