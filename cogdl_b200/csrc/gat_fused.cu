// gat_fused.cu -- fused GAT forward for sm_100a ("next" row).
//
//   e[p,h]   = leakyrelu(h_l[i,h] + h_r[colind[p],h])         p in row i
//   a[p,:]   = softmax over the row's edges, per head
//   out[i,h] = sum_p a[p,h] * feat[colind[p],h,:]
//
// Replaces the unfused chain of cogdl/layers/gat_layer.py:73-77 (index h_l[row]+h_r[col] ->
// LeakyReLU -> edge_softmax -> mh_spmm: three [nnz,H] round trips plus two [nnz,H] gathers) and
// the stale dgNN binding (cogdl/operators/fused_gat.py:17-19; design precedent
// third_party/dgNN/.../fused_gatconv_kernel.cu:26-133, which is one block per (row, head)).
//
// Work decomposition is the multi-head SpMM's: a GROUP of lanes owns one (row, 512-byte slice of
// the [H,F] block).  Phase 1: the TEAM of lanes that share a head strides over the row's edges
// and builds the online-softmax statistics (max, sum of exp) -- 4 bytes of h_r per edge per
// head, never written to memory.  Phase 2: the feature gather of the SpMM with the attention
// weight recomputed in registers from the same h_r sector.  att_out (optional) receives the
// normalised attention for the backward pass.
// Hub rows (degree > plan chunk) are not fused: a block per hub row writes their attention into
// att_out and the multi-head SpMM runs on the plan's chunk items only (mhspmm_run, rows_too =
// false); the fused kernel skips them.
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
  const float *feat;
  float slope;
  float *out;
  float *att_out;  // nullable
  int64_t n_rows;
  int H;
  int HFV, FVL, S;
  int team;        // lanes sharing one head inside a group (power of two, 1 = no cooperation)
  int hub_T;
  const int *hub_rows;
};

__device__ __forceinline__ float lrelu(float z, float slope) { return z > 0.f ? z : z * slope; }

// merge two online-softmax states
__device__ __forceinline__ void os_merge(float &m, float &s, float m2, float s2) {
  const float M = fmaxf(m, m2);
  if (M == -CUDART_INF_F) { s = 0.f; m = M; return; }
  s = s * expf(m - M) + s2 * expf(m2 - M);
  m = M;
}

template <typename VecT> __device__ __forceinline__ VecT gat_zero();
template <> __device__ __forceinline__ float4 gat_zero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ float gat_zero<float>() { return 0.f; }
__device__ __forceinline__ void gat_fma(float &acc, float a, const float &x) { acc = fmaf(a, x, acc); }
__device__ __forceinline__ void gat_fma(float4 &acc, float a, const float4 &x) {
  acc.x = fmaf(a, x.x, acc.x); acc.y = fmaf(a, x.y, acc.y); acc.z = fmaf(a, x.z, acc.z); acc.w = fmaf(a, x.w, acc.w);
}

template <typename VecT, int GROUP>
__global__ void __launch_bounds__(256) gat_fused_kernel(const GatParams p) {
  constexpr int U0 = 8;
  constexpr int U = U0 < GROUP ? U0 : GROUP;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (GROUP - 1);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gidx = tid / GROUP;
  const int64_t row = gidx / p.S;
  const int slice = (int)(gidx - row * p.S);
  int lb = 0, hb = 0;
  bool hubrow = false;
  if (row < p.n_rows) {
    lb = __ldg(p.rowptr + row);
    hb = __ldg(p.rowptr + row + 1);
    hubrow = p.hub_T > 0 && hb - lb > p.hub_T;  // hub row: handled by the unfused hub path
    if (hubrow) hb = lb;
  }
  const int cv = slice * GROUP + gl;
  const bool colok = cv < p.HFV && row < p.n_rows;
  const int head = colok ? cv / p.FVL : 0;
  const float hl = colok ? __ldg(p.h_l + row * p.H + head) : 0.f;

  // ---- phase 1: online softmax statistics of this lane's head over the row's edges
  float m = -CUDART_INF_F, s = 0.f;
  if (colok) {
    const int tl = gl & (p.team - 1);
    for (int e = lb + tl; e < hb; e += p.team) {
      const int c = ld_stream(p.colind + e);
      const float z = lrelu(hl + __ldg(p.h_r + (int64_t)c * p.H + head), p.slope);
      if (z > m) {
        s = s * expf(m - z) + 1.f;   // m = -inf first time: s = 0 * 0 + 1
        m = z;
      } else {
        s += expf(z - m);
      }
    }
  }
  for (int st = p.team >> 1; st > 0; st >>= 1) {   // team is kernel-uniform: uniform shuffles
    const float m2 = __shfl_xor_sync(FULL, m, st);
    const float s2 = __shfl_xor_sync(FULL, s, st);
    os_merge(m, s, m2, s2);
  }
  const float inv = s > 0.f ? 1.f / s : 0.f;
  const bool writer = p.att_out && colok && (cv % p.FVL == 0);  // one lane per (row, head)

  // ---- phase 2: weighted gather
  const VecT *X = reinterpret_cast<const VecT *>(p.feat);
  int maxdeg = hb - lb;
  if (GROUP < 32) maxdeg = warp_max(maxdeg);
  VecT acc = gat_zero<VecT>();
  for (int off = 0; off < maxdeg; off += GROUP) {
    const int cnt = min(GROUP, hb - lb - off);
    int c = 0;
    if (gl < cnt) c = ld_stream(p.colind + lb + off + gl);
#pragma unroll 1
    for (int j = 0; j < GROUP; j += U) {
      if (GROUP == 32) {
        if (j >= cnt) break;
      } else {
        if (!__any_sync(FULL, j < cnt)) break;
      }
      VecT x[U];
      float z[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int cj = __shfl_sync(FULL, c, j + u, GROUP);
        if (j + u < cnt && colok) {
          x[u] = ld_gather(X + (int64_t)cj * p.HFV + cv);
          z[u] = __ldg(p.h_r + (int64_t)cj * p.H + head);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (j + u < cnt && colok) {
          const float a = expf(lrelu(hl + z[u], p.slope) - m) * inv;
          gat_fma(acc, a, x[u]);
          if (writer) p.att_out[(int64_t)(lb + off + j + u) * p.H + head] = a;
        }
    }
  }
  if (colok && !hubrow)
    st_stream(reinterpret_cast<VecT *>(p.out) + row * p.HFV + cv, acc);
}

// Block per hub row: attention of the hub's edges -> att_out.  H = 2^k <= 32 (256 % H == 0, so a
// thread's head is loop-invariant); pairs (edge, head) are strided over the block.
__global__ void __launch_bounds__(256) gat_hub_att_kernel(const GatParams p) {
  __shared__ float sm_m[8 * 32], sm_s[8 * 32];
  const int row = __ldg(p.hub_rows + blockIdx.x);
  const int lb = __ldg(p.rowptr + row), hb = __ldg(p.rowptr + row + 1);
  const int64_t n = (int64_t)(hb - lb) * p.H;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int head = tid & (p.H - 1);
  const float hl = __ldg(p.h_l + (int64_t)row * p.H + head);
  float m = -CUDART_INF_F, s = 0.f;
  for (int64_t k = tid; k < n; k += 256) {
    const int c = __ldg(p.colind + lb + k / p.H);
    const float z = lrelu(hl + __ldg(p.h_r + (int64_t)c * p.H + head), p.slope);
    if (z > m) { s = s * expf(m - z) + 1.f; m = z; } else { s += expf(z - m); }
  }
  for (int st = 16; st >= p.H; st >>= 1) {
    const float m2 = __shfl_xor_sync(FULL, m, st), s2 = __shfl_xor_sync(FULL, s, st);
    os_merge(m, s, m2, s2);
  }
  sm_m[wid * 32 + lane] = m;
  sm_s[wid * 32 + lane] = s;
  __syncthreads();
  m = sm_m[lane]; s = sm_s[lane];
  for (int q = 1; q < 8; ++q) os_merge(m, s, sm_m[q * 32 + lane], sm_s[q * 32 + lane]);
  const float inv = s > 0.f ? 1.f / s : 0.f;
  for (int64_t k = tid; k < n; k += 256) {
    const int c = __ldg(p.colind + lb + k / p.H);
    const float z = lrelu(hl + __ldg(p.h_r + (int64_t)c * p.H + head), p.slope);
    p.att_out[(int64_t)lb * p.H + k] = expf(z - m) * inv;
  }
}

template <typename VecT, int GROUP>
static int launch_gat(const GatParams &p, cudaStream_t stream) {
  const int64_t blocks = ceil_div(p.n_rows * p.S * GROUP, 256);
  if (blocks == 0) return COGDL_B200_OK;
  if (blocks > 0x7fffffffLL) return set_error(COGDL_B200_EINVAL, "gat_fwd: problem too large for one launch");
  gat_fused_kernel<VecT, GROUP><<<(unsigned)blocks, 256, 0, stream>>>(p);
  CB_LAUNCH_CHECK();
  return COGDL_B200_OK;
}

template <typename VecT>
static int dispatch_gat(GatParams &p, cudaStream_t s) {
  int g = 1;
  while (g < 32 && g < p.HFV) g <<= 1;
  p.S = (int)ceil_div(p.HFV, g);
  // lanes of a group that share a head
  if (p.FVL % g == 0) p.team = g;
  else if (p.FVL < g && (p.FVL & (p.FVL - 1)) == 0) p.team = p.FVL;
  else p.team = 1;
  switch (g) {
    case 1: return launch_gat<VecT, 1>(p, s);
    case 2: return launch_gat<VecT, 2>(p, s);
    case 4: return launch_gat<VecT, 4>(p, s);
    case 8: return launch_gat<VecT, 8>(p, s);
    case 16: return launch_gat<VecT, 16>(p, s);
    default: return launch_gat<VecT, 32>(p, s);
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
  CB_REQUIRE(n_rows < 0x7fffffffLL && H * F < 0x7fffffffLL, "cogdl_b200_gat_fwd_f32: sizes must fit int32");
  cudaStream_t s = (cudaStream_t)stream;
  GatParams p;
  p.rowptr = rowptr; p.colind = colind; p.h_l = h_l; p.h_r = h_r; p.feat = feat; p.slope = negative_slope;
  p.out = out; p.att_out = att_out; p.n_rows = n_rows; p.H = (int)H; p.S = 1; p.team = 1;
  p.hub_T = 0; p.hub_rows = nullptr;
  const bool pow2 = H <= 32 && (H & (H - 1)) == 0;
  const bool hubs = plan && plan->chunk_edges > 0 && plan->n_chunks > 0 && pow2;
  if (hubs) {
    CB_REQUIRE(att_out, "cogdl_b200_gat_fwd_f32: att_out is required when the hub plan has hub rows");
    int rc = check_plan(plan, (int64_t)plan->n_chunks * H * F * (int64_t)sizeof(float));
    if (rc) return rc;
    p.hub_T = plan->chunk_edges; p.hub_rows = plan->hub_rows;
    gat_hub_att_kernel<<<(unsigned)plan->n_hub_rows, 256, 0, s>>>(p);
    CB_LAUNCH_CHECK();
    rc = mhspmm_run(rowptr, colind, nullptr, att_out, feat, out, n_rows, H, F, plan, stream, false);
    if (rc) return rc;
  }
  const bool vec = (F % 4 == 0) && aligned16(feat) && aligned16(out);
  if (vec) {
    p.HFV = (int)(H * F / 4);
    p.FVL = (int)(F / 4);
    return dispatch_gat<float4>(p, s);
  }
  p.HFV = (int)(H * F);
  p.FVL = (int)F;
  return dispatch_gat<float>(p, s);
}
