// mhspmm.cu -- multi-head SpMM (GAT aggregation) for sm_100a.
//
//   out[i,h,:] = sum_{p in row i} att[P(p),h] * feat[colind[p],h,:]      P(p) = perm ? perm[p] : p
//
// Replaces mhspmmSimple / mhspmm_1 (cogdl/operators/spmm/multiheadSpmm.cu:6-51): grid (V,H) with
// F-thread blocks => rowptr/colind re-read H times, one dependent colind->feat load chain per
// edge, F=16 gives half-warp blocks.  `perm` additionally fuses the backward's mhtranspose
// (mhTranspose.cu:6-49, mhspmm.py:60-61) so the CSC pass never materialises att[perm].
//
// A node's [H,F] block is one contiguous run of H*F floats.  It is cut into slices of GROUP
// 16-byte vectors (512 B for GROUP = 32); a GROUP of lanes owns one (row-or-hub-chunk, slice)
// item, slices of the same row sit in adjacent warps so the index/attention lines they share
// are L1 hits.  The lane's head is fixed (its column / F), its attention scalar is fetched per
// edge next to the feature gather (same 32-byte sector for all lanes of a head).  Accumulation
// order and rounding are those of the SpMM (CSR order, separate fp32 mul and add), i.e. the
// reference's CPU fallback (one spmm_cpu per head, spmm_utils.py:216-225).
#include "common.cuh"
#include "stream.cuh"

namespace cogdl_b200 {

struct MhParams {
  const int *rowptr;
  const int *colind;
  const int *perm;   // nullable
  const float *att;  // [nnz, H]
  const float *feat; // [n_src, H*F]
  float *out;        // [n_rows, H*F]
  int64_t n_rows;
  int H;
  int HFV;           // H*F in vector units
  int FVL;           // F in vector units  (head = col / FVL)
  int S;             // slices per row
  HubView hub;
};

template <typename VecT> __device__ __forceinline__ VecT mh_zero();
template <> __device__ __forceinline__ float4 mh_zero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ float mh_zero<float>() { return 0.f; }
__device__ __forceinline__ void mh_add(float &a, const float &b) { a = __fadd_rn(a, b); }
__device__ __forceinline__ void mh_add(float4 &a, const float4 &b) { add_rn(a, b); }

template <typename VecT, int GROUP, bool HAS_PERM>
__global__ void __launch_bounds__(256) mhspmm_kernel(const MhParams p) {
  constexpr int U0 = 8;
  constexpr int U = U0 < GROUP ? U0 : GROUP;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (GROUP - 1);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gidx = tid / GROUP;
  const int64_t item = gidx / p.S;
  const int slice = (int)(gidx - item * p.S);
  const WorkItem w = decode_item(item, p.n_rows, p.rowptr, p.hub);
  const bool warp_has_chunk = ((tid - lane) / GROUP) / p.S < p.hub.n_chunks;

  const int cv = slice * GROUP + gl;
  const bool colok = cv < p.HFV;
  const int head = colok ? cv / p.FVL : 0;
  const VecT *X = reinterpret_cast<const VecT *>(p.feat);
  VecT *Y = reinterpret_cast<VecT *>(p.out);
  VecT *P = reinterpret_cast<VecT *>(p.hub.partials);

  int maxdeg = w.hb - w.lb;
  if (GROUP < 32) maxdeg = warp_max(maxdeg);

  VecT acc = mh_zero<VecT>();
  int c = 0, pe = 0;
  {
    const int e = w.lb + gl;
    if (e < w.hb) {
      c = ld_stream(p.colind + e);
      pe = HAS_PERM ? ld_stream(p.perm + e) : e;
    }
  }
  for (int off = 0; off < maxdeg; off += GROUP) {
    const int cnt = min(GROUP, w.hb - w.lb - off);
    int cn = 0, pn = 0;
    {
      const int e = w.lb + off + GROUP + gl;
      if (e < w.hb) {
        cn = ld_stream(p.colind + e);
        pn = HAS_PERM ? ld_stream(p.perm + e) : e;
      }
    }
#pragma unroll 1
    for (int j = 0; j < GROUP; j += U) {
      if (GROUP == 32) {
        if (j >= cnt) break;
      } else {
        if (!__any_sync(FULL, j < cnt)) break;
      }
      VecT x[U];
      float a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int cj = __shfl_sync(FULL, c, j + u, GROUP);
        const int pj = __shfl_sync(FULL, pe, j + u, GROUP);
        if (j + u < cnt && colok) {
          x[u] = ld_gather(X + (int64_t)cj * p.HFV + cv);
          a[u] = __ldg(p.att + (int64_t)pj * p.H + head);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (j + u < cnt && colok) axpy_rn(acc, a[u], x[u]);
    }
    c = cn;
    pe = pn;
  }

  if (!w.is_chunk) {
    if (w.active && colok) st_stream(Y + (int64_t)w.row * p.HFV + cv, acc);
  } else if (colok) {
    st_cg(P + (int64_t)w.slot * p.HFV + cv, acc);
  }
  if (warp_has_chunk) {
    // all S slice-groups of every chunk of the row arrive on one counter; the last one of all
    // combines the whole row (every slice) in chunk order
    WorkItem wa = w;
    wa.n_row_chunks = w.n_row_chunks * p.S;
    if (hub_arrive_last<GROUP>(wa, p.hub, gl)) {
      for (int k = gl; k < p.HFV; k += GROUP) {
        const VecT *pp = P + (int64_t)w.first * p.HFV + k;
        VecT s = ld_cg(pp);
        for (int q = 1; q < w.n_row_chunks; ++q) mh_add(s, ld_cg(pp + (int64_t)q * p.HFV));
        st_stream(Y + (int64_t)w.row * p.HFV + k, s);
      }
    }
  }
}

template <typename VecT, int GROUP>
static int launch_mh(const MhParams &p, cudaStream_t stream) {
  const int64_t groups = ((int64_t)p.hub.n_chunks + p.n_rows) * p.S;
  const int64_t blocks = ceil_div(groups * GROUP, 256);
  if (blocks == 0) return COGDL_B200_OK;
  if (blocks > 0x7fffffffLL) return set_error(COGDL_B200_EINVAL, "mhspmm: problem too large for one launch");
  if (p.perm)
    mhspmm_kernel<VecT, GROUP, true><<<(unsigned)blocks, 256, 0, stream>>>(p);
  else
    mhspmm_kernel<VecT, GROUP, false><<<(unsigned)blocks, 256, 0, stream>>>(p);
  CB_LAUNCH_CHECK();
  return COGDL_B200_OK;
}

template <typename VecT>
static int dispatch_mh(MhParams &p, cudaStream_t s) {
  const int n = p.HFV;
  int g = 1;
  while (g < 32 && g < n) g <<= 1;
  p.S = (int)ceil_div(n, g);
  if (g == 32 && p.hub.n_segs > 0) {
    // row-stream form: one warp per (segment or hub chunk, 512-byte slice)
    StreamParams q;
    q.rowptr = p.rowptr; q.colind = p.colind; q.val = nullptr; q.att = p.att; q.perm = p.perm;
    q.X0 = p.feat; q.X1 = p.feat; q.n0 = INT64_MAX; q.Y = p.out; q.ldv = p.HFV; q.H = p.H; q.FVL = p.FVL;
    q.S = p.S; q.hub = p.hub; q.n_peers = 0; q.peer_shift = 0;
    if (p.n_rows == 0) q.hub.n_segs = 0;   // chunks-only call (fused GAT hub path)
    return launch_stream<VecT, 1, 8, 3>(q, MODE_MULTIHEAD, s);
  }
  switch (g) {
    case 1: return launch_mh<VecT, 1>(p, s);
    case 2: return launch_mh<VecT, 2>(p, s);
    case 4: return launch_mh<VecT, 4>(p, s);
    case 8: return launch_mh<VecT, 8>(p, s);
    case 16: return launch_mh<VecT, 16>(p, s);
    default: return launch_mh<VecT, 32>(p, s);
  }
}

// rows_too == false processes only the hub-chunk items of the plan.
int mhspmm_run(const int32_t *rowptr, const int32_t *colind, const int32_t *perm, const float *att,
               const float *feat, float *out, int64_t n_rows, int64_t H, int64_t F,
               const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream, bool rows_too) {
  MhParams p;
  p.rowptr = rowptr; p.colind = colind; p.perm = perm; p.att = att; p.feat = feat; p.out = out;
  p.n_rows = rows_too ? n_rows : 0; p.H = (int)H; p.hub = hub_view(plan); p.S = 1;
  const bool vec = (F % 4 == 0) && aligned16(feat) && aligned16(out) &&
                   (p.hub.n_chunks == 0 || aligned16(p.hub.partials));
  if (vec) {
    p.HFV = (int)(H * F / 4);
    p.FVL = (int)(F / 4);
    return dispatch_mh<float4>(p, (cudaStream_t)stream);
  }
  p.HFV = (int)(H * F);
  p.FVL = (int)F;
  return dispatch_mh<float>(p, (cudaStream_t)stream);
}

}  // namespace cogdl_b200

using namespace cogdl_b200;

extern "C" int cogdl_b200_mhspmm_f32(const int32_t *rowptr, const int32_t *colind, const int32_t *perm,
                                     const float *att, const float *feat, float *out, int64_t n_rows,
                                     int64_t H, int64_t F, const cogdl_b200_hub_plan_t *plan,
                                     cogdl_b200_stream_t stream) {
  CB_REQUIRE(n_rows >= 0 && H >= 0 && F >= 0, "cogdl_b200_mhspmm_f32: negative size");
  if (n_rows == 0 || H == 0 || F == 0) return COGDL_B200_OK;
  CB_REQUIRE(rowptr && colind && att && feat && out, "cogdl_b200_mhspmm_f32: null pointer");
  CB_REQUIRE(n_rows < 0x7fffffffLL && H * F < 0x7fffffffLL, "cogdl_b200_mhspmm_f32: sizes must fit int32");
  int rc = check_plan(plan, (plan ? (int64_t)plan->n_chunks : 0) * H * F * (int64_t)sizeof(float));
  if (rc) return rc;
  return mhspmm_run(rowptr, colind, perm, att, feat, out, n_rows, H, F, plan, stream, true);
}

// ------------------------------------------------------------------------------------------
// GAT forward:   e[p,h] = leakyrelu(h_l[i,h] + h_r[colind[p],h])   (p in row i)
//                              a[p,:] = softmax over the row's edges, per head
//                              out[i,h] = sum_p a[p,h] * feat[colind[p],h,:]
// Replaces the unfused chain of cogdl/layers/gat_layer.py:73-77 (two [nnz,H] gathers, an add, a LeakyReLU and a
// softmax, each a kernel with its own [nnz,H] temporary) and the stale dgNN binding (cogdl/operators/fused_gat.py:
// 17-19; precedent third_party/dgNN/.../fused_gatconv_kernel.cu:26-133).  Two stages on the caller's stream:
//   1. attention (edge_softmax.cu, MODE 2): logits formed from h_l / h_r while loading and soft-maxed per row in
//      the same pass; only the normalised attention [nnz,H] is written -- it is also what the backward saves;
//   2. the row-stream multi-head SpMM above consumes it (0.97 of the HBM roofline at H=8, F=128, so a single-pass
//      fusion could only remove the 8*H-byte attention round trip per edge against 4*H*F bytes of gather: 1.6 %;
//      round 1's per-(row, slice) fully fused kernel recomputed the statistics per slice and was 4x slower).
// The backward is native as well: see cogdl_b200_gat_attn_bwd_f32 (edge_softmax.cu) and operators/fused_gat.py.
// ------------------------------------------------------------------------------------------
namespace cogdl_b200 {
int gat_attention(const int32_t *rowptr, const int32_t *colind, const float *h_l, const float *h_r, float slope,
                  float *att, int64_t n_rows, int64_t H, const cogdl_b200_hub_plan_t *plan, cudaStream_t s);
}

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
  rc = gat_attention(rowptr, colind, h_l, h_r, negative_slope, att_out, n_rows, H, plan, (cudaStream_t)stream);
  if (rc) return rc;
  return mhspmm_run(rowptr, colind, nullptr, att_out, feat, out, n_rows, H, F, plan, stream, true);
}
