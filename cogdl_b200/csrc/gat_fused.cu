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
// Two stages:
//   1. attention (edge_softmax.cu, MODE 2): logits are formed in registers / shared memory from h_l,
//      h_r and soft-maxed per row in the same pass; only the normalised attention [nnz,H] is written
//      (it is also what the backward saves).  Same row tiers as the edge softmax: segments of short
//      rows, warp, block, and an 8-CTA cluster with a DSMEM combine for giant hub rows.
//   2. the row-stream multi-head SpMM (stream.cuh) consumes it.
// A single-kernel fusion would save the attention write+read: 8*H bytes per edge against 4*H*F + ...
// bytes of feature gather (1.6 % at H=8, F=128) -- not worth a second copy of the stream kernel.
// (The first version of this file was a per-(row, slice) fused kernel that recomputed the softmax
// statistics in every slice: 4.6 ms vs 1.2 ms for the two-stage form on the arxiv shape.)
#include "common.cuh"

namespace cogdl_b200 {

int gat_attention(const int32_t *rowptr, const int32_t *colind, const float *h_l, const float *h_r, float slope,
                  float *att, int64_t n_rows, int64_t H, const cogdl_b200_hub_plan_t *plan, cudaStream_t s);

int mhspmm_run(const int32_t *rowptr, const int32_t *colind, const int32_t *perm, const float *att,
               const float *feat, float *out, int64_t n_rows, int64_t H, int64_t F,
               const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream, bool rows_too);

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
  rc = gat_attention(rowptr, colind, h_l, h_r, negative_slope, att_out, n_rows, H, plan, (cudaStream_t)stream);
  if (rc) return rc;
  return mhspmm_run(rowptr, colind, nullptr, att_out, feat, out, n_rows, H, F, plan, stream, true);
}
