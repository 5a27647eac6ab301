/*
 * cogdl_b200.h -- C ABI of libcogdl_b200.so: B200 (sm_100a) kernels for CogDL's sparse
 * message-passing hot path (CSR SpMM / SDDMM / edge-softmax / multi-head SpMM / scatter_max).
 *
 * This is the drop-in boundary.  Each entry point replaces one function that the reference
 * exposes today through a pybind11/torch extension (THUDM/CogDL @ 281f4742; paths relative to
 * the CogDL tree).  Differences from the reference ABI, all deliberate:
 *   - plain C: raw device pointers + int64 sizes + an opaque stream handle, no torch types;
 *   - the caller owns every buffer (outputs included); no device memory is allocated here (the
 *     reference leaks a cusparseHandle and cudaMalloc's per call, spmm_kernel.cu:517-531); the library
 *     keeps no streams or events of its own;
 *   - work is enqueued on the caller's stream (the reference uses the legacy default stream);
 *   - errors are returned (0 = ok, <0 = COGDL_B200_E*), text via cogdl_b200_last_error();
 *     the reference `assert`s / exit(1)s (spmm.cpp:28-39, computeUtil.h:13-27);
 *   - offsets are 64-bit (the reference overflows int32 at N*F >= 2^31, spmm_kernel.cu:393).
 * Index dtype is int32 and feature dtype fp32, exactly as the reference kernels take them
 * (`rowptr.int(), colind.int()`, cogdl/utils/spmm_utils.py:106).  All arrays are contiguous,
 * row-major; float arrays must be 4-byte aligned (16-byte alignment enables the vector path).
 *
 * No CPU fallback exists in this library: every call needs a CUDA device of compute
 * capability 10.x and fails with COGDL_B200_EDEVICE otherwise.
 */
#ifndef COGDL_B200_H_
#define COGDL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COGDL_B200_ABI_VERSION 5

#define COGDL_B200_OK 0
#define COGDL_B200_EINVAL (-1)   /* bad argument (null pointer, negative size, unsupported shape) */
#define COGDL_B200_ECUDA (-2)    /* a CUDA runtime call or kernel launch failed */
#define COGDL_B200_EDEVICE (-3)  /* no sm_100-class device / kernel image not loadable */
#define COGDL_B200_ESCRATCH (-4) /* hub-plan scratch too small for this call */

/* Exported-symbol marker (the library is built with -fvisibility=hidden). */
#if defined(COGDL_B200_BUILD)
#define COGDL_B200_API __attribute__((visibility("default")))
#else
#define COGDL_B200_API
#endif

/* cudaStream_t, passed as an opaque pointer (NULL = legacy default stream). */
typedef void *cogdl_b200_stream_t;

COGDL_B200_API int cogdl_b200_abi_version(void);
/* Thread-local text of the last error returned on this thread ("" if none). */
COGDL_B200_API const char *cogdl_b200_last_error(void);
/* 0 if the current device can run this library (compute capability 10.x), else EDEVICE. */
COGDL_B200_API int cogdl_b200_check_device(void);
/* Number of kernels this library has launched in this process (all threads). */
COGDL_B200_API int64_t cogdl_b200_launch_count(void);
/* Thread-local description of the kernel instantiation chosen by the last dispatching call on this
 * thread (e.g. "stream_kernel<float4,NV=1,weighted,SRC_ONE,U=4>"); "" before the first call.  Lets a
 * harness check that a profile it cites is of the kernel that actually ran. */
COGDL_B200_API const char *cogdl_b200_last_kernel(void);
/* Experiment knobs (environment variables COGDL_B200_*: kernel variants, tile floors) are read once and cached;
 * this drops the cache so that a tuning sweep can change them inside one process.  Not needed in normal use. */
COGDL_B200_API void cogdl_b200_reload_tuning(void);
/* What the knob `name` currently resolves to: the cached value, else the environment, else `dflt`. */
COGDL_B200_API int cogdl_b200_tuning_value(const char *name, int dflt);

/* ---------------------------------------------------------------------------------------
 * Hub plan: how rows with more than `chunk_edges` edges are cut into fixed-size edge chunks
 * so that a power-law hub does not serialise on one warp (the reference serialises it:
 * one warp per row, spmm_kernel.cu:370-440).  Built once per CSR structure, reused by every
 * op below; optional (plan == NULL => every row is processed whole, in CSR order).
 *
 *   hub_rows[n_hub_rows]        rows with degree > chunk_edges (any order)
 *   chunks[2*n_chunks]          (row, first_slot) per chunk slot; a row's chunks occupy the
 *                               contiguous slots [first_slot, first_slot + ceil(deg/chunk_edges))
 *                               and chunk j covers edges rowptr[row] + j*chunk_edges ...
 *   counters[n_chunks]          zero on entry; left zero on exit (arrival counters)
 *   partials                    scratch for per-chunk partial results; `partials_bytes` capacity.
 *                               needs n_chunks*F*4 bytes (SpMM, mh-SpMM: F := H*F) or
 *                               n_chunks*F*8 bytes (scatter_max: value + argmax).
 * counters/partials are per-stream scratch: do not share one plan between concurrent streams.
 * Partial results of a hub row are combined in chunk order by the last chunk to finish, so
 * results are deterministic run to run.
 * ------------------------------------------------------------------------------------- */
typedef struct cogdl_b200_hub_plan {
  int32_t chunk_edges;
  int32_t n_hub_rows;
  int32_t n_chunks;
  int32_t n_empty_rows;      /* rows of degree 0 (the row-stream kernels zero-fill them) */
  const int32_t *hub_rows;
  const int32_t *chunks;
  int32_t *counters;
  void *partials;
  int64_t partials_bytes;
  /* Row-stream segments (optional; segs == NULL => one warp per row).  A segment is a run of
   * consecutive NON-hub rows [row_begin, row_end) of roughly seg_cost rows+edges; its edges are
   * contiguous in colind, so one warp streams them in coalesced 32-edge slabs with always-full
   * gather batches and flushes its accumulator at row ends.  edge_row[p] = row owning edge p
   * (the COO row array) lets the kernel find row ends with one ballot per slab. */
  int32_t seg_cost;
  int32_t n_segs;
  const int32_t *segs;     /* [2*n_segs] (row_begin, row_end), any order */
  const int32_t *edge_row; /* [nnz] */
  /* Optional: HOST array of the hub rows' degrees, in hub_rows order, sorted DESCENDING (the caller
   * sorts hub_rows accordingly).  Lets row-tiered ops (edge softmax, GAT attention) launch exactly
   * as many blocks / clusters as there are rows in a tier.  NULL => such ops treat every hub row
   * with the per-warp tier. */
  const int32_t *hub_degrees_host;
} cogdl_b200_hub_plan_t;

/* Layout self-check for bindings that mirror the struct (ctypes, cgo, JNI): writes sizeof(cogdl_b200_hub_plan_t)
 * to out[0] and the byte offset of each field, in declaration order, to out[1..]; returns the number of values
 * written (at most n). */
COGDL_B200_API int cogdl_b200_hub_plan_layout(int64_t *out, int n);

/* Pass 1: counts_dev[0..3] = #rows with degree > chunk_edges, #chunks, #empty rows, #segments
 * (for seg_cost; pass seg_cost <= 0 to skip segments).  counts_dev: 4-int device buffer. */
COGDL_B200_API int cogdl_b200_hub_plan_count(const int32_t *rowptr, int64_t n_rows, int32_t chunk_edges,
                              int32_t seg_cost, int32_t *counts_dev, cogdl_b200_stream_t stream);
/* Pass 2: fills hub_rows[counts[0]], chunks[2*counts[1]] and segs[2*counts[3]] (caller-allocated
 * from pass 1's counts; segs may be NULL with seg_cost <= 0); counts_dev is reused as allocator. */
COGDL_B200_API int cogdl_b200_hub_plan_fill(const int32_t *rowptr, int64_t n_rows, int32_t chunk_edges,
                             int32_t seg_cost, int32_t *counts_dev, int32_t *hub_rows, int32_t *chunks,
                             int32_t *segs, cogdl_b200_stream_t stream);
/* edge_row[p] = row owning CSR position p (CSR -> COO row expansion; reference keeps the same
 * array as Adjacency.row, cogdl/data/data.py:136). */
COGDL_B200_API int cogdl_b200_edge_rows(const int32_t *rowptr, int64_t n_rows, int64_t nnz, int32_t *edge_row,
                         cogdl_b200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * CSR SpMM   Y[i,:] = sum_{p in row i} val[p] * X[colind[p],:]        (val == NULL => 1)
 * Replaces  spmm.csr_spmm(rowptr, colind, val, B)          cogdl/operators/spmm/spmm.cpp:22-45
 *           spmm.csr_spmm_no_edge_value(rowptr, colind, B) cogdl/operators/spmm/spmm.cpp:47-70
 *           (kernels spmm_test{0,1,2}, topo*SPMMKernel     .../spmm/spmm_kernel.cu:7-512)
 *           spmm_cpu.csr_spmm_cpu                          .../spmm/spmm_cpu.cpp:39-58 (semantics)
 * X is [n_src, F], Y is [n_rows, F]; colind values index rows of X.  Rows not split by a hub
 * plan are accumulated in CSR order with separate fp32 multiply and add, i.e. bit-identical
 * to the reference CPU SpMM (spmm_cpu.cpp:24-36).
 * ------------------------------------------------------------------------------------- */
COGDL_B200_API int cogdl_b200_spmm_csr_f32(const int32_t *rowptr, const int32_t *colind, const float *val,
                            const float *X, float *Y, int64_t n_rows, int64_t F,
                            const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream);

/* fp16 storage (X, val, Y are IEEE half), fp32 accumulation (the reference accumulates in
 * half, spmm_kernel.cu:222-250,311-368,442-512 -- documented divergence). */
COGDL_B200_API int cogdl_b200_spmm_csr_f16(const int32_t *rowptr, const int32_t *colind, const void *val,
                            const void *X, void *Y, int64_t n_rows, int64_t F,
                            const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream);

/* Two-source variant used by the node-range partitioned (multi-GPU) SpMM: column c reads
 * X0[c] when c < n0, else X1[c - n0] (X1 = halo rows received from / mapped on peers).
 * New functionality, no reference counterpart. */
COGDL_B200_API int cogdl_b200_spmm_csr_f32_2src(const int32_t *rowptr, const int32_t *colind, const float *val,
                                 const float *X0, int64_t n0, const float *X1, float *Y,
                                 int64_t n_rows, int64_t F, const cogdl_b200_hub_plan_t *plan,
                                 cogdl_b200_stream_t stream);

/* Peer form of the partitioned SpMM: the gather over remote feature rows is fused into the kernel.
 * A column c < n_local reads X_local[c]; a column c >= n_local encodes r = c - n_local with
 * owner = r >> owner_shift and row = r & ((1 << owner_shift) - 1), and is read DIRECTLY from
 * peer_ptrs[owner] -- that rank's feature shard mapped into this process (symmetric memory /
 * CUDA IPC), i.e. ld.global over NVLink 5 / NVSwitch inside the SpMM kernel: no pack kernel, no
 * all-to-all, no halo buffer.  peer_ptrs is a HOST array of n_peers (<= 8) device pointers.
 * Needs a hub plan with segments (row-stream kernel), F % 4 == 0, F <= 512.  The caller orders
 * the peers' writes to their shards before this call (barrier) -- see cogdl_b200/dist.py. */
COGDL_B200_API int cogdl_b200_spmm_csr_f32_peers(const int32_t *rowptr, const int32_t *colind, const float *val,
                                  const float *X_local, int64_t n_local, const float *const *peer_ptrs,
                                  int32_t n_peers, int32_t owner_shift, float *Y, int64_t n_rows, int64_t F,
                                  const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * CSR SDDMM   out[p] = < D1[row(p),:], D2[colind[p],:] >
 * Replaces  sddmm.csr_sddmm(rowptr, colind, D1, D2)   cogdl/operators/spmm/sddmm.cpp:47-70
 *           (kernels sddmmCSR{2,1}Scale               .../spmm/sddmm_kernel.cu:249-417)
 * ------------------------------------------------------------------------------------- */
COGDL_B200_API int cogdl_b200_sddmm_csr_f32(const int32_t *rowptr, const int32_t *colind, const float *D1,
                             const float *D2, float *out, int64_t n_rows, int64_t F,
                             const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * CSR -> CSC with permutation (integer only).
 * Replaces  spmm.csr2csc(rowptr, colind, val)   cogdl/operators/spmm/spmm.cpp:72-90
 *           (cuSPARSE cusparseCsr2cscEx2        .../spmm/spmm_kernel.cu:514-532,596-613)
 *           mhtranspose.csr2csc                 .../spmm/mhTranspose.cu:51-111
 * perm[q] = CSR position of CSC entry q; stable (inside a column, ascending row).  Values
 * are moved with cogdl_b200_gather_rows_f32(perm, ...) -- exact for any nnz (the reference
 * round-trips edge ids through fp32, wrong beyond 2^24 edges, mhspmm.py:55-59).
 * `workspace` is caller-provided device scratch of at least
 * cogdl_b200_csr2csc_workspace_bytes(nnz, n_cols) bytes.
 * ------------------------------------------------------------------------------------- */
COGDL_B200_API int64_t cogdl_b200_csr2csc_workspace_bytes(int64_t nnz, int64_t n_cols);
COGDL_B200_API int cogdl_b200_csr2csc(const int32_t *rowptr, const int32_t *colind, int64_t n_rows,
                       int64_t n_cols, int64_t nnz, int32_t *colptr, int32_t *rowind,
                       int32_t *perm, void *workspace, int64_t workspace_bytes,
                       cogdl_b200_stream_t stream);

/* out[q,:] = in[perm[q],:] over [nnz, H] fp32.
 * Replaces  mhtranspose.mhtranspose(perm, att)   cogdl/operators/spmm/mhTranspose.cu:6-49 */
COGDL_B200_API int cogdl_b200_gather_rows_f32(const int32_t *perm, const float *in, float *out, int64_t nnz,
                               int64_t H, cogdl_b200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Edge softmax over the edges of each destination row, per head; [nnz, H] fp32.
 * Replaces  edge_softmax.edge_softmax(rowptr, w)               cogdl/operators/edge_softmax/edge_softmax.cc:16-31
 *           edge_softmax.edge_softmax_backward(rowptr, y, g)   .../edge_softmax.cc:33-49
 *           (kernels                                           .../edge_softmax.cu:7-98)
 * fwd: out = exp(in - max_row) / sum_row exp(in - max_row).  bwd: gin = y * (g - sum_row y*g).
 * Degree-0 rows own no entries, so nothing is written for them.
 * ------------------------------------------------------------------------------------- */
/* With a hub plan that has hub rows, the calls below (and the GAT attention stages) need per-call scratch
 * in plan->partials: at least cogdl_b200_edge_softmax_scratch_bytes(plan->n_chunks, H) bytes (per hub
 * chunk and head: a (max, sum) pair for the two-pass split-row softmax + one partial row sum). */
COGDL_B200_API int64_t cogdl_b200_edge_softmax_scratch_bytes(int64_t n_chunks, int64_t H);
COGDL_B200_API int cogdl_b200_edge_softmax_fwd_f32(const int32_t *rowptr, const float *in, float *out,
                                    int64_t n_rows, int64_t H,
                                    const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream);
COGDL_B200_API int cogdl_b200_edge_softmax_bwd_f32(const int32_t *rowptr, const float *y, const float *g,
                                    float *gin, int64_t n_rows, int64_t H,
                                    const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Multi-head SpMM   out[i,h,:] = sum_p att[P(p),h] * feat[colind[p],h,:],  P(p) = perm ? perm[p] : p
 * Replaces  mhspmm.mhspmm(rowptr, colind, att, feat)   cogdl/operators/spmm/multiheadSpmm.cpp:8-34
 *           (kernels mhspmmSimple / mhspmm_1           .../spmm/multiheadSpmm.cu:6-51)
 * feat/out are [n, H, F]; att is [nnz, H].  `perm` fuses the backward's mhtranspose
 * (mhspmm.py:60-61) into the CSC pass.  Accumulation order/rounding as for SpMM.
 * ------------------------------------------------------------------------------------- */
COGDL_B200_API int cogdl_b200_mhspmm_f32(const int32_t *rowptr, const int32_t *colind, const int32_t *perm,
                          const float *att, const float *feat, float *out, int64_t n_rows,
                          int64_t H, int64_t F, const cogdl_b200_hub_plan_t *plan,
                          cogdl_b200_stream_t stream);

/* Multi-head SDDMM   out[p,h] = < grad[row(p),h,:], feat[colind[p],h,:] >
 * Replaces  mhsddmm.mhsddmm(rowptr, colind, grad, feat)   cogdl/operators/spmm/multiheadSddmm.cu:6-113 */
COGDL_B200_API int cogdl_b200_mhsddmm_f32(const int32_t *rowptr, const int32_t *colind, const float *grad,
                           const float *feat, float *out, int64_t n_rows, int64_t H, int64_t F,
                           const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * scatter_max (neighbour max + argmax) and its backward.
 * Replaces  scatter_max.scatter_max_fp(rowptr, colind, x) -> [out, max_id]   cogdl/operators/scatter_max/scatter_max.cc:6-22
 *           scatter_max.scatter_max_bp(grad, max_id)                          .../scatter_max.cc:24-38
 *           (kernels                                                          .../scatter_max.cu:5-42)
 * fwd: out[i,f] = max_p X[colind[p],f]; argmax[i,f] = colind of the first edge (CSR order)
 *      attaining it.  Fixed semantics (documented divergences from the reference bugs):
 *      the running max starts at -inf (reference: FLT_MIN, scatter_max.cu:16); degree-0 rows
 *      give out = 0, argmax = -1 (reference: uninitialised max_id).  Identical to the
 *      reference whenever every row has a neighbour value > FLT_MIN.
 * bwd: gx[argmax[i,f], f] += g[i,f]; gx [n_src, F] is zero-filled by the call (the reference
 *      forgets to, scatter_max.cu:70); argmax < 0 is skipped.
 * ------------------------------------------------------------------------------------- */
COGDL_B200_API int cogdl_b200_scatter_max_fwd_f32(const int32_t *rowptr, const int32_t *colind, const float *X,
                                   float *out, int32_t *argmax, int64_t n_rows, int64_t F,
                                   const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream);
COGDL_B200_API int cogdl_b200_scatter_max_bwd_f32(const float *g, const int32_t *argmax, float *gx,
                                   int64_t n_rows, int64_t n_src, int64_t F,
                                   cogdl_b200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * "Next" rows.
 *
 * GAT forward: e = leakyrelu(h_l[i,h] + h_r[col,h]); a = softmax_row(e); out = sum a*feat.
 * The logits are formed and soft-maxed in registers (no [nnz,H] gather / add / activation
 * temporaries); the normalised attention is written once to att_out ([nnz,H], required: it is the
 * saved tensor of the backward and the operand of the row-stream multi-head SpMM that follows).
 * Replaces the unfused chain cogdl/layers/gat_layer.py:73-77 and the stale dgNN binding
 * cogdl/operators/fused_gat.py:17-19.
 * ------------------------------------------------------------------------------------- */
COGDL_B200_API int cogdl_b200_gat_fwd_f32(const int32_t *rowptr, const int32_t *colind, const float *h_l,
                           const float *h_r, const float *feat, float negative_slope, float *out,
                           float *att_out, int64_t n_rows, int64_t H, int64_t F,
                           const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream);

/* GAT attention backward (autograd of cogdl/layers/gat_layer.py:73-74, precedent
 * third_party/dgNN/dgNN/src/fused_gatconv/fused_gatconv_kernel.cu:636-790): softmax backward, LeakyReLU'
 * and the row sums in ONE pass over the edge-aligned tensors --
 *   d_edge[p,h] = att*(d_att - sum_row att*d_att) * leakyrelu'(h_l[row,h] + h_r[col,h])   [nnz,H]
 *   g_row[i,h]  = sum_{p in row i} d_edge[p,h]                                            [n_rows,H]
 * d_att is the multi-head SDDMM of (d out, feat).  g_col = cogdl_b200_edge_colsum_f32 over the transpose.
 * Scratch in plan->partials as for the edge softmax. */
COGDL_B200_API int cogdl_b200_gat_attn_bwd_f32(const int32_t *rowptr, const int32_t *colind, const float *att,
                                const float *d_att, const float *h_l, const float *h_r, float negative_slope,
                                float *d_edge, float *g_row, int64_t n_rows, int64_t H,
                                const cogdl_b200_hub_plan_t *plan, cogdl_b200_stream_t stream);
/* out[j,h] = sum_{q in column j of the transpose} e[perm[q], h]: column sums of an edge-aligned [nnz,H]
 * tensor through the cached CSC (colptr, perm).  Deterministic (no atomics; hub columns via the plan's chunks). */
COGDL_B200_API int cogdl_b200_edge_colsum_f32(const int32_t *colptr, const int32_t *perm, const float *e, float *out,
                               int64_t n_cols, int64_t H, const cogdl_b200_hub_plan_t *plan /* of the transpose; scratch n_chunks*H*4 B */,
                               cogdl_b200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Fused GCN layer: OUT = act( (A.X).W^T + (A.1) b^T ), A = (rowptr, colind, val).
 * Equal to the reference layer  act( A.(X.W^T + 1 b^T) )  (cogdl/layers/gcn_layer.py:51-64, bias inside the
 * aggregation) up to fp32 rounding; the products are re-associated so the gather runs on the raw features
 * and the dense transform runs on the aggregated 128-row tile while it is on chip: bf16x3-split operands
 * (24 mantissa bits, six products), tcgen05.mma into TMEM, fp32 accumulate -- error vs fp64 ~1e-7 of the row
 * scale.  X [n_src,K], W [Fout,K] (torch.nn.Linear.weight layout), bias [Fout] or NULL, out [n_rows,Fout].
 * Needs K == 128, Fout <= 128 (cogdl_b200_gcn_fused_supported), a hub plan with edge_row, 16-byte aligned
 * X / W / out.  hub_agg: scratch [n_rows, K] fp32, required iff the plan has hub rows (only those rows are
 * written); plan->partials as for cogdl_b200_spmm_csr_f32 (n_chunks*K*4 bytes).  rowsum: optional [n_rows]
 * = A.1 (sum of each row's edge values), worth caching when the edge weights are fixed (GCN); NULL =>
 * the kernel sums the weights itself (one warp per hub row: slow on 10^4-edge hubs).
 * ------------------------------------------------------------------------------------- */
COGDL_B200_API int cogdl_b200_gcn_fused_supported(int64_t K, int64_t Fout);
COGDL_B200_API int cogdl_b200_gcn_fused_f32(const int32_t *rowptr, const int32_t *colind, const float *val, const float *X,
                             const float *W, const float *bias, const float *rowsum, float *out, float *hub_agg, int64_t n_rows,
                             int64_t K, int64_t Fout, int32_t relu, const cogdl_b200_hub_plan_t *plan,
                             cogdl_b200_stream_t stream);

/* Device COO -> CSR (stable): row_ptr[num_nodes+1] and reindex[nnz] (CSR slot -> COO position),
 * int64 as cogdl.data.Graph stores them.
 * Replaces  sampler.coo2csr_cpu_index(row, col, num_nodes)   cogdl/operators/sample/sample.cpp:234-270
 * (single-thread CPU counting sort even for CUDA graphs, cogdl/utils/graph_utils.py:133-142). */
COGDL_B200_API int64_t cogdl_b200_coo2csr_workspace_bytes(int64_t nnz, int64_t num_nodes);
COGDL_B200_API int cogdl_b200_coo2csr_index(const int64_t *row, int64_t nnz, int64_t num_nodes, int64_t *row_ptr,
                             int64_t *reindex, void *workspace, int64_t workspace_bytes,
                             cogdl_b200_stream_t stream);

/* int64 -> int32 narrowing of CSR arrays (fails with EINVAL semantics left to the caller: values
 * must fit).  Replaces the per-call `.int()` casts of cogdl/utils/spmm_utils.py:106. */
COGDL_B200_API int cogdl_b200_narrow_i64_i32(const int64_t *in, int32_t *out, int64_t n, cogdl_b200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Neighbour sampling and induced subgraphs on the device.  int64 CSR in / int64 out, as
 * cogdl.data.Graph stores it (row_ptr, col).
 * Replaces  sampler.sample_adj(indptr, indices, node_idx, num_neighbors, replace)
 *             -> (out_indptr, out_indices, out_nodes, out_edges)   cogdl/operators/sample/sample.cpp:6-146
 *           sampler.subgraph(indptr, indices, node_idx)
 *             -> (out_indptr, out_indices, arange, out_edges)      cogdl/operators/sample/sample.cpp:148-188
 * (single-thread host loops called per mini-batch by Graph.sample_adj / csr_subgraph, data.py:792-874).
 *
 * Results are IDENTICAL to the reference's sequential loops for any emitted edge list: out_nodes =
 * batch nodes, then new source nodes in order of first appearance; out_indices = their positions.
 * Which edges are emitted: num_neighbors < 0 -> every edge of each batch row (CSR order);
 * replace != 0 -> num_neighbors draws with replacement per row (none for a degree-0 row; the
 * reference divides by zero there); else min(deg, num_neighbors) DISTINCT edges by Floyd's algorithm
 * in insertion order (the reference's variant draws `rand() % j` and is biased -- fixed, see
 * sampler.cu).  Every draw is cogdl_b200_sample_draw(seed, batch slot, draw index): a counter-based
 * generator, so batches are reproducible and bit-checkable (the reference uses unseeded libc rand()).
 *
 * Two phases because the output sizes are data dependent and the caller owns every buffer:
 *   1. ..._count : out_indptr[n_batch+1] (device).  Read out_indptr[n_batch] = n_edges on the host.
 *   2. ..._fill  : out_indices / out_edges [n_edges], out_nodes [capacity n_batch + n_edges],
 *                  *n_out_nodes_dev (device int64) = number of out_nodes entries written.
 * `assoc`: device int32[num_nodes] scratch, every entry COGDL_B200_SAMPLE_UNSEEN on entry; restored
 * before the fill returns (subgraph: marked by _count, restored by _fill -- always call both).
 * node_idx entries must be distinct.  n_batch + n_edges must fit int32.
 * `workspace`: >= cogdl_b200_sample_workspace_bytes(n_batch, n_edges) bytes (n_edges = 0 for _count).
 * ------------------------------------------------------------------------------------- */
#define COGDL_B200_SAMPLE_UNSEEN 0x7fffffff
COGDL_B200_API uint64_t cogdl_b200_sample_draw(uint64_t seed, int64_t slot, int64_t k);
COGDL_B200_API int64_t cogdl_b200_sample_workspace_bytes(int64_t n_batch, int64_t n_edges);
COGDL_B200_API int cogdl_b200_sample_adj_count(const int64_t *indptr, const int64_t *node_idx, int64_t n_batch,
                                int64_t num_neighbors, int32_t replace, int64_t *out_indptr, void *workspace,
                                int64_t workspace_bytes, cogdl_b200_stream_t stream);
COGDL_B200_API int cogdl_b200_sample_adj_fill(const int64_t *indptr, const int64_t *indices, const int64_t *node_idx,
                               int64_t n_batch, int64_t num_nodes, int64_t num_neighbors, int32_t replace,
                               uint64_t seed, const int64_t *out_indptr, int64_t n_edges, int32_t *assoc,
                               int64_t *out_indices, int64_t *out_edges, int64_t *out_nodes,
                               int64_t *n_out_nodes_dev, void *workspace, int64_t workspace_bytes,
                               cogdl_b200_stream_t stream);
COGDL_B200_API int cogdl_b200_subgraph_count(const int64_t *indptr, const int64_t *indices, const int64_t *node_idx,
                              int64_t n_sub, int32_t *assoc, int64_t *out_indptr, void *workspace,
                              int64_t workspace_bytes, cogdl_b200_stream_t stream);
COGDL_B200_API int cogdl_b200_subgraph_fill(const int64_t *indptr, const int64_t *indices, const int64_t *node_idx,
                             int64_t n_sub, int32_t *assoc, const int64_t *out_indptr, int64_t *out_indices,
                             int64_t *out_edges, cogdl_b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* COGDL_B200_H_ */
