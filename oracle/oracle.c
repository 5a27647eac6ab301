/*
 * oracle.c -- CPU restatement of CogDL's sparse message-passing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (cogdl_b200/) may
 * import, link or call this file.  Allowed callers: tests/, __graft_entry__.smoke()
 * and the cpu_baseline / --impl reference legs of bench.py.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the CogDL tree, THUDM/CogDL @ 281f4742).  Arithmetic notes:
 *   - oracle_spmm_csr_f32 keeps the reference CPU operation order exactly
 *     (CSR order inside a row, one fp32 multiply then one fp32 add, no FMA):
 *     build with -ffp-contract=off.  It is pinned bit-for-bit against the
 *     reference's own spmm_cpu.cpp compiled from /root/reference (oracle/_ref)
 *     and against golden vectors produced by importing the reference package.
 *   - ops that exist only as CUDA in the reference (edge_softmax, mhspmm,
 *     sddmm, mhsddmm, scatter_max) are restated from the kernel source; the
 *     fp32 summation order of a GPU butterfly is not part of their contract,
 *     so they accumulate in fp64 and round once ("tightest" answer).  Their
 *     pinning: the reference's python CPU fallbacks (golden vectors, logits<=10)
 *     and the reference's own CUDA kernels compiled for sm_100a (oracle/_ref,
 *     run only on the GPU box by the -m gpu tests).
 *   - indices are int32 as in the reference ABI; offsets are widened to 64 bit
 *     (the reference overflows int32 at N*F >= 2^31 -- documented divergence).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORACLE_API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------
 * Weighted / unweighted CSR SpMM:  Y[i,:] = sum_p val[p] * X[col[p],:]
 * Restates cogdl/operators/spmm/spmm_cpu.cpp:24-36 (spmm_cpu): rows in
 * parallel (OpenMP dynamic), edges of a row in CSR order, `out += val*dense`
 * as a separate multiply and add.  val == NULL restates the unweighted
 * GE-SpMM (spmm_kernel.cu:7-153, csr_spmm_no_edge_value) as val = 1.
 * ---------------------------------------------------------------------- */
ORACLE_API void oracle_spmm_csr_f32(const int32_t *rowptr, const int32_t *colind,
                                    const float *val, const float *X, float *Y,
                                    int64_t n_rows, int64_t F) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n_rows; ++i) {
    float *y = Y + i * F;
    for (int64_t t = 0; t < F; ++t) y[t] = 0.0f;
    for (int32_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
      const float *x = X + (int64_t)colind[p] * F;
      const float v = val ? val[p] : 1.0f;
      for (int64_t t = 0; t < F; ++t) {
        float prod = v * x[t];
        y[t] = y[t] + prod;
      }
    }
  }
}

/* ------------------------------------------------------------------------
 * CSR SDDMM: out[e] = < D1[row(e),:], D2[col[e],:] >
 * Restates cogdl/operators/spmm/sddmm_kernel.cu:249-417 (sddmmCSR{2,1}Scale);
 * row(e) is the CSR row owning edge e (the kernel recovers it with findRow,
 * computeUtil.h:36-53).  Used for grad_edge_weight in spmm.py:70.
 * ---------------------------------------------------------------------- */
ORACLE_API void oracle_sddmm_csr_f32(const int32_t *rowptr, const int32_t *colind,
                                     const float *D1, const float *D2, float *out,
                                     int64_t n_rows, int64_t F) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n_rows; ++i) {
    const float *a = D1 + i * F;
    for (int32_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
      const float *b = D2 + (int64_t)colind[p] * F;
      double acc = 0.0;
      for (int64_t t = 0; t < F; ++t) acc += (double)a[t] * (double)b[t];
      out[p] = (float)acc;
    }
  }
}

/* ------------------------------------------------------------------------
 * Edge softmax over the edges of each destination row, per head.
 * Restates cogdl/operators/edge_softmax/edge_softmax.cu:7-60: m = max_p v,
 * s = sum_p exp(v - m), out = exp(v - m) / s.  Rows of degree 0 write nothing.
 * (The reference seeds the max with -1e8; we use the true max, identical for
 * any finite logit > -1e8.)
 * ---------------------------------------------------------------------- */
ORACLE_API void oracle_edge_softmax_fwd_f32(const int32_t *rowptr, const float *in,
                                            float *out, int64_t n_rows, int64_t H) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n_rows; ++i) {
    const int32_t lb = rowptr[i], hb = rowptr[i + 1];
    if (hb <= lb) continue;
    for (int64_t h = 0; h < H; ++h) {
      float m = in[(int64_t)lb * H + h];
      for (int32_t p = lb + 1; p < hb; ++p) {
        float v = in[(int64_t)p * H + h];
        if (v > m) m = v;
      }
      double s = 0.0;
      for (int32_t p = lb; p < hb; ++p) s += exp((double)(in[(int64_t)p * H + h] - m));
      for (int32_t p = lb; p < hb; ++p)
        out[(int64_t)p * H + h] = (float)(exp((double)(in[(int64_t)p * H + h] - m)) / s);
    }
  }
}

/* Restates edge_softmax.cu:63-98: g_in = y * (g - sum_row y*g). */
ORACLE_API void oracle_edge_softmax_bwd_f32(const int32_t *rowptr, const float *y,
                                            const float *g, float *gin, int64_t n_rows,
                                            int64_t H) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n_rows; ++i) {
    const int32_t lb = rowptr[i], hb = rowptr[i + 1];
    for (int64_t h = 0; h < H; ++h) {
      double s = 0.0;
      for (int32_t p = lb; p < hb; ++p)
        s += (double)y[(int64_t)p * H + h] * (double)g[(int64_t)p * H + h];
      for (int32_t p = lb; p < hb; ++p) {
        int64_t k = (int64_t)p * H + h;
        gin[k] = (float)((double)y[k] * ((double)g[k] - s));
      }
    }
  }
}

/* ------------------------------------------------------------------------
 * Multi-head SpMM: out[i,h,:] = sum_p att[p,h] * feat[col[p],h,:]
 * Restates cogdl/operators/spmm/multiheadSpmm.cu:6-51 (mhspmmSimple/mhspmm_1)
 * in the operation order of the reference's CPU fallback (spmm_utils.py:216-225:
 * one spmm_cpu per head => CSR order, fp32 multiply then add).  `perm` (nullable)
 * redirects the attention lookup, att[perm[p],h]: that is the fused form of the
 * backward's mhtranspose (mhTranspose.cu:6-28) followed by mhspmm on the CSC.
 * ---------------------------------------------------------------------- */
ORACLE_API void oracle_mhspmm_f32(const int32_t *rowptr, const int32_t *colind,
                                  const int32_t *perm, const float *att, const float *feat,
                                  float *out, int64_t n_rows, int64_t H, int64_t F) {
  const int64_t HF = H * F;
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n_rows; ++i) {
    float *y = out + i * HF;
    for (int64_t t = 0; t < HF; ++t) y[t] = 0.0f;
    for (int32_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
      const float *x = feat + (int64_t)colind[p] * HF;
      const float *a = att + (int64_t)(perm ? perm[p] : p) * H;
      for (int64_t h = 0; h < H; ++h) {
        const float v = a[h];
        for (int64_t f = 0; f < F; ++f) {
          float prod = v * x[h * F + f];
          y[h * F + f] = y[h * F + f] + prod;
        }
      }
    }
  }
}

/* Multi-head SDDMM: out[e,h] = < grad[row(e),h,:], feat[col[e],h,:] >
 * Restates cogdl/operators/spmm/multiheadSddmm.cu:6-93. */
ORACLE_API void oracle_mhsddmm_f32(const int32_t *rowptr, const int32_t *colind,
                                   const float *grad, const float *feat, float *out,
                                   int64_t n_rows, int64_t H, int64_t F) {
  const int64_t HF = H * F;
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n_rows; ++i) {
    const float *a = grad + i * HF;
    for (int32_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
      const float *b = feat + (int64_t)colind[p] * HF;
      for (int64_t h = 0; h < H; ++h) {
        double acc = 0.0;
        for (int64_t f = 0; f < F; ++f) acc += (double)a[h * F + f] * (double)b[h * F + f];
        out[(int64_t)p * H + h] = (float)acc;
      }
    }
  }
}

/* out[e,:] = in[perm[e],:]  -- restates mhTranspose.cu:6-28 (mhtranspose). */
ORACLE_API void oracle_gather_rows_f32(const int32_t *perm, const float *in, float *out,
                                       int64_t nnz, int64_t H) {
#pragma omp parallel for
  for (int64_t e = 0; e < nnz; ++e)
    for (int64_t h = 0; h < H; ++h) out[e * H + h] = in[(int64_t)perm[e] * H + h];
}

/* ------------------------------------------------------------------------
 * scatter_max forward: out[i,f] = max_p X[col[p],f], argmax[i,f] = the col of
 * the FIRST edge (CSR order) attaining it (strict `<` update).
 * Restates cogdl/operators/scatter_max/scatter_max.cu:5-28.
 *   reference_semantics != 0 reproduces the reference literally: the running
 *     max starts at FLT_MIN (smallest positive normal, scatter_max.cu:16), so
 *     the result is max(FLT_MIN, ...) and argmax is left unset (we report -1)
 *     when no neighbour exceeds FLT_MIN.
 *   reference_semantics == 0 is the fixed semantics the CUDA path implements:
 *     start at -inf.  Both agree whenever every row has a neighbour value
 *     > FLT_MIN (e.g. strictly positive features).
 * Degree-0 rows give out = 0, argmax = -1 in both modes.
 * ---------------------------------------------------------------------- */
ORACLE_API void oracle_scatter_max_fwd_f32(const int32_t *rowptr, const int32_t *colind,
                                           const float *X, float *out, int32_t *argmax,
                                           int64_t n_rows, int64_t F, int reference_semantics) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n_rows; ++i) {
    const int32_t lb = rowptr[i], hb = rowptr[i + 1];
    for (int64_t f = 0; f < F; ++f) {
      float acc = (hb > lb) ? (reference_semantics ? FLT_MIN : -INFINITY) : 0.0f;
      int32_t id = -1;
      for (int32_t p = lb; p < hb; ++p) {
        const int32_t c = colind[p];
        const float x = X[(int64_t)c * F + f];
        if (acc < x) {
          acc = x;
          id = c;
        }
      }
      out[i * F + f] = acc;
      argmax[i * F + f] = id;
    }
  }
}

/* scatter_max backward: gx[argmax[i,f], f] += g[i,f]; gx zero-initialised here
 * (the reference forgets to, scatter_max.cu:70 -- documented divergence).
 * Restates scatter_max.cu:30-42.  Sequential => deterministic fp32 order
 * (rows ascending); the GPU uses atomics, compared with a tolerance. */
ORACLE_API void oracle_scatter_max_bwd_f32(const float *g, const int32_t *argmax, float *gx,
                                           int64_t n_rows, int64_t n_src, int64_t F) {
  memset(gx, 0, (size_t)(n_src * F) * sizeof(float));
  for (int64_t i = 0; i < n_rows; ++i)
    for (int64_t f = 0; f < F; ++f) {
      const int32_t id = argmax[i * F + f];
      if (id >= 0) gx[(int64_t)id * F + f] += g[i * F + f];
    }
}

/* ------------------------------------------------------------------------
 * CSR -> CSC with the edge permutation (stable: inside a column, entries keep
 * CSR order, i.e. ascending row).  The reference gets this from cuSPARSE
 * cusparseCsr2cscEx2 ALG1 (spmm_kernel.cu:514-532, 596-613) which is stable as
 * well; integer-only, no arithmetic.  perm[q] = CSR position of CSC entry q.
 * ---------------------------------------------------------------------- */
ORACLE_API void oracle_csr2csc(const int32_t *rowptr, const int32_t *colind, int32_t *colptr,
                               int32_t *rowind, int32_t *perm, int64_t n_rows, int64_t n_cols) {
  const int64_t nnz = rowptr[n_rows];
  for (int64_t c = 0; c <= n_cols; ++c) colptr[c] = 0;
  for (int64_t e = 0; e < nnz; ++e) colptr[colind[e] + 1]++;
  for (int64_t c = 0; c < n_cols; ++c) colptr[c + 1] += colptr[c];
  int32_t *cursor = (int32_t *)malloc((size_t)(n_cols + 1) * sizeof(int32_t));
  memcpy(cursor, colptr, (size_t)(n_cols + 1) * sizeof(int32_t));
  for (int64_t i = 0; i < n_rows; ++i)
    for (int32_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
      const int32_t q = cursor[colind[e]]++;
      rowind[q] = (int32_t)i;
      perm[q] = e;
    }
  free(cursor);
}

/* ------------------------------------------------------------------------
 * COO -> CSR row pointer + stable permutation (int64, as Graph stores them).
 * Restates cogdl/operators/sample/sample.cpp:234-270 (coo2csr_cpu_index):
 * counting sort by row; reindex[q] = COO position of CSR entry q, original
 * order kept inside a row.  This defines the edge order every edge-aligned
 * tensor ([E], [E,H]) lives in.
 * ---------------------------------------------------------------------- */
ORACLE_API void oracle_coo2csr_index(const int64_t *row, int64_t nnz, int64_t num_nodes,
                                     int64_t *row_ptr, int64_t *reindex) {
  for (int64_t i = 0; i <= num_nodes; ++i) row_ptr[i] = 0;
  for (int64_t e = 0; e < nnz; ++e) row_ptr[row[e] + 1]++;
  for (int64_t i = 0; i < num_nodes; ++i) row_ptr[i + 1] += row_ptr[i];
  int64_t *cursor = (int64_t *)malloc((size_t)(num_nodes + 1) * sizeof(int64_t));
  memcpy(cursor, row_ptr, (size_t)(num_nodes + 1) * sizeof(int64_t));
  for (int64_t e = 0; e < nnz; ++e) reindex[cursor[row[e]]++] = e;
  free(cursor);
}

/* ------------------------------------------------------------------------
 * Fused GAT forward (the "next" row), unfused definition from
 * cogdl/layers/gat_layer.py:63-77 with attn_drop = 0:
 *   e[p,h]   = leakyrelu(h_l[row(p),h] + h_r[col[p],h], slope)
 *   a[p,:]   = edge_softmax(e)                    (edge_softmax.cu:7-60)
 *   out[i,h] = sum_p a[p,h] * feat[col[p],h,:]    (multiheadSpmm.cu:6-27)
 * fp64 inside, rounded once.
 * ---------------------------------------------------------------------- */
ORACLE_API void oracle_gat_fwd_f32(const int32_t *rowptr, const int32_t *colind,
                                   const float *h_l, const float *h_r, const float *feat,
                                   float slope, float *out, float *att_out /* nullable */,
                                   int64_t n_rows, int64_t H, int64_t F) {
  const int64_t HF = H * F;
#pragma omp parallel
  {
    double *acc = (double *)malloc((size_t)F * sizeof(double));
#pragma omp for schedule(dynamic, 64)
    for (int64_t i = 0; i < n_rows; ++i) {
      const int32_t lb = rowptr[i], hb = rowptr[i + 1];
      for (int64_t h = 0; h < H; ++h) {
        for (int64_t f = 0; f < F; ++f) acc[f] = 0.0;
        if (hb > lb) {
          float m = -INFINITY;
          for (int32_t p = lb; p < hb; ++p) {
            float e = h_l[i * H + h] + h_r[(int64_t)colind[p] * H + h];
            e = e > 0.0f ? e : e * slope;
            if (e > m) m = e;
          }
          double s = 0.0;
          for (int32_t p = lb; p < hb; ++p) {
            float e = h_l[i * H + h] + h_r[(int64_t)colind[p] * H + h];
            e = e > 0.0f ? e : e * slope;
            s += exp((double)(e - m));
          }
          for (int32_t p = lb; p < hb; ++p) {
            float e = h_l[i * H + h] + h_r[(int64_t)colind[p] * H + h];
            e = e > 0.0f ? e : e * slope;
            const double a = exp((double)(e - m)) / s;
            if (att_out) att_out[(int64_t)p * H + h] = (float)a;
            const float *x = feat + (int64_t)colind[p] * HF + h * F;
            for (int64_t f = 0; f < F; ++f) acc[f] += a * (double)x[f];
          }
        }
        for (int64_t f = 0; f < F; ++f) out[i * HF + h * F + f] = (float)acc[f];
      }
    }
    free(acc);
  }
}

/* ------------------------------------------------------------------------
 * Neighbour sampling and induced subgraphs.
 * Restates cogdl/operators/sample/sample.cpp:6-146 (sample_adj) and :148-188
 * (subgraph_cpu) as the sequential host loops they are:
 *   - out_nodes = batch nodes, then every new source node in order of first
 *     appearance (the running `num_nodes` counter, sample.cpp:44-47 etc.);
 *   - out_indices[q] = id of edge q's source in out_nodes; out_edges[q] = the
 *     global CSR position of the edge; out_indptr = running edge count.
 * Randomness: the reference uses libc rand() (unseeded, not reproducible).  The
 * oracle and the CUDA path share a COUNTER-BASED generator instead: draw k of
 * batch slot i = mix64(mix64(seed + i*0xD1342543DE82EF95) + k), mix64 = the
 * splitmix64 finaliser.  `floyd_variant` selects the without-replacement rule:
 *   0: Robert Floyd's algorithm, t ~ U{0..j}   (what the CUDA path implements)
 *   1: the reference's loop literally, t = draw % j  (sample.cpp:101-104), kept so
 *      tests can show its bias (degree 2, k = 1 always yields the first edge).
 * Edges of a row are emitted in insertion order (the reference iterates a
 * std::unordered_set: unspecified order).  replace on a degree-0 row emits
 * nothing (the reference would divide by zero).
 * Returns the number of out_nodes; out_nodes needs room for n_batch + n_edges.
 * ---------------------------------------------------------------------- */
static uint64_t oracle_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
ORACLE_API uint64_t oracle_sample_draw(uint64_t seed, int64_t slot, int64_t k) {
  return oracle_mix64(oracle_mix64(seed + (uint64_t)slot * 0xD1342543DE82EF95ull) + (uint64_t)k);
}

ORACLE_API int64_t oracle_sample_adj(const int64_t *indptr, const int64_t *indices, const int64_t *node_idx,
                                     int64_t n_batch, int64_t num_nodes, int64_t k, int replace, uint64_t seed,
                                     int floyd_variant, int64_t *out_indptr, int64_t *out_indices,
                                     int64_t *out_edges, int64_t *out_nodes, int64_t edge_capacity) {
  int64_t *assoc = (int64_t *)malloc((size_t)(num_nodes > 0 ? num_nodes : 1) * sizeof(int64_t));
  for (int64_t v = 0; v < num_nodes; ++v) assoc[v] = -1;
  for (int64_t i = 0; i < n_batch; ++i) {
    assoc[node_idx[i]] = i;
    out_nodes[i] = node_idx[i];
  }
  int64_t n_nodes = n_batch, n_edges = 0;
  out_indptr[0] = 0;
  for (int64_t i = 0; i < n_batch; ++i) {
    const int64_t node = node_idx[i];
    const int64_t rs = indptr[node], deg = indptr[node + 1] - rs;
    int64_t cnt;
    if (k < 0) cnt = deg;
    else if (replace) cnt = deg > 0 ? k : 0;
    else cnt = deg < k ? deg : k;
    if (n_edges + cnt > edge_capacity) { free(assoc); return -1; }
    int64_t *picks = out_edges + n_edges;
    if (k < 0 || (!replace && deg <= k)) {
      for (int64_t t = 0; t < cnt; ++t) picks[t] = rs + t;
    } else if (replace) {
      for (int64_t t = 0; t < cnt; ++t) picks[t] = rs + (int64_t)(oracle_sample_draw(seed, i, t) % (uint64_t)deg);
    } else {
      int64_t have = 0;
      for (int64_t j = deg - k; j < deg; ++j, ++have) {
        const uint64_t r = oracle_sample_draw(seed, i, have);
        const int64_t t = floyd_variant ? (j > 0 ? (int64_t)(r % (uint64_t)j) : 0) : (int64_t)(r % (uint64_t)(j + 1));
        int taken = 0;
        for (int64_t q = 0; q < have; ++q) taken |= (picks[q] == rs + t);
        picks[have] = rs + (taken ? j : t);
      }
    }
    for (int64_t t = 0; t < cnt; ++t) {
      const int64_t src = indices[picks[t]];
      if (assoc[src] == -1) {
        assoc[src] = n_nodes;
        out_nodes[n_nodes++] = src;
      }
      out_indices[n_edges + t] = assoc[src];
    }
    n_edges += cnt;
    out_indptr[i + 1] = n_edges;
  }
  free(assoc);
  return n_nodes;
}

/* Restates sample.cpp:148-188 (subgraph_cpu): edges between listed nodes, relabelled to list positions,
 * CSR order kept.  Returns the number of edges; outputs need room for the rows' total degree. */
ORACLE_API int64_t oracle_subgraph(const int64_t *indptr, const int64_t *indices, const int64_t *node_idx,
                                   int64_t n_sub, int64_t num_nodes, int64_t *out_indptr, int64_t *out_indices,
                                   int64_t *out_edges) {
  int64_t *assoc = (int64_t *)malloc((size_t)(num_nodes > 0 ? num_nodes : 1) * sizeof(int64_t));
  for (int64_t v = 0; v < num_nodes; ++v) assoc[v] = -1;
  for (int64_t i = 0; i < n_sub; ++i) assoc[node_idx[i]] = i;
  int64_t n_edges = 0;
  out_indptr[0] = 0;
  for (int64_t i = 0; i < n_sub; ++i) {
    const int64_t node = node_idx[i];
    for (int64_t e = indptr[node]; e < indptr[node + 1]; ++e) {
      const int64_t a = assoc[indices[e]];
      if (a > -1) {
        out_indices[n_edges] = a;
        out_edges[n_edges] = e;
        ++n_edges;
      }
    }
    out_indptr[i + 1] = n_edges;
  }
  free(assoc);
  return n_edges;
}

/* ------------------------------------------------------------------------
 * GAT attention backward (the autograd of cogdl/layers/gat_layer.py:73-74 written out; the reference
 * gets it from PyTorch autograd over h_l[row] + h_r[col] -> LeakyReLU -> edge_softmax):
 *   s[i,h]      = sum_{p in row i} att[p,h] * d_att[p,h]          (edge_softmax.cu:63-98)
 *   d_edge[p,h] = att[p,h] * (d_att[p,h] - s) * (z > 0 ? 1 : slope),  z = h_l[row,h] + h_r[col,h]
 *   g_row[i,h]  = sum_{p in row i} d_edge[p,h];   g_col[j,h] = sum_{p: col[p] = j} d_edge[p,h]
 * fp64 inside, rounded once.
 * ---------------------------------------------------------------------- */
ORACLE_API void oracle_gat_attn_bwd_f32(const int32_t *rowptr, const int32_t *colind, const float *att,
                                        const float *d_att, const float *h_l, const float *h_r, float slope,
                                        float *d_edge, float *g_row, float *g_col, int64_t n_rows,
                                        int64_t n_cols, int64_t H) {
  double *gc = (double *)calloc((size_t)(n_cols * H > 0 ? n_cols * H : 1), sizeof(double));
  for (int64_t i = 0; i < n_rows; ++i) {
    const int32_t lb = rowptr[i], hb = rowptr[i + 1];
    for (int64_t h = 0; h < H; ++h) {
      double s = 0.0;
      for (int32_t p = lb; p < hb; ++p) s += (double)att[(int64_t)p * H + h] * (double)d_att[(int64_t)p * H + h];
      double gr = 0.0;
      for (int32_t p = lb; p < hb; ++p) {
        const int64_t k = (int64_t)p * H + h;
        const float z = h_l[i * H + h] + h_r[(int64_t)colind[p] * H + h];
        const double v = (double)att[k] * ((double)d_att[k] - s) * (z > 0.0f ? 1.0 : (double)slope);
        d_edge[k] = (float)v;
        gr += v;
        gc[(int64_t)colind[p] * H + h] += v;
      }
      g_row[i * H + h] = (float)gr;
    }
  }
  for (int64_t t = 0; t < n_cols * H; ++t) g_col[t] = (float)gc[t];
  free(gc);
}
