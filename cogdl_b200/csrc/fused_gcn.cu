// fused_gcn.cu -- GCN layer with the dense transform fused onto the aggregated tile:
//
//     OUT[i,:] = act( (sum_p val[p] * X[col[p],:]) . W^T  +  (sum_p val[p]) * b )          i.e.  (A.X).W^T + (A.1) b^T
//
// which equals the reference's  A.(X.W^T + 1 b^T)  (cogdl/layers/gcn_layer.py:51-64: `support = self.linear(x);
// out = spmm(graph, support)`; bias INSIDE the aggregation, then activation) up to fp32 rounding -- the order of
// the two products is swapped so that the sparse gather runs on the raw features and the GEMM runs on the
// aggregated 128-row tile while it is still on chip.  Behind a flag in the Python layer because it reorders.
// This is the only tensor-core work on the hot path (north star: "tensor cores only where a dense
// feature x weight GEMM is fused onto the aggregated output").
//
// One persistent CTA per SM made of TWO independent 8-warp groups; each group owns a tile of 64 consecutive
// destination rows (its own operand buffer, TMEM accumulator, mbarrier and named barrier) and both share W^T.
// The groups drift apart by construction, so while one waits for its tensor-core products and writes its output
// the other keeps the SM's load pipeline busy (a single 128-row tile per SM left it idle ~40 % of the time: ncu).
//   1. AGGREGATE (the group's warps): the tile's rows are dealt to the warps in contiguous runs of ~equal COST
//      (rowptr of the tile staged one tile ahead); each warp streams the edges of its rows like the row-stream SpMM
//      (stream.cuh: 32-edge slabs, ballot row-end masks, U independent 512-byte gathers in flight, CSR order
//      with separate fp32 mul / add => the aggregated tile is bit-identical to the SpMM output).  At a row end
//      the fp32 row is NOT stored to global memory: it is split into three bf16 terms a = a1 + a2 + a3
//      (a1 = bf16(a), a2 = bf16(a - a1), a3 = bf16(a - a1 - a2): 24 mantissa bits) and written into shared memory
//      in the tcgen05 K-major SWIZZLE_128B operand layout.  Hub rows (degree > plan chunk) were aggregated
//      beforehand by the hub-chunk items of the row-stream kernel (deterministic in-order combine) and are
//      picked up from a scratch matrix.
//   2. MMA (one thread): D[64 x Fout] (fp32, TMEM) = sum over the six split products a_i . w_j with i + j <= 4
//      (a3.w1, a2.w2, a1.w3, a2.w1, a1.w2, a1.w1 -- small terms first), each 8 x tcgen05.mma.kind::f16 of
//      K = 16, W^T resident in shared memory (split the same way once per CTA).  Dropped terms are <= 2^-24
//      relative: measured error vs an fp64 product ~1e-7 of the row scale, i.e. better than an fp32 FFMA GEMM
//      (TF32 would give 1e-3, plain bf16 4e-3; tests/test_cpu_oracle_and_host.py emulates the split in numpy).
//   3. EPILOGUE (the group's warps): tcgen05.ld the accumulator (data path = row), add rowsum * bias,
//      ReLU, store.
// No reference counterpart as a kernel; callers cogdl/layers/gcn_layer.py:51-64 (and sage_layer.py:69-87 for the
// aggregate-then-linear order).  K (input width) must be 128, Fout <= 128.
#include "common.cuh"
#include "stream.cuh"

#include <cuda_bf16.h>

namespace cogdl_b200 {

int spmm_hub_rows_only(const int32_t *rowptr, const int32_t *colind, const float *val, const float *X, float *Y,
                       int64_t F, const cogdl_b200_hub_plan_t *plan, cudaStream_t stream);

namespace fg {

constexpr int TILE_M = 64;                            // rows per tile (UMMA M = 64); two tiles are in flight per CTA
constexpr int GROUPS = 2;                             // independent warp groups, one tile each, W^T shared
constexpr int GWARPS = 8;                             // warps per group
constexpr int WARPS = GROUPS * GWARPS;
constexpr int KDIM = 128;
constexpr int U = 8;                                  // gathers per batch; two batches in flight per warp (software pipelined):
                                                      // 16 warps x 16 x 512 B = 128 KB of feature rows in flight per SM
constexpr int SLAB_BYTES_A = TILE_M * 128;            // one K-slab (64 bf16 = 128 B per row) of a tile: 8 KB
constexpr int A_TILE_BYTES = 3 * 2 * SLAB_BYTES_A;    // 3 splits x 2 K-slabs = 48 KB per group
constexpr int W_MAX_BYTES = 3 * 2 * 128 * 128;        // 96 KB at Fout = 128
constexpr int OFF_W = GROUPS * A_TILE_BYTES;
constexpr int OFF_BIAS = OFF_W + W_MAX_BYTES;
constexpr int OFF_TMEM = OFF_BIAS + 128 * 4;
constexpr int OFF_GROUP = OFF_TMEM + 16;              // per group: rowsum[64] | tile_rp[68] | tile_cost[68] | mbarrier
constexpr int GROUP_BYTES = TILE_M * 4 + 2 * (TILE_M + 4) * 4 + 16;
constexpr int OFF_SLAB = OFF_GROUP + GROUPS * GROUP_BYTES;      // per warp: 32 x (column, value) + 32 x row
constexpr int SMEM_BYTES = OFF_SLAB + WARPS * 32 * 12 + 1024;   // + slack to align the base to 1024 B (SWIZZLE_128B atoms)

struct Params {
  const int *rowptr;
  const int *colind;
  const float *val;        // nullable
  const int *edge_row;     // plan->edge_row
  const float *X;          // [n_src, 128]
  const float *W;          // [Fout, 128]  (nn.Linear weight layout = K-major B operand)
  const float *bias;       // nullable [Fout]
  const float *hub_agg;    // [n_rows, 128]: rows with degree > chunk_edges hold (A.X)[row]
  const float *rowsum;     // nullable [n_rows]: (A.1)[row], cached by the caller for fixed edge weights
  float *out;              // [n_rows, Fout]
  int n_rows;
  int Fout;
  int Npad;                // Fout rounded up to 16 (UMMA N)
  int chunk_edges;         // 0: no hub rows anywhere
  int relu;
  int n_tiles;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int g) {     // named barrier of one warp group (ids 1, 2; 0 is __syncthreads)
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GWARPS * 32) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, bf16 inputs, fp32 accumulate, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
// all tcgen05 ops issued so far by this thread -> one arrival on the mbarrier when they have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Feature-row gathers bypass the L1: this kernel leaves the SM ~16 KB of L1 (212 KB of shared memory), and an L1-allocating
// load holds a line per 128 bytes in flight -- 128 lines cap the SM at ~32 row gathers in flight, whatever the warp count
// or unroll depth (measured: three different schedules all landed on the same time).  ld.global.cg is tracked outside the
// L1 data array.
__device__ __forceinline__ float4 fg_gather(const float4 *p) { return __ldcg(p); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1)
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B between 8-row groups) | [46,48) version = 1
//   [61,64) layout type = 2 (SWIZZLE_128B).  The slab base is 1024-byte aligned (base_offset = 0).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 [4,6) = 1, A = B = bf16 [7,10) = [10,13) = 1,
// both K-major (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

// byte offset of bf16 element (row r, k) inside one split of a [rows x 128] K-major SW128 operand whose K-slabs
// (64 elements = 128 B per row) are `slab_bytes` apart
__device__ __forceinline__ uint32_t sw128_offset(int r, int k, int slab_bytes) {
  const int slab = k >> 6, kk = k & 63;
  const int chunk = (kk >> 3) ^ (r & 7);
  return (uint32_t)(slab * slab_bytes + (r >> 3) * 1024 + (r & 7) * 128 + chunk * 16 + (kk & 7) * 2);
}

// a = a1 + a2 + a3 (bf16 each, round to nearest even; the residuals are exact in fp32)
__device__ __forceinline__ void split3(float a, __nv_bfloat16 &a1, __nv_bfloat16 &a2, __nv_bfloat16 &a3) {
  a1 = __float2bfloat16_rn(a);
  const float r1 = a - __bfloat162float(a1);
  a2 = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(a2);
  a3 = __float2bfloat16_rn(r2);
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 lo, __nv_bfloat16 hi) {
  return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

// write one aggregated row (lane owns columns 4*lane .. 4*lane+3) into the three split operands of a tile
__device__ __forceinline__ void store_row(unsigned char *tileA, int r, int lane, const float4 &acc) {
  __nv_bfloat16 s1[4], s2[4], s3[4];
  split3(acc.x, s1[0], s2[0], s3[0]);
  split3(acc.y, s1[1], s2[1], s3[1]);
  split3(acc.z, s1[2], s2[2], s3[2]);
  split3(acc.w, s1[3], s2[3], s3[3]);
  const uint32_t off = sw128_offset(r, 4 * lane, SLAB_BYTES_A);
  *reinterpret_cast<uint2 *>(tileA + 0 * 2 * SLAB_BYTES_A + off) = make_uint2(pack2(s1[0], s1[1]), pack2(s1[2], s1[3]));
  *reinterpret_cast<uint2 *>(tileA + 1 * 2 * SLAB_BYTES_A + off) = make_uint2(pack2(s2[0], s2[1]), pack2(s2[2], s2[3]));
  *reinterpret_cast<uint2 *>(tileA + 2 * 2 * SLAB_BYTES_A + off) = make_uint2(pack2(s3[0], s3[1]), pack2(s3[2], s3[3]));
}

// rowptr and cost prefix of one 64-row tile -> shared memory (one warp: lane L takes rows 2L, 2L+1).  cost(row) =
// 1 + (hub ? 4 : deg): hub rows are one 512-byte read here, their edges were consumed by the hub pre-pass
__device__ __forceinline__ void stage_tile(const Params &p, int row0, int lane, int *tile_rp, int *tile_cost) {
  int rp[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) rp[t] = __ldg(p.rowptr + min(row0 + 2 * lane + t, p.n_rows));
  int c[2], tot = 0;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int d = rp[t + 1] - rp[t];
    c[t] = (row0 + 2 * lane + t < p.n_rows) ? 1 + ((p.chunk_edges > 0 && d > p.chunk_edges) ? 4 : d) : 0;
    tot += c[t];
  }
  int incl = tot;
#pragma unroll
  for (int st = 1; st < 32; st <<= 1) {
    const int o = __shfl_up_sync(FULL, incl, st);
    if (lane >= st) incl += o;
  }
  int run = incl - tot;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    tile_rp[2 * lane + t] = rp[t];
    tile_cost[2 * lane + t] = run;
    run += c[t];
  }
  if (lane == 31) { tile_rp[TILE_M] = rp[2]; tile_cost[TILE_M] = run; }
}

__global__ void __launch_bounds__(WARPS * 32, 1) gcn_fused_kernel(const Params p) {
  extern __shared__ unsigned char fg_smem_raw[];
  // SWIZZLE_128B atoms need a 1024-byte aligned base
  unsigned char *smem = fg_smem_raw + ((1024u - (smem_u32(fg_smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = warp / GWARPS, gw = warp % GWARPS, gtid = tid - g * GWARPS * 32;   // my group, warp / thread inside it
  unsigned char *tileA = smem + g * A_TILE_BYTES;
  unsigned char *smemW = smem + OFF_W;
  float *bias_s = reinterpret_cast<float *>(smem + OFF_BIAS);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_TMEM);
  unsigned char *gbase = smem + OFF_GROUP + g * GROUP_BYTES;
  float *rowsum_s = reinterpret_cast<float *>(gbase);
  int *tile_rp = reinterpret_cast<int *>(gbase + TILE_M * 4);
  int *tile_cost = tile_rp + (TILE_M + 4);
  const uint32_t bar = smem_u32(gbase + TILE_M * 4 + 2 * (TILE_M + 4) * 4);
  int2 *s_cv = reinterpret_cast<int2 *>(smem + OFF_SLAB) + warp * 32;
  int *s_r = reinterpret_cast<int *>(smem + OFF_SLAB + WARPS * 32 * 8) + warp * 32;
  const int w_slab_bytes = p.Npad * 128;
  const int tile_stride = (int)gridDim.x * GROUPS;
  const int first_tile = (int)blockIdx.x * GROUPS + g;

  // ---- one-time setup: TMEM (one 128-column accumulator per group), barriers, W^T split into three bf16
  //      operands (SW128 K-major), bias
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (gtid == 32) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int idx = tid; idx < p.Npad * (KDIM / 4); idx += WARPS * 32) {     // 4 consecutive k per thread
    const int n = idx / (KDIM / 4), k = (idx - n * (KDIM / 4)) * 4;
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < p.Fout) w = __ldg(reinterpret_cast<const float4 *>(p.W + (int64_t)n * KDIM + k));
    __nv_bfloat16 s1[4], s2[4], s3[4];
    split3(w.x, s1[0], s2[0], s3[0]);
    split3(w.y, s1[1], s2[1], s3[1]);
    split3(w.z, s1[2], s2[2], s3[2]);
    split3(w.w, s1[3], s2[3], s3[3]);
    const uint32_t off = sw128_offset(n, k, w_slab_bytes);
    *reinterpret_cast<uint2 *>(smemW + 0 * 2 * w_slab_bytes + off) = make_uint2(pack2(s1[0], s1[1]), pack2(s1[2], s1[3]));
    *reinterpret_cast<uint2 *>(smemW + 1 * 2 * w_slab_bytes + off) = make_uint2(pack2(s2[0], s2[1]), pack2(s2[2], s2[3]));
    *reinterpret_cast<uint2 *>(smemW + 2 * 2 * w_slab_bytes + off) = make_uint2(pack2(s3[0], s3[1]), pack2(s3[2], s3[3]));
  }
  if (tid < 128) bias_s[tid] = (p.bias && tid < p.Fout) ? __ldg(p.bias + tid) : 0.f;
  if (gw == 1 && first_tile < p.n_tiles) stage_tile(p, first_tile * TILE_M, lane, tile_rp, tile_cost);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // M = 64 accumulators occupy data paths {0-15, 32-47, 64-79, 96-111} (cute tmem_frg, UMMA_1SM M_MMA = 64);
  // group 1 takes the next 128 columns
  const uint32_t tmem_acc = *tmem_slot + (uint32_t)(g * 128);
  const uint32_t idesc = make_idesc(p.Npad);
  const float4 *X4 = reinterpret_cast<const float4 *>(p.X);
  const float4 *H4 = reinterpret_cast<const float4 *>(p.hub_agg);
  uint32_t phase = 0;

  // ===== the two groups run this loop independently (named barriers): while one is in its MMA / epilogue, the other
  //       keeps the SM's memory pipeline busy with its gather
  for (int tile = first_tile; tile < p.n_tiles; tile += tile_stride) {
    const int row0 = tile * TILE_M;
    // =========================================================== 1. aggregate my rows into shared memory
    // rows are dealt to the group's 8 warps by COST (contiguous runs; hub rows count as 4 edges): on power-law
    // rows an equal-rows split leaves most warps idle while one finishes (binary search in the cost prefix)
    {
      const int rows_here = min(TILE_M, p.n_rows - row0);
      const int c_tot = tile_cost[rows_here];
      auto split = [&](int w) {          // first row whose cost prefix >= share w
        if (w <= 0) return 0;
        if (w >= GWARPS) return rows_here;
        const int target = (int)(((long long)c_tot * w) / GWARPS);
        int lo = 0, hi = rows_here;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (tile_cost[mid] < target) lo = mid + 1; else hi = mid;
        }
        return lo;
      };
      const int r_begin = row0 + split(gw), r_end = row0 + split(gw + 1);
      int next_row = r_begin;      // rows < next_row of my range have been written
      auto zero_rows = [&](int upto) {   // rows [next_row, upto): no edges -> zeros
        for (; next_row < upto; ++next_row) {
          store_row(tileA, next_row - row0, lane, make_float4(0.f, 0.f, 0.f, 0.f));
          if (lane == 0) rowsum_s[next_row - row0] = 0.f;
        }
      };
      int r = r_begin;
      while (r < r_end) {
        const int lb = tile_rp[r - row0], hb = tile_rp[r - row0 + 1];
        if (p.chunk_edges > 0 && hb - lb > p.chunk_edges) {
          // hub row: aggregated beforehand (hub chunks of the row-stream kernel)
          const float4 a = __ldg(H4 + (int64_t)r * (KDIM / 4) + lane);
          float s = 0.f;
          if (p.rowsum) {
            s = __ldg(p.rowsum + r);
          } else if (p.val) {      // no cached row sums: one warp sums the hub's weights, 8 loads in flight per lane
            float sa[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            int e = lb + lane;
            for (; e + 7 * 32 < hb; e += 256) {
#pragma unroll
              for (int t = 0; t < 8; ++t) sa[t] += ld_stream(p.val + e + 32 * t);
            }
            for (; e < hb; e += 32) sa[0] += ld_stream(p.val + e);
            s = ((sa[0] + sa[1]) + (sa[2] + sa[3])) + ((sa[4] + sa[5]) + (sa[6] + sa[7]));
            for (int st = 16; st > 0; st >>= 1) s += __shfl_xor_sync(FULL, s, st);
          } else {
            s = (float)(hb - lb);
          }
          store_row(tileA, r - row0, lane, a);
          if (lane == 0) rowsum_s[r - row0] = s;
          next_row = r + 1;
          ++r;
          continue;
        }
        // maximal run of non-hub rows [r, rb)
        int rb = r + 1;
        while (rb < r_end && !(p.chunk_edges > 0 && tile_rp[rb - row0 + 1] - tile_rp[rb - row0] > p.chunk_edges)) ++rb;
        const int e_begin = lb, e_end = tile_rp[rb - row0];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float rs = 0.f;
        auto flush = [&](int rj) {                       // last edge of row rj
          zero_rows(rj);
          store_row(tileA, rj - row0, lane, acc);
          if (lane == 0) rowsum_s[rj - row0] = rs;
          next_row = rj + 1;
          acc = make_float4(0.f, 0.f, 0.f, 0.f);
          rs = 0.f;
        };
        // slab registers (column, value, row, row of the next edge); the NEXT slab is fetched while this one is used
        int c = 0, rid = -1, rnx = -1;
        float v = 0.f;
        auto load_slab = [&](int e0, int &c_, float &v_, int &rid_, int &rnx_) {
          const int q = e0 + lane;
          c_ = 0; v_ = 0.f; rid_ = -1; rnx_ = -1;
          if (q < e_end) {
            c_ = ld_stream(p.colind + q);
            v_ = p.val ? ld_stream(p.val + q) : 1.f;
            rid_ = ld_stream(p.edge_row + q);
            if (q + 1 < e_end) rnx_ = __ldg(p.edge_row + q + 1);
          }
        };
        load_slab(e_begin, c, v, rid, rnx);
        for (int e = e_begin; e < e_end; e += 32) {
          const int cnt = min(32, e_end - e);
          const unsigned endmask = __ballot_sync(FULL, lane < cnt && rid != rnx);
          __syncwarp();                                  // previous slab fully consumed
          s_cv[lane] = make_int2(c, __float_as_int(v));  // slab parked in shared memory: broadcast LDS.64 instead of
          s_r[lane] = rid;                               // three shuffles + convergence checks per edge (stream.cuh)
          __syncwarp();
          int cn, ridn, rnxn;
          float vn;
          load_slab(e + 32, cn, vn, ridn, rnxn);
          // two batches of U gathers in flight: batch b+1 is issued before batch b is consumed
          float4 xa[U], xb[U];
          auto load_batch = [&](float4 (&x)[U], int j) {
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (j + u < cnt) x[u] = fg_gather(X4 + (int64_t)s_cv[j + u].x * (KDIM / 4) + lane);
          };
          auto consume = [&](const float4 (&x)[U], int j) {
            const unsigned em = (endmask >> j) & ((1u << U) - 1u);
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (j + u < cnt) {
                const float vj = __int_as_float(s_cv[j + u].y);
                if (p.val) axpy_rn(acc, vj, x[u]); else add_rn(acc, x[u]);
                rs += vj;
                if (em != 0u && ((em >> u) & 1u)) flush(s_r[j + u]);
              }
            }
          };
          load_batch(xa, 0);
#pragma unroll 1
          for (int j = 0; j < cnt; j += 2 * U) {
            if (j + U < cnt) load_batch(xb, j + U);
            consume(xa, j);
            if (j + 2 * U < cnt) load_batch(xa, j + 2 * U);
            if (j + U < cnt) consume(xb, j + U);
          }
          c = cn; v = vn; rid = ridn; rnx = rnxn;
        }
        zero_rows(rb);     // trailing empty rows of the run
        r = rb;
      }
      // rows of the tile beyond n_rows (last tile): zeros, dealt round-robin
      for (int rr = rows_here + gw; rr < TILE_M; rr += GWARPS) {
        store_row(tileA, rr, lane, make_float4(0.f, 0.f, 0.f, 0.f));
        if (lane == 0) rowsum_s[rr] = 0.f;
      }
    }
    fence_async_smem();          // generic-proxy writes of the tile -> visible to the tensor core (async proxy)
    tc_fence_before();
    group_sync(g);
    // =========================================================== 2. six split products into TMEM (one thread)
    if (gtid == 0) {
      tc_fence_after();
      const uint32_t a_base = smem_u32(tileA), w_base = smem_u32(smemW);
      // (i, j): split of A x split of W, smallest terms first
      const int ai[6] = {2, 1, 0, 1, 0, 0};
      const int wj[6] = {0, 1, 2, 0, 1, 0};
      uint32_t accumulate = 0;
#pragma unroll
      for (int t = 0; t < 6; ++t) {
#pragma unroll
        for (int ks = 0; ks < KDIM / 16; ++ks) {            // K = 16 bf16 = 32 bytes per instruction
          const uint32_t ka = a_base + ai[t] * 2 * SLAB_BYTES_A + (ks >> 2) * SLAB_BYTES_A + (ks & 3) * 32;
          const uint32_t kb = w_base + wj[t] * 2 * w_slab_bytes + (ks >> 2) * w_slab_bytes + (ks & 3) * 32;
          umma_bf16(tmem_acc, make_desc(ka), make_desc(kb), idesc, accumulate);
          accumulate = 1;
        }
      }
      umma_commit(bar);
    }
    // rowptr + cost prefix of my NEXT tile while the tensor core works (read again only after the barrier below)
    if (gw == 1 && tile + tile_stride < p.n_tiles) stage_tile(p, (tile + tile_stride) * TILE_M, lane, tile_rp, tile_cost);
    // =========================================================== 3. epilogue: TMEM -> registers -> global
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
    {
      // warp gw may read TMEM data paths [32 q, 32 q + 32), q = gw % 4; an M = 64 accumulator keeps rows
      // 16 q .. 16 q + 15 in the FIRST 16 of them (lanes 16..31 of the warp idle here); the group's two warps with the
      // same q split the 128 columns: four 32-column loads cover 64 columns each
      const int q = gw & 3;
      const int r = q * 16 + (lane & 15);
      const int grow = row0 + r;
      const bool live = (lane < 16) && grow < p.n_rows;
      const float rsum = rowsum_s[r];
      const bool vec = (p.Fout % 4 == 0);
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = (gw >> 2) * 64 + cc * 32;
        if (c0 < p.Npad) {                       // warp-uniform
          uint32_t v[32];
          tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
          if (live) {
            float *o = p.out + (int64_t)grow * p.Fout + c0;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              float4 y;
              y.x = __uint_as_float(v[c + 0]) + rsum * bias_s[c0 + c + 0];
              y.y = __uint_as_float(v[c + 1]) + rsum * bias_s[c0 + c + 1];
              y.z = __uint_as_float(v[c + 2]) + rsum * bias_s[c0 + c + 2];
              y.w = __uint_as_float(v[c + 3]) + rsum * bias_s[c0 + c + 3];
              if (p.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
              if (vec && c0 + c + 3 < p.Fout) {
                __stcs(reinterpret_cast<float4 *>(o + c), y);
              } else {
                if (c0 + c + 0 < p.Fout) o[c + 0] = y.x;
                if (c0 + c + 1 < p.Fout) o[c + 1] = y.y;
                if (c0 + c + 2 < p.Fout) o[c + 2] = y.z;
                if (c0 + c + 3 < p.Fout) o[c + 3] = y.w;
              }
            }
          }
        }
      }
    }
    tc_fence_before();
    group_sync(g);               // accumulator, operand tile and staged rowptr are free / ready for the next tile
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(256) : "memory");
}

}  // namespace fg
}  // namespace cogdl_b200

using namespace cogdl_b200;

extern "C" int cogdl_b200_gcn_fused_supported(int64_t K, int64_t Fout) { return (K == fg::KDIM && Fout >= 1 && Fout <= 128) ? 1 : 0; }

extern "C" int cogdl_b200_gcn_fused_f32(const int32_t *rowptr, const int32_t *colind, const float *val, const float *X,
                                        const float *W, const float *bias, const float *rowsum, float *out, float *hub_agg,
                                        int64_t n_rows,
                                        int64_t K, int64_t Fout, int32_t relu, const cogdl_b200_hub_plan_t *plan,
                                        cogdl_b200_stream_t stream) {
  const char *who = "cogdl_b200_gcn_fused_f32";
  CB_REQUIRE(n_rows >= 0 && K >= 0 && Fout >= 0, "%s: negative size", who);
  if (n_rows == 0 || Fout == 0) return COGDL_B200_OK;
  CB_REQUIRE(cogdl_b200_gcn_fused_supported(K, Fout), "%s: needs K == 128 and 1 <= Fout <= 128 (got K=%lld, Fout=%lld)", who,
             (long long)K, (long long)Fout);
  CB_REQUIRE(rowptr && colind && X && W && out, "%s: null pointer", who);
  CB_REQUIRE(n_rows < 0x7fffff00LL, "%s: n_rows must fit int32", who);
  CB_REQUIRE(plan && plan->chunk_edges > 0 && plan->edge_row, "%s: needs a hub plan with edge_row (row-stream form)", who);
  CB_REQUIRE(aligned16(X) && aligned16(W) && aligned16(out), "%s: X, W and out must be 16-byte aligned", who);
  int rc = check_plan(plan, (int64_t)plan->n_chunks * K * (int64_t)sizeof(float));
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (plan->n_chunks > 0) {
    CB_REQUIRE(hub_agg && aligned16(hub_agg), "%s: hub_agg scratch [n_rows, 128] is required when the plan has hub rows", who);
    rc = spmm_hub_rows_only(rowptr, colind, val, X, hub_agg, K, plan, s);
    if (rc) return rc;
  }
  fg::Params p;
  p.rowptr = rowptr; p.colind = colind; p.val = val; p.edge_row = plan->edge_row; p.X = X; p.W = W; p.bias = bias;
  p.hub_agg = hub_agg ? hub_agg : X; p.rowsum = rowsum; p.out = out; p.n_rows = (int)n_rows; p.Fout = (int)Fout;
  p.Npad = (int)((Fout + 15) / 16 * 16); p.chunk_edges = plan->chunk_edges; p.relu = relu;
  p.n_tiles = (int)ceil_div(n_rows, fg::TILE_M);
  static unsigned long long attr_done = 0;
  int dev = 0, n_sms = 148;
  if (first_use_on_device(attr_done, &dev))
    CB_CUDA(cudaFuncSetAttribute(fg::gcn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fg::SMEM_BYTES));
  CB_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  const int want = (p.n_tiles + fg::GROUPS - 1) / fg::GROUPS;
  const int grid = want < n_sms ? want : n_sms;
  note_kernel("cogdl_b200::fg::gcn_fused_kernel<2 groups x tile 64x%d, K=128, bf16x3 split, tcgen05.mma kind::f16>", p.Npad);
  fg::gcn_fused_kernel<<<grid, fg::WARPS * 32, fg::SMEM_BYTES, s>>>(p);
  CB_LAUNCH_CHECK();
  return COGDL_B200_OK;
}
