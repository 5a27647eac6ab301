// common.cuh -- shared device/host helpers for libcogdl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cogdl_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcogdl_b200 is written for sm_100a (B200) only"
#endif

namespace cogdl_b200 {

// ------------------------------------------------------------------ host side
int set_error(int code, const char *fmt, ...);
void count_launch(int n = 1);
// Records which kernel instantiation the last dispatching entry point on this thread chose
// (cogdl_b200_last_kernel(); bench.py checks it against the profiled kernel's name).
void note_kernel(const char *fmt, ...);
void append_kernel_note(const char *fmt, ...);   // launch-shape suffix after the instantiation name
// Experiment knobs read from the environment (COGDL_B200_*), cached after the first read;
// cogdl_b200_reload_tuning() drops the cache so a tuning sweep can change them inside one process.
int tuning(const char *env_name, int dflt);

#define CB_REQUIRE(cond, ...)                                             \
  do {                                                                    \
    if (!(cond)) return ::cogdl_b200::set_error(COGDL_B200_EINVAL, __VA_ARGS__); \
  } while (0)

#define CB_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::cogdl_b200::set_error(_e == cudaErrorNoKernelImageForDevice ? COGDL_B200_EDEVICE \
                                                                           : COGDL_B200_ECUDA, \
                                     "%s: %s", #expr, cudaGetErrorString(_e));              \
  } while (0)

// Every launch goes through this: counts it and turns a launch failure into a status code.
#define CB_LAUNCH_CHECK()                  \
  do {                                     \
    ::cogdl_b200::count_launch();          \
    CB_CUDA(cudaPeekAtLastError());        \
  } while (0)

// cudaFuncSetAttribute is per device: a once-flag per (kernel, device), device < 64.  Racing first calls from two
// host threads set the attribute twice, which is harmless.  Returns the current device in *dev.
static inline bool first_use_on_device(unsigned long long &mask, int *dev) {
  *dev = 0;
  if (cudaGetDevice(dev) != cudaSuccess || *dev < 0 || *dev >= 64) return true;
  const unsigned long long bit = 1ull << *dev;
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Plain-old-data view of the hub plan handed to kernels.
struct HubView {
  int chunk_edges;       // 0 => no plan: rows are never skipped / split
  int n_chunks;
  const int2 *chunks;    // (row, first_slot)
  int *counters;
  void *partials;
  int n_segs;            // 0 => no row-stream segments
  const int2 *segs;      // (row_begin, row_end), hub-free
  const int *edge_row;
  int n_empty_rows;
};

static inline HubView hub_view(const cogdl_b200_hub_plan_t *plan) {
  HubView h{0, 0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0};
  if (plan && plan->chunk_edges > 0) {
    h.chunk_edges = plan->chunk_edges;
    h.n_chunks = plan->n_chunks;
    h.chunks = reinterpret_cast<const int2 *>(plan->chunks);
    h.counters = plan->counters;
    h.partials = plan->partials;
    if (plan->segs && plan->edge_row && plan->n_segs > 0) {
      h.n_segs = plan->n_segs;
      h.segs = reinterpret_cast<const int2 *>(plan->segs);
      h.edge_row = plan->edge_row;
      h.n_empty_rows = plan->n_empty_rows;
    }
  }
  return h;
}

int check_plan(const cogdl_b200_hub_plan_t *plan, int64_t need_partial_bytes);

// ------------------------------------------------------------------ device side
#ifdef __CUDACC__

constexpr unsigned FULL = 0xffffffffu;

// Streaming (read-once) index/value loads: keep them out of L1 so the gathered feature rows
// own it.  ld.global.nc.L1::no_allocate.
__device__ __forceinline__ int ld_stream(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// Gathered feature rows: read-only path, default caching (hub columns do get reused).
__device__ __forceinline__ float4 ld_gather(const float4 *p) { return __ldg(p); }
__device__ __forceinline__ float ld_gather(const float *p) { return __ldg(p); }

// L2 eviction-priority hints (createpolicy + ld/st .L2::cache_hint): the gathered matrix X is the
// only data with reuse, everything else is touched once.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ld_gather_hint(const float4 *p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ld_gather_hint(const float *p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}

// Output rows are written once and never re-read by the kernel: streaming store.
__device__ __forceinline__ void st_stream(float4 *p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(int *p, int v) { __stcs(p, v); }

// L2-coherent accesses for inter-CTA partial results (L1 is not coherent across SMs).
__device__ __forceinline__ float4 ld_cg(const float4 *p) { return __ldcg(p); }
__device__ __forceinline__ float ld_cg(const float *p) { return __ldcg(p); }
__device__ __forceinline__ int ld_cg(const int *p) { return __ldcg(p); }
__device__ __forceinline__ void st_cg(float4 *p, float4 v) { __stcg(p, v); }
__device__ __forceinline__ void st_cg(float *p, float v) { __stcg(p, v); }
__device__ __forceinline__ void st_cg(int *p, int v) { __stcg(p, v); }

// acc += v*x with a separate multiply and add (no FMA contraction): this is what makes the
// SpMM bit-identical to the reference CPU loop `out[ik+t] += val * dense[j+t]`
// (cogdl/operators/spmm/spmm_cpu.cpp:31-33, built without -O => no contraction).
__device__ __forceinline__ float mul_add_rn(float acc, float v, float x) {
  return __fadd_rn(acc, __fmul_rn(v, x));
}
// in-place forms, one per vector type
__device__ __forceinline__ void axpy_rn(float &acc, float v, const float &x) { acc = mul_add_rn(acc, v, x); }
__device__ __forceinline__ void axpy_rn(float4 &acc, float v, const float4 &x) {
  acc.x = mul_add_rn(acc.x, v, x.x);
  acc.y = mul_add_rn(acc.y, v, x.y);
  acc.z = mul_add_rn(acc.z, v, x.z);
  acc.w = mul_add_rn(acc.w, v, x.w);
}
__device__ __forceinline__ void add_rn(float4 &acc, const float4 &x) {
  acc.x = __fadd_rn(acc.x, x.x);
  acc.y = __fadd_rn(acc.y, x.y);
  acc.z = __fadd_rn(acc.z, x.z);
  acc.w = __fadd_rn(acc.w, x.w);
}

template <int WIDTH>
__device__ __forceinline__ int group_max(int v) {
#pragma unroll
  for (int s = WIDTH / 2; s > 0; s >>= 1) v = max(v, __shfl_xor_sync(FULL, v, s));
  return v;
}
__device__ __forceinline__ int warp_max(int v) { return group_max<32>(v); }

// What a group of lanes works on: either one whole row, or one fixed-size edge chunk of a hub
// row.  Items [0, n_chunks) are hub chunks (scheduled first: longest work first), items
// [n_chunks, n_chunks + n_rows) are rows; hub rows are skipped in the row range.
struct WorkItem {
  int row;        // destination row
  int lb, hb;     // edge range
  int slot;       // chunk slot (valid if is_chunk)
  int first;      // first slot of this row's chunks
  int n_row_chunks;
  bool active;
  bool is_chunk;
};

__device__ __forceinline__ WorkItem decode_item(int64_t item, int64_t n_rows, const int *__restrict__ rowptr,
                                                const HubView &hub) {
  WorkItem w;
  w.row = 0; w.lb = 0; w.hb = 0; w.slot = 0; w.first = 0; w.n_row_chunks = 0;
  w.active = false; w.is_chunk = false;
  if (item < hub.n_chunks) {
    const int2 c = __ldg(hub.chunks + item);
    const int rb = __ldg(rowptr + c.x), re = __ldg(rowptr + c.x + 1);
    w.row = c.x;
    w.first = c.y;
    w.slot = (int)item;
    w.lb = rb + ((int)item - c.y) * hub.chunk_edges;
    w.hb = min(w.lb + hub.chunk_edges, re);
    w.n_row_chunks = (re - rb + hub.chunk_edges - 1) / hub.chunk_edges;
    w.active = true;
    w.is_chunk = true;
  } else if (item - hub.n_chunks < n_rows) {
    w.row = (int)(item - hub.n_chunks);
    w.lb = __ldg(rowptr + w.row);
    w.hb = __ldg(rowptr + w.row + 1);
    w.active = !(hub.chunk_edges > 0 && (w.hb - w.lb) > hub.chunk_edges);
    if (!w.active) w.hb = w.lb;
  }
  return w;
}

// Arrival protocol for the chunks of one hub row: every lane has stored its part of the
// partial (st.cg); returns true in all lanes of the group that arrived last.
template <int GROUP>
__device__ __forceinline__ bool hub_arrive_last(const WorkItem &w, const HubView &hub, int gl) {
  __threadfence();   // my partial is visible device-wide before the counter moves
  __syncwarp();
  int old = 0;
  if (w.is_chunk && gl == 0) old = atomicAdd(hub.counters + w.first, 1);
  old = __shfl_sync(FULL, old, 0, GROUP);
  const bool last = w.is_chunk && (old == w.n_row_chunks - 1);
  if (last) {
    if (gl == 0) hub.counters[w.first] = 0;   // leave the counter ready for the next call
    __threadfence();                          // acquire side: partials read after this point
  }
  return last;
}

#endif  // __CUDACC__

}  // namespace cogdl_b200
