// cabi.cu -- C-ABI plumbing of libcogdl_b200: error text, device check, launch counter and the
// hub plan (which rows are cut into edge chunks).  See include/cogdl_b200.h.
#include "common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace cogdl_b200 {

static thread_local char g_err[512] = {0};
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_plan(const cogdl_b200_hub_plan_t *plan, int64_t need_partial_bytes) {
  if (!plan || plan->chunk_edges <= 0) return COGDL_B200_OK;
  if (plan->n_chunks < 0 || plan->n_hub_rows < 0)
    return set_error(COGDL_B200_EINVAL, "hub plan: negative counts");
  if (plan->n_chunks == 0) return COGDL_B200_OK;
  if (!plan->chunks || !plan->counters || !plan->hub_rows)
    return set_error(COGDL_B200_EINVAL, "hub plan: null chunk/counter/hub_rows array with n_chunks=%d", plan->n_chunks);
  if (need_partial_bytes > 0 && (!plan->partials || plan->partials_bytes < need_partial_bytes))
    return set_error(COGDL_B200_ESCRATCH, "hub plan: partials scratch too small (have %lld bytes, need %lld)",
                     (long long)plan->partials_bytes, (long long)need_partial_bytes);
  return COGDL_B200_OK;
}

// counts[0] += #hub rows, counts[1] += #chunks
__global__ void hub_count_kernel(const int *__restrict__ rowptr, int64_t n_rows, int T, int *counts) {
  int hubs = 0, chunks = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int deg = rowptr[r + 1] - rowptr[r];
    if (deg > T) {
      hubs += 1;
      chunks += (deg + T - 1) / T;
    }
  }
  for (int s = 16; s > 0; s >>= 1) {
    hubs += __shfl_xor_sync(FULL, hubs, s);
    chunks += __shfl_xor_sync(FULL, chunks, s);
  }
  if ((threadIdx.x & 31) == 0 && hubs) {
    atomicAdd(counts + 0, hubs);
    atomicAdd(counts + 1, chunks);
  }
}

__global__ void hub_fill_kernel(const int *__restrict__ rowptr, int64_t n_rows, int T, int *counts,
                                int *hub_rows, int2 *chunks) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int deg = rowptr[r + 1] - rowptr[r];
    if (deg > T) {
      const int n = (deg + T - 1) / T;
      const int h = atomicAdd(counts + 0, 1);
      const int first = atomicAdd(counts + 1, n);
      hub_rows[h] = (int)r;
      for (int j = 0; j < n; ++j) chunks[first + j] = make_int2((int)r, first);
    }
  }
}

// seg_starts[k] = first row r with rowptr[r] + r >= k*Q
__global__ void seg_starts_kernel(const int *__restrict__ rowptr, int64_t n_rows, int Q, int n_segs, int *seg_starts) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= n_segs; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t target = k * (int64_t)Q;
    int64_t lo = 0, hi = n_rows;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)__ldg(rowptr + mid) + mid < target) lo = mid + 1; else hi = mid;
    }
    seg_starts[k] = (int)lo;
  }
}

}  // namespace cogdl_b200

using namespace cogdl_b200;

extern "C" int cogdl_b200_abi_version(void) { return COGDL_B200_ABI_VERSION; }

extern "C" const char *cogdl_b200_last_error(void) { return g_err; }

extern "C" int64_t cogdl_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int cogdl_b200_check_device(void) {
  int dev = 0;
  CB_CUDA(cudaGetDevice(&dev));
  int major = 0;
  CB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10)
    return set_error(COGDL_B200_EDEVICE, "device %d has compute capability %d.x; libcogdl_b200 is built for sm_100a only", dev, major);
  return COGDL_B200_OK;
}

static int plan_grid(int64_t n_rows) {
  int64_t b = ceil_div(n_rows, 256);
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int cogdl_b200_hub_plan_count(const int32_t *rowptr, int64_t n_rows, int32_t chunk_edges,
                                         int32_t *counts_dev, cogdl_b200_stream_t stream) {
  CB_REQUIRE(rowptr && counts_dev, "cogdl_b200_hub_plan_count: null pointer");
  CB_REQUIRE(n_rows >= 0 && chunk_edges > 0, "cogdl_b200_hub_plan_count: bad size");
  cudaStream_t s = (cudaStream_t)stream;
  CB_CUDA(cudaMemsetAsync(counts_dev, 0, 2 * sizeof(int32_t), s));
  if (n_rows == 0) return COGDL_B200_OK;
  hub_count_kernel<<<plan_grid(n_rows), 256, 0, s>>>(rowptr, n_rows, chunk_edges, counts_dev);
  CB_LAUNCH_CHECK();
  return COGDL_B200_OK;
}

extern "C" int cogdl_b200_hub_plan_fill(const int32_t *rowptr, int64_t n_rows, int32_t chunk_edges,
                                        int32_t *counts_dev, int32_t *hub_rows, int32_t *chunks,
                                        cogdl_b200_stream_t stream) {
  CB_REQUIRE(rowptr && counts_dev && hub_rows && chunks, "cogdl_b200_hub_plan_fill: null pointer");
  CB_REQUIRE(n_rows >= 0 && chunk_edges > 0, "cogdl_b200_hub_plan_fill: bad size");
  cudaStream_t s = (cudaStream_t)stream;
  CB_CUDA(cudaMemsetAsync(counts_dev, 0, 2 * sizeof(int32_t), s));
  if (n_rows == 0) return COGDL_B200_OK;
  hub_fill_kernel<<<plan_grid(n_rows), 256, 0, s>>>(rowptr, n_rows, chunk_edges, counts_dev, hub_rows,
                                                    reinterpret_cast<int2 *>(chunks));
  CB_LAUNCH_CHECK();
  return COGDL_B200_OK;
}

extern "C" int cogdl_b200_hub_plan_segments(const int32_t *rowptr, int64_t n_rows, int32_t seg_cost,
                                            int32_t n_segs, int32_t *seg_starts, cogdl_b200_stream_t stream) {
  CB_REQUIRE(rowptr && seg_starts, "cogdl_b200_hub_plan_segments: null pointer");
  CB_REQUIRE(n_rows >= 0 && seg_cost > 0 && n_segs >= 0, "cogdl_b200_hub_plan_segments: bad size");
  seg_starts_kernel<<<plan_grid((int64_t)n_segs + 1), 256, 0, (cudaStream_t)stream>>>(rowptr, n_rows, seg_cost, n_segs,
                                                                                       seg_starts);
  CB_LAUNCH_CHECK();
  return COGDL_B200_OK;
}
