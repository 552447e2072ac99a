// Error reporting, device query and the small row gather / scatter-add helpers of the C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace pc {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

namespace {

// out[r, :] = table[index[r], :]; one warp per row, float4 per lane, `width4` float4 per row.
__global__ void __launch_bounds__(256)
rows_gather_kernel(const float4* __restrict__ table, const int64_t* __restrict__ index, int64_t n, int width4,
                   float4* __restrict__ out) {
  const int64_t r = int64_t(blockIdx.x) * 8 + warp_id();
  if (r >= n) return;
  const float4* src = table + index[r] * width4;
  float4* dst = out + r * width4;
  for (int c = lane_id(); c < width4; c += 32) dst[c] = ldg4(src + c);
}

// table[index[r], :] += rows[r, :]; index unique within one call => race free and deterministic.
__global__ void __launch_bounds__(256)
rows_scatter_add_kernel(const float4* __restrict__ rows, const int64_t* __restrict__ index, int64_t n, int width4,
                        float4* __restrict__ table) {
  const int64_t r = int64_t(blockIdx.x) * 8 + warp_id();
  if (r >= n) return;
  const float4* src = rows + r * width4;
  float4* dst = table + index[r] * width4;
  for (int c = lane_id(); c < width4; c += 32) {
    float4 a = dst[c];
    const float4 b = ldg4(src + c);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    dst[c] = a;
  }
}

// out[r, :] = sum over e in [rowptr[r], rowptr[r+1]) of rows[col[e], :] in ascending e (zeros for empty rows):
// the deterministic transpose of a row gather (gradient of table[index] without float atomics).
__global__ void __launch_bounds__(256)
rows_segment_sum_kernel(const float4* __restrict__ rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                        int64_t n, int width4, float4* __restrict__ out) {
  const int64_t r = int64_t(blockIdx.x) * 8 + warp_id();
  if (r >= n) return;
  const int64_t beg = rowptr[r], end = rowptr[r + 1];
  for (int c = lane_id(); c < width4; c += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t e = beg; e < end; ++e) {
      const float4 v = ldg4(rows + int64_t(col[e]) * width4 + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[r * width4 + c] = acc;
  }
}

}  // namespace
}  // namespace pc

using namespace pc;

extern "C" int pc_abi_version(void) { return PC_ABI_VERSION; }

extern "C" const char* pc_last_error(void) { return g_error; }

extern "C" int pc_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host) {
  int dev = 0;
  PC_CUDA(cudaGetDevice(&dev));
  if (sm_count_host) PC_CUDA(cudaDeviceGetAttribute(sm_count_host, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major_host) PC_CUDA(cudaDeviceGetAttribute(cc_major_host, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor_host) PC_CUDA(cudaDeviceGetAttribute(cc_minor_host, cudaDevAttrComputeCapabilityMinor, dev));
  return PC_OK;
}

extern "C" int pc_rows_gather(const float* table, const int64_t* index, int64_t n, int width, float* out,
                              pc_stream_t stream) {
  PC_REQUIRE(n >= 0 && width > 0 && width % 4 == 0, PC_ERR_INVALID, "rows_gather: bad n=%lld width=%d", (long long)n, width);
  if (n == 0) return PC_OK;
  PC_REQUIRE(table && index && out, PC_ERR_INVALID, "rows_gather: null pointer");
  rows_gather_kernel<<<unsigned(ceil_div(n, 8)), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(table), index, n, width / 4, reinterpret_cast<float4*>(out));
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_rows_scatter_add(const float* rows, const int64_t* index, int64_t n, int width, float* table,
                                   pc_stream_t stream) {
  PC_REQUIRE(n >= 0 && width > 0 && width % 4 == 0, PC_ERR_INVALID, "rows_scatter_add: bad n=%lld width=%d", (long long)n, width);
  if (n == 0) return PC_OK;
  PC_REQUIRE(table && index && rows, PC_ERR_INVALID, "rows_scatter_add: null pointer");
  rows_scatter_add_kernel<<<unsigned(ceil_div(n, 8)), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(rows), index, n, width / 4, reinterpret_cast<float4*>(table));
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_rows_segment_sum(const float* rows, const int64_t* rowptr, const int32_t* col, int64_t n, int width,
                                   float* out, pc_stream_t stream) {
  PC_REQUIRE(n >= 0 && width > 0 && width % 4 == 0, PC_ERR_INVALID, "rows_segment_sum: bad n=%lld width=%d", (long long)n, width);
  if (n == 0) return PC_OK;
  PC_REQUIRE(rowptr && out, PC_ERR_INVALID, "rows_segment_sum: null pointer");
  rows_segment_sum_kernel<<<unsigned(ceil_div(n, 8)), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(rows), rowptr, col, n, width / 4, reinterpret_cast<float4*>(out));
  PC_LAUNCH_CHECK();
  return PC_OK;
}
