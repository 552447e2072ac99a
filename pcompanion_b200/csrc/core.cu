// Error reporting, device query and the small row gather / scatter-add helpers of the C ABI.
#include <stdarg.h>
#include <string.h>

#include <algorithm>

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
  static int cached[64] = {};
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached[dev] = n;
    else
      return 148;
  }
  return cached[dev];
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
                        int64_t n, int width4, const float* __restrict__ scale, float4* __restrict__ out) {
  const int64_t r = int64_t(blockIdx.x) * 8 + warp_id();
  if (r >= n) return;
  const int64_t beg = rowptr[r], end = rowptr[r + 1];
  const float sc = scale ? scale[0] : 1.f;
  for (int c = lane_id(); c < width4; c += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t e = beg; e < end; ++e) {
      const float4 v = ldg4(rows + int64_t(col[e]) * width4 + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[r * width4 + c] = scale ? make_float4(acc.x * sc, acc.y * sc, acc.z * sc, acc.w * sc) : acc;
  }
}

// keys[s] = index[s] << 32 | s: a stable sort on the index bytes groups the slots of one table row, slots ascending
__global__ void index_slot_keys_kernel(const int64_t* __restrict__ index, int64_t n, uint64_t* __restrict__ keys) {
  const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s < n) keys[s] = (uint64_t(index[s]) << 32) | uint64_t(uint32_t(s));
}

// table[r, :] += sum over peers p = 0..world-1 (in that order) of rows[slot[p * n + r], :] where slot >= 0: the
// owner-side reduction of returned halo partials as ONE pass over the local rows (fixed order => deterministic).
__global__ void __launch_bounds__(256)
rows_reduce_peers_kernel(const float4* __restrict__ rows, const int32_t* __restrict__ slot, int world, int64_t n, int width4,
                         float4* __restrict__ table) {
  const int64_t r = int64_t(blockIdx.x) * 8 + warp_id();
  if (r >= n) return;
  int32_t mine = lane_id() < world ? slot[int64_t(lane_id()) * n + r] : -1;
  if (__ballot_sync(0xffffffffu, mine >= 0) == 0) return;
  float4* dst = table + r * width4;
  for (int c = lane_id(); c < width4; c += 32) {
    float4 a = dst[c];
    for (int p = 0; p < world; ++p) {
      const int32_t s = __shfl_sync(0xffffffffu, mine, p);
      if (s < 0) continue;
      const float4 b = ld_stream4(rows + int64_t(s) * width4 + c);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    dst[c] = a;
  }
}

// Fused halo pack + exchange: every row of the send list is read from the local table (through `index`, or a
// contiguous run when index == nullptr) and stored straight into the owner / consumer GPU's table over NVLink.
struct HaloPushArgs {
  int64_t row_off[PC_MAX_PEERS + 1];
  float4* base[PC_MAX_PEERS];
  int64_t src_row0[PC_MAX_PEERS];
  int64_t dst_row0[PC_MAX_PEERS];
  int64_t first_row;
  int world;
};

__global__ void __launch_bounds__(256)
halo_push_kernel(const float4* __restrict__ table, int64_t ld4, const int64_t* __restrict__ index, int width4,
                 const __grid_constant__ HaloPushArgs a) {
  const int64_t total = a.row_off[a.world];
  for (int64_t v = int64_t(blockIdx.x) * 8 + warp_id(); v < total; v += int64_t(gridDim.x) * 8) {
    int64_t j = v + a.first_row;          // cyclic walk: every rank starts with a different peer's rows
    if (j >= total) j -= total;
    int p = 0;
    while (j >= a.row_off[p + 1]) ++p;
    const int64_t k = j - a.row_off[p];
    const float4* src = table + (index ? index[j] : a.src_row0[p] + k) * ld4;
    float4* dst = a.base[p] + (a.dst_row0[p] + k) * width4;
    for (int c = lane_id(); c < width4; c += 32) dst[c] = ldg4(src + c);
  }
}

}  // namespace
}  // namespace pc

using namespace pc;

extern "C" int pc_rows_reduce_peers(const float* rows, const int32_t* slot, int world, int64_t n, int width, float* table,
                                    pc_stream_t stream) {
  PC_REQUIRE(world >= 1 && world <= PC_MAX_PEERS, PC_ERR_UNSUPPORTED, "rows_reduce_peers: world=%d outside [1,%d]", world,
             PC_MAX_PEERS);
  PC_REQUIRE(n >= 0 && width > 0 && width % 4 == 0, PC_ERR_INVALID, "rows_reduce_peers: bad n=%lld width=%d", (long long)n, width);
  if (n == 0) return PC_OK;
  PC_REQUIRE(rows && slot && table, PC_ERR_INVALID, "rows_reduce_peers: null pointer");
  rows_reduce_peers_kernel<<<unsigned(ceil_div(n, 8)), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(rows), slot, world, n, width / 4, reinterpret_cast<float4*>(table));
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_halo_push(const float* table, int64_t ld, const int64_t* index, int world, const int64_t* row_off,
                            float* const* peer_base, const int64_t* src_row0, const int64_t* dst_row0, int64_t first_row,
                            int width, pc_stream_t stream) {
  PC_REQUIRE(world >= 1 && world <= PC_MAX_PEERS, PC_ERR_UNSUPPORTED, "halo_push: world=%d outside [1,%d]", world, PC_MAX_PEERS);
  PC_REQUIRE(width > 0 && width % 4 == 0 && ld % 4 == 0 && ld >= width, PC_ERR_INVALID, "halo_push: bad width=%d ld=%lld", width,
             (long long)ld);
  PC_REQUIRE(row_off && peer_base && dst_row0 && (index || src_row0), PC_ERR_INVALID, "halo_push: null pointer");
  HaloPushArgs a;
  a.world = world;
  a.row_off[0] = row_off[0];
  PC_REQUIRE(row_off[0] == 0, PC_ERR_INVALID, "halo_push: row_off[0] must be 0");
  for (int p = 0; p < world; ++p) {
    PC_REQUIRE(row_off[p + 1] >= row_off[p], PC_ERR_INVALID, "halo_push: row_off not ascending at peer %d", p);
    PC_REQUIRE(row_off[p + 1] == row_off[p] || peer_base[p], PC_ERR_INVALID, "halo_push: rows for peer %d but no base pointer", p);
    a.row_off[p + 1] = row_off[p + 1];
    a.base[p] = reinterpret_cast<float4*>(peer_base[p]);
    a.src_row0[p] = src_row0 ? src_row0[p] : 0;
    a.dst_row0[p] = dst_row0[p];
  }
  const int64_t total = row_off[world];
  if (total == 0) return PC_OK;
  PC_REQUIRE(table, PC_ERR_INVALID, "halo_push: null table");
  PC_REQUIRE(first_row >= 0 && first_row <= total, PC_ERR_INVALID, "halo_push: first_row=%lld outside [0,%lld]",
             (long long)first_row, (long long)total);
  a.first_row = first_row == total ? 0 : first_row;
  const int64_t blocks = std::min<int64_t>(ceil_div(total, 8), int64_t(sm_count()) * 32);
  halo_push_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(table), ld / 4, index,
                                                                    width / 4, a);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

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

static int segment_sum_scaled(const float* rows, const int64_t* rowptr, const int32_t* col, int64_t n, int width, const float* scale,
                              float* out, pc_stream_t stream) {
  rows_segment_sum_kernel<<<unsigned(ceil_div(n, 8)), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(rows), rowptr, col, n, width / 4, scale, reinterpret_cast<float4*>(out));
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_rows_segment_sum(const float* rows, const int64_t* rowptr, const int32_t* col, int64_t n, int width,
                                   float* out, pc_stream_t stream) {
  PC_REQUIRE(n >= 0 && width > 0 && width % 4 == 0, PC_ERR_INVALID, "rows_segment_sum: bad n=%lld width=%d", (long long)n, width);
  if (n == 0) return PC_OK;
  PC_REQUIRE(rowptr && out, PC_ERR_INVALID, "rows_segment_sum: null pointer");
  return segment_sum_scaled(rows, rowptr, col, n, width, nullptr, out, stream);
}


// Dense gradient of a row gather table[index] without float atomics, as ONE call: slots keyed by table row, stable radix
// sort on the row bytes, CSR over the table rows, per-row sum in slot order (pc_rows_segment_sum).
extern "C" size_t pc_rows_index_grad_workspace_bytes(int64_t slots, int64_t n_rows) {
  if (slots <= 0) return 0;
  return align_up(size_t(slots) * 8, 256) + align_up(pc_sort_keys_workspace_bytes(slots), 256) +
         align_up(size_t(n_rows + 1) * 8, 256) + align_up(size_t(slots) * 4, 256);
}

extern "C" int pc_rows_index_grad(const float* rows, const int64_t* index, int64_t slots, int64_t n_rows, int width, const float* scale,
                                  float* out, void* workspace, size_t workspace_bytes, pc_stream_t stream) {
  PC_REQUIRE(slots >= 0 && n_rows >= 0 && n_rows < (int64_t(1) << 31) && slots < (int64_t(1) << 31), PC_ERR_INVALID,
             "rows_index_grad: bad sizes");
  PC_REQUIRE(width > 0 && width % 4 == 0, PC_ERR_INVALID, "rows_index_grad: bad width=%d", width);
  if (n_rows == 0) return PC_OK;
  PC_REQUIRE(out, PC_ERR_INVALID, "rows_index_grad: null output");
  cudaStream_t st = as_stream(stream);
  if (slots == 0) {
    PC_CUDA(cudaMemsetAsync(out, 0, size_t(n_rows) * width * sizeof(float), st));
    return PC_OK;
  }
  PC_REQUIRE(rows && index && workspace, PC_ERR_INVALID, "rows_index_grad: null pointer");
  PC_REQUIRE(workspace_bytes >= pc_rows_index_grad_workspace_bytes(slots, n_rows), PC_ERR_WORKSPACE, "rows_index_grad: workspace too small");
  char* ws = reinterpret_cast<char*>(workspace);
  uint64_t* keys = reinterpret_cast<uint64_t*>(ws);          ws += align_up(size_t(slots) * 8, 256);
  void* sort_ws = ws;                                         const size_t sort_bytes = pc_sort_keys_workspace_bytes(slots);
  ws += align_up(sort_bytes, 256);
  int64_t* rowptr = reinterpret_cast<int64_t*>(ws);           ws += align_up(size_t(n_rows + 1) * 8, 256);
  int32_t* col = reinterpret_cast<int32_t*>(ws);
  index_slot_keys_kernel<<<unsigned(ceil_div(slots, 256)), 256, 0, st>>>(index, slots, keys);
  PC_LAUNCH_CHECK();
  uint32_t mask = 0;
  for (int b = 0; b < 4; ++b)
    if ((uint64_t(n_rows > 1 ? n_rows - 1 : 1) >> (8 * b)) != 0) mask |= 1u << (4 + b);
  if (int rc = pc_sort_keys(keys, slots, mask, sort_ws, sort_bytes, stream)) return rc;
  if (int rc = pc_csr_from_sorted_keys(keys, slots, n_rows, rowptr, col, stream)) return rc;
  return segment_sum_scaled(rows, rowptr, col, n_rows, width, scale, out, stream);
}
