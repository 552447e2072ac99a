// Behaviour-product-graph construction on the device: edge keys, LSD radix sort, unique,
// sorted-set intersection / difference, CSR and transposed (CSC) edge lists.
//
// Replaces the Python dict-of-sets graph of /root/reference/src/data/bpg.py:7-38 (add_edge's
// set-insert deduplication, the O(E) get_neighbors scan) and the inline set algebra of
// /root/reference/src/data/synthetic_data.py:89-90,110-128 ((Bcv n Bpv) - Bcp and
// Bcp - (Bpv u Bcv)).  An edge is the 64-bit key src<<32 | dst; everything is integer work and
// bit-exact.  All kernels are HBM-bound streaming passes; only integer atomics are used (their
// results do not depend on arrival order), so outputs are deterministic.
#include "common.cuh"

namespace pc {
namespace {

// ------------------------------------------------------------------ pack / unpack
__global__ void pack_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int64_t n,
                            uint64_t* __restrict__ keys) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = (uint64_t(uint32_t(src[i])) << 32) | uint64_t(uint32_t(dst[i]));
}
__global__ void unpack_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ src,
                              int32_t* __restrict__ dst) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint64_t k = keys[i];
    src[i] = int32_t(k >> 32);
    dst[i] = int32_t(k & 0xffffffffu);
  }
}

// ------------------------------------------------------------------ exclusive scan of uint32 (reduce-then-scan)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* smem /*[33]*/) {
  const int lane = lane_id(), w = warp_id();
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) smem[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < (SCAN_THREADS / 32) ? smem[lane] : 0;
    uint32_t si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(FULL, si, o);
      if (lane >= o) si += y;
    }
    if (lane < SCAN_THREADS / 32) smem[lane] = si - s;
    if (lane == 31) smem[32] = si;
  }
  __syncthreads();
  const uint32_t res = smem[w] + inc - v;
  if (total) *total = smem[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t smem[33];
  const int64_t base = int64_t(blockIdx.x) * SCAN_TILE + int64_t(threadIdx.x) * SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) s += in[base + i];
  uint32_t total;
  block_exclusive_scan(s, &total, smem);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of `m` tile sums in place, total -> *grand (optional)
__global__ void __launch_bounds__(SCAN_THREADS)
scan_tiles_kernel(uint32_t* __restrict__ tile_sums, int64_t m, uint32_t* __restrict__ grand) {
  __shared__ uint32_t smem[33];
  uint32_t carry = 0;
  for (int64_t base = 0; base < m; base += SCAN_THREADS) {
    const int64_t i = base + threadIdx.x;
    const uint32_t v = i < m ? tile_sums[i] : 0;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, &total, smem);
    if (i < m) tile_sums[i] = carry + ex;
    carry += total;
  }
  if (grand && threadIdx.x == 0) *grand = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const uint32_t* in, int64_t n, const uint32_t* __restrict__ tile_offsets, uint32_t* out) {
  __shared__ uint32_t smem[33];
  const int64_t base = int64_t(blockIdx.x) * SCAN_TILE + int64_t(threadIdx.x) * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = base + i < n ? in[base + i] : 0;
    s += v[i];
  }
  uint32_t run = block_exclusive_scan(s, nullptr, smem) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

// exclusive scan of n uint32 (in -> out, may alias); scratch holds ceil(n/SCAN_TILE) tile sums (recursively scanned)
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* scratch, cudaStream_t st) {
  if (n == 0) return PC_OK;
  const int64_t tiles = ceil_div(n, SCAN_TILE);
  scan_reduce_kernel<<<unsigned(tiles), SCAN_THREADS, 0, st>>>(in, n, scratch);
  PC_LAUNCH_CHECK();
  scan_tiles_kernel<<<1, SCAN_THREADS, 0, st>>>(scratch, tiles, nullptr);
  PC_LAUNCH_CHECK();
  scan_apply_kernel<<<unsigned(tiles), SCAN_THREADS, 0, st>>>(in, n, scratch, out);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

// ------------------------------------------------------------------ radix sort (64-bit keys, 8-bit digits)
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ROUNDS = 16;                               // keys per thread
constexpr int SORT_WARP_KEYS = 32 * SORT_ROUNDS;              // 512 consecutive keys per warp
constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS;         // 4096 keys per CTA

// per-CTA digit histogram -> hist[digit * n_tiles + tile]  (digit-major so one scan gives the offsets)
__global__ void __launch_bounds__(SORT_THREADS)
sort_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, uint32_t n_tiles,
                 uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = int64_t(blockIdx.x) * SORT_TILE;
#pragma unroll 4
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    const int64_t i = base + r * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xff], 1u);
  }
  __syncthreads();
  hist[size_t(threadIdx.x) * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// stable scatter: warp w of the CTA owns keys [w*512, (w+1)*512) of the tile and walks them in
// 16 rounds of 32 consecutive keys; rank inside a round comes from __match_any_sync, the running
// per-warp digit counters live in shared memory.
__global__ void __launch_bounds__(SORT_THREADS)
sort_scatter_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, uint32_t n_tiles,
                    const uint32_t* __restrict__ offsets, uint64_t* __restrict__ out) {
  __shared__ uint32_t cnt[SORT_WARPS][256];
  const int lane = lane_id(), w = warp_id();
  for (int d = lane; d < 256; d += 32) cnt[w][d] = 0;
  __syncwarp();
  const int64_t warp_base = int64_t(blockIdx.x) * SORT_TILE + int64_t(w) * SORT_WARP_KEYS;
  uint64_t key[SORT_ROUNDS];
  uint16_t rank[SORT_ROUNDS];
#pragma unroll
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    const int64_t i = warp_base + r * 32 + lane;
    key[r] = i < n ? keys[i] : ~uint64_t(0);
  }
#pragma unroll
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    const int64_t i = warp_base + r * 32 + lane;
    const bool valid = i < n;
    const uint32_t digit = valid ? uint32_t((key[r] >> shift) & 0xff) : 256u;
    const uint32_t peers = __match_any_sync(FULL, digit);
    const uint32_t before = __popc(peers & ((1u << lane) - 1u));
    uint32_t base = 0;
    if (valid) base = cnt[w][digit];
    __syncwarp();
    if (valid && before == 0) cnt[w][digit] = base + __popc(peers);
    __syncwarp();
    rank[r] = uint16_t(base + before);
  }
  __syncthreads();
  // thread d: exclusive prefix over warps of digit d, plus the global offset of (digit d, this tile)
  {
    const int d = threadIdx.x;
    uint32_t run = offsets[size_t(d) * n_tiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww) {
      const uint32_t c = cnt[ww][d];
      cnt[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    const int64_t i = warp_base + r * 32 + lane;
    if (i < n) {
      const uint32_t digit = uint32_t((key[r] >> shift) & 0xff);
      out[cnt[w][digit] + rank[r]] = key[r];
    }
  }
}

// ------------------------------------------------------------------ compaction (flag -> scan -> scatter)
enum FlagMode { FLAG_UNIQUE = 0, FLAG_IN_B = 1, FLAG_NOT_IN_B = 2 };

__device__ __forceinline__ bool contains_sorted(const uint64_t* __restrict__ b, int64_t nb, uint64_t x) {
  int64_t lo = 0, hi = nb;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(b + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo < nb && __ldg(b + lo) == x;
}

template <int MODE>
__global__ void flag_kernel(const uint64_t* __restrict__ a, int64_t na, const uint64_t* __restrict__ b, int64_t nb,
                            uint32_t* __restrict__ flags) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= na) return;
  bool f;
  if (MODE == FLAG_UNIQUE) f = (i == 0) || (a[i] != a[i - 1]);
  else f = contains_sorted(b, nb, a[i]) == (MODE == FLAG_IN_B);
  flags[i] = f ? 1u : 0u;
}

__global__ void compact_scatter_kernel(const uint64_t* __restrict__ a, int64_t na, const uint32_t* __restrict__ flags,
                                       const uint32_t* __restrict__ pos, uint64_t* __restrict__ out,
                                       int64_t* __restrict__ n_out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= na) return;
  if (flags[i]) out[pos[i]] = a[i];
  if (i == na - 1) *n_out = int64_t(pos[i]) + int64_t(flags[i]);
}

__global__ void set_zero_i64(int64_t* p) { *p = 0; }

template <int MODE>
int compact(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb, uint64_t* out, int64_t* n_out, void* ws,
            size_t ws_bytes, cudaStream_t st) {
  PC_REQUIRE(na >= 0 && na < (int64_t(1) << 31), PC_ERR_UNSUPPORTED, "compact: n=%lld outside [0, 2^31)", (long long)na);
  PC_REQUIRE(n_out, PC_ERR_INVALID, "compact: null n_out");
  if (na == 0) {
    set_zero_i64<<<1, 1, 0, st>>>(n_out);
    PC_LAUNCH_CHECK();
    return PC_OK;
  }
  PC_REQUIRE(a && out && ws, PC_ERR_INVALID, "compact: null pointer");
  PC_REQUIRE(ws_bytes >= pc_compact_workspace_bytes(na), PC_ERR_WORKSPACE, "compact: workspace %zu < %zu", ws_bytes,
             pc_compact_workspace_bytes(na));
  uint32_t* flags = reinterpret_cast<uint32_t*>(ws);
  uint32_t* pos = flags + align_up(size_t(na), 64);
  uint32_t* scratch = pos + align_up(size_t(na), 64);
  flag_kernel<MODE><<<unsigned(ceil_div(na, 256)), 256, 0, st>>>(a, na, b, nb, flags);
  PC_LAUNCH_CHECK();
  if (int rc = exclusive_scan_u32(flags, pos, na, scratch, st)) return rc;
  compact_scatter_kernel<<<unsigned(ceil_div(na, 256)), 256, 0, st>>>(a, na, flags, pos, out, n_out);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

// ------------------------------------------------------------------ CSR
__global__ void csr_rowptr_kernel(const uint64_t* __restrict__ keys, int64_t n_edges, int64_t n_rows,
                                  int64_t* __restrict__ rowptr) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  const uint64_t target = uint64_t(r) << 32;  // first key whose src >= r
  int64_t lo = 0, hi = n_edges;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < target) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = lo;
}
__global__ void csr_col_kernel(const uint64_t* __restrict__ keys, int64_t n_edges, int32_t* __restrict__ col) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e < n_edges) col[e] = int32_t(keys[e] & 0xffffffffu);
}
// one warp per row: keys_t[e] = col[e] << 32 | row
__global__ void __launch_bounds__(256)
transpose_keys_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                      uint64_t* __restrict__ keys_t) {
  const int64_t r = int64_t(blockIdx.x) * 8 + warp_id();
  if (r >= n_rows) return;
  const int64_t beg = rowptr[r], end = rowptr[r + 1];
  for (int64_t e = beg + lane_id(); e < end; e += 32)
    keys_t[e] = (uint64_t(uint32_t(col[e])) << 32) | uint64_t(uint32_t(r));
}

}  // namespace
}  // namespace pc

using namespace pc;

extern "C" int pc_edge_keys_pack(const int32_t* src, const int32_t* dst, int64_t n, uint64_t* keys,
                                 pc_stream_t stream) {
  PC_REQUIRE(n >= 0, PC_ERR_INVALID, "edge_keys_pack: negative n");
  if (n == 0) return PC_OK;
  PC_REQUIRE(src && dst && keys, PC_ERR_INVALID, "edge_keys_pack: null pointer");
  pack_kernel<<<unsigned(ceil_div(n, 256)), 256, 0, as_stream(stream)>>>(src, dst, n, keys);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_edge_keys_unpack(const uint64_t* keys, int64_t n, int32_t* src, int32_t* dst, pc_stream_t stream) {
  PC_REQUIRE(n >= 0, PC_ERR_INVALID, "edge_keys_unpack: negative n");
  if (n == 0) return PC_OK;
  PC_REQUIRE(src && dst && keys, PC_ERR_INVALID, "edge_keys_unpack: null pointer");
  unpack_kernel<<<unsigned(ceil_div(n, 256)), 256, 0, as_stream(stream)>>>(keys, n, src, dst);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" size_t pc_sort_keys_workspace_bytes(int64_t n) {
  if (n <= 0) return 0;
  const size_t tiles = size_t(ceil_div(n, SORT_TILE));
  const size_t hist = align_up(tiles * 256 * sizeof(uint32_t), 256);
  const size_t scan_scratch = align_up(size_t(ceil_div(int64_t(tiles) * 256, SCAN_TILE)) * sizeof(uint32_t), 256);
  return align_up(size_t(n) * sizeof(uint64_t), 256) + hist + scan_scratch;
}

extern "C" int pc_sort_keys(uint64_t* keys, int64_t n, uint32_t digit_mask, void* workspace, size_t workspace_bytes,
                            pc_stream_t stream) {
  PC_REQUIRE(n >= 0 && n < (int64_t(1) << 31), PC_ERR_UNSUPPORTED, "sort_keys: n=%lld outside [0, 2^31)", (long long)n);
  if (n <= 1 || (digit_mask & 0xff) == 0) return PC_OK;
  PC_REQUIRE(keys && workspace, PC_ERR_INVALID, "sort_keys: null pointer");
  PC_REQUIRE(workspace_bytes >= pc_sort_keys_workspace_bytes(n), PC_ERR_WORKSPACE, "sort_keys: workspace %zu < %zu",
             workspace_bytes, pc_sort_keys_workspace_bytes(n));
  cudaStream_t st = as_stream(stream);
  const uint32_t tiles = uint32_t(ceil_div(n, SORT_TILE));
  char* ws = reinterpret_cast<char*>(workspace);
  uint64_t* tmp = reinterpret_cast<uint64_t*>(ws);
  uint32_t* hist = reinterpret_cast<uint32_t*>(ws + align_up(size_t(n) * sizeof(uint64_t), 256));
  uint32_t* scratch = hist + align_up(size_t(tiles) * 256 * sizeof(uint32_t), 256) / sizeof(uint32_t);
  uint64_t* in = keys;
  uint64_t* out = tmp;
  for (int d = 0; d < 8; ++d) {
    if (!(digit_mask & (1u << d))) continue;
    sort_hist_kernel<<<tiles, SORT_THREADS, 0, st>>>(in, n, d * 8, tiles, hist);
    PC_LAUNCH_CHECK();
    if (int rc = exclusive_scan_u32(hist, hist, int64_t(tiles) * 256, scratch, st)) return rc;
    sort_scatter_kernel<<<tiles, SORT_THREADS, 0, st>>>(in, n, d * 8, tiles, hist, out);
    PC_LAUNCH_CHECK();
    uint64_t* t = in; in = out; out = t;
  }
  if (in != keys) PC_CUDA(cudaMemcpyAsync(keys, in, size_t(n) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
  return PC_OK;
}

extern "C" size_t pc_compact_workspace_bytes(int64_t n) {
  if (n <= 0) return 0;
  return (2 * align_up(size_t(n), 64) + align_up(size_t(ceil_div(n, SCAN_TILE)), 64)) * sizeof(uint32_t);
}

extern "C" int pc_unique_sorted_keys(const uint64_t* keys, int64_t n, uint64_t* out, int64_t* n_out, void* workspace,
                                     size_t workspace_bytes, pc_stream_t stream) {
  return compact<FLAG_UNIQUE>(keys, n, nullptr, 0, out, n_out, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int pc_set_filter_sorted(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb, int keep_if_present,
                                    uint64_t* out, int64_t* n_out, void* workspace, size_t workspace_bytes,
                                    pc_stream_t stream) {
  PC_REQUIRE(nb >= 0 && (nb == 0 || b), PC_ERR_INVALID, "set_filter_sorted: bad b");
  if (keep_if_present)
    return compact<FLAG_IN_B>(a, na, b, nb, out, n_out, workspace, workspace_bytes, as_stream(stream));
  return compact<FLAG_NOT_IN_B>(a, na, b, nb, out, n_out, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int pc_csr_from_sorted_keys(const uint64_t* keys, int64_t n_edges, int64_t n_rows, int64_t* rowptr,
                                       int32_t* col, pc_stream_t stream) {
  PC_REQUIRE(n_edges >= 0 && n_rows >= 0, PC_ERR_INVALID, "csr_from_sorted_keys: negative size");
  PC_REQUIRE(rowptr && (n_edges == 0 || (keys && col)), PC_ERR_INVALID, "csr_from_sorted_keys: null pointer");
  cudaStream_t st = as_stream(stream);
  csr_rowptr_kernel<<<unsigned(ceil_div(n_rows + 1, 256)), 256, 0, st>>>(keys, n_edges, n_rows, rowptr);
  PC_LAUNCH_CHECK();
  if (n_edges > 0) {
    csr_col_kernel<<<unsigned(ceil_div(n_edges, 256)), 256, 0, st>>>(keys, n_edges, col);
    PC_LAUNCH_CHECK();
  }
  return PC_OK;
}

extern "C" int pc_csr_transpose_keys(const int64_t* rowptr, const int32_t* col, int64_t n_rows, int64_t n_edges,
                                     uint64_t* keys_t, pc_stream_t stream) {
  PC_REQUIRE(n_rows >= 0 && n_edges >= 0, PC_ERR_INVALID, "csr_transpose_keys: negative size");
  if (n_rows == 0 || n_edges == 0) return PC_OK;
  PC_REQUIRE(rowptr && col && keys_t, PC_ERR_INVALID, "csr_transpose_keys: null pointer");
  transpose_keys_kernel<<<unsigned(ceil_div(n_rows, 8)), 256, 0, as_stream(stream)>>>(rowptr, col, n_rows, keys_t);
  PC_LAUNCH_CHECK();
  return PC_OK;
}
