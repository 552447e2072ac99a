// BatchNorm1d + tanh of the Product2Vec FFN as streaming kernels, and the row-select split of the
// backward pass.
//
// Replaces native_batch_norm / tanh / native_batch_norm_backward / tanh_backward / where that the
// reference issues through nn.BatchNorm1d, nn.Tanh (/root/reference/src/models/product2vec.py:16-17)
// and the `if neighbors ...` branch (:76).  The per-column reductions over the (up to millions of) rows
// accumulate in float64 and are summed in a fixed order (per-CTA partials, then CTA order), so the
// statistics are bit-reproducible and accurate to fp32 rounding; in the node-partitioned multi-GPU
// run the [2, n] partial sums are what gets all-reduced (SyncBN).
#include "common.cuh"

namespace pc {
namespace {

constexpr int RED_THREADS = 256;

// sums[0, c] = sum_r f0(r, c), sums[1, c] = sum_r f1(r, c) over rows, n4 = n / 4 column groups.
// MODE 0: f0 = x, f1 = x^2.   MODE 1: f0 = dy, f1 = dy * (x - mean) * rstd.
template <int MODE>
__global__ void __launch_bounds__(RED_THREADS)
col_reduce_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb, int64_t m, int n4,
                  const float* __restrict__ mean, const float* __restrict__ rstd, const int64_t* __restrict__ rowptr,
                  double* __restrict__ partial) {
  extern __shared__ double sm[];  // [row_lanes][2][n4 * 4]
  const int row_lanes = RED_THREADS / n4;
  const int cg = threadIdx.x % n4, rl = threadIdx.x / n4;
  double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  if (rl < row_lanes) {
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = mu;
    if (MODE == 1) {
      mu = *reinterpret_cast<const float4*>(mean + cg * 4);
      rs = *reinterpret_cast<const float4*>(rstd + cg * 4);
    }
    const int64_t rows_per_cta = (m + gridDim.x - 1) / gridDim.x;
    const int64_t r_beg = int64_t(blockIdx.x) * rows_per_cta;
    const int64_t r_end = min(m, r_beg + rows_per_cta);
    for (int64_t r = r_beg + rl; r < r_end; r += row_lanes) {
      const float4 x = ld_stream4(reinterpret_cast<const float4*>(a + r * lda) + cg);
      if (MODE == 0) {
        s0[0] += x.x; s0[1] += x.y; s0[2] += x.z; s0[3] += x.w;
        s1[0] += double(x.x) * x.x; s1[1] += double(x.y) * x.y; s1[2] += double(x.z) * x.z; s1[3] += double(x.w) * x.w;
      } else {
        const float4 y = ld_stream4(reinterpret_cast<const float4*>(b + r * ldb) + cg);
        s0[0] += x.x; s0[1] += x.y; s0[2] += x.z; s0[3] += x.w;
        s1[0] += double(x.x) * double((y.x - mu.x) * rs.x); s1[1] += double(x.y) * double((y.y - mu.y) * rs.y);
        s1[2] += double(x.z) * double((y.z - mu.z) * rs.z); s1[3] += double(x.w) * double((y.w - mu.w) * rs.w);
      }
    }
  }
  const int n = n4 * 4;
  if (rl < row_lanes) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sm[(rl * 2 + 0) * n + cg * 4 + j] = s0[j];
      sm[(rl * 2 + 1) * n + cg * 4 + j] = s1[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * n; i += RED_THREADS) {
    double acc = 0;
    for (int l = 0; l < row_lanes; ++l) acc += sm[l * 2 * n + i];
    partial[int64_t(blockIdx.x) * 2 * n + i] = acc;
  }
}

__global__ void col_reduce_final_kernel(const double* __restrict__ partial, int parts, int n2, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  double acc = 0;
  for (int p = 0; p < parts; ++p) acc += partial[int64_t(p) * n2 + i];
  out[i] = acc;
}

// y = tanh(x * scale[c] + shift[c])  (apply_tanh) or x * scale[c] + shift[c]
__global__ void __launch_bounds__(256)
scale_shift_tanh_kernel(const float* __restrict__ x, int64_t ldx, int64_t m, int n4, const float* __restrict__ scale,
                        const float* __restrict__ shift, int apply_tanh, float* __restrict__ y, int64_t ldy) {
  const int64_t total = m * n4;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / n4;
    const int c = int(i - r * n4);
    const float4 v = ld_stream4(reinterpret_cast<const float4*>(x + r * ldx) + c);
    const float4 s = *reinterpret_cast<const float4*>(scale + c * 4);
    const float4 t = *reinterpret_cast<const float4*>(shift + c * 4);
    float4 o = make_float4(fmaf(v.x, s.x, t.x), fmaf(v.y, s.y, t.y), fmaf(v.z, s.z, t.z), fmaf(v.w, s.w, t.w));
    if (apply_tanh) o = make_float4(tanhf(o.x), tanhf(o.y), tanhf(o.z), tanhf(o.w));
    reinterpret_cast<float4*>(y + r * ldy)[c] = o;
  }
}

// out = ca[c] * a + cb[c] * b + cc[c]   (BatchNorm input gradient with folded coefficients)
__global__ void __launch_bounds__(256)
affine2_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb, int64_t m, int n4,
               const float* __restrict__ ca, const float* __restrict__ cb, const float* __restrict__ cc,
               float* __restrict__ out, int64_t ldo) {
  const int64_t total = m * n4;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / n4;
    const int c = int(i - r * n4);
    const float4 x = ld_stream4(reinterpret_cast<const float4*>(a + r * lda) + c);
    const float4 y = ld_stream4(reinterpret_cast<const float4*>(b + r * ldb) + c);
    const float4 p = *reinterpret_cast<const float4*>(ca + c * 4);
    const float4 q = *reinterpret_cast<const float4*>(cb + c * 4);
    const float4 s = *reinterpret_cast<const float4*>(cc + c * 4);
    reinterpret_cast<float4*>(out + r * ldo)[c] =
        make_float4(fmaf(p.x, x.x, fmaf(q.x, y.x, s.x)), fmaf(p.y, x.y, fmaf(q.y, y.y, s.y)),
                    fmaf(p.z, x.z, fmaf(q.z, y.z, s.z)), fmaf(p.w, x.w, fmaf(q.w, y.w, s.w)));
  }
}

// rows with neighbours: kept = g, rest = 0; rows without: kept = 0, rest = g   (backward of the row select)
__global__ void __launch_bounds__(256)
mask_split_kernel(const float* __restrict__ g, int64_t m, int n4, const int64_t* __restrict__ rowptr,
                  float* __restrict__ kept, float* __restrict__ rest) {
  const int64_t total = m * n4;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / n4;
    const bool has = rowptr[r + 1] > rowptr[r];
    const float4 v = ld_stream4(reinterpret_cast<const float4*>(g) + i);
    reinterpret_cast<float4*>(kept)[i] = has ? v : z;
    reinterpret_cast<float4*>(rest)[i] = has ? z : v;
  }
}

// Column sums over the rows WITHOUT neighbours.  Lane l of a warp tests row base + l (two coalesced rowptr loads per warp);
// only the empty rows are then read, all lanes together, in ascending row order; per-CTA float64 partials in shared memory,
// summed in CTA order by col_reduce_final_kernel.  On a graph with few isolated nodes this reads little more than rowptr.
__global__ void __launch_bounds__(RED_THREADS)
col_sum_unselected_kernel(const float* __restrict__ x, int64_t ldx, int64_t m, int n, const int64_t* __restrict__ rowptr,
                          double* __restrict__ partial) {
  extern __shared__ double sm[];   // [warps][n]
  const int lane = lane_id(), w = warp_id(), warps = RED_THREADS / 32;
  double* mine = sm + w * n;
  for (int c = lane; c < n; c += 32) mine[c] = 0.0;
  __syncwarp();
  // 32-row blocks are dealt round-robin over all warps of the grid: isolated nodes tend to cluster (the tail of a src < dst
  // graph), a contiguous slab per CTA would leave all of them to a few CTAs.  The assignment is fixed: deterministic.
  const int64_t r_end = m;
  for (int64_t base = (int64_t(blockIdx.x) * warps + w) * 32; base < r_end; base += int64_t(gridDim.x) * warps * 32) {
    const int64_t r = base + lane;
    uint32_t empty = __ballot_sync(FULL, r < r_end && rowptr[r + 1] == rowptr[r]);
    while (empty) {
      const int l = __ffs(empty) - 1;
      empty &= empty - 1;
      const float* row = x + (base + l) * ldx;
      for (int c = lane; c < n; c += 32) mine[c] += double(row[c]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n; c += RED_THREADS) {
    double acc = 0;
    for (int ww = 0; ww < warps; ++ww) acc += sm[ww * n + c];
    partial[int64_t(blockIdx.x) * 2 * n + c] = acc;
    partial[int64_t(blockIdx.x) * 2 * n + n + c] = 0.0;
  }
}

int reduce_grid(int64_t m) {
  const int64_t want = (m + 63) / 64;
  const int64_t cap = int64_t(sm_count()) * 4;
  return int(want < cap ? (want < 1 ? 1 : want) : cap);
}

int elementwise_grid(int64_t total) {
  const int64_t want = (total + 255) / 256;
  const int64_t cap = int64_t(sm_count()) * 16;
  return int(want < cap ? (want < 1 ? 1 : want) : cap);
}

template <int MODE>
int col_reduce(const float* a, int64_t lda, const float* b, int64_t ldb, int64_t m, int n, const float* mean,
               const float* rstd, const int64_t* rowptr, double* sums, void* ws, size_t ws_bytes, cudaStream_t st) {
  PC_REQUIRE(m > 0 && n >= 4 && n % 4 == 0 && n <= 1024 && RED_THREADS % (n / 4) == 0, PC_ERR_UNSUPPORTED,
             "col_reduce: n=%d must divide 1024 and be a multiple of 4", n);
  PC_REQUIRE(a && sums && ws && lda % 4 == 0 && ldb % 4 == 0, PC_ERR_INVALID, "col_reduce: bad pointer / leading dimension");
  const int grid = reduce_grid(m);
  PC_REQUIRE(ws_bytes >= size_t(grid) * 2 * n * sizeof(double), PC_ERR_WORKSPACE, "col_reduce: workspace too small");
  const int n4 = n / 4, row_lanes = RED_THREADS / n4;
  const size_t smem = size_t(row_lanes) * 2 * n * sizeof(double);
  col_reduce_kernel<MODE><<<grid, RED_THREADS, smem, st>>>(a, lda, b, ldb, m, n4, mean, rstd, rowptr, reinterpret_cast<double*>(ws));
  PC_LAUNCH_CHECK();
  col_reduce_final_kernel<<<(2 * n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const double*>(ws), grid, 2 * n, sums);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

}  // namespace
}  // namespace pc

using namespace pc;

extern "C" size_t pc_col_reduce_workspace_bytes(int n) { return size_t(sm_count()) * 4 * 2 * size_t(n) * sizeof(double); }

extern "C" int pc_col_stats(const float* x, int64_t m, int n, int64_t ldx, double* sums, void* workspace,
                            size_t workspace_bytes, pc_stream_t stream) {
  return col_reduce<0>(x, ldx, nullptr, 4, m, n, nullptr, nullptr, nullptr, sums, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int pc_col_sum_unselected(const float* x, int64_t m, int n, int64_t ldx, const int64_t* rowptr, double* sums, void* workspace,
                                     size_t workspace_bytes, pc_stream_t stream) {
  PC_REQUIRE(rowptr && x && sums && workspace, PC_ERR_INVALID, "col_sum_unselected: null pointer");
  PC_REQUIRE(m > 0 && n >= 1 && n <= 1024, PC_ERR_UNSUPPORTED, "col_sum_unselected: bad shape m=%lld n=%d", (long long)m, n);
  const int grid = reduce_grid(m);
  PC_REQUIRE(workspace_bytes >= size_t(grid) * 2 * n * sizeof(double), PC_ERR_WORKSPACE, "col_sum_unselected: workspace too small");
  cudaStream_t st = as_stream(stream);
  col_sum_unselected_kernel<<<grid, RED_THREADS, size_t(RED_THREADS / 32) * n * sizeof(double), st>>>(x, ldx, m, n, rowptr,
                                                                                                      reinterpret_cast<double*>(workspace));
  PC_LAUNCH_CHECK();
  col_reduce_final_kernel<<<(2 * n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const double*>(workspace), grid, 2 * n, sums);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_bn_bwd_reduce(const float* dy, int64_t ld_dy, const float* x, int64_t ldx, int64_t m, int n,
                                const float* mean, const float* rstd, double* sums, void* workspace,
                                size_t workspace_bytes, pc_stream_t stream) {
  PC_REQUIRE(x && mean && rstd, PC_ERR_INVALID, "bn_bwd_reduce: null pointer");
  return col_reduce<1>(dy, ld_dy, x, ldx, m, n, mean, rstd, nullptr, sums, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int pc_scale_shift_tanh(const float* x, int64_t ldx, int64_t m, int n, const float* scale, const float* shift,
                                   int apply_tanh, float* y, int64_t ldy, pc_stream_t stream) {
  PC_REQUIRE(m >= 0 && n >= 4 && n % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, PC_ERR_INVALID, "scale_shift_tanh: bad shape");
  if (m == 0) return PC_OK;
  PC_REQUIRE(x && scale && shift && y, PC_ERR_INVALID, "scale_shift_tanh: null pointer");
  scale_shift_tanh_kernel<<<elementwise_grid(m * (n / 4)), 256, 0, as_stream(stream)>>>(x, ldx, m, n / 4, scale, shift,
                                                                                        apply_tanh, y, ldy);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_affine2(const float* a, int64_t lda, const float* b, int64_t ldb, int64_t m, int n, const float* ca,
                          const float* cb, const float* cc, float* out, int64_t ldo, pc_stream_t stream) {
  PC_REQUIRE(m >= 0 && n >= 4 && n % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldo % 4 == 0, PC_ERR_INVALID, "affine2: bad shape");
  if (m == 0) return PC_OK;
  PC_REQUIRE(a && b && ca && cb && cc && out, PC_ERR_INVALID, "affine2: null pointer");
  affine2_kernel<<<elementwise_grid(m * (n / 4)), 256, 0, as_stream(stream)>>>(a, lda, b, ldb, m, n / 4, ca, cb, cc, out, ldo);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_mask_split(const float* g, int64_t m, int n, const int64_t* rowptr, float* kept, float* rest,
                             pc_stream_t stream) {
  PC_REQUIRE(m >= 0 && n >= 4 && n % 4 == 0, PC_ERR_INVALID, "mask_split: bad shape");
  if (m == 0) return PC_OK;
  PC_REQUIRE(g && rowptr && kept && rest, PC_ERR_INVALID, "mask_split: null pointer");
  mask_split_kernel<<<elementwise_grid(m * (n / 4)), 256, 0, as_stream(stream)>>>(g, m, n / 4, rowptr, kept, rest);
  PC_LAUNCH_CHECK();
  return PC_OK;
}
