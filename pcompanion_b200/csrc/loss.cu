// Fused hinge losses (forward + backward).
//
// Row hinge replaces F.pairwise_distance x2 + expand + mean + relu + mean of
// /root/reference/src/models/product2vec.py:137-154 (triplet, sign as written there) and
// torch.norm x2 + clamp + mean of /root/reference/src/models/p_companion.py:105-119 (item loss).
// Type hinge replaces the advanced-index gathers + clamp + mean of p_companion.py:95-103.
// One warp per row; all reductions are fixed-order (xor butterflies, one strided block
// reduction for the mean), so the loss value is bit-reproducible.
#include <math.h>

#include "common.cuh"

namespace pc {
namespace {

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
  return x;
}

// || a - b + eps ||_2 over dim4 float4 chunks, whole warp cooperates; every lane gets the result
__device__ __forceinline__ float row_distance(const float4* a, const float4* b, int dim4, float eps) {
  float acc = 0.f;
  for (int c = lane_id(); c < dim4; c += 32) {
    const float4 x = ldg4(a + c), y = ldg4(b + c);
    const float dx = x.x - y.x + eps, dy = x.y - y.y + eps, dz = x.z - y.z + eps, dw = x.w - y.w + eps;
    acc = fmaf(dw, dw, fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, acc))));
  }
  return sqrtf(warp_sum(acc));
}

__global__ void __launch_bounds__(256)
hinge_rows_fwd_kernel(const float4* __restrict__ A, const float4* __restrict__ P, const float4* __restrict__ N,
                      int64_t rows, int a_per_group, int kneg, int dim4, float margin, float eps,
                      float* __restrict__ per_row) {
  const int64_t r = int64_t(blockIdx.x) * 8 + warp_id();
  if (r >= rows) return;
  const int64_t g = r / a_per_group;
  const float4* a = A + r * dim4;
  const float dpos = row_distance(a, P + g * dim4, dim4, eps);
  float dneg = 0.f;
  for (int k = 0; k < kneg; ++k) dneg += row_distance(a, N + (g * kneg + k) * dim4, dim4, eps);
  dneg /= float(kneg);
  if (lane_id() == 0) per_row[r] = fmaxf(margin - dpos + dneg, 0.f);
}

// out[0] = sum(x) / n with a fixed summation order (single CTA).
__global__ void __launch_bounds__(1024) mean_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float part[32];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += x[i];
  acc = warp_sum(acc);
  if (lane_id() == 0) part[warp_id()] = acc;
  __syncthreads();
  if (warp_id() == 0) {
    float v = part[lane_id()];
    v = warp_sum(v);
    if (lane_id() == 0) out[0] = v / float(n);
  }
}

__global__ void __launch_bounds__(256)
hinge_rows_bwd_kernel(const float4* __restrict__ A, const float4* __restrict__ P, const float4* __restrict__ N,
                      int64_t rows, int a_per_group, int kneg, int dim4, float margin, float eps,
                      const float* __restrict__ grad_loss, float4* __restrict__ dA, float4* __restrict__ dP,
                      float4* __restrict__ dN) {
  const int64_t r = int64_t(blockIdx.x) * 8 + warp_id();
  if (r >= rows) return;
  const int lane = lane_id();
  const int64_t g = r / a_per_group;
  const float4* a = A + r * dim4;
  const float4* p = P + g * dim4;
  const float dpos = row_distance(a, p, dim4, eps);
  float my_dneg = 0.f, dneg_sum = 0.f;  // lane k keeps ||a - n_k||
  for (int k = 0; k < kneg; ++k) {
    const float d = row_distance(a, N + (g * kneg + k) * dim4, dim4, eps);
    if (lane == k) my_dneg = d;
    dneg_sum += d;
  }
  const bool active = (margin - dpos + dneg_sum / float(kneg)) > 0.f;
  const float gscale = active ? grad_loss[0] / float(rows) : 0.f;
  const float wpos = dpos > 0.f ? gscale / dpos : 0.f;  // torch.norm backward: 0 at norm == 0
  for (int c = lane; c < dim4; c += 32) {
    const float4 x = ldg4(a + c), y = ldg4(p + c);
    float4 up = make_float4((x.x - y.x + eps) * wpos, (x.y - y.y + eps) * wpos, (x.z - y.z + eps) * wpos,
                            (x.w - y.w + eps) * wpos);
    if (dP) dP[g * dim4 + c] = up;
    dA[r * dim4 + c] = make_float4(-up.x, -up.y, -up.z, -up.w);
  }
  for (int k = 0; k < kneg; ++k) {
    const float dk = __shfl_sync(FULL, my_dneg, k);
    const float wneg = dk > 0.f ? gscale / (dk * float(kneg)) : 0.f;
    const float4* nk = N + (g * kneg + k) * dim4;
    for (int c = lane; c < dim4; c += 32) {
      const float4 x = ldg4(a + c), y = ldg4(nk + c);
      const float4 un = make_float4((x.x - y.x + eps) * wneg, (x.y - y.y + eps) * wneg, (x.z - y.z + eps) * wneg,
                                    (x.w - y.w + eps) * wneg);
      float4 acc = dA[r * dim4 + c];
      acc.x += un.x; acc.y += un.y; acc.z += un.z; acc.w += un.w;
      dA[r * dim4 + c] = acc;
      if (dN) dN[(g * kneg + k) * dim4 + c] = make_float4(-un.x, -un.y, -un.z, -un.w);
    }
  }
}

// Triplet hinge on rows of ONE table picked by index, forward and (unscaled) gradient rows in one pass: warp b reads
// a = T[idx[b]], p = T[idx[B + b]], n_k = T[idx[2B + b K + k]] straight from the table (no gathered copy), writes the hinge of
// triplet b and - when `grads` is given - the gradient rows of its 2 + K slots for an upstream gradient of ONE
// (d loss / d slot row; the scalar upstream gradient is applied by the segment sum that builds d_table).  Same
// arithmetic as hinge_rows_fwd / hinge_rows_bwd, the anchor gradient accumulated in registers.
constexpr int TRI_MAXC = 4;   // float4 per lane: dim <= 512
__global__ void __launch_bounds__(256)
triplet_indexed_kernel(const float4* __restrict__ T, const int64_t* __restrict__ idx, int64_t batch, int kneg, int dim4,
                       float margin, float eps, float* __restrict__ per_row, float4* __restrict__ grads) {
  const int64_t b = int64_t(blockIdx.x) * 8 + warp_id();
  if (b >= batch) return;
  const int lane = lane_id();
  const float4* a = T + idx[b] * dim4;
  const float4* p = T + idx[batch + b] * dim4;
  const int64_t n0 = 2 * batch + b * kneg;
  const float dpos = row_distance(a, p, dim4, eps);
  float my_dneg = 0.f, dneg_sum = 0.f;  // lane k keeps ||a - n_k||
  for (int k = 0; k < kneg; ++k) {
    const float d = row_distance(a, T + idx[n0 + k] * dim4, dim4, eps);
    if (lane == k) my_dneg = d;
    dneg_sum += d;
  }
  const float h = margin - dpos + dneg_sum / float(kneg);
  if (lane == 0) per_row[b] = fmaxf(h, 0.f);
  if (!grads) return;
  const float gscale = h > 0.f ? 1.f / float(batch) : 0.f;
  const float wpos = dpos > 0.f ? gscale / dpos : 0.f;  // torch.norm backward: 0 at norm == 0
  float4 da[TRI_MAXC];
#pragma unroll
  for (int i = 0; i < TRI_MAXC; ++i) {
    const int c = lane + 32 * i;
    if (c < dim4) {
      const float4 x = ldg4(a + c), y = ldg4(p + c);
      const float4 up = make_float4((x.x - y.x + eps) * wpos, (x.y - y.y + eps) * wpos, (x.z - y.z + eps) * wpos,
                                    (x.w - y.w + eps) * wpos);
      grads[(batch + b) * dim4 + c] = up;
      da[i] = make_float4(-up.x, -up.y, -up.z, -up.w);
    }
  }
  for (int k = 0; k < kneg; ++k) {
    const float dk = __shfl_sync(FULL, my_dneg, k);
    const float wneg = dk > 0.f ? gscale / (dk * float(kneg)) : 0.f;
    const float4* nk = T + idx[n0 + k] * dim4;
#pragma unroll
    for (int i = 0; i < TRI_MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < dim4) {
        const float4 x = ldg4(a + c), y = ldg4(nk + c);
        const float4 un = make_float4((x.x - y.x + eps) * wneg, (x.y - y.y + eps) * wneg, (x.z - y.z + eps) * wneg,
                                      (x.w - y.w + eps) * wneg);
        da[i].x += un.x; da[i].y += un.y; da[i].z += un.z; da[i].w += un.w;
        grads[(n0 + k) * dim4 + c] = make_float4(-un.x, -un.y, -un.z, -un.w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TRI_MAXC; ++i) {
    const int c = lane + 32 * i;
    if (c < dim4) grads[b * dim4 + c] = da[i];
  }
}

__global__ void hinge_type_fwd_kernel(const float* __restrict__ S, const int64_t* __restrict__ pos,
                                      const int64_t* __restrict__ neg, int64_t rows, int64_t n_types, float margin,
                                      float* __restrict__ per_row) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  per_row[i] = fmaxf(margin - S[i * n_types + pos[i]] + S[i * n_types + neg[i]], 0.f);
}

__global__ void hinge_type_bwd_kernel(const float* __restrict__ S, const int64_t* __restrict__ pos,
                                      const int64_t* __restrict__ neg, int64_t rows, int64_t n_types, float margin,
                                      const float* __restrict__ grad_loss, float* __restrict__ dS) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float sp = S[i * n_types + pos[i]], sn = S[i * n_types + neg[i]];
  const float g = (margin - sp + sn) > 0.f ? grad_loss[0] / float(rows) : 0.f;
  // pos == neg: the two contributions cancel exactly, as index_put with accumulate does in autograd
  if (pos[i] == neg[i]) return;
  dS[i * n_types + pos[i]] = -g;
  dS[i * n_types + neg[i]] = g;
}

}  // namespace
}  // namespace pc

using namespace pc;

static int check_rows(const float* a, const float* p, const float* n, int64_t rows, int a_per_group, int kneg, int dim) {
  PC_REQUIRE(rows > 0, PC_ERR_INVALID, "hinge_rows: rows must be positive (mean of an empty batch is undefined)");
  PC_REQUIRE(a && p && n, PC_ERR_INVALID, "hinge_rows: null pointer");
  PC_REQUIRE(a_per_group >= 1 && rows % a_per_group == 0, PC_ERR_INVALID, "hinge_rows: rows %% a_per_group != 0");
  PC_REQUIRE(kneg >= 1 && kneg <= 32, PC_ERR_UNSUPPORTED, "hinge_rows: kneg=%d outside [1,32]", kneg);
  PC_REQUIRE(dim > 0 && dim % 4 == 0, PC_ERR_UNSUPPORTED, "hinge_rows: dim=%d must be a multiple of 4", dim);
  return PC_OK;
}

extern "C" int pc_hinge_rows_fwd(const float* a, const float* p, const float* n, int64_t rows, int a_per_group,
                                 int kneg, int dim, float margin, float eps, float* per_row, float* loss,
                                 pc_stream_t stream) {
  if (int rc = check_rows(a, p, n, rows, a_per_group, kneg, dim)) return rc;
  PC_REQUIRE(per_row && loss, PC_ERR_INVALID, "hinge_rows_fwd: null output");
  cudaStream_t st = as_stream(stream);
  hinge_rows_fwd_kernel<<<unsigned(ceil_div(rows, 8)), 256, 0, st>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(p), reinterpret_cast<const float4*>(n), rows,
      a_per_group, kneg, dim / 4, margin, eps, per_row);
  PC_LAUNCH_CHECK();
  mean_kernel<<<1, 1024, 0, st>>>(per_row, rows, loss);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_hinge_rows_bwd(const float* a, const float* p, const float* n, int64_t rows, int a_per_group,
                                 int kneg, int dim, float margin, float eps, const float* grad_loss, float* d_a,
                                 float* d_p, float* d_n, pc_stream_t stream) {
  if (int rc = check_rows(a, p, n, rows, a_per_group, kneg, dim)) return rc;
  PC_REQUIRE(grad_loss && d_a, PC_ERR_INVALID, "hinge_rows_bwd: null pointer");
  PC_REQUIRE(a_per_group == 1 || (!d_p && !d_n), PC_ERR_INVALID,
             "hinge_rows_bwd: d_p/d_n must be NULL when a_per_group > 1");
  hinge_rows_bwd_kernel<<<unsigned(ceil_div(rows, 8)), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(p), reinterpret_cast<const float4*>(n), rows,
      a_per_group, kneg, dim / 4, margin, eps, grad_loss, reinterpret_cast<float4*>(d_a),
      reinterpret_cast<float4*>(d_p), reinterpret_cast<float4*>(d_n));
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_hinge_type_fwd(const float* sims, const int64_t* pos, const int64_t* neg, int64_t rows,
                                 int64_t n_types, float margin, float* per_row, float* loss, pc_stream_t stream) {
  PC_REQUIRE(rows > 0 && n_types > 0, PC_ERR_INVALID, "hinge_type: empty input");
  PC_REQUIRE(sims && pos && neg && per_row && loss, PC_ERR_INVALID, "hinge_type_fwd: null pointer");
  cudaStream_t st = as_stream(stream);
  hinge_type_fwd_kernel<<<unsigned(ceil_div(rows, 256)), 256, 0, st>>>(sims, pos, neg, rows, n_types, margin, per_row);
  PC_LAUNCH_CHECK();
  mean_kernel<<<1, 1024, 0, st>>>(per_row, rows, loss);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_hinge_type_bwd(const float* sims, const int64_t* pos, const int64_t* neg, int64_t rows,
                                 int64_t n_types, float margin, const float* grad_loss, float* d_sims,
                                 pc_stream_t stream) {
  PC_REQUIRE(rows > 0 && n_types > 0, PC_ERR_INVALID, "hinge_type: empty input");
  PC_REQUIRE(sims && pos && neg && grad_loss && d_sims, PC_ERR_INVALID, "hinge_type_bwd: null pointer");
  hinge_type_bwd_kernel<<<unsigned(ceil_div(rows, 256)), 256, 0, as_stream(stream)>>>(sims, pos, neg, rows, n_types,
                                                                                       margin, grad_loss, d_sims);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_triplet_indexed(const float* table, const int64_t* slot_rows, int64_t batch, int kneg, int dim, float margin,
                                  float eps, float* per_row, float* loss, float* slot_grads, pc_stream_t stream) {
  PC_REQUIRE(batch > 0, PC_ERR_INVALID, "triplet_indexed: batch must be positive (mean of an empty batch is undefined)");
  PC_REQUIRE(table && slot_rows && per_row && loss, PC_ERR_INVALID, "triplet_indexed: null pointer");
  PC_REQUIRE(kneg >= 1 && kneg <= 32, PC_ERR_UNSUPPORTED, "triplet_indexed: kneg=%d outside [1,32]", kneg);
  PC_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 128 * TRI_MAXC, PC_ERR_UNSUPPORTED, "triplet_indexed: dim=%d must be a multiple of 4 up to %d", dim,
             128 * TRI_MAXC);
  cudaStream_t st = as_stream(stream);
  triplet_indexed_kernel<<<unsigned(ceil_div(batch, 8)), 256, 0, st>>>(reinterpret_cast<const float4*>(table), slot_rows, batch, kneg,
                                                                      dim / 4, margin, eps, per_row, reinterpret_cast<float4*>(slot_grads));
  PC_LAUNCH_CHECK();
  mean_kernel<<<1, 1024, 0, st>>>(per_row, batch, loss);
  PC_LAUNCH_CHECK();
  return PC_OK;
}
