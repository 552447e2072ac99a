// Small dense layers of the P-Companion joint model as fused kernels.
//
// Replaces, for /root/reference/src/models/type_transition.py:15-20, the chain
//   nn.Embedding gather -> Linear(64, 32) -> relu -> dropout -> Linear(32, 64)      (and its autograd)
// with one forward and one backward kernel (pc_mlp2_fwd / pc_mlp2_bwd); for
// /root/reference/src/models/item_prediction.py:33-38 the broadcast product
//   item_projection(q)[:, None, :] * type_projection(T)                              (pc_item_combine_fwd / _bwd;
// the two projections themselves are tcgen05 GEMMs, gemm.cu); and the gradient of the type hinge
// (/root/reference/src/models/p_companion.py:95-103) taken straight to the factors of S = base . W^T
// (pc_hinge_type_factored_bwd).  These layers are a few hundred rows per step in the reference's configuration: one warp
// per row, weights staged in shared memory, all reductions in a fixed order (bit-reproducible, no float atomics).
#include <math.h>

#include "common.cuh"

namespace pc {
namespace {

constexpr int ML_WARPS = 8;

__device__ __forceinline__ uint32_t mlp_hash(uint64_t seed, uint32_t row, uint32_t unit) {
  uint32_t x = row * 0x9E3779B1u + uint32_t(seed);
  x ^= x >> 15;
  x *= 0x2C1B3C6Du;
  x ^= unit * 0x85EBCA77u + uint32_t(seed >> 32);
  x ^= x >> 13;
  x *= 0x297A2D39u;
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  return x;
}

// out = W2 . drop(relu(W1 . x + b1)) + b2 per row; x = table[idx[row]] (idx != null) or table[row].
// Shared memory: W1t [d_in][hid], W2t [hid][d_out] (transposed: consecutive lanes read consecutive words), b1, b2,
// per-warp x and hidden rows.
__global__ void __launch_bounds__(ML_WARPS * 32)
mlp2_fwd_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx, int64_t rows, int d_in, int hid, int d_out,
                const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                const float* __restrict__ b2, uint32_t drop_threshold, float inv_keep, uint64_t seed,
                const uint64_t* __restrict__ seed_dev, float* __restrict__ hidden, float* __restrict__ out) {
  if (seed_dev) seed += *seed_dev;      // per-replay seed of a CUDA-graph-captured step (the graph increments the counter)
  extern __shared__ float sm[];
  float* w1t = sm;                         // [d_in][hid]
  float* w2t = w1t + d_in * hid;           // [hid][d_out]
  float* b1s = w2t + hid * d_out;
  float* b2s = b1s + hid;
  float* xs = b2s + d_out;                 // [ML_WARPS][d_in]
  float* hs = xs + ML_WARPS * d_in;        // [ML_WARPS][hid]
  for (int i = threadIdx.x; i < d_in * hid; i += blockDim.x) {
    const int j = i / d_in, k = i - j * d_in;          // W1 [hid, d_in] row-major
    w1t[k * hid + j] = W1[i];
  }
  for (int i = threadIdx.x; i < hid * d_out; i += blockDim.x) {
    const int o = i / hid, j = i - o * hid;            // W2 [d_out, hid] row-major
    w2t[j * d_out + o] = W2[i];
  }
  for (int i = threadIdx.x; i < hid; i += blockDim.x) b1s[i] = b1 ? b1[i] : 0.f;
  for (int i = threadIdx.x; i < d_out; i += blockDim.x) b2s[i] = b2 ? b2[i] : 0.f;
  __syncthreads();
  const int lane = lane_id(), w = warp_id();
  float* x = xs + w * d_in;
  float* h = hs + w * hid;
  for (int64_t r = int64_t(blockIdx.x) * ML_WARPS + w; r < rows; r += int64_t(gridDim.x) * ML_WARPS) {
    const float* src = table + (idx ? idx[r] : r) * d_in;
    for (int k = lane; k < d_in; k += 32) x[k] = src[k];
    __syncwarp();
    for (int j = lane; j < hid; j += 32) {
      float acc = b1s[j];
      for (int k = 0; k < d_in; ++k) acc = fmaf(w1t[k * hid + j], x[k], acc);
      acc = fmaxf(acc, 0.f);
      if (drop_threshold) acc = mlp_hash(seed, uint32_t(r), uint32_t(j)) >= drop_threshold ? acc * inv_keep : 0.f;
      h[j] = acc;
      hidden[r * hid + j] = acc;
    }
    __syncwarp();
    for (int o = lane; o < d_out; o += 32) {
      float acc = b2s[o];
      for (int j = 0; j < hid; ++j) acc = fmaf(w2t[j * d_out + o], h[j], acc);
      out[r * d_out + o] = acc;
    }
    __syncwarp();
  }
}

// Backward of mlp2_fwd.  Phase A (warp per row): d_hidden, d_pre (relu and dropout folded: hidden != 0 <=> unit alive), d_x.
// Phase B (whole CTA): the batch of ML_WARPS rows is accumulated into the CTA's weight-gradient partials in shared
// memory, every element by its one owner thread in row order.  A second kernel sums the CTA partials in CTA order.
__global__ void __launch_bounds__(ML_WARPS * 32)
mlp2_bwd_kernel(const float* __restrict__ d_out_g, const float* __restrict__ table, const int64_t* __restrict__ idx,
                const float* __restrict__ hidden, int64_t rows, int d_in, int hid, int d_out,
                const float* __restrict__ W1, const float* __restrict__ W2, float inv_keep, float* __restrict__ d_x,
                float* __restrict__ partial) {
  extern __shared__ float sm[];
  const int n1 = hid * d_in, n2 = d_out * hid;
  float* w1s = sm;                          // [hid][d_in]
  float* w2s = w1s + n1;                    // [d_out][hid]
  float* acc = w2s + n2;                    // [n1 | hid | n2 | d_out] partial gradients
  const int nacc = n1 + hid + n2 + d_out;
  float* xs = acc + nacc;                   // [ML_WARPS][d_in]
  float* hs = xs + ML_WARPS * d_in;         // [ML_WARPS][hid]
  float* ps = hs + ML_WARPS * hid;          // [ML_WARPS][hid]  d_pre
  float* gs = ps + ML_WARPS * hid;          // [ML_WARPS][d_out]
  for (int i = threadIdx.x; i < n1; i += blockDim.x) w1s[i] = W1[i];
  for (int i = threadIdx.x; i < n2; i += blockDim.x) w2s[i] = W2[i];
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int lane = lane_id(), w = warp_id();
  float* x = xs + w * d_in;
  float* h = hs + w * hid;
  float* dp = ps + w * hid;
  float* g = gs + w * d_out;
  const int64_t batches = (rows + ML_WARPS - 1) / ML_WARPS;
  for (int64_t b = blockIdx.x; b < batches; b += gridDim.x) {
    const int64_t r = b * ML_WARPS + w;
    const bool live = r < rows;
    if (live) {
      const float* src = table + (idx ? idx[r] : r) * d_in;
      for (int k = lane; k < d_in; k += 32) x[k] = src[k];
      for (int j = lane; j < hid; j += 32) h[j] = hidden[r * hid + j];
      for (int o = lane; o < d_out; o += 32) g[o] = d_out_g[r * d_out + o];
    } else {
      for (int k = lane; k < d_in; k += 32) x[k] = 0.f;
      for (int j = lane; j < hid; j += 32) h[j] = 0.f;
      for (int o = lane; o < d_out; o += 32) g[o] = 0.f;
    }
    __syncwarp();
    for (int j = lane; j < hid; j += 32) {
      float a = 0.f;
      for (int o = 0; o < d_out; ++o) a = fmaf(g[o], w2s[o * hid + j], a);
      dp[j] = h[j] != 0.f ? a * inv_keep : 0.f;
    }
    __syncwarp();
    if (live && d_x) {
      for (int k = lane; k < d_in; k += 32) {
        float a = 0.f;
        for (int j = 0; j < hid; ++j) a = fmaf(dp[j], w1s[j * d_in + k], a);
        d_x[r * d_in + k] = a;
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nacc; e += blockDim.x) {
      float a = acc[e];
      if (e < n1) {                                    // dW1[j, k] += d_pre[j] * x[k]
        const int j = e / d_in, k = e - j * d_in;
        for (int rr = 0; rr < ML_WARPS; ++rr) a = fmaf(ps[rr * hid + j], xs[rr * d_in + k], a);
      } else if (e < n1 + hid) {                       // db1[j] += d_pre[j]
        const int j = e - n1;
        for (int rr = 0; rr < ML_WARPS; ++rr) a += ps[rr * hid + j];
      } else if (e < n1 + hid + n2) {                  // dW2[o, j] += d_out[o] * hidden[j]
        const int q = e - n1 - hid;
        const int o = q / hid, j = q - o * hid;
        for (int rr = 0; rr < ML_WARPS; ++rr) a = fmaf(gs[rr * d_out + o], hs[rr * hid + j], a);
      } else {                                         // db2[o] += d_out[o]
        const int o = e - n1 - hid - n2;
        for (int rr = 0; rr < ML_WARPS; ++rr) a += gs[rr * d_out + o];
      }
      acc[e] = a;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) partial[int64_t(blockIdx.x) * nacc + i] = acc[i];
}

// out[i] = sum_c partial[c, i] in ascending c (fixed order)
__global__ void sum_partials_kernel(const float* __restrict__ partial, int parts, int n, float* __restrict__ o0, int n0,
                                    float* __restrict__ o1, int n1, float* __restrict__ o2, int n2, float* __restrict__ o3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  for (int c = 0; c < parts; ++c) a += partial[int64_t(c) * n + i];
  if (i < n0) { if (o0) o0[i] = a; }
  else if (i < n0 + n1) { if (o1) o1[i - n0] = a; }
  else if (i < n0 + n1 + n2) { if (o2) o2[i - n0 - n1] = a; }
  else if (o3) o3[i - n0 - n1 - n2] = a;
}

// out[b, t, :] = pi[b, :] * tp[b * kt + t, :]
__global__ void __launch_bounds__(256)
item_combine_fwd_kernel(const float4* __restrict__ pi, const float4* __restrict__ tp, int64_t rows, int kt, int d4,
                        float4* __restrict__ out) {
  const int64_t total = rows * kt * d4;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t bt = i / d4;
    const int c = int(i - bt * d4);
    const float4 a = pi[(bt / kt) * d4 + c], t = tp[i];
    out[i] = make_float4(a.x * t.x, a.y * t.y, a.z * t.z, a.w * t.w);
  }
}

// d_pi[b, :] = sum_t d_out[b, t, :] * tp[b, t, :] (t ascending); d_tp[b, t, :] = d_out[b, t, :] * pi[b, :]
__global__ void __launch_bounds__(256)
item_combine_bwd_kernel(const float4* __restrict__ d_out, const float4* __restrict__ pi, const float4* __restrict__ tp,
                        int64_t rows, int kt, int d4, float4* __restrict__ d_pi, float4* __restrict__ d_tp) {
  const int64_t total = rows * d4;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = i / d4;
    const int c = int(i - b * d4);
    const float4 a = pi[i];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < kt; ++t) {
      const int64_t j = (b * kt + t) * d4 + c;
      const float4 g = d_out[j], v = tp[j];
      s.x = fmaf(g.x, v.x, s.x); s.y = fmaf(g.y, v.y, s.y); s.z = fmaf(g.z, v.z, s.z); s.w = fmaf(g.w, v.w, s.w);
      d_tp[j] = make_float4(g.x * a.x, g.y * a.y, g.z * a.z, g.w * a.w);
    }
    d_pi[i] = s;
  }
}

// c_i = (per_i > 0 && pos_i != neg_i) ? grad / rows : 0;  d_base[i] = c_i (W[neg_i] - W[pos_i]);
// vals[i] = -c_i base[i] (slot of pos_i), vals[rows + i] = +c_i base[i] (slot of neg_i).  One warp per row.
__global__ void __launch_bounds__(256)
hinge_type_factored_bwd_kernel(const float* __restrict__ per, const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
                               const float* __restrict__ grad, const float4* __restrict__ base, const float4* __restrict__ W,
                               int64_t rows, int l4, float4* __restrict__ d_base, float4* __restrict__ vals) {
  const int64_t r = int64_t(blockIdx.x) * 8 + warp_id();
  if (r >= rows) return;
  const int64_t p = pos[r], n = neg[r];
  const float c = (per[r] > 0.f && p != n) ? grad[0] / float(rows) : 0.f;
  for (int k = lane_id(); k < l4; k += 32) {
    const float4 wn = W[n * l4 + k], wp = W[p * l4 + k], b = base[r * l4 + k];
    if (d_base) d_base[r * l4 + k] = make_float4(c * (wn.x - wp.x), c * (wn.y - wp.y), c * (wn.z - wp.z), c * (wn.w - wp.w));
    if (vals) {
      const float4 v = make_float4(c * b.x, c * b.y, c * b.z, c * b.w);
      vals[r * l4 + k] = make_float4(-v.x, -v.y, -v.z, -v.w);
      vals[(rows + r) * l4 + k] = v;
    }
  }
}

int mlp_grid(int64_t rows) {
  const int64_t want = ceil_div(rows, ML_WARPS);
  const int64_t cap = int64_t(sm_count()) * 2;
  return int(want < cap ? (want < 1 ? 1 : want) : cap);
}

int mlp_check(int64_t rows, int d_in, int hid, int d_out) {
  PC_REQUIRE(rows >= 0, PC_ERR_INVALID, "mlp2: negative row count");
  PC_REQUIRE(d_in >= 4 && d_in <= 256 && hid >= 1 && hid <= 128 && d_out >= 1 && d_out <= 256, PC_ERR_UNSUPPORTED,
             "mlp2: sizes (%d, %d, %d) outside d_in <= 256, hidden <= 128, d_out <= 256", d_in, hid, d_out);
  return PC_OK;
}

}  // namespace
}  // namespace pc

using namespace pc;

extern "C" int pc_mlp2_fwd(const float* table, const int64_t* idx, int64_t rows, int d_in, int hid, int d_out,
                           const float* w1, const float* b1, const float* w2, const float* b2, float dropout_p,
                           uint64_t seed, const uint64_t* seed_dev, float* hidden, float* out, pc_stream_t stream) {
  if (int rc = mlp_check(rows, d_in, hid, d_out)) return rc;
  if (rows == 0) return PC_OK;
  PC_REQUIRE(table && w1 && w2 && hidden && out, PC_ERR_INVALID, "mlp2_fwd: null pointer");
  PC_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PC_ERR_INVALID, "mlp2_fwd: dropout_p=%f outside [0,1)", dropout_p);
  const double t = double(dropout_p) * 4294967296.0;
  const uint32_t thr = dropout_p > 0.f ? (t >= 4294967295.0 ? 4294967295u : (uint32_t(t) ? uint32_t(t) : 1u)) : 0u;
  const size_t smem = size_t(d_in * hid + hid * d_out + hid + d_out + ML_WARPS * (d_in + hid)) * sizeof(float);
  PC_CUDA(cudaFuncSetAttribute(mlp2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  PC_REQUIRE(smem <= 200 * 1024, PC_ERR_UNSUPPORTED, "mlp2_fwd: shared memory budget exceeded");
  mlp2_fwd_kernel<<<mlp_grid(rows), ML_WARPS * 32, smem, as_stream(stream)>>>(table, idx, rows, d_in, hid, d_out, w1, b1, w2, b2, thr,
                                                                             1.f / (1.f - dropout_p), seed, seed_dev, hidden, out);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" size_t pc_mlp2_bwd_workspace_bytes(int d_in, int hid, int d_out) {
  return size_t(sm_count()) * 2 * size_t(hid * d_in + hid + d_out * hid + d_out) * sizeof(float);
}

extern "C" int pc_mlp2_bwd(const float* d_out_rows, const float* table, const int64_t* idx, const float* hidden, int64_t rows,
                           int d_in, int hid, int d_out, const float* w1, const float* w2, float dropout_p, float* d_x,
                           float* d_w1, float* d_b1, float* d_w2, float* d_b2, void* workspace, size_t workspace_bytes,
                           pc_stream_t stream) {
  if (int rc = mlp_check(rows, d_in, hid, d_out)) return rc;
  PC_REQUIRE(rows > 0, PC_ERR_INVALID, "mlp2_bwd: need at least one row");
  PC_REQUIRE(d_out_rows && table && hidden && w1 && w2 && workspace, PC_ERR_INVALID, "mlp2_bwd: null pointer");
  PC_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PC_ERR_INVALID, "mlp2_bwd: dropout_p=%f outside [0,1)", dropout_p);
  PC_REQUIRE(workspace_bytes >= pc_mlp2_bwd_workspace_bytes(d_in, hid, d_out), PC_ERR_WORKSPACE, "mlp2_bwd: workspace too small");
  const int n1 = hid * d_in, n2 = d_out * hid, nacc = n1 + hid + n2 + d_out;
  const size_t smem = size_t(n1 + n2 + nacc + ML_WARPS * (d_in + 2 * hid + d_out)) * sizeof(float);
  PC_CUDA(cudaFuncSetAttribute(mlp2_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  PC_REQUIRE(smem <= 200 * 1024, PC_ERR_UNSUPPORTED, "mlp2_bwd: shared memory budget exceeded");
  const int grid = mlp_grid(rows);
  cudaStream_t st = as_stream(stream);
  float* partial = reinterpret_cast<float*>(workspace);
  mlp2_bwd_kernel<<<grid, ML_WARPS * 32, smem, st>>>(d_out_rows, table, idx, hidden, rows, d_in, hid, d_out, w1, w2,
                                                     1.f / (1.f - dropout_p), d_x, partial);
  PC_LAUNCH_CHECK();
  sum_partials_kernel<<<(nacc + 255) / 256, 256, 0, st>>>(partial, grid, nacc, d_w1, n1, d_b1, hid, d_w2, n2, d_b2);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_item_combine_fwd(const float* pi, const float* tp, int64_t rows, int kt, int dim, float* out,
                                   pc_stream_t stream) {
  PC_REQUIRE(rows >= 0 && kt >= 1 && dim >= 4 && dim % 4 == 0, PC_ERR_INVALID, "item_combine_fwd: bad shape");
  if (rows == 0) return PC_OK;
  PC_REQUIRE(pi && tp && out, PC_ERR_INVALID, "item_combine_fwd: null pointer");
  const int64_t total = rows * kt * (dim / 4);
  const int64_t cap = int64_t(sm_count()) * 16;
  const int64_t want = ceil_div(total, 256);
  item_combine_fwd_kernel<<<unsigned(want < cap ? want : cap), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(pi), reinterpret_cast<const float4*>(tp), rows, kt, dim / 4, reinterpret_cast<float4*>(out));
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_item_combine_bwd(const float* d_out, const float* pi, const float* tp, int64_t rows, int kt, int dim,
                                   float* d_pi, float* d_tp, pc_stream_t stream) {
  PC_REQUIRE(rows >= 0 && kt >= 1 && dim >= 4 && dim % 4 == 0, PC_ERR_INVALID, "item_combine_bwd: bad shape");
  if (rows == 0) return PC_OK;
  PC_REQUIRE(d_out && pi && tp && d_pi && d_tp, PC_ERR_INVALID, "item_combine_bwd: null pointer");
  const int64_t total = rows * (dim / 4);
  const int64_t cap = int64_t(sm_count()) * 16;
  const int64_t want = ceil_div(total, 256);
  item_combine_bwd_kernel<<<unsigned(want < cap ? want : cap), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(d_out), reinterpret_cast<const float4*>(pi), reinterpret_cast<const float4*>(tp), rows, kt,
      dim / 4, reinterpret_cast<float4*>(d_pi), reinterpret_cast<float4*>(d_tp));
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_hinge_type_factored_bwd(const float* per_row, const int64_t* pos, const int64_t* neg, const float* grad,
                                          const float* base, const float* weight, int64_t rows, int width, float* d_base,
                                          float* vals, pc_stream_t stream) {
  PC_REQUIRE(rows >= 0 && width >= 4 && width % 4 == 0, PC_ERR_INVALID, "hinge_type_factored_bwd: bad shape");
  if (rows == 0) return PC_OK;
  PC_REQUIRE(per_row && pos && neg && grad && base && weight, PC_ERR_INVALID, "hinge_type_factored_bwd: null pointer");
  hinge_type_factored_bwd_kernel<<<unsigned(ceil_div(rows, 8)), 256, 0, as_stream(stream)>>>(
      per_row, pos, neg, grad, reinterpret_cast<const float4*>(base), reinterpret_cast<const float4*>(weight), rows, width / 4,
      reinterpret_cast<float4*>(d_base), reinterpret_cast<float4*>(vals));
  PC_LAUNCH_CHECK();
  return PC_OK;
}
