// Product2Vec graph attention over a CSR: forward, dst-major backward, src-major backward.
//
// Replaces the attention core of nn.MultiheadAttention as the reference calls it
// (/root/reference/src/models/product2vec.py:24-29,60; torch need_weights branch:
//  q*sqrt(1/dh) -> bmm -> softmax -> dropout -> bmm) and its autograd.
//
// Layout: one warp per CSR row.  A 128-float row is one float4 per lane, so a K or V row is a
// single fully-coalesced 512-byte warp load, and head h owns the 32/H consecutive lanes
// [h*32/H, (h+1)*32/H): per-head dot products are __shfl_xor reductions inside that lane group.
// U edges are fetched back to back before any arithmetic so every lane keeps 2*U independent
// 16-byte loads in flight (the kernels are HBM-latency/bandwidth bound).
// Softmax is computed online in base 2 (logits pre-multiplied by log2 e); the log2-sum-exp is
// kept per (row, head) so the backward kernels recompute the attention weights instead of
// storing 32 B/edge.  Nothing uses float atomics: every output element is produced by exactly
// one warp in a fixed edge order, so results are bit-reproducible run to run.
#include <math.h>

#include "common.cuh"

namespace pc {
namespace {

constexpr int ROW4 = 32;  // float4 per 128-float row
constexpr int KV4 = 64;   // float4 per K|V row
constexpr int WARPS = 8;  // warps (rows) per CTA
constexpr float LOG2E = 1.4426950408889634f;

struct DropArgs {
  uint64_t seed;
  uint32_t threshold;  // drop when hash < threshold
  float inv_keep;
};

template <int H, int U, bool DROP>
__global__ void __launch_bounds__(WARPS * 32)
gat_fwd_kernel(const float4* __restrict__ Q, const float4* __restrict__ KV, const int64_t* __restrict__ rowptr,
               const int32_t* __restrict__ col, int64_t n_dst, int64_t ldq4, float scale_log2e, DropArgs drop,
               float4* __restrict__ O, float* __restrict__ stats) {
  constexpr int G = 32 / H;
  const int lane = lane_id();
  const int64_t i = int64_t(blockIdx.x) * WARPS + warp_id();
  if (i >= n_dst) return;
  const int head = lane / G;
  const float4 q = scale4(ldg4(Q + i * ldq4 + lane), scale_log2e);
  const int64_t beg = rowptr[i], end = rowptr[i + 1];
  float m = -INFINITY, l = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t base = beg; base < end; base += 32) {
    const int cnt = int(min(int64_t(32), end - base));
    const int my_col = lane < cnt ? col[base + lane] : 0;
    for (int t = 0; t < cnt; t += U) {
      float4 k[U], v[U];
      int src[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        src[u] = __shfl_sync(FULL, my_col, (t + u) & 31);
        if (t + u < cnt) {
          const float4* rowp = KV + int64_t(src[u]) * KV4 + lane;
          k[u] = ldg4(rowp);
          v[u] = ldg4(rowp + ROW4);
        }
      }
      float s[U];
      float cm = -INFINITY;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        s[u] = -INFINITY;
        if (t + u < cnt) s[u] = group_sum<G>(dot4(q, k[u]));
        cm = fmaxf(cm, s[u]);
      }
      const float m_new = fmaxf(m, cm);
      const float corr = exp2f(m - m_new);
      l *= corr;
      acc = scale4(acc, corr);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (t + u < cnt) {
          const float p = exp2f(s[u] - m_new);
          l += p;
          float pv = p;
          if (DROP) pv *= keep_scale(drop.seed, uint32_t(i), uint32_t(src[u]), head, drop.threshold, drop.inv_keep);
          fma4(acc, pv, v[u]);
        }
      }
      m = m_new;
    }
  }
  float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
  float lse2 = 0.f;
  if (end > beg) {
    const float inv = 1.f / l;
    out = scale4(acc, inv);
    lse2 = m + log2f(l);
  }
  O[i * ROW4 + lane] = out;
  if (lane % G == 0) stats[i * (2 * H) + head] = lse2;
}

template <int H, int U, bool DROP>
__global__ void __launch_bounds__(WARPS * 32)
gat_bwd_dst_kernel(const float4* __restrict__ Q, const float4* __restrict__ KV, const int64_t* __restrict__ rowptr,
                   const int32_t* __restrict__ col, int64_t n_dst, int64_t ldq4, int64_t lddo4, int64_t lddq4, float scale,
                   DropArgs drop, const float4* __restrict__ O, const float4* __restrict__ dO, float* __restrict__ stats,
                   float4* __restrict__ dQ) {
  constexpr int G = 32 / H;
  const int lane = lane_id();
  const int64_t i = int64_t(blockIdx.x) * WARPS + warp_id();
  if (i >= n_dst) return;
  const int head = lane / G;
  const float4 q = scale4(ldg4(Q + i * ldq4 + lane), scale * LOG2E);
  const float4 go = ldg4(dO + i * lddo4 + lane);
  const float4 o = ldg4(O + i * ROW4 + lane);
  const float delta = group_sum<G>(dot4(go, o));
  const float lse2 = stats[i * (2 * H) + head];
  if (lane % G == 0) stats[i * (2 * H) + H + head] = delta;
  const int64_t beg = rowptr[i], end = rowptr[i + 1];
  float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t base = beg; base < end; base += 32) {
    const int cnt = int(min(int64_t(32), end - base));
    const int my_col = lane < cnt ? col[base + lane] : 0;
    for (int t = 0; t < cnt; t += U) {
      float4 k[U], v[U];
      int src[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        src[u] = __shfl_sync(FULL, my_col, (t + u) & 31);
        if (t + u < cnt) {
          const float4* rowp = KV + int64_t(src[u]) * KV4 + lane;
          k[u] = ldg4(rowp);
          v[u] = ldg4(rowp + ROW4);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (t + u < cnt) {
          const float s2 = group_sum<G>(dot4(q, k[u]));
          const float p = exp2f(s2 - lse2);
          float da = group_sum<G>(dot4(go, v[u]));
          if (DROP) da *= keep_scale(drop.seed, uint32_t(i), uint32_t(src[u]), head, drop.threshold, drop.inv_keep);
          fma4(dq, p * (da - delta), k[u]);
        }
      }
    }
  }
  dQ[i * lddq4 + lane] = scale4(dq, scale);
}

template <int H, int U, bool DROP>
__global__ void __launch_bounds__(WARPS * 32)
gat_bwd_src_kernel(const float4* __restrict__ Q, const float4* __restrict__ KV, const int64_t* __restrict__ colptr,
                   const int32_t* __restrict__ row, int64_t n_src, int64_t ldq4, int64_t lddo4, int64_t lddkv4, float scale,
                   DropArgs drop, const float4* __restrict__ dO, const float* __restrict__ stats, float4* __restrict__ dKV) {
  constexpr int G = 32 / H;
  const int lane = lane_id();
  const int64_t j = int64_t(blockIdx.x) * WARPS + warp_id();
  if (j >= n_src) return;
  const int head = lane / G;
  const int64_t beg = colptr[j], end = colptr[j + 1];
  float4 dk = make_float4(0.f, 0.f, 0.f, 0.f), dv = dk;
  if (end > beg) {
    const float4 k = ldg4(KV + j * KV4 + lane);
    const float4 v = ldg4(KV + j * KV4 + ROW4 + lane);
    const float scale_log2e = scale * LOG2E;
    for (int64_t base = beg; base < end; base += 32) {
      const int cnt = int(min(int64_t(32), end - base));
      const int my_row = lane < cnt ? row[base + lane] : 0;
      for (int t = 0; t < cnt; t += U) {
        float4 qi[U], gi[U];
        float lse2[U], delta[U];
        int dst[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          dst[u] = __shfl_sync(FULL, my_row, (t + u) & 31);
          if (t + u < cnt) {
            const int64_t i = dst[u];
            qi[u] = ldg4(Q + i * ldq4 + lane);
            gi[u] = ldg4(dO + i * lddo4 + lane);
            lse2[u] = __ldg(stats + i * (2 * H) + head);
            delta[u] = __ldg(stats + i * (2 * H) + H + head);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (t + u < cnt) {
            // same arithmetic as the forward (q pre-scaled, then dot) so p sums to 1 over the row
            const float4 q2 = scale4(qi[u], scale_log2e);
            const float s2 = group_sum<G>(dot4(q2, k));
            const float p = exp2f(s2 - lse2[u]);
            float da = group_sum<G>(dot4(gi[u], v));
            float pk = p;
            if (DROP) {
              const float ks = keep_scale(drop.seed, uint32_t(dst[u]), uint32_t(j), head, drop.threshold, drop.inv_keep);
              da *= ks;
              pk *= ks;
            }
            fma4(dk, p * (da - delta[u]), qi[u]);
            fma4(dv, pk, gi[u]);
          }
        }
      }
    }
    dk = scale4(dk, scale);
  }
  dKV[j * lddkv4 + lane] = dk;
  dKV[j * lddkv4 + ROW4 + lane] = dv;
}

// stats[i, 1, h] = dO_i . O_i per head (the softmax-gradient row constant); lets the src-major backward run
// before the dst-major one (the multi-GPU path starts its reverse halo exchange as early as possible)
template <int H>
__global__ void __launch_bounds__(WARPS * 32)
gat_delta_kernel(const float4* __restrict__ O, const float4* __restrict__ dO, int64_t lddo4, int64_t n, float* __restrict__ stats) {
  constexpr int G = 32 / H;
  const int lane = lane_id();
  const int64_t i = int64_t(blockIdx.x) * WARPS + warp_id();
  if (i >= n) return;
  const float delta = group_sum<G>(dot4(ldg4(dO + i * lddo4 + lane), ldg4(O + i * ROW4 + lane)));
  if (lane % G == 0) stats[i * (2 * H) + H + lane / G] = delta;
}

DropArgs make_drop(float p, uint64_t seed) {
  DropArgs d;
  d.seed = seed;
  double t = double(p) * 4294967296.0;
  d.threshold = t >= 4294967295.0 ? 4294967295u : uint32_t(t);
  d.inv_keep = 1.f / (1.f - p);
  return d;
}

int check_common(const void* a, const void* b, const void* c, const void* d, int64_t n, int heads, float p) {
  PC_REQUIRE(n >= 0, PC_ERR_INVALID, "gat: negative row count");
  PC_REQUIRE(n == 0 || (a && b && c && d), PC_ERR_INVALID, "gat: null pointer argument");
  PC_REQUIRE(heads == 1 || heads == 2 || heads == 4 || heads == 8, PC_ERR_UNSUPPORTED,
             "gat: heads=%d unsupported (embed dim 128, heads in {1,2,4,8})", heads);
  PC_REQUIRE(p >= 0.f && p < 1.f, PC_ERR_INVALID, "gat: dropout_p=%f outside [0,1)", p);
  PC_REQUIRE(n < (int64_t(1) << 31) * WARPS, PC_ERR_UNSUPPORTED, "gat: too many rows");
  return PC_OK;
}

#define PC_DISPATCH_HEADS(heads, DROPV, CALL) \
  switch (heads) {                            \
    case 1: { constexpr int H = 1; constexpr bool DROP = DROPV; CALL; } break; \
    case 2: { constexpr int H = 2; constexpr bool DROP = DROPV; CALL; } break; \
    case 4: { constexpr int H = 4; constexpr bool DROP = DROPV; CALL; } break; \
    default: { constexpr int H = 8; constexpr bool DROP = DROPV; CALL; } break; \
  }

constexpr int UNROLL = 4;

}  // namespace
}  // namespace pc

using namespace pc;

extern "C" int pc_gat_fwd(const float* q, int64_t ld_q, const float* kv, const int64_t* rowptr, const int32_t* col,
                          int64_t n_dst, int heads, float dropout_p, uint64_t seed, float* o, float* stats,
                          pc_stream_t stream) {
  if (int rc = check_common(q, rowptr, o, stats, n_dst, heads, dropout_p)) return rc;
  PC_REQUIRE(ld_q >= 128 && ld_q % 4 == 0, PC_ERR_INVALID, "gat_fwd: ld_q=%lld must be a multiple of 4 and >= 128", (long long)ld_q);
  if (n_dst == 0) return PC_OK;
  const float scale_log2e = sqrtf(1.f / float(128 / heads)) * LOG2E;
  const DropArgs drop = make_drop(dropout_p, seed);
  const unsigned grid = unsigned(ceil_div(n_dst, WARPS));
  cudaStream_t st = as_stream(stream);
#define CALL_FWD                                                                                   \
  gat_fwd_kernel<H, UNROLL, DROP><<<grid, WARPS * 32, 0, st>>>(                                    \
      reinterpret_cast<const float4*>(q), reinterpret_cast<const float4*>(kv), rowptr, col, n_dst, \
      ld_q / 4, scale_log2e, drop, reinterpret_cast<float4*>(o), stats)
  if (dropout_p > 0.f) {
    PC_DISPATCH_HEADS(heads, true, CALL_FWD)
  } else {
    PC_DISPATCH_HEADS(heads, false, CALL_FWD)
  }
#undef CALL_FWD
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_gat_bwd_dst(const float* q, int64_t ld_q, const float* kv, const int64_t* rowptr, const int32_t* col,
                              int64_t n_dst, int heads, float dropout_p, uint64_t seed, const float* o,
                              const float* d_o, int64_t ld_do, float* stats, float* dq, int64_t ld_dq,
                              pc_stream_t stream) {
  if (int rc = check_common(q, rowptr, o, stats, n_dst, heads, dropout_p)) return rc;
  PC_REQUIRE(n_dst == 0 || (d_o && dq), PC_ERR_INVALID, "gat_bwd_dst: null pointer argument");
  PC_REQUIRE(ld_q >= 128 && ld_q % 4 == 0 && ld_do >= 128 && ld_do % 4 == 0 && ld_dq >= 128 && ld_dq % 4 == 0, PC_ERR_INVALID,
             "gat_bwd_dst: leading dimensions must be multiples of 4 and >= 128");
  if (n_dst == 0) return PC_OK;
  const float scale = sqrtf(1.f / float(128 / heads));
  const DropArgs drop = make_drop(dropout_p, seed);
  const unsigned grid = unsigned(ceil_div(n_dst, WARPS));
  cudaStream_t st = as_stream(stream);
#define CALL_BD                                                                                          \
  gat_bwd_dst_kernel<H, UNROLL, DROP><<<grid, WARPS * 32, 0, st>>>(                                      \
      reinterpret_cast<const float4*>(q), reinterpret_cast<const float4*>(kv), rowptr, col, n_dst,       \
      ld_q / 4, ld_do / 4, ld_dq / 4, scale,                                                             \
      drop, reinterpret_cast<const float4*>(o), reinterpret_cast<const float4*>(d_o), stats,             \
      reinterpret_cast<float4*>(dq))
  if (dropout_p > 0.f) {
    PC_DISPATCH_HEADS(heads, true, CALL_BD)
  } else {
    PC_DISPATCH_HEADS(heads, false, CALL_BD)
  }
#undef CALL_BD
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_gat_bwd_src(const float* q, int64_t ld_q, const float* kv, const int64_t* colptr, const int32_t* row,
                              int64_t n_src, int heads, float dropout_p, uint64_t seed, const float* d_o, int64_t ld_do,
                              const float* stats, float* dkv, int64_t ld_dkv, pc_stream_t stream) {
  if (int rc = check_common(kv, colptr, dkv, stats, n_src, heads, dropout_p)) return rc;
  PC_REQUIRE(ld_q >= 128 && ld_q % 4 == 0 && ld_do >= 128 && ld_do % 4 == 0 && ld_dkv >= 256 && ld_dkv % 4 == 0, PC_ERR_INVALID,
             "gat_bwd_src: leading dimensions must be multiples of 4 (q, d_o >= 128, dkv >= 256)");
  if (n_src == 0) return PC_OK;
  const float scale = sqrtf(1.f / float(128 / heads));
  const DropArgs drop = make_drop(dropout_p, seed);
  const unsigned grid = unsigned(ceil_div(n_src, WARPS));
  cudaStream_t st = as_stream(stream);
#define CALL_BS                                                                                          \
  gat_bwd_src_kernel<H, UNROLL, DROP><<<grid, WARPS * 32, 0, st>>>(                                      \
      reinterpret_cast<const float4*>(q), reinterpret_cast<const float4*>(kv), colptr, row, n_src,       \
      ld_q / 4, ld_do / 4, ld_dkv / 4, scale,                                                            \
      drop, reinterpret_cast<const float4*>(d_o), stats, reinterpret_cast<float4*>(dkv))
  if (dropout_p > 0.f) {
    PC_DISPATCH_HEADS(heads, true, CALL_BS)
  } else {
    PC_DISPATCH_HEADS(heads, false, CALL_BS)
  }
#undef CALL_BS
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" int pc_gat_delta(const float* o, const float* d_o, int64_t ld_do, int64_t n, int heads, float* stats,
                            pc_stream_t stream) {
  if (int rc = check_common(o, d_o, stats, stats, n, heads, 0.f)) return rc;
  PC_REQUIRE(ld_do >= 128 && ld_do % 4 == 0, PC_ERR_INVALID, "gat_delta: bad leading dimension");
  if (n == 0) return PC_OK;
  const unsigned grid = unsigned(ceil_div(n, WARPS));
  cudaStream_t st = as_stream(stream);
  switch (heads) {
    case 1: gat_delta_kernel<1><<<grid, WARPS * 32, 0, st>>>(reinterpret_cast<const float4*>(o), reinterpret_cast<const float4*>(d_o), ld_do / 4, n, stats); break;
    case 2: gat_delta_kernel<2><<<grid, WARPS * 32, 0, st>>>(reinterpret_cast<const float4*>(o), reinterpret_cast<const float4*>(d_o), ld_do / 4, n, stats); break;
    case 4: gat_delta_kernel<4><<<grid, WARPS * 32, 0, st>>>(reinterpret_cast<const float4*>(o), reinterpret_cast<const float4*>(d_o), ld_do / 4, n, stats); break;
    default: gat_delta_kernel<8><<<grid, WARPS * 32, 0, st>>>(reinterpret_cast<const float4*>(o), reinterpret_cast<const float4*>(d_o), ld_do / 4, n, stats); break;
  }
  PC_LAUNCH_CHECK();
  return PC_OK;
}
