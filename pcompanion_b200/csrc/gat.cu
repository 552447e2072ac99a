// Product2Vec graph attention over a CSR: forward, dst-major backward, src-major backward.
//
// Replaces the attention core of nn.MultiheadAttention as the reference calls it
// (/root/reference/src/models/product2vec.py:24-29,60; torch need_weights branch:
//  q*sqrt(1/dh) -> bmm -> softmax -> dropout -> bmm) and its autograd.
//
// Layout: one warp per CSR row (a warp walks a chunk of 4 consecutive rows).  A 128-float row is
// one float4 per lane and head h owns the 32/H consecutive lanes [h*32/H, (h+1)*32/H): per-head dot
// products are __shfl_xor reductions inside that lane group.  Neighbour rows are fetched by 1 KB
// bulk async copies into a per-warp shared-memory ring, 8 rows ahead of the arithmetic (the
// kernels are HBM-latency/bandwidth bound).
// Softmax is computed online in base 2 (logits pre-multiplied by log2 e); the log2-sum-exp is
// kept per (row, head) so the backward kernels recompute the attention weights instead of
// storing 32 B/edge.  Nothing uses float atomics: every output element is produced by exactly
// one warp in a fixed edge order, so results are bit-reproducible run to run.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

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

// ---------------------------------------------------------------------------------------------------------------
// The gathered rows travel global -> shared memory as 1 KB bulk async copies (cp.async.bulk, completion on an
// mbarrier), RING rows per warp permanently in flight, requested across the row boundaries of the warp's chunk of
// CSR rows.  In-flight data lives in shared memory instead of registers, so the bytes in flight per SM do not
// collapse while a warp does its softmax arithmetic: a first version that loaded 4 neighbours at a time into
// registers kept ~60 KB/SM in flight and ran HBM at 4.4-4.9 TB/s (fwd 4.2 ms on C2); this one runs at 3.4 ms.
constexpr int RING = 8;          // rows in flight per warp
constexpr int CHUNK = 4;         // CSR rows (or CSC columns) per warp
constexpr int SLOT_KV = 1024;    // one K|V row
constexpr int SLOT_QG = 1024;        // Q row | dO row; the stats (lse2[H], delta[H]) of the slot live in a separate 64-byte cell
constexpr int STAT_CELL = 64;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

template <int BYTES>
__device__ __forceinline__ void cp_async_small(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES) : "memory");
}

// Per-warp prefetch stream over the edges of the warp's chunk; positions are 32-bit offsets from the chunk's first
// edge (`ids` already points there).  Invariant: edges [0, pe) have been requested, pe = consumed + RING (clamped),
// so a batch that consumed nb edges is followed by nb new requests, one per lane (no loops, no shuffles).
template <int SLOT>
struct EdgeRing {
  static_assert((RING & (RING - 1)) == 0, "RING must be a power of two");
  uint32_t slots = 0, bars = 0;
  int n_edges = 0, pe = 0, nid = 0;   // nid: ids[pe + lane], loaded one batch ahead
  const int32_t* ids = nullptr;
  __device__ __forceinline__ void init(uint8_t* smem, int warp, int lane, const int32_t* ids_, int n, int arrivals = 1) {
    slots = smem_u32(smem) + uint32_t(warp) * RING * SLOT;
    bars = smem_u32(smem) + uint32_t(WARPS) * RING * SLOT + uint32_t(warp) * RING * 8;
    if (lane == 0) {
      for (int s = 0; s < RING; ++s) mbar_init(bars + 8 * s, arrivals);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    ids = ids_; n_edges = n; pe = 0;
    nid = lane < n ? ids[lane] : 0;
  }
  __device__ __forceinline__ uint32_t slot_of(int e) const { return slots + uint32_t(e & (RING - 1)) * SLOT; }
  __device__ __forceinline__ void wait(int e) const { mbar_wait(bars + 8 * uint32_t(e & (RING - 1)), uint32_t(e / RING) & 1u); }
  // after `count` requests have been made by lanes [0, count): advance and fetch the ids of the next ones
  __device__ __forceinline__ void advance(int count, int lane) {
    pe += count;
    nid = (lane < RING && pe + lane < n_edges) ? ids[pe + lane] : 0;
  }
};

// lanes [0, count) request edge pe + lane: one 1 KB K|V row each
#define PC_RING_ISSUE_KV(count_expr)                                                                   \
  {                                                                                                    \
    const int count_ = min(int(count_expr), ring.n_edges - ring.pe);                                   \
    if (lane < count_) {                                                                               \
      const uint32_t n_ = uint32_t((ring.pe + lane) & (RING - 1));                                     \
      mbar_arrive_expect_tx(ring.bars + 8 * n_, SLOT_KV);                                              \
      bulk_g2s(ring.slots + n_ * SLOT_KV, KV + int64_t(ring.nid) * KV4, SLOT_KV, ring.bars + 8 * n_);  \
    }                                                                                                  \
    if (count_ > 0) ring.advance(count_, lane);                                                        \
  }

template <int H, int U, bool DROP>
__global__ void __launch_bounds__(WARPS * 32, 3)
gat_fwd_kernel(const float4* __restrict__ Q, const float4* __restrict__ KV, const int64_t* __restrict__ rowptr,
                    const int32_t* __restrict__ col, int64_t n_dst, int64_t ldq4, float scale_log2e, DropArgs drop,
                    const int32_t* __restrict__ dst_ids, float4* __restrict__ O, float* __restrict__ stats) {
  extern __shared__ __align__(128) uint8_t ring_smem[];
  constexpr int G = 32 / H;
  const int lane = lane_id(), warp = warp_id();
  const int head = lane / G;
  const int64_t r0 = (int64_t(blockIdx.x) * WARPS + warp) * CHUNK;
  if (r0 >= n_dst) return;
  const int rows = int(min(int64_t(CHUNK), n_dst - r0));
  const int64_t my_ptr = lane <= rows ? rowptr[r0 + lane] : 0;
  const int64_t e0 = __shfl_sync(FULL, my_ptr, 0);
  const int my_rel = int(my_ptr - e0);                      // lanes 0..rows: row boundaries relative to the chunk
  col += e0;
  EdgeRing<SLOT_KV> ring;
  ring.init(ring_smem, warp, lane, col, __shfl_sync(FULL, my_rel, rows));
  PC_RING_ISSUE_KV(RING)
  float4 q_next = ldg4(Q + r0 * ldq4 + lane);
  for (int r = 0; r < rows; ++r) {
    const int64_t i = r0 + r;
    const uint32_t i_id = DROP ? uint32_t(dst_ids ? dst_ids[i] : i) : 0u;   // virtual rows of a split hub keep the hub's id
    const float4 q = scale4(q_next, scale_log2e);
    if (r + 1 < rows) q_next = ldg4(Q + (i + 1) * ldq4 + lane);
    const int beg = __shfl_sync(FULL, my_rel, r), end = __shfl_sync(FULL, my_rel, r + 1);
    float m = -INFINITY, l = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      uint32_t my_keep = 0;       // lane l: keep bits (one per head) of edge base + l
      if (DROP && lane < cnt) my_keep = keep_bits<H>(drop.seed, i_id, uint32_t(col[base + lane]), drop.threshold);
      for (int t = 0; t < cnt; t += U) {
        float4 k[U], v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (t + u < cnt) {
            const int e = base + t + u;
            ring.wait(e);
            const uint32_t sl = ring.slot_of(e) + lane * 16;
            k[u] = lds128(sl);
            v[u] = lds128(sl + 512);
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
            if (DROP) pv *= ((__shfl_sync(FULL, my_keep, (t + u) & 31) >> head) & 1u) ? drop.inv_keep : 0.f;
            fma4(acc, pv, v[u]);
          }
        }
        m = m_new;
        __syncwarp();                                      // every lane has used its k / v: the slots may be refilled
        PC_RING_ISSUE_KV(min(U, cnt - t))
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
}

template <int H, int U, bool DROP>
__global__ void __launch_bounds__(WARPS * 32, 3)
gat_bwd_dst_kernel(const float4* __restrict__ Q, const float4* __restrict__ KV, const int64_t* __restrict__ rowptr,
                        const int32_t* __restrict__ col, int64_t n_dst, int64_t ldq4, int64_t lddo4, int64_t lddq4, float scale,
                        DropArgs drop, const int32_t* __restrict__ dst_ids, const float4* __restrict__ O,
                        const float4* __restrict__ dO, float* __restrict__ stats, float4* __restrict__ dQ) {
  extern __shared__ __align__(128) uint8_t ring_smem[];
  constexpr int G = 32 / H;
  const int lane = lane_id(), warp = warp_id();
  const int head = lane / G;
  const int64_t r0 = (int64_t(blockIdx.x) * WARPS + warp) * CHUNK;
  if (r0 >= n_dst) return;
  const int rows = int(min(int64_t(CHUNK), n_dst - r0));
  const int64_t my_ptr = lane <= rows ? rowptr[r0 + lane] : 0;
  const int64_t e0 = __shfl_sync(FULL, my_ptr, 0);
  const int my_rel = int(my_ptr - e0);
  col += e0;
  EdgeRing<SLOT_KV> ring;
  ring.init(ring_smem, warp, lane, col, __shfl_sync(FULL, my_rel, rows));
  PC_RING_ISSUE_KV(RING)
  for (int r = 0; r < rows; ++r) {
    const int64_t i = r0 + r;
    const uint32_t i_id = DROP ? uint32_t(dst_ids ? dst_ids[i] : i) : 0u;
    const float4 q = scale4(ldg4(Q + i * ldq4 + lane), scale * LOG2E);
    const float4 go = ldg4(dO + i * lddo4 + lane);
    const float4 o = ldg4(O + i * ROW4 + lane);
    const float lse2 = stats[i * (2 * H) + head];
    const float delta = group_sum<G>(dot4(go, o));
    if (lane % G == 0) stats[i * (2 * H) + H + head] = delta;
    const int beg = __shfl_sync(FULL, my_rel, r), end = __shfl_sync(FULL, my_rel, r + 1);
    float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      uint32_t my_keep = 0;
      if (DROP && lane < cnt) my_keep = keep_bits<H>(drop.seed, i_id, uint32_t(col[base + lane]), drop.threshold);
      for (int t = 0; t < cnt; t += U) {
        float4 k[U], v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (t + u < cnt) {
            const int e = base + t + u;
            ring.wait(e);
            const uint32_t sl = ring.slot_of(e) + lane * 16;
            k[u] = lds128(sl);
            v[u] = lds128(sl + 512);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (t + u < cnt) {
            const float s2 = group_sum<G>(dot4(q, k[u]));
            const float p = exp2f(s2 - lse2);
            float da = group_sum<G>(dot4(go, v[u]));
            if (DROP) da *= ((__shfl_sync(FULL, my_keep, (t + u) & 31) >> head) & 1u) ? drop.inv_keep : 0.f;
            fma4(dq, p * (da - delta), k[u]);
          }
        }
        __syncwarp();
        PC_RING_ISSUE_KV(min(U, cnt - t))
      }
    }
    dQ[i * lddq4 + lane] = scale4(dq, scale);
  }
}

// packed: the dO row sits right behind the Q row (Q | dO table of the fused layer) -> one 1 KB copy per edge
template <int H, int U, bool DROP>
__global__ void __launch_bounds__(WARPS * 32, 3)
gat_bwd_src_kernel(const float4* __restrict__ Q, const float4* __restrict__ KV, const int64_t* __restrict__ colptr,
                        const int32_t* __restrict__ row, int64_t n_src, int64_t ldq4, int64_t lddo4, int64_t lddkv4, float scale,
                        DropArgs drop, const float4* __restrict__ dO, const float* __restrict__ stats, float4* __restrict__ dKV,
                        int packed, int64_t src_base, const int32_t* __restrict__ src_ids) {
  extern __shared__ __align__(128) uint8_t ring_smem[];
  constexpr int G = 32 / H;
  // the 2H floats of stats ride along as 16-byte (8 for H = 1) cp.async copies issued by the requesting lane and
  // complete on the slot's mbarrier (cp.async.mbarrier.arrive.noinc)
  constexpr int ST_BYTES = 2 * H * 4, ST_CHUNK = ST_BYTES < 16 ? ST_BYTES : 16, ST_LANES = ST_BYTES / ST_CHUNK;
  constexpr uint32_t TX = 1024;
  const int lane = lane_id(), warp = warp_id();
  const int head = lane / G;
  const int64_t c0 = (int64_t(blockIdx.x) * WARPS + warp) * CHUNK;
  if (c0 >= n_src) return;
  const int cols = int(min(int64_t(CHUNK), n_src - c0));
  const int64_t my_ptr = lane <= cols ? colptr[c0 + lane] : 0;
  const int64_t e0 = __shfl_sync(FULL, my_ptr, 0);
  const int my_rel = int(my_ptr - e0);
  row += e0;
  EdgeRing<SLOT_QG> ring;
  ring.init(ring_smem, warp, lane, row, __shfl_sync(FULL, my_rel, cols), 2);   // expect_tx arrive + cp.async arrive
  const uint32_t cells = smem_u32(ring_smem) + uint32_t(WARPS) * RING * (SLOT_QG + 8) + uint32_t(warp) * RING * STAT_CELL;
#define PC_RING_ISSUE_QG(count_expr)                                                                   \
  {                                                                                                    \
    const int count_ = min(int(count_expr), ring.n_edges - ring.pe);                                   \
    if (lane < count_) {                                                                               \
      const uint32_t n_ = uint32_t((ring.pe + lane) & (RING - 1));                                     \
      const uint32_t bar_ = ring.bars + 8 * n_, dst_ = ring.slots + n_ * SLOT_QG;                      \
      const int64_t i_ = ring.nid;                                                                     \
      const char* st_ = reinterpret_cast<const char*>(stats + i_ * (2 * H));                           \
      _Pragma("unroll") for (int c_ = 0; c_ < ST_LANES; ++c_)                                          \
        cp_async_small<ST_CHUNK>(cells + n_ * STAT_CELL + c_ * ST_CHUNK, st_ + c_ * ST_CHUNK);         \
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_) : "memory");     \
      mbar_arrive_expect_tx(bar_, TX);                                                                 \
      if (packed) {                                                                                    \
        bulk_g2s(dst_, Q + i_ * ldq4, 1024, bar_);                                                     \
      } else {                                                                                         \
        bulk_g2s(dst_, Q + i_ * ldq4, 512, bar_);                                                      \
        bulk_g2s(dst_ + 512, dO + i_ * lddo4, 512, bar_);                                              \
      }                                                                                                \
    }                                                                                                  \
    if (count_ > 0) ring.advance(count_, lane);                                                        \
  }
  PC_RING_ISSUE_QG(RING)
  const float scale_log2e = scale * LOG2E;
  for (int r = 0; r < cols; ++r) {
    const int64_t j = c0 + r;
    const uint32_t j_id = DROP ? uint32_t(src_ids ? int64_t(src_ids[j]) : j + src_base) : 0u;
    const int beg = __shfl_sync(FULL, my_rel, r), end = __shfl_sync(FULL, my_rel, r + 1);
    float4 dk = make_float4(0.f, 0.f, 0.f, 0.f), dv = dk;
    if (end > beg) {
      const float4 k = ldg4(KV + j * KV4 + lane);
      const float4 v = ldg4(KV + j * KV4 + ROW4 + lane);
      const float4 k2 = scale4(k, scale_log2e);              // the logit scale applied once per column instead of once per edge
      for (int base = beg; base < end; base += 32) {
        const int cnt = min(32, end - base);
        uint32_t my_keep = 0;
        if (DROP && lane < cnt) my_keep = keep_bits<H>(drop.seed, uint32_t(row[base + lane]), j_id, drop.threshold);
        for (int t = 0; t < cnt; t += U) {
          float4 qi[U], gi[U];
          float lse2[U], delta[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (t + u < cnt) {
              const int e = base + t + u;
              ring.wait(e);
              const uint32_t sl = ring.slot_of(e);
              qi[u] = lds128(sl + lane * 16);
              gi[u] = lds128(sl + 512 + lane * 16);
              const uint32_t cell = cells + uint32_t(e & (RING - 1)) * STAT_CELL;
              lse2[u] = lds32(cell + head * 4);
              delta[u] = lds32(cell + (H + head) * 4);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (t + u < cnt) {
              const float s2 = group_sum<G>(dot4(qi[u], k2));
              const float p = exp2f(s2 - lse2[u]);
              float da = group_sum<G>(dot4(gi[u], v));
              float pk = p;
              if (DROP) {
                const float ks = ((__shfl_sync(FULL, my_keep, (t + u) & 31) >> head) & 1u) ? drop.inv_keep : 0.f;
                da *= ks;
                pk *= ks;
              }
              fma4(dk, p * (da - delta[u]), qi[u]);
              fma4(dv, pk, gi[u]);
            }
          }
          __syncwarp();
          PC_RING_ISSUE_QG(min(U, cnt - t))
        }
      }
      dk = scale4(dk, scale);
    }
    dKV[j * lddkv4 + lane] = dk;
    dKV[j * lddkv4 + ROW4 + lane] = dv;
  }
#undef PC_RING_ISSUE_QG
}

constexpr int RING_SMEM_KV = WARPS * RING * (SLOT_KV + 8);
constexpr int RING_SMEM_QG = WARPS * RING * (SLOT_QG + 8 + STAT_CELL);

template <typename K>
int ring_smem_attr(K kernel, int bytes) {
  PC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return PC_OK;
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

// Split hub rows: virtual rows [seg_ptr[h], seg_ptr[h+1]) are consecutive slices of hub row hub_rows[h]'s neighbour list,
// each attended on its own (normalised output O_v, log2-sum-exp lse_v).  The row's softmax over the whole list is
// O = sum_v O_v 2^(lse_v - lse), lse = log2 sum_v 2^(lse_v), accumulated in ascending v (fixed order).  One warp per hub.
template <int H>
__global__ void __launch_bounds__(WARPS * 32)
gat_merge_segments_kernel(const float4* __restrict__ Oseg, const float* __restrict__ stats_seg, const int64_t* __restrict__ seg_ptr,
                          const int64_t* __restrict__ hub_rows, int64_t n_hubs, float4* __restrict__ O, float* __restrict__ stats) {
  constexpr int G = 32 / H;
  const int lane = lane_id();
  const int head = lane / G;
  const int64_t hh = int64_t(blockIdx.x) * WARPS + warp_id();
  if (hh >= n_hubs) return;
  const int64_t v0 = seg_ptr[hh], v1 = seg_ptr[hh + 1], i = hub_rows[hh];
  float m = -INFINITY;
  for (int64_t v = v0; v < v1; ++v) m = fmaxf(m, stats_seg[v * (2 * H) + head]);
  float l = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t v = v0; v < v1; ++v) {
    const float w = exp2f(stats_seg[v * (2 * H) + head] - m);
    l += w;
    fma4(acc, w, ldg4(Oseg + v * ROW4 + lane));
  }
  O[i * ROW4 + lane] = scale4(acc, 1.f / l);
  if (lane % G == 0) stats[i * (2 * H) + head] = m + log2f(l);
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

constexpr int RU_FWD = 3, RU_DST = 3, RU_SRC = 3;   // edges per arithmetic batch in the ring kernels (data is already in shared memory)

}  // namespace
}  // namespace pc

using namespace pc;

extern "C" int pc_gat_fwd(const float* q, int64_t ld_q, const float* kv, const int64_t* rowptr, const int32_t* col,
                          int64_t n_dst, int heads, float dropout_p, uint64_t seed, const int32_t* dst_ids, float* o,
                          float* stats, pc_stream_t stream) {
  if (int rc = check_common(q, rowptr, o, stats, n_dst, heads, dropout_p)) return rc;
  PC_REQUIRE(ld_q >= 128 && ld_q % 4 == 0, PC_ERR_INVALID, "gat_fwd: ld_q=%lld must be a multiple of 4 and >= 128", (long long)ld_q);
  if (n_dst == 0) return PC_OK;
  const float scale_log2e = sqrtf(1.f / float(128 / heads)) * LOG2E;
  const DropArgs drop = make_drop(dropout_p, seed);
  const unsigned grid = unsigned(ceil_div(n_dst, int64_t(WARPS) * CHUNK));
  cudaStream_t st = as_stream(stream);
#define CALL_FWD                                                                                   \
  if (int rc = ring_smem_attr(gat_fwd_kernel<H, RU_FWD, DROP>, RING_SMEM_KV)) return rc;      \
  gat_fwd_kernel<H, RU_FWD, DROP><<<grid, WARPS * 32, RING_SMEM_KV, st>>>(                    \
      reinterpret_cast<const float4*>(q), reinterpret_cast<const float4*>(kv), rowptr, col, n_dst, \
      ld_q / 4, scale_log2e, drop, dst_ids, reinterpret_cast<float4*>(o), stats)
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
                              int64_t n_dst, int heads, float dropout_p, uint64_t seed, const int32_t* dst_ids,
                              const float* o, const float* d_o, int64_t ld_do, float* stats, float* dq, int64_t ld_dq,
                              pc_stream_t stream) {
  if (int rc = check_common(q, rowptr, o, stats, n_dst, heads, dropout_p)) return rc;
  PC_REQUIRE(n_dst == 0 || (d_o && dq), PC_ERR_INVALID, "gat_bwd_dst: null pointer argument");
  PC_REQUIRE(ld_q >= 128 && ld_q % 4 == 0 && ld_do >= 128 && ld_do % 4 == 0 && ld_dq >= 128 && ld_dq % 4 == 0, PC_ERR_INVALID,
             "gat_bwd_dst: leading dimensions must be multiples of 4 and >= 128");
  if (n_dst == 0) return PC_OK;
  const float scale = sqrtf(1.f / float(128 / heads));
  const DropArgs drop = make_drop(dropout_p, seed);
  const unsigned grid = unsigned(ceil_div(n_dst, int64_t(WARPS) * CHUNK));
  cudaStream_t st = as_stream(stream);
#define CALL_BD                                                                                          \
  if (int rc = ring_smem_attr(gat_bwd_dst_kernel<H, RU_DST, DROP>, RING_SMEM_KV)) return rc;        \
  gat_bwd_dst_kernel<H, RU_DST, DROP><<<grid, WARPS * 32, RING_SMEM_KV, st>>>(                      \
      reinterpret_cast<const float4*>(q), reinterpret_cast<const float4*>(kv), rowptr, col, n_dst,       \
      ld_q / 4, ld_do / 4, ld_dq / 4, scale,                                                             \
      drop, dst_ids, reinterpret_cast<const float4*>(o), reinterpret_cast<const float4*>(d_o), stats,    \
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
                              const float* stats, float* dkv, int64_t ld_dkv, int64_t src_base, const int32_t* src_ids,
                              pc_stream_t stream) {
  if (int rc = check_common(kv, colptr, dkv, stats, n_src, heads, dropout_p)) return rc;
  PC_REQUIRE(ld_q >= 128 && ld_q % 4 == 0 && ld_do >= 128 && ld_do % 4 == 0 && ld_dkv >= 256 && ld_dkv % 4 == 0, PC_ERR_INVALID,
             "gat_bwd_src: leading dimensions must be multiples of 4 (q, d_o >= 128, dkv >= 256)");
  if (n_src == 0) return PC_OK;
  const float scale = sqrtf(1.f / float(128 / heads));
  const DropArgs drop = make_drop(dropout_p, seed);
  const unsigned grid = unsigned(ceil_div(n_src, int64_t(WARPS) * CHUNK));
  cudaStream_t st = as_stream(stream);
  const int packed = (d_o == q + 128 && ld_do == ld_q) ? 1 : 0;   // Q | dO rows of the fused layer: one 1 KB copy per edge
  PC_REQUIRE((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(d_o) | reinterpret_cast<uintptr_t>(stats)) % 16 == 0,
             PC_ERR_INVALID, "gat_bwd_src: q, d_o and stats must be 16-byte aligned (bulk copies)");
#define CALL_BS                                                                                          \
  if (int rc = ring_smem_attr(gat_bwd_src_kernel<H, RU_SRC, DROP>, RING_SMEM_QG)) return rc;        \
  gat_bwd_src_kernel<H, RU_SRC, DROP><<<grid, WARPS * 32, RING_SMEM_QG, st>>>(                      \
      reinterpret_cast<const float4*>(q), reinterpret_cast<const float4*>(kv), colptr, row, n_src,       \
      ld_q / 4, ld_do / 4, ld_dkv / 4, scale,                                                            \
      drop, reinterpret_cast<const float4*>(d_o), stats, reinterpret_cast<float4*>(dkv), packed, src_base, src_ids)
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

extern "C" int pc_gat_merge_segments(const float* o_seg, const float* stats_seg, const int64_t* seg_ptr, const int64_t* hub_rows,
                                     int64_t n_hubs, int heads, float* o, float* stats, pc_stream_t stream) {
  if (int rc = check_common(o_seg, stats_seg, o, stats, n_hubs, heads, 0.f)) return rc;
  if (n_hubs == 0) return PC_OK;
  PC_REQUIRE(seg_ptr && hub_rows, PC_ERR_INVALID, "gat_merge_segments: null pointer");
  const unsigned grid = unsigned(ceil_div(n_hubs, WARPS));
  cudaStream_t st = as_stream(stream);
#define CALL_MERGE(HH) gat_merge_segments_kernel<HH><<<grid, WARPS * 32, 0, st>>>(reinterpret_cast<const float4*>(o_seg), stats_seg, seg_ptr, hub_rows, n_hubs, reinterpret_cast<float4*>(o), stats)
  switch (heads) {
    case 1: CALL_MERGE(1); break;
    case 2: CALL_MERGE(2); break;
    case 4: CALL_MERGE(4); break;
    default: CALL_MERGE(8); break;
  }
#undef CALL_MERGE
  PC_LAUNCH_CHECK();
  return PC_OK;
}
