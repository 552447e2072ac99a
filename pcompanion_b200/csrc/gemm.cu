// Dense row projections on the 5th-generation tensor cores: Y = epi(A . W^T + bias), fp32-faithful.
//
// Replaces the addmm calls behind nn.Linear in /root/reference/src/models/product2vec.py:14-21
// (FFN), the packed in-projection / out-projection of nn.MultiheadAttention (:24-29, :60) and
// their autograd dgrad GEMMs (dX = dY . W, run as the same kernel on the transposed weight).
//
// Precision: the reference computes these GEMMs in fp32 and BASELINE.json asks for 1e-5 relative
// parity, which single-pass TF32/BF16 tensor-core math (~1e-3) cannot give.  Each operand is
// therefore split on the fly into hi = trunc_tf32(x) (what the tensor core reads of the raw
// fp32 word anyway) and lo = rn_tf32(x - hi) (weights: hi = rn_tf32(w)), and three products
// (kind::tf32, fp32 accumulation in TMEM) are issued per K step: lo.hi + hi.lo + hi.hi.
// The dropped lo.lo term is < 2^-20 relative (2^-22 on average) and the rounding of lo 2^-21.  The tensor core
// accumulates with truncation (measured: -2.4e-8 relative per MMA on same-sign data,
// profiles/gemm_accuracy.py), so the hi.hi chain and the small cross terms go to two separate
// TMEM accumulators (the cross terms would otherwise cost a truncation at the full magnitude
// each) and are added once, rounded, in the epilogue.
//
// Structure (one persistent CTA per SM, 640 threads, warp-specialised):
//   warp 0 / 3    TMA producers: A [128 x 32] fp32 tiles (3 stages, HBM) / pre-split W_hi | W_lo tiles (3 stages, L2)
//   warps 8-15    compute the lo tiles of the landed activation tiles (the raw tile itself is the hi operand)
//   warp 1        one elected lane issues two tcgen05.mma per K step (a_hi . [W_hi | W_lo] 256 wide, a_lo . W_hi);
//                 tcgen05.commit frees stages
//   warps 4-7, 16-19  two epilogue sets (alternate 32-column chunks): tcgen05.ld both accumulators (double-buffered
//                 in TMEM), bias / tanh / tanh-gradient / add / row-select - the aux operand of the gradient
//                 epilogues arrives by TMA in the set's staging tile -, swizzled staging tile, TMA store
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace pc {
namespace {

constexpr int MAX_BN = 128;             // columns per tile: 2 accumulators x 2 buffers x 128 = the 512 TMEM columns
constexpr int A_BYTES = BM * BK * 4;    // 16 KB
constexpr int B_BYTES = MAX_BN * BK * 4;  // 16 KB
constexpr int A_STAGES = 3;             // activation tiles come from HBM (the raw tile is the hi operand); the fourth
                                        // stage was traded for a second output staging tile (two epilogue warp sets)
constexpr int LO_STAGES = 2;            // lo tiles of A live only between the split and the MMA
constexpr int B_STAGES = 3;             // weight tiles (pre-split hi | lo) come from L2
constexpr int OFF_LO = A_STAGES * A_BYTES;
constexpr int OFF_B = OFF_LO + LO_STAGES * A_BYTES;
constexpr int OFF_OUT = OFF_B + B_STAGES * 2 * B_BYTES;   // two [128 rows x 32 cols] output staging tiles for the TMA stores
constexpr int OUT_BYTES = BM * 32 * 4;
constexpr int OPERAND_BYTES = OFF_OUT + 2 * OUT_BYTES;    // 48 + 32 + 96 + 32 = 208 KB
constexpr int GEMM_THREADS = 512;
constexpr int LIN_THREADS = 640;        // linear kernel: + warps 16-19 = second epilogue set
constexpr int SPLIT_THREADS = 256;      // warps 8-15
constexpr int SMEM_MISC = 4096;         // barriers, tmem pointer, bias
constexpr int STATS_MAX_N = 256;        // fused column statistics: per-CTA float64 accumulators [2][n] + quarter partials of both sets
constexpr int SMEM_STATS = 2 * STATS_MAX_N * 8 + 2 * 4 * 64 * 4;
constexpr int GEMM_SMEM = OPERAND_BYTES + 1024 + SMEM_MISC + SMEM_STATS;

enum Epilogue { EPI_BIAS = 0, EPI_BIAS_TANH = 1, EPI_TANH_GRAD = 2, EPI_BIAS_SELECT = 3, EPI_BIAS_ADD = 4,
                EPI_ROWMASK = 5,          // rows without neighbours (rowptr) give 0: backward of the row select, no mask pass
                EPI_ADD_UNSELECTED = 6 }; // rows without neighbours add their aux row (they kept ffn(x) in the forward)

struct LinearParams {
  int64_t m;
  int n, k;          // output columns, reduction length
  int bn, n_tiles;   // columns per tile, tiles along n
  const float* bias; // [n] or null
  float* out0; int ld0; int split;   // columns [0, split) -> out0
  float* out1; int ld1;              // columns [split, n) -> out1 (may be null when split == n)
  int epilogue;
  const float* aux; int ld_aux;      // EPI_TANH_GRAD: tanh output t (y = acc * (1 - t^2)); EPI_BIAS_SELECT / ADD_UNSELECTED: rows without neighbours
  const int64_t* rowptr;             // EPI_BIAS_SELECT: row keeps acc + bias iff rowptr[r+1] > rowptr[r]
  // TOPK instantiation (pc_type_scores_topk): a CTA owns a CONTIGUOUS run of tiles (m-tile major), every epilogue thread
  // keeps the best tk_k columns of its row over the run and flushes them to list (segment, set) of the row
  int tk_k;                          // entries per list (<= TK_MAX)
  int tk_lists;                      // lists per row = 2 * max segments
  int64_t tiles_per_cta;
  double* tk_s;                      // [m, tk_lists, tk_k]
  int64_t* tk_i;                     // [m, tk_lists, tk_k], pre-set to -1
  int store_out;                     // write the [m, n] matrix as well
  // fused BatchNorm statistics (EPI_BIAS, n <= STATS_MAX_N): per-CTA column sums of y and y^2 over the rows < m
  double* col_partial;               // [grid, 2, n] or null
};
constexpr int TK_MAX = 4;

// w_hi = rn_tf32(w), w_lo = rn_tf32(w - w_hi): the weight operand is split once per call, not once per tile
__global__ void split_tf32_kernel(const float4* __restrict__ w, int64_t n4, float4* __restrict__ hi, float4* __restrict__ lo) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = w[i];
  float4 h, l;
  h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
  l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y); l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
  hi[i] = h;
  lo[i] = l;
}

template <bool TOPK>
__global__ void __launch_bounds__(LIN_THREADS, 1)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_whi,
                     const __grid_constant__ CUtensorMap map_wlo,
                     const __grid_constant__ CUtensorMap map_out0,
                     const __grid_constant__ CUtensorMap map_out1, const __grid_constant__ CUtensorMap map_aux,
                     const LinearParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* misc = smem + OPERAND_BYTES;
  // barriers: a_full[5] a_ready[5] a_empty[5] lo_empty[2] b_full[3] b_empty[3] tmem_full[2] tmem_empty[2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 256);
  float* bias_s = reinterpret_cast<float*>(misc + 512);     // up to 768 floats
  double* stat_acc = reinterpret_cast<double*>(misc + SMEM_MISC);                       // [2][n]
  float* stat_part = reinterpret_cast<float*>(misc + SMEM_MISC + 2 * STATS_MAX_N * 8);  // [2 sets][4 quarters][64]
  const uint32_t a_full = smem_u32(bars + 0), a_ready = smem_u32(bars + A_STAGES), a_empty = smem_u32(bars + 2 * A_STAGES);
  const uint32_t lo_empty = smem_u32(bars + 3 * A_STAGES), b_full = smem_u32(bars + 3 * A_STAGES + LO_STAGES);
  const uint32_t b_empty = smem_u32(bars + 3 * A_STAGES + LO_STAGES + B_STAGES);
  const uint32_t tfull_bar = smem_u32(bars + 3 * A_STAGES + LO_STAGES + 2 * B_STAGES), tempty_bar = tfull_bar + 16;
  const uint32_t aux_bar = tempty_bar + 16;   // [2]: the aux chunk of an epilogue set has landed in its staging tile
  const int warp = warp_id(), lane = lane_id();

  if (threadIdx.x == 0) {
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(a_full + 8 * s, 1);
      mbar_init(a_ready + 8 * s, SPLIT_THREADS / 32);
      mbar_init(a_empty + 8 * s, 1);
    }
    for (int s = 0; s < LO_STAGES; ++s) mbar_init(lo_empty + 8 * s, 1);
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + 8 * s, 1);
      mbar_init(tempty_bar + 8 * s, 256);
      mbar_init(aux_bar + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!TOPK)
    for (int i = threadIdx.x; i < p.n; i += LIN_THREADS) bias_s[i] = p.bias ? p.bias[i] : 0.f;
  if (!TOPK && p.col_partial)
    for (int i = threadIdx.x; i < 2 * p.n; i += LIN_THREADS) stat_acc[i] = 0.0;
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t m_tiles = (p.m + BM - 1) / BM;
  const int64_t all_tiles = m_tiles * p.n_tiles;
  const int k_blocks = p.k / BK;
  const uint32_t b_tile_bytes = uint32_t(p.bn) * BK * 4;
  // tile schedule: round-robin over the CTAs, or (TOPK) one contiguous run per CTA; every role walks the same sequence
  const int64_t t_first = TOPK ? int64_t(blockIdx.x) * p.tiles_per_cta : int64_t(blockIdx.x);
  const int64_t tiles = TOPK ? (t_first + p.tiles_per_cta < all_tiles ? t_first + p.tiles_per_cta : all_tiles) : all_tiles;
  const int64_t t_step = TOPK ? 1 : int64_t(gridDim.x);

  if (warp == 0) {
    // ---------------- TMA producer, activations (HBM)
    if (lane == 0) {
      Ring<A_STAGES> ra;
      for (int64_t t = t_first; t < tiles; t += t_step) {
        const int m0 = int((t / p.n_tiles) * BM);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(a_empty + 8 * ra.stage, ra.phase ^ 1);
          mbar_arrive_expect_tx(a_full + 8 * ra.stage, A_BYTES);
          tma_load_2d(smem_u32(smem + ra.stage * A_BYTES), &map_a, kb * BK, m0, a_full + 8 * ra.stage);
          ra.advance();
        }
      }
    }
  } else if (warp == 3) {
    // ---------------- TMA producer, pre-split weights (L2)
    if (lane == 0) {
      Ring<B_STAGES> rb;
      for (int64_t t = t_first; t < tiles; t += t_step) {
        const int n0 = int(t % p.n_tiles) * p.bn;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(b_empty + 8 * rb.stage, rb.phase ^ 1);
          uint8_t* st = smem + OFF_B + rb.stage * 2 * B_BYTES;
          mbar_arrive_expect_tx(b_full + 8 * rb.stage, 2 * b_tile_bytes);
          tma_load_2d(smem_u32(st), &map_whi, kb * BK, n0, b_full + 8 * rb.stage);
          tma_load_2d(smem_u32(st + B_BYTES), &map_wlo, kb * BK, n0, b_full + 8 * rb.stage);
          rb.advance();
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    Ring<A_STAGES> ra;
    Ring<LO_STAGES> rl;
    Ring<B_STAGES> rb;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t idesc = instr_desc_tf32(p.bn);
    // full-width tiles: W_hi and W_lo sit back to back in the stage and so do the two accumulators, so
    // a_hi . [W_hi | W_lo] is ONE 256-wide MMA (a_hi is fetched from shared memory once instead of twice)
    const bool wide = p.bn == MAX_BN;
    const uint32_t idesc_wide = instr_desc_tf32(2 * MAX_BN);
    for (int64_t t = t_first; t < tiles; t += t_step) {
      mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_main = tmem_base + uint32_t(acc * 2 * MAX_BN), d_cross = d_main + MAX_BN;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(b_full + 8 * rb.stage, rb.phase);
        mbar_wait(a_ready + 8 * ra.stage, ra.phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t base = smem_u32(smem);
          const uint64_t a_hi = smem_desc_k_sw128(base + ra.stage * A_BYTES);
          const uint64_t a_lo = smem_desc_k_sw128(base + OFF_LO + rl.stage * A_BYTES);
          const uint64_t b_hi = smem_desc_k_sw128(base + OFF_B + rb.stage * 2 * B_BYTES);
          const uint64_t b_lo = smem_desc_k_sw128(base + OFF_B + rb.stage * 2 * B_BYTES + B_BYTES);
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t adv = uint64_t(kk * 32 >> 4);  // 8 tf32 = 32 bytes along K inside the swizzle row
            if (wide) {
              umma_tf32(d_main, a_hi + adv, b_hi + adv, idesc_wide, (kb | kk) != 0);   // main | cross = a_hi . [W_hi | W_lo]
              umma_tf32(d_cross, a_lo + adv, b_hi + adv, idesc, 1);
            } else {
              umma_tf32(d_cross, a_lo + adv, b_hi + adv, idesc, (kb | kk) != 0);
              umma_tf32(d_cross, a_hi + adv, b_lo + adv, idesc, 1);
              umma_tf32(d_main, a_hi + adv, b_hi + adv, idesc, (kb | kk) != 0);
            }
          }
          umma_commit(a_empty + 8 * ra.stage);
          umma_commit(lo_empty + 8 * rl.stage);
          umma_commit(b_empty + 8 * rb.stage);
          if (kb == k_blocks - 1) umma_commit(tfull_bar + 8 * acc);
        }
        __syncwarp();
        ra.advance();
        rl.advance();
        rb.advance();
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 8 && warp < 16) {
    // ---------------- activation split.  The raw fp32 tile IS the hi operand: under kind::tf32 the tensor core
    // ignores the low 13 mantissa bits, so hi = trunc_tf32(x) needs no rewrite; only lo = rn_tf32(x - hi) is
    // written (to the lo ring).  x = hi + (x - hi) exactly and |lo| < 2^-10 |x|.
    Ring<A_STAGES> ra;
    Ring<LO_STAGES> rl;
    const int tid = threadIdx.x - 256;
    for (int64_t t = t_first; t < tiles; t += t_step) {
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(a_full + 8 * ra.stage, ra.phase);
        mbar_wait(lo_empty + 8 * rl.stage, rl.phase ^ 1);
        const uint32_t hi_base = smem_u32(smem + ra.stage * A_BYTES);
        const uint32_t lo_base = smem_u32(smem + OFF_LO + rl.stage * A_BYTES);
        float4 v[A_BYTES / 16 / SPLIT_THREADS];
#pragma unroll
        for (int it = 0; it < A_BYTES / 16 / SPLIT_THREADS; ++it) v[it] = lds128(hi_base + (tid + it * SPLIT_THREADS) * 16);
#pragma unroll
        for (int it = 0; it < A_BYTES / 16 / SPLIT_THREADS; ++it) {
          float4 l;
          l.x = to_tf32(v[it].x - trunc_tf32(v[it].x)); l.y = to_tf32(v[it].y - trunc_tf32(v[it].y));
          l.z = to_tf32(v[it].z - trunc_tf32(v[it].z)); l.w = to_tf32(v[it].w - trunc_tf32(v[it].w));
          sts128(lo_base + (tid + it * SPLIT_THREADS) * 16, l);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready + 8 * ra.stage);
        ra.advance();
        rl.advance();
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: TMEM -> registers -> (bias / activation) -> swizzled smem tile -> TMA store
    // two sets of four warps (4-7 and 16-19; warp % 4 = TMEM lane quarter): set s takes the 32-column chunks
    // s, s + 2, ... of every tile and has its own staging tile and named barrier.  One warp per scheduler could not
    // hide its own latencies in the tanh / aux epilogues (profiles/gemm_epilogue_bench.py).
    int acc = 0;
    uint32_t acc_phase = 0;
    const int set = warp >= 16 ? 1 : 0;
    const int quad = warp & 3;
    const int trow = quad * 32 + lane;                      // row of the tile this thread owns
    const bool issuer = quad == 0 && lane == 0;             // first thread of the set issues its bulk stores
    uint8_t* out_tile = smem + OFF_OUT + set * OUT_BYTES;
    const uint32_t bias_addr = smem_u32(bias_s), out_addr = smem_u32(out_tile);
    // tanh-gradient / add epilogues read a whole [128 x 32] tile of `aux` per chunk.  As per-thread row loads that is
    // 32 L1 wavefronts per instruction (0.95 vs 0.38 ms for 128 -> 256 with / without aux); instead the issuer TMA-loads
    // the aux chunk into the set's staging tile (same swizzle as the output), every thread combines its own row in
    // place and the tile goes out again as the result.  The rows of the NEXT tile are pulled into L2 a tile ahead.
    // The row-select epilogue only needs aux for rows without neighbours and keeps plain loads.
    const bool aux_tile = p.epilogue == EPI_TANH_GRAD || p.epilogue == EPI_BIAS_ADD;
    const bool row_aux = p.epilogue == EPI_BIAS_SELECT || p.epilogue == EPI_ADD_UNSELECTED;   // aux only for rows without neighbours
    const bool has_aux = aux_tile || row_aux;
    uint32_t aux_phase = 0;
    // TOPK: best tk_k (score, column) of my row over this CTA's run of tiles; columns arrive ascending, so the strict
    // comparison keeps the lower column on ties
    float tk_v[TK_MAX];
    int tk_c[TK_MAX];
    int64_t tk_mtile = -1;
    auto tk_reset = [&]() {
#pragma unroll
      for (int j = 0; j < TK_MAX; ++j) { tk_v[j] = -INFINITY; tk_c[j] = -1; }
    };
    auto tk_flush = [&]() {
      if (!TOPK || tk_mtile < 0) return;
      const int64_t frow = tk_mtile * BM + trow;
      if (frow >= p.m) return;
      const int64_t first_cta = (tk_mtile * p.n_tiles) / p.tiles_per_cta;      // CTA that holds the m-tile's first n-tile
      const int64_t list = (int64_t(blockIdx.x) - first_cta) * 2 + set;
      double* os = p.tk_s + (frow * p.tk_lists + list) * p.tk_k;
      int64_t* oi = p.tk_i + (frow * p.tk_lists + list) * p.tk_k;
#pragma unroll
      for (int j = 0; j < TK_MAX; ++j)
        if (j < p.tk_k) { os[j] = double(tk_v[j]); oi[j] = int64_t(tk_c[j]); }
    };
    if (TOPK) tk_reset();
    auto prefetch_aux = [&](int64_t tile) {
      if (!has_aux || tile >= tiles) return;
      const int64_t prow = (tile / p.n_tiles) * BM + trow;
      if (prow < p.m && (!row_aux || p.rowptr[prow + 1] == p.rowptr[prow]))   // select reads aux only for empty rows
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.aux + prow * p.ld_aux + (tile % p.n_tiles) * p.bn),
                     "r"(uint32_t(p.bn) * 4u)
                     : "memory");
    };
    prefetch_aux(t_first);
    for (int64_t t = t_first; t < tiles; t += t_step) {
      const int m0 = int((t / p.n_tiles) * BM);
      const int64_t row = int64_t(m0) + trow;
      const int n0 = int(t % p.n_tiles) * p.bn;
      prefetch_aux(t + t_step);
      if (TOPK && t / p.n_tiles != tk_mtile) {
        tk_flush();
        tk_reset();
        tk_mtile = t / p.n_tiles;
      }
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      bool keep = true;
      if ((row_aux || p.epilogue == EPI_ROWMASK) && row < p.m) keep = p.rowptr[row + 1] > p.rowptr[row];
      for (int c0 = set * 32; c0 < p.bn; c0 += 64) {
        if (aux_tile && issuer) {   // the staging tile is free once the previous store has read it
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive_expect_tx(aux_bar + 8 * set, OUT_BYTES);
          tma_load_2d(out_addr, &map_aux, n0 + c0, m0, aux_bar + 8 * set);
        }
        uint32_t r[32], rc[32];
        tmem_ld32_async(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * 2 * MAX_BN + c0), r);
        tmem_ld32_async(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * 2 * MAX_BN + MAX_BN + c0), rc);
        tmem_ld_wait32(r);
        tmem_ld_wait32(rc);
        const int n = n0 + c0;
        float y[32];
        if (TOPK) {
          if (n >= p.n) continue;                     // chunk past the last column (n need not be a multiple of 128)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            y[j] = __uint_as_float(r[j]) + __uint_as_float(rc[j]);
            if (n + j < p.n && y[j] > tk_v[TK_MAX - 1]) {      // unused tail entries of the list stay at -inf
              float v = y[j];
              int c = n + j;
              bool placed = false;                     // once placed, the displaced entries shift down unconditionally
#pragma unroll                                         // (a strict comparison there would reorder equal scores)
              for (int q = 0; q < TK_MAX; ++q) {
                if (placed || v > tk_v[q]) {
                  const float tv = tk_v[q]; const int tc = tk_c[q];
                  tk_v[q] = v; tk_c[q] = c; v = tv; c = tc;
                  placed = true;
                }
              }
            }
          }
          if (!p.store_out) continue;
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = lds128(bias_addr + (n + j) * 4);
            y[j] = __uint_as_float(r[j]) + __uint_as_float(rc[j]) + b4.x;
            y[j + 1] = __uint_as_float(r[j + 1]) + __uint_as_float(rc[j + 1]) + b4.y;
            y[j + 2] = __uint_as_float(r[j + 2]) + __uint_as_float(rc[j + 2]) + b4.z;
            y[j + 3] = __uint_as_float(r[j + 3]) + __uint_as_float(rc[j + 3]) + b4.w;
          }
        }
        if (p.epilogue == EPI_BIAS_TANH) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = tanhf(y[j]);
        } else if (aux_tile) {
          mbar_wait(aux_bar + 8 * set, aux_phase);   // implies the issuer saw the previous store finish reading the tile
          aux_phase ^= 1;
#pragma unroll
          for (int j = 0; j < 8; ++j) {              // my own row of the aux chunk, same swizzle as the output below
            const float4 a = lds128(out_addr + trow * 128 + ((j ^ (trow & 7)) << 4));
            if (p.epilogue == EPI_TANH_GRAD) {
              y[4 * j] *= 1.f - a.x * a.x; y[4 * j + 1] *= 1.f - a.y * a.y; y[4 * j + 2] *= 1.f - a.z * a.z; y[4 * j + 3] *= 1.f - a.w * a.w;
            } else {
              y[4 * j] += a.x; y[4 * j + 1] += a.y; y[4 * j + 2] += a.z; y[4 * j + 3] += a.w;
            }
          }
        } else if (row_aux && !keep && row < p.m) {
          const float* aux = p.aux + row * p.ld_aux + n;
          const bool add = p.epilogue == EPI_ADD_UNSELECTED;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 a = *reinterpret_cast<const float4*>(aux + j);
            y[j] = (add ? y[j] : 0.f) + a.x; y[j + 1] = (add ? y[j + 1] : 0.f) + a.y;
            y[j + 2] = (add ? y[j + 2] : 0.f) + a.z; y[j + 3] = (add ? y[j + 3] : 0.f) + a.w;
          }
        } else if (p.epilogue == EPI_ROWMASK && !keep) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = 0.f;
        }
        if (!aux_tile) {
          // the previous chunk's bulk store must have finished reading the staging tile
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)   // 128-byte swizzle: 16-byte chunk j of row r lives at chunk j ^ (r & 7)
          sts128(out_addr + trow * 128 + ((j ^ (trow & 7)) << 4), make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]));
        fence_proxy_async();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
        if (!TOPK && p.col_partial) {
          // column sums of this [128 x 32] chunk straight from the staging tile: thread (quarter, column) adds its 32 rows,
          // the four quarters are combined in fixed order into the CTA's float64 accumulators (pc_col_stats fused away)
          const int t128 = quad * 32 + lane, cc = t128 & 31, qtr = t128 >> 5;
          float s1 = 0.f, s2 = 0.f;
          const int64_t left = p.m - m0 - qtr * 32;           // rows of this quarter that exist (only the last tile is short)
          const int valid = left >= 32 ? 32 : (left > 0 ? int(left) : 0);
          const uint32_t col_off = uint32_t((cc & 3) << 2), chunk = uint32_t(cc >> 2);
          if (valid == 32) {
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {                 // rw & 7 == rr & 7: the quarter starts at a multiple of 32
              const float v = lds32(out_addr + uint32_t(qtr * 32 + rr) * 128 + (((chunk ^ uint32_t(rr & 7)) << 4) | col_off));
              s1 += v;
              s2 = fmaf(v, v, s2);
            }
          } else {
            for (int rr = 0; rr < valid; ++rr) {
              const float v = lds32(out_addr + uint32_t(qtr * 32 + rr) * 128 + (((chunk ^ uint32_t(rr & 7)) << 4) | col_off));
              s1 += v;
              s2 = fmaf(v, v, s2);
            }
          }
          float* part = stat_part + (set * 4 + qtr) * 64;
          part[cc] = s1;
          part[32 + cc] = s2;
          asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
          if (t128 < 64) {
            const int which = t128 >> 5;
            const float* pp = stat_part + set * 4 * 64 + which * 32 + cc;
            const float tot = ((pp[0] + pp[64]) + pp[128]) + pp[192];
            stat_acc[which * p.n + n + cc] += double(tot);
          }
        }
        if (issuer) {
          const bool first = n < p.split;
          const CUtensorMap* map = first ? &map_out0 : &map_out1;
          const int cn = first ? n : n - p.split;
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(cn),
                       "r"(m0), "r"(smem_u32(out_tile))
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + 8 * acc);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    tk_flush();
    if (!TOPK && p.col_partial) {
      asm volatile("bar.sync 3, 256;" ::: "memory");       // both epilogue sets are done accumulating
      const int t256 = set * 128 + quad * 32 + lane;
      for (int i = t256; i < 2 * p.n; i += 256) p.col_partial[int64_t(blockIdx.x) * 2 * p.n + i] = stat_acc[i];
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ================================================================ weight gradient
// dW[n, k] = sum_m dY[m, n] * X[m, k]  (+ db[n] = sum_m dY[m, n]): the reduction runs over the
// rows of two row-major activations, so both MMA operands are MN-major.  A stage holds 16 rows
// of a 128-column half of dY and of X as [32-column group][16 rows][128 bytes] blocks, which is
// the canonical MN-major layout for 32-bit operands (SWIZZLE_128B_BASE32B: 32-byte chunks XOR
// row mod 4) with LBO = 2048 (next 32 columns) and SBO = 512 (next 4 rows).
// CTA c owns column half (c mod halves) of dY and every (grid/halves)-th 16-row block.  Because
// the tensor core accumulates with truncation, a chain is cut every 32 blocks: the two TMEM
// accumulators (hi.hi and cross terms) are added, rounded, into the CTA's fp32 partial in
// global memory.  A second kernel sums the partials in CTA order (deterministic, no atomics).
constexpr int WG_ROWS = 16;                        // reduction rows per stage (2 x UMMA_K)
constexpr int WG_STAGES = 4;                       // a stage = raw dY | lo dY | raw X | lo X.  The raw tiles are the hi
                                                   // operands as they are (the tensor core ignores the low 13 mantissa
                                                   // bits); the lo tiles are written next to them by the split warps, so
                                                   // the only ring is TMA -> split -> MMA -> free (profiles/r1_gemm_stage_probe.md)
constexpr int WG_MAX_K = 256;
constexpr int WG_DY_BYTES = BM * WG_ROWS * 4;      // 8 KB: 128 columns x 16 rows
constexpr int WG_X_BYTES = WG_MAX_K * WG_ROWS * 4; // 16 KB
constexpr int WG_OFF_X = 2 * WG_DY_BYTES;          // raw X starts here; its lo tile follows at + k * WG_ROWS * 4
constexpr int WG_STAGE_BYTES = 2 * (WG_DY_BYTES + WG_X_BYTES);     // 48 KB
constexpr int WG_OUT_BYTES = BM * 32 * 4;            // [128 x 32] staging tile of the flush (TMA store / reduce-add)
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + WG_OUT_BYTES + 1024 + SMEM_MISC;   // 192 + 16 KB + misc
constexpr int WG_GROUP_BYTES = WG_ROWS * 128;      // one 32-column group of a stage
constexpr int WG_FLUSH = 32;                       // blocks per accumulation chain

struct WgradParams {
  int64_t m;
  int n, k;            // dY columns (rows of dW), X columns (cols of dW)
  int halves;          // n / 128
  float* partial_w;    // [grid / halves, n, k]
  float* partial_b;    // [2 * grid / halves, n] or null
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// MN-major 32-bit operand: the only legal layout is SWIZZLE_128B_BASE32B (atom = 128 bytes of MN x 4 K rows).
// LBO = stride between 32-element MN groups, SBO = stride between 4-row K groups (an 8-deep tf32 MMA spans two).
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t addr) {
  return uint64_t((addr & 0x3FFFFu) >> 4) | (uint64_t(WG_GROUP_BYTES >> 4) << 16) | (uint64_t(512 >> 4) << 32) |
         (uint64_t(1) << 46) | (uint64_t(1) << 61);
}
__device__ __forceinline__ uint32_t instr_desc_tf32_mn(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | (uint32_t(n >> 3) << 17) | (uint32_t(BM >> 4) << 24);
}

using WgPipe = Ring<WG_STAGES>;

__global__ void __launch_bounds__(GEMM_THREADS, 1)
wgrad_tf32x3_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                    const __grid_constant__ CUtensorMap map_pw, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wg_out = smem + WG_STAGES * WG_STAGE_BYTES;
  uint8_t* misc = wg_out + WG_OUT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);   // full[S], ready[S], empty[S], lo_empty[L], tmem_full, tmem_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 256);
  const uint32_t full_bar = smem_u32(bars + 0), ready_bar = smem_u32(bars + WG_STAGES), empty_bar = smem_u32(bars + 2 * WG_STAGES);
  const uint32_t tfull_bar = smem_u32(bars + 3 * WG_STAGES), tempty_bar = tfull_bar + 8;
  const int warp = warp_id(), lane = lane_id();
  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(ready_bar + 8 * s, SPLIT_THREADS / 32);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int half = blockIdx.x % p.halves;
  const int subset = blockIdx.x / p.halves, subsets = gridDim.x / p.halves;
  const int64_t blocks = (p.m + WG_ROWS - 1) / WG_ROWS;
  const int64_t my_blocks = blocks > subset ? (blocks - subset + subsets - 1) / subsets : 0;
  const uint32_t x_bytes = uint32_t(p.k) * WG_ROWS * 4;

  if (warp == 0) {
    if (lane == 0) {
      WgPipe pipe;
      for (int64_t i = 0; i < my_blocks; ++i) {
        const int row0 = int((subset + i * subsets) * WG_ROWS);
        mbar_wait(empty_bar + 8 * pipe.stage, pipe.phase ^ 1);
        uint8_t* st = smem + pipe.stage * WG_STAGE_BYTES;
        mbar_arrive_expect_tx(full_bar + 8 * pipe.stage, WG_DY_BYTES + x_bytes);
        tma_load_3d(smem_u32(st), &map_dy, 0, row0, half * 4, full_bar + 8 * pipe.stage);
        tma_load_3d(smem_u32(st + WG_OFF_X), &map_x, 0, row0, 0, full_bar + 8 * pipe.stage);
        pipe.advance();
      }
    }
  } else if (warp == 1) {
    WgPipe pipe;
    const uint32_t idesc = instr_desc_tf32_mn(p.k);
    // k <= 128: X_hi | X_lo (adjacent in the stage) form ONE operand of width 2k and main | cross are adjacent in
    // TMEM, so dY_hi . [X_hi | X_lo] is a single MMA (dY_hi is fetched from shared memory once instead of twice)
    const bool wide = 2 * p.k <= 256;
    const uint32_t idesc_wide = instr_desc_tf32_mn(2 * p.k);
    const uint32_t d_main = tmem_base, d_cross = tmem_base + uint32_t(p.k);
    uint32_t flush_phase = 0;
    for (int64_t i = 0; i < my_blocks; ++i) {
      const int in_chain = int(i % WG_FLUSH);
      if (in_chain == 0) {
        mbar_wait(tempty_bar, flush_phase ^ 1);   // previous chain drained by the epilogue warps
        tc_fence_after();
      }
      mbar_wait(ready_bar + 8 * pipe.stage, pipe.phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t st = smem_u32(smem + pipe.stage * WG_STAGE_BYTES);
#pragma unroll
        for (int ks = 0; ks < WG_ROWS / 8; ++ks) {
          const uint64_t a_hi = smem_desc_mn_sw128(st + ks * 1024), a_lo = smem_desc_mn_sw128(st + WG_DY_BYTES + ks * 1024);
          const uint64_t b_hi = smem_desc_mn_sw128(st + WG_OFF_X + ks * 1024);
          const uint64_t b_lo = smem_desc_mn_sw128(st + WG_OFF_X + x_bytes + ks * 1024);
          const uint32_t accum = (in_chain | ks) != 0;
          if (wide) {
            umma_tf32(d_main, a_hi, b_hi, idesc_wide, accum);     // main | cross = dY_hi . [X_hi | X_lo]
            umma_tf32(d_cross, a_lo, b_hi, idesc, 1);
          } else {
            umma_tf32(d_cross, a_lo, b_hi, idesc, accum);
            umma_tf32(d_cross, a_hi, b_lo, idesc, 1);
            umma_tf32(d_main, a_hi, b_hi, idesc, accum);
          }
        }
        umma_commit(empty_bar + 8 * pipe.stage);
        if (in_chain == WG_FLUSH - 1 || i == my_blocks - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
      if (in_chain == WG_FLUSH - 1) flush_phase ^= 1;
      pipe.advance();
    }
  } else if (warp >= 8) {
    // split hi/lo; a thread owns one (32-column group, 16-byte chunk, 8-row half) item of dY or X, so the dY
    // column sums stay in its registers
    WgPipe pipe;
    const int tid = threadIdx.x - 256;
    const int dy_items = (BM / 4) * 2, x_items = (p.k / 4) * 2;   // 64 + up to 128 <= 256 threads
    const bool active = tid < dy_items + x_items;
    const bool is_dy = tid < dy_items;
    const int item = is_dy ? tid : tid - dy_items;
    // lane -> 16-byte chunk c of a 128-byte row: the 8 lanes of a quarter-warp cover one whole row (no bank conflicts)
    const int c = item & 7, r0 = ((item >> 3) & 1) * 8, group = item >> 4;
    const uint32_t op_off = (is_dy ? 0u : uint32_t(WG_OFF_X)) + uint32_t(group) * WG_GROUP_BYTES;
    const uint32_t lo_off = is_dy ? uint32_t(WG_DY_BYTES) : x_bytes;     // the lo tile sits right behind its raw tile
    float4 colsum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = 0; i < my_blocks; ++i) {
      mbar_wait(full_bar + 8 * pipe.stage, pipe.phase);
      if (active) {
        const uint32_t base = smem_u32(smem + pipe.stage * WG_STAGE_BYTES + op_off);
        const uint32_t lo_base = base + lo_off;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int r = r0 + rr;
          const uint32_t off = uint32_t(r * 128 + (((((c >> 1) ^ (r & 3)) << 1) | (c & 1)) << 4));
          const float4 v = lds128(base + off);
          float4 l;
          l.x = to_tf32(v.x - trunc_tf32(v.x)); l.y = to_tf32(v.y - trunc_tf32(v.y));
          l.z = to_tf32(v.z - trunc_tf32(v.z)); l.w = to_tf32(v.w - trunc_tf32(v.w));
          sts128(lo_base + off, l);
          colsum.x += v.x; colsum.y += v.y; colsum.z += v.z; colsum.w += v.w;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(ready_bar + 8 * pipe.stage);
      pipe.advance();
    }
    if (p.partial_b && is_dy) {
      // logical chunk c of group g covers columns g*32 + 4c .. +3 (the swizzle only permutes positions)
      float* dst = p.partial_b + int64_t(subset * 2 + ((item >> 3) & 1)) * p.n + half * BM + group * 32 + c * 4;
      *reinterpret_cast<float4*>(dst) = colsum;
    }
  } else if (warp >= 4) {
    const int quad = warp - 4;
    const int row = half * BM + quad * 32 + lane;
    float* out = p.partial_w + (int64_t(subset) * p.n + row) * p.k;
    const int64_t chains = (my_blocks + WG_FLUSH - 1) / WG_FLUSH;
    uint32_t phase = 0;
    if (chains == 0) {
      for (int c0 = 0; c0 < p.k; c0 += 4) *reinterpret_cast<float4*>(out + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // A chain's [128 x k] result leaves through a swizzled staging tile and TMA: a plain store for the first chain,
    // a reduce-add (cp.reduce.async.bulk.tensor .add) into the CTA's partial for the following ones.  (As per-thread
    // row read-modify-writes every load / store instruction touched 32 different rows.)  One chain's operations are
    // complete before the next chain's are issued, so the additions happen in chain order: deterministic.
    const int trow = quad * 32 + lane;
    const bool issuer = threadIdx.x == 128;
    const uint32_t out_addr = smem_u32(wg_out);
    const int prow0 = subset * p.n + half * BM;               // first row of this CTA's tile in the [subsets * n, k] partials
    for (int64_t ch = 0; ch < chains; ++ch) {
      mbar_wait(tfull_bar, phase);
      tc_fence_after();
      if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // previous chain fully applied
      for (int c0 = 0; c0 < p.k; c0 += 32) {
        uint32_t r[32], rc[32];
        tmem_ld32_async(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(c0), r);
        tmem_ld32_async(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(p.k + c0), rc);
        tmem_ld_wait32(r);
        tmem_ld_wait32(rc);
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile free again
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j)   // 128-byte swizzle: 16-byte chunk j of row r lives at chunk j ^ (r & 7)
          sts128(out_addr + trow * 128 + ((j ^ (trow & 7)) << 4),
                 make_float4(__uint_as_float(r[4 * j]) + __uint_as_float(rc[4 * j]), __uint_as_float(r[4 * j + 1]) + __uint_as_float(rc[4 * j + 1]),
                             __uint_as_float(r[4 * j + 2]) + __uint_as_float(rc[4 * j + 2]), __uint_as_float(r[4 * j + 3]) + __uint_as_float(rc[4 * j + 3])));
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          if (ch == 0)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map_pw), "r"(c0),
                         "r"(prow0), "r"(out_addr)
                         : "memory");
          else
            asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map_pw),
                         "r"(c0), "r"(prow0), "r"(out_addr)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar);
      phase ^= 1;
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// out[i] = sum_c partial[c, i] in ascending c (fixed order)
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int parts, int64_t n, float* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int c = 0; c < parts; ++c) acc += partial[int64_t(c) * n + i];
  out[i] = acc;
}

__global__ void reduce_partials_f64_kernel(const double* __restrict__ partial, int parts, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = 0.0;
  for (int c = 0; c < parts; ++c) acc += partial[int64_t(c) * n + i];
  out[i] = acc;
}

// ---------------------------------------------------------------- host side
}  // namespace
}  // namespace pc

using namespace pc;

extern "C" size_t pc_linear_workspace_bytes(int n, int k) {
  return align_up(size_t(n) * k * sizeof(float), 256) * 2 + size_t(sm_count()) * 2 * size_t(n) * sizeof(double);
}

extern "C" int pc_linear_tf32x3(const float* a, int64_t m, int k, int64_t lda, const float* w, int n, const float* bias,
                                int epilogue, const float* aux, int64_t ld_aux, const int64_t* rowptr, float* out0,
                                int64_t ld0, int split, float* out1, int64_t ld1, double* col_sums, void* workspace,
                                size_t workspace_bytes, pc_stream_t stream) {
  PC_REQUIRE(m >= 0, PC_ERR_INVALID, "linear: negative row count");
  PC_REQUIRE(!col_sums || (m > 0 && epilogue == EPI_BIAS && n <= STATS_MAX_N), PC_ERR_UNSUPPORTED,
             "linear: fused column statistics need epilogue 0, n <= %d and at least one row", STATS_MAX_N);
  if (m == 0) return PC_OK;
  PC_REQUIRE(a && w && out0 && workspace, PC_ERR_INVALID, "linear: null pointer");
  PC_REQUIRE(k >= BK && k % BK == 0 && k <= 4096, PC_ERR_UNSUPPORTED, "linear: k=%d must be a multiple of %d", k, BK);
  PC_REQUIRE(n >= 32 && n % 32 == 0 && n <= 768, PC_ERR_UNSUPPORTED, "linear: n=%d must be a multiple of 32 in [32, 768]", n);
  PC_REQUIRE(split > 0 && split <= n && split % 32 == 0 && (split == n || out1), PC_ERR_INVALID, "linear: bad output split");
  PC_REQUIRE(lda % 4 == 0 && ld0 % 4 == 0 && (split == n || ld1 % 4 == 0), PC_ERR_INVALID, "linear: leading dimensions must be multiples of 4 floats");
  PC_REQUIRE((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out0) | reinterpret_cast<uintptr_t>(out1)) % 16 == 0, PC_ERR_INVALID,
             "linear: A / outputs must be 16-byte aligned (TMA)");
  PC_REQUIRE(epilogue >= EPI_BIAS && epilogue <= EPI_ADD_UNSELECTED, PC_ERR_INVALID, "linear: unknown epilogue %d", epilogue);
  PC_REQUIRE((epilogue != EPI_TANH_GRAD && epilogue != EPI_BIAS_SELECT && epilogue != EPI_BIAS_ADD && epilogue != EPI_ADD_UNSELECTED) ||
                 (aux && ld_aux % 4 == 0),
             PC_ERR_INVALID, "linear: epilogue needs aux");
  PC_REQUIRE((epilogue != EPI_BIAS_SELECT && epilogue != EPI_ROWMASK && epilogue != EPI_ADD_UNSELECTED) || rowptr, PC_ERR_INVALID,
             "linear: row-select epilogues need rowptr");
  PC_REQUIRE(workspace_bytes >= pc_linear_workspace_bytes(n, k), PC_ERR_WORKSPACE, "linear: workspace too small");
  LinearParams p;
  p.m = m; p.n = n; p.k = k;
  p.n_tiles = (n + MAX_BN - 1) / MAX_BN;
  while (n % p.n_tiles != 0 || (n / p.n_tiles) % 32 != 0) ++p.n_tiles;
  p.bn = n / p.n_tiles;
  p.bias = bias;
  p.out0 = out0; p.ld0 = int(ld0); p.split = split; p.out1 = out1; p.ld1 = int(ld1);
  p.epilogue = epilogue; p.aux = aux; p.ld_aux = int(ld_aux); p.rowptr = rowptr;
  cudaStream_t st = as_stream(stream);
  float* w_hi = reinterpret_cast<float*>(workspace);
  float* w_lo = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align_up(size_t(n) * k * sizeof(float), 256));
  const int64_t n4 = int64_t(n) * k / 4;
  {
    split_tf32_kernel<<<unsigned((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(w), n4,
                                                                 reinterpret_cast<float4*>(w_hi), reinterpret_cast<float4*>(w_lo));
    PC_LAUNCH_CHECK();
  }
  // (splitting the weight tile inside the kernel instead was measured ~20 % slower: the operand-split warps sit on
  // the critical path, the extra L2 -> SM traffic of the pre-split lo tile does not)
  CUtensorMap map_a, map_whi, map_wlo;
  // consecutive K blocks of a row are adjacent in memory: let L2 fetch 256 B per miss so the next block's
  // request hits, halving the DRAM page activations of the strided [128 x 32] activation boxes
  if (int rc = make_map(&map_a, a, m, k, lda, BM, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return rc;
  if (int rc = make_map(&map_whi, w_hi, n, k, k, p.bn)) return rc;
  if (int rc = make_map(&map_wlo, w_lo, n, k, k, p.bn)) return rc;
  CUtensorMap map_out0, map_out1;
  if (int rc = make_map(&map_out0, out0, m, split, ld0, BM)) return rc;
  if (split < n) {
    if (int rc = make_map(&map_out1, out1, m, n - split, ld1, BM)) return rc;
  } else {
    map_out1 = map_out0;
  }
  CUtensorMap map_aux = map_out0;
  if (epilogue == EPI_TANH_GRAD || epilogue == EPI_BIAS_ADD) {
    PC_REQUIRE(reinterpret_cast<uintptr_t>(aux) % 16 == 0, PC_ERR_INVALID, "linear: aux must be 16-byte aligned (TMA)");
    if (int rc = make_map(&map_aux, aux, m, n, ld_aux, BM)) return rc;
  }
  static bool configured[64] = {};
  if (first_use_on_device(configured))
    PC_CUDA(cudaFuncSetAttribute(linear_tf32x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  const int64_t tiles = ((m + BM - 1) / BM) * p.n_tiles;
  const int grid = int(tiles < sm_count() ? tiles : sm_count());
  p.tk_k = 0; p.tk_lists = 0; p.tiles_per_cta = 0; p.tk_s = nullptr; p.tk_i = nullptr; p.store_out = 1;
  p.col_partial = col_sums ? reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 2 * align_up(size_t(n) * k * sizeof(float), 256)) : nullptr;
  linear_tf32x3_kernel<false><<<grid, LIN_THREADS, GEMM_SMEM, st>>>(map_a, map_whi, map_wlo, map_out0, map_out1, map_aux, p);
  PC_LAUNCH_CHECK();
  if (col_sums) {   // CTA partials summed in CTA order: bit-reproducible
    reduce_partials_f64_kernel<<<(2 * n + 255) / 256, 256, 0, st>>>(p.col_partial, grid, 2 * n, col_sums);
    PC_LAUNCH_CHECK();
  }
  return PC_OK;
}

// ---------------------------------------------------------------- [B, L] x [T, L]^T scoring with the row top-k in the epilogue
namespace {
struct ScoreSchedule {
  int64_t m_tiles, n_tiles, tiles, per;
  int grid, max_segs;
};
ScoreSchedule score_schedule(int64_t m, int n) {
  ScoreSchedule sc;
  sc.m_tiles = (m + BM - 1) / BM;
  sc.n_tiles = (n + MAX_BN - 1) / MAX_BN;
  sc.tiles = sc.m_tiles * sc.n_tiles;
  const int64_t sms = sm_count();
  sc.per = (sc.tiles + sms - 1) / sms;
  if (sc.per < 1) sc.per = 1;
  sc.grid = int((sc.tiles + sc.per - 1) / sc.per);
  sc.max_segs = int((sc.n_tiles - 1) / sc.per + 2);      // CTAs whose runs can intersect one m-tile's n-tiles
  return sc;
}
}  // namespace

extern "C" size_t pc_type_scores_topk_workspace_bytes(int64_t m, int n, int k, int topk) {
  if (m <= 0 || n <= 0 || k <= 0 || topk <= 0) return 0;
  const ScoreSchedule sc = score_schedule(m, n);
  return 2 * align_up(size_t(n) * k * sizeof(float), 256) + size_t(m) * size_t(2 * sc.max_segs) * size_t(topk) * 16;
}

extern "C" int pc_type_scores_topk(const float* a, int64_t m, int k, int64_t lda, const float* w, int n, float* out,
                                   int64_t ld_out, int topk, double* out_scores, int64_t* out_idx, void* workspace,
                                   size_t workspace_bytes, pc_stream_t stream) {
  PC_REQUIRE(m >= 0, PC_ERR_INVALID, "type_scores_topk: negative row count");
  if (m == 0) return PC_OK;
  PC_REQUIRE(a && w && out_scores && out_idx && workspace, PC_ERR_INVALID, "type_scores_topk: null pointer");
  PC_REQUIRE(k >= BK && k % BK == 0 && k <= 4096, PC_ERR_UNSUPPORTED, "type_scores_topk: k=%d must be a multiple of %d", k, BK);
  PC_REQUIRE(n >= 1 && n < (1 << 30), PC_ERR_UNSUPPORTED, "type_scores_topk: bad n=%d", n);
  PC_REQUIRE(topk >= 1 && topk <= TK_MAX && topk <= n, PC_ERR_UNSUPPORTED, "type_scores_topk: topk=%d outside [1,%d] (or > n)", topk, TK_MAX);
  PC_REQUIRE(lda % 4 == 0 && reinterpret_cast<uintptr_t>(a) % 16 == 0, PC_ERR_INVALID, "type_scores_topk: A must be 16-byte aligned, lda % 4 == 0");
  PC_REQUIRE(!out || (ld_out % 4 == 0 && ld_out >= n && reinterpret_cast<uintptr_t>(out) % 16 == 0), PC_ERR_INVALID,
             "type_scores_topk: out must be 16-byte aligned with ld_out % 4 == 0 (TMA)");
  PC_REQUIRE(workspace_bytes >= pc_type_scores_topk_workspace_bytes(m, n, k, topk), PC_ERR_WORKSPACE, "type_scores_topk: workspace too small");
  const ScoreSchedule sc = score_schedule(m, n);
  LinearParams p;
  p.m = m; p.n = n; p.k = k; p.bn = MAX_BN; p.n_tiles = int(sc.n_tiles);
  p.bias = nullptr; p.out0 = out; p.ld0 = int(ld_out); p.split = n; p.out1 = nullptr; p.ld1 = 0;
  p.epilogue = EPI_BIAS; p.aux = nullptr; p.ld_aux = 0; p.rowptr = nullptr;
  p.tk_k = topk; p.tk_lists = 2 * sc.max_segs; p.tiles_per_cta = sc.per; p.store_out = out ? 1 : 0;
  p.col_partial = nullptr;
  cudaStream_t st = as_stream(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  const size_t wbytes = align_up(size_t(n) * k * sizeof(float), 256);
  float* w_hi = reinterpret_cast<float*>(ws);
  float* w_lo = reinterpret_cast<float*>(ws + wbytes);
  const size_t entries = size_t(m) * p.tk_lists * topk;
  p.tk_s = reinterpret_cast<double*>(ws + 2 * wbytes);
  p.tk_i = reinterpret_cast<int64_t*>(p.tk_s + entries);
  PC_REQUIRE((int64_t(n) * k) % 4 == 0, PC_ERR_UNSUPPORTED, "type_scores_topk: n * k must be a multiple of 4");
  const int64_t n4 = int64_t(n) * k / 4;
  split_tf32_kernel<<<unsigned((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(w), n4,
                                                               reinterpret_cast<float4*>(w_hi), reinterpret_cast<float4*>(w_lo));
  PC_LAUNCH_CHECK();
  PC_CUDA(cudaMemsetAsync(p.tk_i, 0xFF, entries * sizeof(int64_t), st));      // -1 = empty slot (lists a CTA never visits)
  CUtensorMap map_a, map_whi, map_wlo, map_out;
  if (int rc = make_map(&map_a, a, m, k, lda, BM, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return rc;
  if (int rc = make_map(&map_whi, w_hi, n, k, k, MAX_BN)) return rc;
  if (int rc = make_map(&map_wlo, w_lo, n, k, k, MAX_BN)) return rc;
  if (out) {
    if (int rc = make_map(&map_out, out, m, n, ld_out, BM)) return rc;
  } else {
    map_out = map_a;
  }
  static bool configured[64] = {};
  if (first_use_on_device(configured))
    PC_CUDA(cudaFuncSetAttribute(linear_tf32x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  linear_tf32x3_kernel<true><<<sc.grid, LIN_THREADS, GEMM_SMEM, st>>>(map_a, map_whi, map_wlo, map_out, map_out, map_out, p);
  PC_LAUNCH_CHECK();
  return pc_topk_merge(p.tk_s, p.tk_i, m, p.tk_lists, topk, out_scores, out_idx, stream);
}

// row-major fp32 [rows, cols] viewed as [cols/32 groups][rows][32]: boxes of [groups, 16 rows, 32 cols]
static int make_map_mn(CUtensorMap* map, const float* base, int64_t rows, int cols, int64_t ld, int box_groups) {
  EncodeTiledFn fn = encode_tiled();
  PC_REQUIRE(fn, PC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[3] = {32, cuuint64_t(rows), cuuint64_t(cols / 32)};
  cuuint64_t gstride[2] = {cuuint64_t(ld) * 4, 128};
  cuuint32_t box[3] = {32, WG_ROWS, cuuint32_t(box_groups)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PC_REQUIRE(r == CUDA_SUCCESS, PC_ERR_CUDA, "cuTensorMapEncodeTiled (3d) failed (%d)", int(r));
  return PC_OK;
}

extern "C" size_t pc_wgrad_workspace_bytes(int n, int k) {
  return size_t(sm_count()) * (size_t(n) * k + 2 * size_t(n)) * sizeof(float);
}

extern "C" int pc_wgrad_tf32x3(const float* dy, int64_t m, int n, int64_t ld_dy, const float* x, int k, int64_t ld_x,
                               float* dw, float* db, void* workspace, size_t workspace_bytes, pc_stream_t stream) {
  PC_REQUIRE(m > 0, PC_ERR_INVALID, "wgrad: need at least one row");
  PC_REQUIRE(dy && x && dw && workspace, PC_ERR_INVALID, "wgrad: null pointer");
  PC_REQUIRE(n >= 128 && n % 128 == 0 && n <= 1024 && k >= 32 && k % 32 == 0 && k <= WG_MAX_K, PC_ERR_UNSUPPORTED,
             "wgrad: n=%d must be a multiple of 128 and k=%d a multiple of 32 up to 256", n, k);
  PC_REQUIRE(ld_dy % 4 == 0 && ld_x % 4 == 0, PC_ERR_INVALID, "wgrad: leading dimensions must be multiples of 4 floats");
  PC_REQUIRE(workspace_bytes >= pc_wgrad_workspace_bytes(n, k), PC_ERR_WORKSPACE, "wgrad: workspace too small");
  const int64_t blocks = (m + WG_ROWS - 1) / WG_ROWS;
  const int halves = n / BM;
  int subsets = sm_count() / halves;
  if (subsets < 1) subsets = 1;
  if (blocks < subsets) subsets = int(blocks);
  const int grid = subsets * halves;
  WgradParams p;
  p.m = m; p.n = n; p.k = k; p.halves = halves;
  p.partial_w = reinterpret_cast<float*>(workspace);
  p.partial_b = db ? p.partial_w + size_t(subsets) * n * k : nullptr;
  CUtensorMap map_dy, map_x, map_pw;
  if (int rc = make_map_mn(&map_dy, dy, m, n, ld_dy, 4)) return rc;
  if (int rc = make_map_mn(&map_x, x, m, k, ld_x, k / 32)) return rc;
  if (int rc = make_map(&map_pw, p.partial_w, int64_t(subsets) * n, k, k, BM)) return rc;
  static bool configured[64] = {};
  if (first_use_on_device(configured))
    PC_CUDA(cudaFuncSetAttribute(wgrad_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
  cudaStream_t st = as_stream(stream);
  wgrad_tf32x3_kernel<<<grid, GEMM_THREADS, WG_SMEM, st>>>(map_dy, map_x, map_pw, p);
  PC_LAUNCH_CHECK();
  const int64_t nk = int64_t(n) * k;
  reduce_partials_kernel<<<unsigned((nk + 255) / 256), 256, 0, st>>>(p.partial_w, subsets, nk, dw);
  PC_LAUNCH_CHECK();
  if (db) {
    reduce_partials_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(p.partial_b, 2 * subsets, n, db);
    PC_LAUNCH_CHECK();
  }
  return PC_OK;
}
