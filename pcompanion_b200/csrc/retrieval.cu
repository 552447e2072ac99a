// Complementary retrieval: exact segment scoring + warp-level top-K, and top-K list merging.
//
// Replaces torch.matmul + torch.topk of /root/reference/src/models/p_companion.py:60-64 and
// /root/reference/src/utils/metrics.py:21,89, and the per-type filter -> matmul -> topk loop
// of /root/reference/inference.py:93-113.  Each score row ranks one contiguous run of catalog
// row ids (the members of its complementary type in a type-sorted permutation, or a plain
// slice of the catalog), which is exactly what the reference's per-type filter computes - the
// masked-out 99.9 % of a dense [R, P] score matrix is never formed.
//
// Scores are float64 sums of exact float32 x float32 products in a fixed order (lane l of a
// warp owns dims 128c + 4l .. 4l+3 sequentially, then a butterfly over lanes at distance
// 16, 8, 4, 2, 1), mirrored by oracle/retrieval.py, so indices match the oracle bit for bit;
// ties rank the lower catalog index first.
#include <math.h>

#include "common.cuh"

namespace pc {
namespace {

constexpr int TK_WARPS = 8;

struct Cand {
  double s;
  int64_t i;  // < 0: empty slot
};

// a ranks strictly before b
__device__ __forceinline__ bool ranks_before(double sa, int64_t ia, double sb, int64_t ib) {
  if (ib < 0) return ia >= 0;
  if (ia < 0) return false;
  return sa > sb || (sa == sb && ia < ib);
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ int64_t shfl_i64(int64_t v, int src) { return __shfl_sync(FULL, (long long)v, src); }

// Warp-distributed sorted list: lane j (< k) holds the j-th best entry.  Insert (s, i) - known to all lanes.
__device__ __forceinline__ void warp_topk_insert(Cand& mine, int k, double s, int64_t i) {
  const int lane = lane_id();
  const bool before = lane < k && ranks_before(s, i, mine.s, mine.i);
  const uint32_t mask = __ballot_sync(FULL, before);
  if (mask == 0) return;
  const int pos = __ffs(mask) - 1;
  const double up_s = __shfl_up_sync(FULL, mine.s, 1);
  const int64_t up_i = __shfl_up_sync(FULL, (long long)mine.i, 1);
  if (lane == pos) {
    mine.s = s;
    mine.i = i;
  } else if (lane > pos) {
    mine.s = up_s;
    mine.i = up_i;
  }
}

// 32 per-lane partial sums (v[j] = this lane's share of product j) -> lane j holds the full sum of
// product j, combining lanes at distance 16, then 8, 4, 2, 1.
__device__ __forceinline__ double transpose_reduce(double (&v)[32]) {
  const int lane = lane_id();
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const double keep = upper ? v[j + half] : v[j];
      const double send = upper ? v[j] : v[j + half];
      v[j] = keep + __shfl_xor_sync(FULL, send, half);
    }
  }
  return v[0];
}

// grid = (splits, rows).  CTA (split, r) ranks its share of row r's segment; warp lists are merged
// by warp 0 and written to part[(r * splits + split) * k ..].
template <bool PRECOMP>
__global__ void __launch_bounds__(TK_WARPS * 32)
topk_segments_kernel(const float* __restrict__ Q, int dim, const float* __restrict__ catalog,
                     const int32_t* __restrict__ members, const int64_t* __restrict__ seg_begin,
                     const int64_t* __restrict__ seg_end, int k, int splits, int64_t index_base,
                     double* __restrict__ part_s, int64_t* __restrict__ part_i) {
  extern __shared__ double q_s[];  // dim doubles (unused when PRECOMP), then TK_WARPS * 32 candidates
  Cand* lists = reinterpret_cast<Cand*>(q_s + (PRECOMP ? 0 : dim));
  const int lane = lane_id(), w = warp_id();
  const int64_t r = blockIdx.y;
  const int split = blockIdx.x;
  if (!PRECOMP) {
    for (int d = threadIdx.x; d < dim; d += blockDim.x) q_s[d] = double(Q[r * dim + d]);
    __syncthreads();
  }
  // PRECOMP: Q is a materialised [rows, dim] fp32 score matrix and every row ranks columns [0, dim)
  const int64_t beg = PRECOMP ? 0 : seg_begin[r], end = PRECOMP ? int64_t(dim) : seg_end[r];
  const int64_t len = end > beg ? end - beg : 0;
  const int64_t per = (ceil_div(len, int64_t(splits)) + 31) / 32 * 32;
  const int64_t sb = beg + per * split;
  const int64_t se = min(end, sb + per);
  Cand mine{-INFINITY, -1};
  const int chunks = dim / 128;
  for (int64_t g = sb + int64_t(w) * 32; g < se; g += TK_WARPS * 32) {
    const int64_t my_pos = g + lane;
    int64_t my_member = -1;
    if (my_pos < se) my_member = members ? int64_t(members[my_pos]) : my_pos;
    double score;
    if (PRECOMP) {
      score = my_member >= 0 ? double(__ldg(Q + r * dim + my_member)) : 0.0;
    } else {
    double v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int64_t m = __shfl_sync(FULL, (long long)my_member, j);
      double acc = 0.0;
      if (m >= 0) {
        const float4* rowp = reinterpret_cast<const float4*>(catalog + m * dim) + lane;
        for (int c = 0; c < chunks; ++c) {
          const float4 x = ld_stream4(rowp + c * 32);
          const double* qq = q_s + c * 128 + lane * 4;
          acc = fma(qq[0], double(x.x), acc);
          acc = fma(qq[1], double(x.y), acc);
          acc = fma(qq[2], double(x.z), acc);
          acc = fma(qq[3], double(x.w), acc);
        }
      }
      v[j] = acc;
    }
    score = transpose_reduce(v);
    }
    const int64_t gidx = my_member >= 0 ? my_member + index_base : -1;
    const double thr_s = shfl_d(mine.s, k - 1);
    const int64_t thr_i = shfl_i64(mine.i, k - 1);
    uint32_t cand = __ballot_sync(FULL, gidx >= 0 && ranks_before(score, gidx, thr_s, thr_i));
    while (cand) {
      const int src = __ffs(cand) - 1;
      cand &= cand - 1;
      warp_topk_insert(mine, k, shfl_d(score, src), shfl_i64(gidx, src));
    }
  }
  lists[w * 32 + lane] = mine;
  __syncthreads();
  if (w == 0) {
    for (int ww = 1; ww < TK_WARPS; ++ww) {
      for (int j = 0; j < k; ++j) {
        const Cand c = lists[ww * 32 + j];
        if (c.i < 0) break;  // lists are sorted: the rest is empty
        warp_topk_insert(mine, k, c.s, c.i);
      }
    }
    if (lane < k) {
      const int64_t o = (r * splits + split) * k + lane;
      part_s[o] = mine.s;
      part_i[o] = mine.i;
    }
  }
}

// one warp per row: merge `lists` sorted-or-not candidate lists of k entries
__global__ void __launch_bounds__(256)
topk_merge_kernel(const double* __restrict__ S, const int64_t* __restrict__ I, int64_t rows, int lists, int k,
                  double* __restrict__ out_s, int64_t* __restrict__ out_i) {
  const int64_t r = int64_t(blockIdx.x) * 8 + warp_id();
  if (r >= rows) return;
  const int lane = lane_id();
  Cand mine{-INFINITY, -1};
  const int64_t total = int64_t(lists) * k;
  for (int64_t base = 0; base < total; base += 32) {
    const int64_t c = base + lane;
    double s = -INFINITY;
    int64_t i = -1;
    if (c < total) {
      s = S[r * total + c];
      i = I[r * total + c];
    }
    uint32_t cand = __ballot_sync(FULL, i >= 0);
    while (cand) {
      const int src = __ffs(cand) - 1;
      cand &= cand - 1;
      warp_topk_insert(mine, k, shfl_d(s, src), shfl_i64(i, src));
    }
  }
  if (lane < k) {
    out_s[r * k + lane] = mine.s;
    out_i[r * k + lane] = mine.i;
  }
}

}  // namespace
}  // namespace pc

using namespace pc;

extern "C" size_t pc_topk_segments_workspace_bytes(int64_t rows, int k, int splits) {
  if (rows <= 0 || k <= 0 || splits <= 1) return 0;
  return size_t(rows) * size_t(splits) * size_t(k) * (sizeof(double) + sizeof(int64_t));
}

extern "C" int pc_topk_segments(const float* q, int64_t rows, int dim, const float* catalog, const int32_t* members,
                                const int64_t* seg_begin, const int64_t* seg_end, int k, int splits,
                                int64_t index_base, double* out_scores, int64_t* out_idx, void* workspace,
                                size_t workspace_bytes, pc_stream_t stream) {
  PC_REQUIRE(rows >= 0, PC_ERR_INVALID, "topk_segments: negative rows");
  if (rows == 0) return PC_OK;
  PC_REQUIRE(q && catalog && seg_begin && seg_end && out_scores && out_idx, PC_ERR_INVALID, "topk_segments: null pointer");
  PC_REQUIRE(k >= 1 && k <= 32, PC_ERR_UNSUPPORTED, "topk_segments: k=%d outside [1,32]", k);
  PC_REQUIRE(dim >= 128 && dim % 128 == 0 && dim <= 2048, PC_ERR_UNSUPPORTED, "topk_segments: dim=%d must be a multiple of 128 (<= 2048)", dim);
  PC_REQUIRE(splits >= 1 && splits <= 65535 && rows <= 65535 * int64_t(32768), PC_ERR_UNSUPPORTED, "topk_segments: bad splits/rows");
  cudaStream_t st = as_stream(stream);
  const size_t smem = size_t(dim) * sizeof(double) + TK_WARPS * 32 * sizeof(Cand);
  double* ps = out_scores;
  int64_t* pi = out_idx;
  if (splits > 1) {
    PC_REQUIRE(workspace && workspace_bytes >= pc_topk_segments_workspace_bytes(rows, k, splits), PC_ERR_WORKSPACE,
               "topk_segments: workspace too small");
    ps = reinterpret_cast<double*>(workspace);
    pi = reinterpret_cast<int64_t*>(ps + size_t(rows) * splits * k);
  }
  // gridDim.y is limited to 65535: walk the rows in slabs
  for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
    const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
    dim3 grid{unsigned(splits), unsigned(nr), 1u};
    topk_segments_kernel<false><<<grid, TK_WARPS * 32, smem, st>>>(q + r0 * dim, dim, catalog, members, seg_begin + r0,
                                                            seg_end + r0, k, splits, index_base,
                                                            ps + r0 * splits * k, pi + r0 * splits * k);
    PC_LAUNCH_CHECK();
  }
  if (splits > 1) return pc_topk_merge(ps, pi, rows, splits, k, out_scores, out_idx, stream);
  return PC_OK;
}

extern "C" int pc_topk_merge(const double* scores, const int64_t* idx, int64_t rows, int lists, int k,
                             double* out_scores, int64_t* out_idx, pc_stream_t stream) {
  PC_REQUIRE(rows >= 0 && lists >= 1, PC_ERR_INVALID, "topk_merge: bad sizes");
  if (rows == 0) return PC_OK;
  PC_REQUIRE(scores && idx && out_scores && out_idx, PC_ERR_INVALID, "topk_merge: null pointer");
  PC_REQUIRE(k >= 1 && k <= 32, PC_ERR_UNSUPPORTED, "topk_merge: k=%d outside [1,32]", k);
  topk_merge_kernel<<<unsigned(ceil_div(rows, 8)), 256, 0, as_stream(stream)>>>(scores, idx, rows, lists, k,
                                                                                out_scores, out_idx);
  PC_LAUNCH_CHECK();
  return PC_OK;
}

extern "C" size_t pc_topk_rows_workspace_bytes(int64_t rows, int k, int splits) {
  return pc_topk_segments_workspace_bytes(rows, k, splits);
}

extern "C" int pc_topk_rows(const float* values, int64_t rows, int64_t cols, int k, int splits, double* out_scores,
                            int64_t* out_idx, void* workspace, size_t workspace_bytes, pc_stream_t stream) {
  PC_REQUIRE(rows >= 0 && cols >= 0 && cols < (int64_t(1) << 31), PC_ERR_INVALID, "topk_rows: bad sizes");
  if (rows == 0) return PC_OK;
  PC_REQUIRE(values && out_scores && out_idx, PC_ERR_INVALID, "topk_rows: null pointer");
  PC_REQUIRE(k >= 1 && k <= 32, PC_ERR_UNSUPPORTED, "topk_rows: k=%d outside [1,32]", k);
  PC_REQUIRE(splits >= 1 && splits <= 65535, PC_ERR_UNSUPPORTED, "topk_rows: bad splits");
  cudaStream_t st = as_stream(stream);
  double* ps = out_scores;
  int64_t* pi = out_idx;
  if (splits > 1) {
    PC_REQUIRE(workspace && workspace_bytes >= pc_topk_rows_workspace_bytes(rows, k, splits), PC_ERR_WORKSPACE,
               "topk_rows: workspace too small");
    ps = reinterpret_cast<double*>(workspace);
    pi = reinterpret_cast<int64_t*>(ps + size_t(rows) * splits * k);
  }
  const size_t smem = TK_WARPS * 32 * sizeof(Cand);
  for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
    const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
    dim3 grid{unsigned(splits), unsigned(nr), 1u};
    topk_segments_kernel<true><<<grid, TK_WARPS * 32, smem, st>>>(values + r0 * cols, int(cols), nullptr, nullptr,
                                                                  nullptr, nullptr, k, splits, 0,
                                                                  ps + r0 * splits * k, pi + r0 * splits * k);
    PC_LAUNCH_CHECK();
  }
  if (splits > 1) return pc_topk_merge(ps, pi, rows, splits, k, out_scores, out_idx, stream);
  return PC_OK;
}
