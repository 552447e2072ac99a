// Device-side negative sampling for the similarity (triplet) batches.
//
// Replaces SimilarityDataset._get_negative_samples of /root/reference/src/data/data_loader.py:27-40
// (per item: rebuild the anchor's similar set by scanning all pairs, then rejection-sample k products
// that are not the anchor, not similar to it and pairwise distinct).  Here the similar set is one CSR row
// (binary search) and every anchor is one thread with a counter-based generator, so a batch of negatives is
// one launch and is reproducible from (seed, batch slot) alone.
#include "common.cuh"

namespace pc {
namespace {

__device__ __forceinline__ uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void sample_negatives_kernel(const int32_t* __restrict__ anchor, const int64_t* __restrict__ sim_rowptr,
                                        const int32_t* __restrict__ sim_col, int64_t batch, int32_t n_nodes, int k,
                                        uint64_t seed, int32_t* __restrict__ out) {
  const int64_t b = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int32_t a = anchor[b];
  const int64_t lo = sim_rowptr[a], hi = sim_rowptr[a + 1];
  int32_t* mine = out + b * k;
  uint64_t state = splitmix(seed ^ (uint64_t(b) * 0xD1B54A32D192ED03ull));
  int got = 0;
  const int max_attempts = 64 * k + 64;
  for (int attempt = 0; got < k && attempt < max_attempts; ++attempt) {
    state = splitmix(state);
    const int32_t c = int32_t((state >> 11) % uint64_t(n_nodes));   // uniform over products (random.choice, :34)
    bool ok = c != a;                                               // :35
    if (ok) {                                                       // :36  not in the anchor's similar set
      int64_t l = lo, h = hi;
      while (l < h) {
        const int64_t mid = (l + h) >> 1;
        if (sim_col[mid] < c) l = mid + 1; else h = mid;
      }
      ok = !(l < hi && sim_col[l] == c);
    }
    for (int j = 0; ok && j < got; ++j) ok = mine[j] != c;          // :37  distinct
    if (ok) mine[got++] = c;
  }
  // degenerate graphs (almost every product similar to the anchor): finish with a linear scan, never spin forever
  for (int32_t c = 0; got < k && c < n_nodes; ++c) {
    bool ok = c != a;
    if (ok) {
      int64_t l = lo, h = hi;
      while (l < h) {
        const int64_t mid = (l + h) >> 1;
        if (sim_col[mid] < c) l = mid + 1; else h = mid;
      }
      ok = !(l < hi && sim_col[l] == c);
    }
    for (int j = 0; ok && j < got; ++j) ok = mine[j] != c;
    if (ok) mine[got++] = c;
  }
  for (; got < k; ++got) mine[got] = -1;                            // fewer than k eligible products exist
}

}  // namespace
}  // namespace pc

using namespace pc;

extern "C" int pc_sample_negatives(const int32_t* anchor, const int64_t* sim_rowptr, const int32_t* sim_col, int64_t batch,
                                   int32_t n_nodes, int k, uint64_t seed, int32_t* out, pc_stream_t stream) {
  PC_REQUIRE(batch >= 0 && n_nodes > 0 && k >= 1 && k <= 64, PC_ERR_INVALID, "sample_negatives: bad sizes");
  if (batch == 0) return PC_OK;
  PC_REQUIRE(anchor && sim_rowptr && out, PC_ERR_INVALID, "sample_negatives: null pointer");
  sample_negatives_kernel<<<unsigned(ceil_div(batch, 128)), 128, 0, as_stream(stream)>>>(anchor, sim_rowptr, sim_col, batch,
                                                                                      n_nodes, k, seed, out);
  PC_LAUNCH_CHECK();
  return PC_OK;
}
