// Host-side planning decisions of the upload path, free of CUDA so that they are unit-tested on a machine without a GPU
// (tests/cpp/plan_test.cpp, tests/test_plan.py): which coarse space a graph gets (ira_coarse.cuh) and how the SELL slices
// are dealt to the thread blocks of the persistent PCG kernels (ira_pcg.cuh).
#pragma once
#include <algorithm>
#include <cstdint>
#include <utility>
#include <vector>

namespace ira {
namespace plan {

constexpr int kSellRows = 32;           // rows per SELL slice (= kSellC, one warp)
constexpr int kCoarseMax = 64;          // dense variant: coarse unknowns per coordinate (3 x 64 x 64 doubles = 96 KB of shared memory)
constexpr int kCoarseMaxRows = 32768;   // larger graphs keep the one-level kernels (148 x 12 warps hold 56 832 rows)
constexpr int kTriMax = 1024;           // tridiagonal variant: coarse unknowns per coordinate (6 x 3 x 1024 doubles = 144 KB)
constexpr int kTriMinBlock = 8;         // rows per block of the partition at least (the window edges must stay within adjacent blocks)

inline int cdiv_i(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Tridiagonal coarse operator: rows per block of the partition (8, 16 or 32: an aligned lane group of one warp once the
// SELL pattern is in index order), or 0 when the graph does not qualify.  Needs more than kTriMinBlock * kCoarseMax free
// nodes (below that the dense 64-block space is as fine), at most kTriMax blocks, and at most 2 % of the edges may
// reach beyond the adjacent block - those are lumped onto the diagonal of the coarse operator.
inline int tri_block_rows(int n, int f, int64_t m, const int32_t* I_pairs) {
  if (n > kCoarseMaxRows || n - f <= kTriMinBlock * kCoarseMax) return 0;
  for (int bsz = kTriMinBlock; bsz <= kSellRows; bsz *= 2) {
    if (cdiv_i(n, bsz) > kTriMax) continue;
    if (cdiv_i(n, bsz) <= kCoarseMax) break;
    int64_t far = 0;
    for (int64_t k = 0; k < m; ++k) {
      const int d = I_pairs[2 * k] / bsz - I_pairs[2 * k + 1] / bsz;
      if (d > 1 || d < -1) ++far;
    }
    if (far * 50 <= m) return bsz;
  }
  return 0;
}

// Dense coarse space: partition into nc <= kCoarseMax contiguous index blocks of bsz rows; false when the graph is too
// small / too large or not a chain in its node numbering.  Long-range edges (loop closures) make a view graph an
// expander: config 2's 9 % of them leave one-level PCG at 64 iterations per solve and the two-level kernel, at 2.5x the
// cost per iteration, loses (measured 19.7 against 9.9 ms); a SLAM stream (0.05 %) or the reference's fixture (none)
// needs thousands of iterations without the coarse space.  Threshold: at most 2 % of the edges span more than two blocks.
inline bool dense_partition(int n, int f, int64_t m, const int32_t* I_pairs, int* bsz_out, int* nc_out) {
  if (n > kCoarseMaxRows || n - f < 128) return false;
  const int bsz = std::max(2, cdiv_i(n, kCoarseMax));
  const int nc = cdiv_i(n, bsz);
  if (nc < 2 || nc > kCoarseMax) return false;
  int64_t far = 0;
  for (int64_t k = 0; k < m; ++k) {
    const int64_t d = (int64_t)I_pairs[2 * k] - (int64_t)I_pairs[2 * k + 1];
    if ((d < 0 ? -d : d) > 2 * (int64_t)bsz) ++far;
  }
  if (far * 50 > m) return false;
  *bsz_out = bsz; *nc_out = nc;
  return true;
}

// Deal of SELL slices to (block, warp) slots so that every block gathers about the same number of entries: slices
// longest-first to the block with the fewest entries that still has a free warp (LPT); ties by block index, so the map
// is a pure function of the widths.  map[b * wpb + w] = slice or -1.  Degree-sorted 1 024-row windows make slice widths
// periodic (32 slices per window); the round-robin deal slice -> block slice % grid aliases with that period and leaves
// the heaviest block 14 % above the mean on config 3 - and every grid barrier waits for the heaviest block.
inline void lpt_slice_map(const int* width, int nslices, int grid, int wpb, std::vector<int>* map_out) {
  std::vector<int> order((size_t)nslices);
  for (int s = 0; s < nslices; ++s) order[(size_t)s] = s;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return width[a] > width[b]; });
  std::vector<int64_t> load((size_t)grid, 0);
  std::vector<int> used((size_t)grid, 0);
  std::vector<int>& map = *map_out;
  map.assign((size_t)grid * wpb, -1);
  typedef std::pair<int64_t, int> Item;                       // (load, block) min-heap
  std::vector<Item> heap;
  for (int b = 0; b < grid; ++b) heap.push_back(Item(0, b));
  auto cmp = [](const Item& a, const Item& b) { return a > b; };
  std::make_heap(heap.begin(), heap.end(), cmp);
  for (int s : order) {
    if (heap.empty()) break;                                  // more slices than slots: the caller checks grid * wpb >= nslices
    std::pop_heap(heap.begin(), heap.end(), cmp);
    const Item it = heap.back();
    heap.pop_back();
    const int b = it.second;
    map[(size_t)b * wpb + used[(size_t)b]] = s;
    used[(size_t)b] += 1;
    load[(size_t)b] += std::max(width[s], 1);
    if (used[(size_t)b] < wpb) { heap.push_back(Item(load[(size_t)b], b)); std::push_heap(heap.begin(), heap.end(), cmp); }
  }
}

}  // namespace plan
}  // namespace ira
