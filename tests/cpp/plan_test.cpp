// CPU unit test of the host-side planning decisions (irotavg_b200/csrc/ira_plan.hpp); built with g++ by
// tests/test_plan.py.  Prints one "ok <name>" line per check, exits non-zero on the first failure.
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>

#include "ira_plan.hpp"

using namespace ira::plan;

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

// view graph in frame order: every view tied to the previous `back` views, plus `loops` long-range edges
static std::vector<int32_t> chain(int n, int back, int loops, unsigned seed) {
  std::vector<int32_t> I;
  for (int v = 1; v < n; ++v)
    for (int b = 1; b <= back && v - b >= 0; ++b) { I.push_back(v - b); I.push_back(v); }
  unsigned s = seed;
  for (int k = 0; k < loops; ++k) {
    s = s * 1664525u + 1013904223u; const int a = (int)(s % (unsigned)n);
    s = s * 1664525u + 1013904223u; const int b = (int)(s % (unsigned)n);
    if (a != b) { I.push_back(a); I.push_back(b); }
  }
  return I;
}

int main() {
  {  // chain-like graphs take the tridiagonal space with the smallest block that keeps <= 1 024 blocks
    std::vector<int32_t> I = chain(9501, 4, 18, 1u);
    CHECK(tri_block_rows(9501, 1, (int64_t)I.size() / 2, I.data()) == 16);
    I = chain(4001, 4, 8, 2u);
    CHECK(tri_block_rows(4001, 1, (int64_t)I.size() / 2, I.data()) == 8);
    I = chain(1832, 5, 0, 3u);                                   // the bundled fixture's shape
    CHECK(tri_block_rows(1832, 1, (int64_t)I.size() / 2, I.data()) == 8);
    I = chain(30000, 4, 10, 4u);
    CHECK(tri_block_rows(30000, 1, (int64_t)I.size() / 2, I.data()) == 32);
    std::puts("ok tri_block_rows: chains");
  }
  {  // too small, too large, too many loop closures, windows wider than a block
    std::vector<int32_t> I = chain(500, 4, 0, 5u);
    CHECK(tri_block_rows(500, 1, (int64_t)I.size() / 2, I.data()) == 0);           // <= 512 free nodes: dense space
    I = chain(600, 4, 0, 5u);
    CHECK(tri_block_rows(600, 100, (int64_t)I.size() / 2, I.data()) == 0);         // 500 free nodes
    I = chain(40000, 4, 0, 6u);
    CHECK(tri_block_rows(40000, 1, (int64_t)I.size() / 2, I.data()) == 0);         // > kCoarseMaxRows
    I = chain(4541, 4, 4645, 7u);                                                  // config 2: ~20 % long-range edges
    CHECK(tri_block_rows(4541, 1, (int64_t)I.size() / 2, I.data()) == 0);
    I = chain(4000, 20, 0, 8u);                                                    // windows of 20: blocks of 8 are too
    const int b = tri_block_rows(4000, 1, (int64_t)I.size() / 2, I.data());        // narrow, 16 / 32 hold them
    CHECK(b == 16 || b == 32);
    std::vector<int32_t> J = chain(4000, 40, 0, 9u);                               // windows of 40: beyond 2 x 32 rows
    CHECK(tri_block_rows(4000, 1, (int64_t)J.size() / 2, J.data()) == 0);
    std::puts("ok tri_block_rows: rejections");
  }
  {  // dense 64-block space
    int bsz = 0, nc = 0;
    std::vector<int32_t> I = chain(1832, 5, 0, 3u);
    CHECK(dense_partition(1832, 1, (int64_t)I.size() / 2, I.data(), &bsz, &nc) && bsz == 29 && nc == 64);
    I = chain(300, 4, 0, 3u);
    CHECK(dense_partition(300, 1, (int64_t)I.size() / 2, I.data(), &bsz, &nc) && bsz == 5 && nc == 60);
    CHECK(!dense_partition(100, 1, (int64_t)I.size() / 2, I.data(), &bsz, &nc));   // < 128 free nodes
    I = chain(4541, 4, 4645, 7u);
    CHECK(!dense_partition(4541, 1, (int64_t)I.size() / 2, I.data(), &bsz, &nc));  // expander
    std::puts("ok dense_partition");
  }
  {  // LPT deal: every slice exactly once, <= wpb per block, balanced, a pure function of the widths
    const int nslices = 3125, grid = 148, wpb = 24;                                // config 3
    std::vector<int> w((size_t)nslices);
    for (int s = 0; s < nslices; ++s) w[(size_t)s] = 44 - (s % 32);                 // degree-sorted windows: period 32
    std::vector<int> map, map2;
    lpt_slice_map(w.data(), nslices, grid, wpb, &map);
    lpt_slice_map(w.data(), nslices, grid, wpb, &map2);
    CHECK(map == map2 && (int)map.size() == grid * wpb);
    std::vector<int> seen((size_t)nslices, 0);
    std::vector<long> load((size_t)grid, 0), rr((size_t)grid, 0);
    for (int b = 0; b < grid; ++b)
      for (int k = 0; k < wpb; ++k) {
        const int s = map[(size_t)b * wpb + k];
        if (s >= 0) { CHECK(s < nslices); seen[(size_t)s] += 1; load[(size_t)b] += w[(size_t)s]; }
      }
    for (int s = 0; s < nslices; ++s) { CHECK(seen[(size_t)s] == 1); rr[(size_t)(s % grid)] += w[(size_t)s]; }
    const double mean = std::accumulate(load.begin(), load.end(), 0.0) / grid;
    const long mx = *std::max_element(load.begin(), load.end()), mx_rr = *std::max_element(rr.begin(), rr.end());
    std::printf("lpt max/mean %.4f, round-robin max/mean %.4f\n", mx / mean, mx_rr / mean);
    CHECK(mx <= 1.03 * mean);
    CHECK(mx_rr >= 1.08 * mean);                                                   // what the LPT deal replaces
    std::puts("ok lpt_slice_map");
  }
  return 0;
}
