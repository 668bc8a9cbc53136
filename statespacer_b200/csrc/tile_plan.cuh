// Host-side planning of the internal state order that makes T tile-sparse (shared by the warp-level loglik kernels).
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace ssgpu {

// ---- state ordering that makes T tile-sparse ------------------------------------------------------
// States i, k are coupled when T[i][k] or T[k][i] is non-zero.  Coupled components are packed whole into the
// 8-state tiles (exact search: at most 23 components, 3 tiles); inside a tile the caller's order is kept.
// statespacer models are block diagonal in T by construction (residuals: zero rows/columns, level, slope,
// trigonometric seasonal / cycle pairs, regressor coefficients: blocks of size 1-2), so they pack; a SARIMA
// companion block larger than a tile does not and runs dense in the caller's order.
struct TilePlan {
  signed char perm[64];
  unsigned long long mask;  // bit rt*TT + ct of the ordered T (TT <= 8)
};

inline bool pack_components(const std::vector<std::vector<int>>& comps, size_t idx, int* cap, int n_bins,
                            std::vector<int>& bin_of, long* budget) {
  if (idx == comps.size()) return true;
  if (--*budget < 0) return false;  // give up on pathological instances: the caller falls back to the dense order
  const int sz = (int)comps[idx].size();
  for (int b = 0; b < n_bins; b++) {
    if (cap[b] < sz) continue;
    bool dup = false;  // bins with equal remaining capacity are interchangeable
    for (int b2 = 0; b2 < b; b2++) dup = dup || (cap[b2] == cap[b]);
    if (dup) continue;
    cap[b] -= sz;
    bin_of[idx] = b;
    if (pack_components(comps, idx + 1, cap, n_bins, bin_of, budget)) return true;
    cap[b] += sz;
  }
  return false;
}

inline unsigned long long tile_mask_of(const std::vector<double>& Th, int diag, int m, int tt, const signed char* perm) {
  unsigned long long mask = 0;
  for (int i = 0; i < m; i++)
    for (int k = 0; k < m; k++)
      if (mat_get(Th.data(), diag, m, perm[i], perm[k]) != 0.0) mask |= 1ull << ((i >> 3) * tt + (k >> 3));
  return mask;
}

inline TilePlan plan_tiles(const std::vector<double>& Th, int diag, int m, int tt, bool allow_reorder) {
  TilePlan pl{};
  for (int i = 0; i < 64; i++) pl.perm[i] = (signed char)(i < m ? i : 0);
  const unsigned long long dense = tt * tt >= 64 ? ~0ull : (1ull << (tt * tt)) - 1ull;
  pl.mask = dense;
  if (!allow_reorder || tt == 1) return pl;
  // components of the coupling graph (union-find)
  std::vector<int> root(m);
  for (int i = 0; i < m; i++) root[i] = i;
  auto find = [&](int x) {
    while (root[x] != x) x = root[x] = root[root[x]];
    return x;
  };
  for (int i = 0; i < m; i++)
    for (int k = 0; k < m; k++)
      if (i != k && mat_get(Th.data(), diag, m, i, k) != 0.0) root[find(i)] = find(k);
  std::vector<std::vector<int>> comps;
  {
    std::vector<int> id(m, -1);
    for (int i = 0; i < m; i++) {
      const int rr = find(i);
      if (id[rr] < 0) {
        id[rr] = (int)comps.size();
        comps.emplace_back();
      }
      comps[id[rr]].push_back(i);
    }
  }
  std::stable_sort(comps.begin(), comps.end(),
                   [](const std::vector<int>& x, const std::vector<int>& y) { return x.size() > y.size(); });
  int cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long budget = 200000;
  const int n_bins = (m + 7) / 8;
  for (int b = 0; b < n_bins; b++) cap[b] = std::min(8, m - 8 * b);
  std::vector<int> bin_of(comps.size(), 0);
  if (!pack_components(comps, 0, cap, n_bins, bin_of, &budget)) return pl;  // a component larger than a tile: dense
  int pos = 0;
  for (int b = 0; b < n_bins; b++) {
    std::vector<int> members;
    for (size_t c = 0; c < comps.size(); c++)
      if (bin_of[c] == b) members.insert(members.end(), comps[c].begin(), comps[c].end());
    std::sort(members.begin(), members.end());
    for (int s : members) pl.perm[pos++] = (signed char)s;
  }
  pl.mask = tile_mask_of(Th, diag, m, tt, pl.perm);
  return pl;
}

}  // namespace ssgpu
