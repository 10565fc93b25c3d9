// domain.cpp -- host-side helpers for the flattening that feeds the path.
//
// wbem_generate_double_nodes_set restates ComputationalDomain<3>::generate_double_nodes_set
// (reference source/computational_domain.cc:258-307): set[i] = {i} for an interior dof, and
// {i} U {j : |x_i - x_j| < tol} (all dofs j) for a boundary dof i.  The reference tests every
// boundary dof against every dof (O(N_b N), serial); here the points are binned in a uniform grid
// of cell size tol, so only the 27 neighbouring cells are inspected: O(N) expected.  Pure host
// code, no GPU involved; the result is the CSR (dn_ptr, dn_idx) wbem_set_topology takes.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <vector>

#include "../../include/wbem.h"

namespace
{
struct CellKey
{
  int64_t x, y, z;
  bool operator==(const CellKey &o) const { return x == o.x && y == o.y && z == o.z; }
};
struct CellHash
{
  size_t operator()(const CellKey &k) const
  {
    uint64_t h = (uint64_t)k.x * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)k.y + 0x7F4A7C15F39CC060ull) * 0xC2B2AE3D27D4EB4Full;
    h = (h << 31) | (h >> 33);
    h ^= ((uint64_t)k.z + 0x165667B19E3779F9ull) * 0x94D049BB133111EBull;
    return (size_t)(h ^ (h >> 29));
  }
};
} // namespace

extern "C" int wbem_generate_double_nodes_set(uint32_t n_dofs, const double *support_points,
                                              const uint8_t *boundary_dofs, double tol, uint32_t *dn_ptr,
                                              uint32_t *dn_idx, uint64_t capacity, uint64_t *needed)
{
  if (!support_points || !dn_ptr || !(tol > 0.0)) return -1;
  const double inv = 1.0 / tol;
  auto key = [&](uint32_t i) {
    return CellKey{(int64_t)std::floor(support_points[3 * (size_t)i] * inv),
                   (int64_t)std::floor(support_points[3 * (size_t)i + 1] * inv),
                   (int64_t)std::floor(support_points[3 * (size_t)i + 2] * inv)};
  };
  std::unordered_map<CellKey, std::vector<uint32_t>, CellHash> grid;
  grid.reserve((size_t)n_dofs * 2);
  for (uint32_t i = 0; i < n_dofs; ++i) grid[key(i)].push_back(i);
  uint64_t total = 0;
  std::vector<uint32_t> set;
  dn_ptr[0] = 0;
  for (uint32_t i = 0; i < n_dofs; ++i)
    {
      set.clear();
      set.push_back(i);
      if (!boundary_dofs || boundary_dofs[i])
        {
          const CellKey c = key(i);
          const double *xi = support_points + 3 * (size_t)i;
          for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy)
              for (int dz = -1; dz <= 1; ++dz)
                {
                  auto it = grid.find(CellKey{c.x + dx, c.y + dy, c.z + dz});
                  if (it == grid.end()) continue;
                  for (uint32_t j : it->second)
                    {
                      if (j == i) continue;
                      const double *xj = support_points + 3 * (size_t)j;
                      const double d = std::sqrt((xi[0] - xj[0]) * (xi[0] - xj[0]) + (xi[1] - xj[1]) * (xi[1] - xj[1]) +
                                                 (xi[2] - xj[2]) * (xi[2] - xj[2])); // Point<3>::distance
                      if (d < tol) set.push_back(j);
                    }
                }
          std::sort(set.begin(), set.end()); // std::set<unsigned int> iteration order
        }
      if (dn_idx && total + set.size() <= capacity) std::copy(set.begin(), set.end(), dn_idx + total);
      total += set.size();
      if (total > 0xffffffffull) return -1;
      dn_ptr[i + 1] = (uint32_t)total;
    }
  if (needed) *needed = total;
  return (dn_idx && total <= capacity) ? 0 : 1; // 1: call again with a buffer of *needed entries
}
