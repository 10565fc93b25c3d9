// plan.cpp -- once-per-mesh tiling plan of the regular-pair assembly kernel.
//
// The reference scatters 8 partial sums per (node i, cell) pair into row i at the cell's
// four dof columns (source/bem_problem.cc:531-537).  On the GPU that scatter is made
// conflict-free and deterministic by tiling the CELLS into compact clusters:
//
//  * every cell belongs to exactly one cluster (no pair is integrated twice);
//  * a cluster touches at most W distinct dofs ("slots"); one CTA owns (row tile x cluster)
//    and accumulates the cluster's contributions to those W columns in shared memory;
//  * clusters that share a dof get different colours; colours are launched one after the
//    other, so two CTAs never update the same matrix entry concurrently;
//  * the first cluster (in launch order) that touches a dof STOREs its column, later ones
//    ADD -- no memset of the N x N matrices, no atomics, bitwise reproducible;
//  * matrix columns are stored in "first-writer" order, so the columns a cluster stores are
//    contiguous in memory (coalesced row segments).  colperm/colpos translate.
#include <algorithm>
#include <cstdio>
#include <queue>

#include "internal.h"

int wbem_build_plan(uint32_t N, uint32_t C, const uint32_t *cell_dofs, uint32_t W_max,
                    uint32_t max_cells, AssemblyPlan *pl)
{
  *pl = AssemblyPlan();
  pl->W = W_max;
  if (W_max < 4 || W_max > 255) return -1;
  // node -> cells adjacency
  std::vector<uint32_t> nptr(N + 1, 0), nadj(4 * (size_t)C);
  for (size_t k = 0; k < 4 * (size_t)C; ++k)
    {
      if (cell_dofs[k] >= N) return -2;
      nptr[cell_dofs[k] + 1]++;
    }
  for (uint32_t i = 0; i < N; ++i) nptr[i + 1] += nptr[i];
  {
    std::vector<uint32_t> fill(nptr.begin(), nptr.end() - 1);
    for (uint32_t c = 0; c < C; ++c)
      for (int j = 0; j < 4; ++j) nadj[fill[cell_dofs[4 * c + j]]++] = c;
  }

  const uint32_t NONE = 0xffffffffu;
  std::vector<uint32_t> cell_cluster(C, NONE);
  std::vector<uint32_t> node_mark(N, NONE);  // cluster id whose node set holds the node
  std::vector<uint32_t> cand_mark(C, NONE);  // cluster id for which cell is a listed candidate
  std::vector<uint32_t> seed_queue;          // cells adjacent to finished clusters (seed order)
  seed_queue.reserve(C);
  size_t seed_head = 0;
  uint32_t next_unassigned = 0;

  pl->cl_cell_ptr.push_back(0);
  pl->cl_slot_ptr.push_back(0);
  pl->cell_order.reserve(C);
  std::vector<uint32_t> cl_nodes_all; // concatenated slot -> dof
  std::vector<uint32_t> cand;

  while (pl->cell_order.size() < C)
    {
      // seed: first queued neighbour of earlier clusters, else lowest unassigned cell
      uint32_t seed = NONE;
      while (seed_head < seed_queue.size())
        {
          uint32_t c = seed_queue[seed_head++];
          if (cell_cluster[c] == NONE) { seed = c; break; }
        }
      if (seed == NONE)
        {
          while (cell_cluster[next_unassigned] != NONE) ++next_unassigned;
          seed = next_unassigned;
        }
      const uint32_t cid = pl->n_clusters++;
      uint32_t n_nodes = 0, n_cells = 0;
      cand.clear();
      const size_t node_base = cl_nodes_all.size();

      auto add_cell = [&](uint32_t c) {
        cell_cluster[c] = cid;
        pl->cell_order.push_back(c);
        ++n_cells;
        for (int j = 0; j < 4; ++j)
          {
            uint32_t d = cell_dofs[4 * c + j];
            if (node_mark[d] != cid)
              {
                node_mark[d] = cid;
                cl_nodes_all.push_back(d);
                ++n_nodes;
                for (uint32_t a = nptr[d]; a < nptr[d + 1]; ++a)
                  {
                    uint32_t nb = nadj[a];
                    if (cell_cluster[nb] == NONE && cand_mark[nb] != cid)
                      {
                        cand_mark[nb] = cid;
                        cand.push_back(nb);
                      }
                  }
              }
          }
      };
      add_cell(seed);
      while (n_cells < max_cells)
        {
          // candidate adding the fewest new dofs; ties: most shared dofs are implied (4-new),
          // then earliest listed (keeps growth compact around the seed)
          int best = -1;
          uint32_t best_new = 5;
          for (size_t k = 0; k < cand.size(); ++k)
            {
              uint32_t c = cand[k];
              if (cell_cluster[c] != NONE) continue;
              uint32_t nn = 0;
              for (int j = 0; j < 4; ++j)
                {
                  uint32_t d = cell_dofs[4 * c + j];
                  bool dup = false;
                  for (int jj = 0; jj < j; ++jj) dup |= (cell_dofs[4 * c + jj] == d);
                  if (!dup && node_mark[d] != cid) ++nn;
                }
              if (nn < best_new)
                {
                  best_new = nn;
                  best = (int)k;
                  if (nn == 0) break;
                }
            }
          if (best < 0 || n_nodes + best_new > W_max) break;
          uint32_t c = cand[best];
          cand[best] = cand.back();
          cand.pop_back();
          add_cell(c);
        }
      // leftover candidates seed later clusters
      for (uint32_t c : cand)
        if (cell_cluster[c] == NONE) seed_queue.push_back(c);
      pl->cl_cell_ptr.push_back((uint32_t)pl->cell_order.size());
      pl->cl_slot_ptr.push_back((uint32_t)cl_nodes_all.size());
      (void)node_base;
    }

  const uint32_t ncl = pl->n_clusters;
  pl->cell_pos.assign(C, 0);
  for (uint32_t p = 0; p < C; ++p) pl->cell_pos[pl->cell_order[p]] = p;

  // colouring of the cluster conflict graph (clusters sharing a dof)
  std::vector<uint32_t> color(ncl, NONE);
  // dof -> clusters touching it (ascending cluster id)
  std::vector<uint32_t> dptr(N + 1, 0);
  for (uint32_t d : cl_nodes_all) dptr[d + 1]++;
  for (uint32_t i = 0; i < N; ++i) dptr[i + 1] += dptr[i];
  std::vector<uint32_t> dcl(cl_nodes_all.size());
  {
    std::vector<uint32_t> fill(dptr.begin(), dptr.end() - 1);
    for (uint32_t k = 0; k < ncl; ++k)
      for (uint32_t s = pl->cl_slot_ptr[k]; s < pl->cl_slot_ptr[k + 1]; ++s)
        dcl[fill[cl_nodes_all[s]]++] = k;
  }
  {
    std::vector<uint32_t> used; // colour -> last cluster that saw it taken
    uint32_t ncolors = 0;
    for (uint32_t k = 0; k < ncl; ++k)
      {
        if (used.size() < ncolors + 1) used.resize(ncolors + 1, NONE);
        for (uint32_t s = pl->cl_slot_ptr[k]; s < pl->cl_slot_ptr[k + 1]; ++s)
          {
            uint32_t d = cl_nodes_all[s];
            for (uint32_t a = dptr[d]; a < dptr[d + 1]; ++a)
              {
                uint32_t o = dcl[a];
                if (o != k && color[o] != NONE) used[color[o]] = k;
              }
          }
        uint32_t c = 0;
        while (c < ncolors && used[c] == k) ++c;
        if (c == ncolors)
          {
            ++ncolors;
            used.resize(ncolors + 1, NONE);
          }
        color[k] = c;
      }
    pl->n_colors = ncolors;
  }
  // launch order: by colour, then cluster id
  pl->color_ptr.assign(pl->n_colors + 1, 0);
  for (uint32_t k = 0; k < ncl; ++k) pl->color_ptr[color[k] + 1]++;
  for (uint32_t c = 0; c < pl->n_colors; ++c) pl->color_ptr[c + 1] += pl->color_ptr[c];
  pl->color_clusters.assign(ncl, 0);
  {
    std::vector<uint32_t> fill(pl->color_ptr.begin(), pl->color_ptr.end() - 1);
    for (uint32_t k = 0; k < ncl; ++k) pl->color_clusters[fill[color[k]]++] = k;
  }
  // predecessors in launch order (see AssemblyPlan::pred)
  {
    std::vector<uint32_t> launch_pos(ncl);
    for (uint32_t idx = 0; idx < ncl; ++idx) launch_pos[pl->color_clusters[idx]] = idx;
    pl->pred_ptr.assign(ncl + 1, 0);
    std::vector<uint32_t> tmp;
    for (uint32_t idx = 0; idx < ncl; ++idx)
      {
        const uint32_t k = pl->color_clusters[idx];
        tmp.clear();
        for (uint32_t s = pl->cl_slot_ptr[k]; s < pl->cl_slot_ptr[k + 1]; ++s)
          {
            const uint32_t d = cl_nodes_all[s];
            for (uint32_t a = dptr[d]; a < dptr[d + 1]; ++a)
              if (dcl[a] != k && launch_pos[dcl[a]] < idx) tmp.push_back(launch_pos[dcl[a]]);
          }
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
        pl->pred.insert(pl->pred.end(), tmp.begin(), tmp.end());
        pl->pred_ptr[idx + 1] = (uint32_t)pl->pred.size();
        pl->max_pred = std::max(pl->max_pred, (uint32_t)tmp.size());
      }
  }
  // storage columns in first-writer order; later writers get the ADD flag.  Inside a cluster
  // the slots are reordered so that the STORE slots come first (in column order) and the ADD
  // slots last: the flush of one CTA then writes one contiguous row segment per row with
  // (almost) warp-uniform STORE / ADD lanes.
  // The columns a cluster stores are grouped by WHO ELSE touches the dof: interior dofs first, then
  // the dofs shared with one neighbour cluster, neighbour by neighbour, then the corners.  A later
  // cluster's ADD columns inside this segment are then one contiguous run (its common edge), and its
  // ADD slots are sorted by column: the RED.ADD.F64 of a warp fill whole 32-byte sectors instead of
  // costing a sector read-modify-write per 8-byte entry.
  pl->colpos.assign(N, NONE);
  pl->colperm.assign(N, 0);
  pl->slot_col.assign(cl_nodes_all.size(), 0);
  uint32_t next_col = 0;
  for (uint32_t idx = 0; idx < ncl; ++idx)
    {
      const uint32_t k = pl->color_clusters[idx];
      const uint32_t s0 = pl->cl_slot_ptr[k], s1 = pl->cl_slot_ptr[k + 1];
      auto mid = std::stable_partition(cl_nodes_all.begin() + s0, cl_nodes_all.begin() + s1,
                                       [&](uint32_t d) { return pl->colpos[d] == NONE; });
      auto signature_less = [&](uint32_t x, uint32_t y) { // other clusters touching the dof, lexicographic
        const uint32_t nx = dptr[x + 1] - dptr[x], ny = dptr[y + 1] - dptr[y];
        if ((nx == 1) != (ny == 1)) return nx == 1; // interior dofs (this cluster only) first
        if (nx != ny) return nx < ny;               // edges before corners
        for (uint32_t a = 0; a < nx; ++a)
          if (dcl[dptr[x] + a] != dcl[dptr[y] + a]) return dcl[dptr[x] + a] < dcl[dptr[y] + a];
        return false;
      };
      std::stable_sort(cl_nodes_all.begin() + s0, mid, signature_less);
      std::stable_sort(mid, cl_nodes_all.begin() + s1, [&](uint32_t x, uint32_t y) { return pl->colpos[x] < pl->colpos[y]; });
      for (uint32_t s = s0; s < s1; ++s)
        {
          const uint32_t d = cl_nodes_all[s];
          if (pl->colpos[d] == NONE)
            {
              pl->colpos[d] = next_col;
              pl->colperm[next_col] = d;
              pl->slot_col[s] = next_col;
              ++next_col;
            }
          else
            pl->slot_col[s] = pl->colpos[d] | WBEM_SLOT_ADD;
        }
    }
  // slot index of each cell's local dofs (after the reordering above)
  pl->cell_slots.assign(4 * (size_t)C, 0);
  {
    std::vector<uint32_t> slot_of(N, NONE);
    for (uint32_t k = 0; k < ncl; ++k)
      {
        for (uint32_t s = pl->cl_slot_ptr[k]; s < pl->cl_slot_ptr[k + 1]; ++s)
          slot_of[cl_nodes_all[s]] = s - pl->cl_slot_ptr[k];
        for (uint32_t p = pl->cl_cell_ptr[k]; p < pl->cl_cell_ptr[k + 1]; ++p)
          for (int j = 0; j < 4; ++j)
            pl->cell_slots[4 * (size_t)p + j] = (uint8_t)slot_of[cell_dofs[4 * pl->cell_order[p] + j]];
      }
  }
  pl->n_cols_written = next_col;
  for (uint32_t k = 0; k < ncl; ++k)
    pl->max_cells = std::max(pl->max_cells, pl->cl_cell_ptr[k + 1] - pl->cl_cell_ptr[k]);
  // dofs that belong to no cell: columns stay zero (zeroed by the caller), placed last
  for (uint32_t d = 0; d < N; ++d)
    if (pl->colpos[d] == NONE)
      {
        pl->colpos[d] = next_col;
        pl->colperm[next_col] = d;
        ++next_col;
      }
  return 0;
}

// Host-only self check of the plan (no GPU needed; used by tests/test_plan.py).
// stats[0..8] = n_clusters, n_colors, max cells/cluster, max slots/cluster, total slots,
//               slots flagged ADD, columns written, 32-byte sectors under the ADD slots,
//               longest predecessor list
// Returns 0 when every invariant holds, a positive code naming the first violated one.
extern "C" int wbem_plan_check(uint32_t N, uint32_t C, const uint32_t *cell_dofs, uint32_t W_max,
                               uint32_t max_cells, double *stats)
{
  AssemblyPlan pl;
  const int rc = wbem_build_plan(N, C, cell_dofs, W_max, max_cells, &pl);
  if (rc) return 100 - rc;
  const uint32_t ncl = pl.n_clusters;
  // 1. every cell appears exactly once in the processing order
  std::vector<int> seen(C, 0);
  for (uint32_t p = 0; p < C; ++p)
    {
      if (pl.cell_order[p] >= C) return 1;
      if (seen[pl.cell_order[p]]++) return 1;
      if (pl.cell_pos[pl.cell_order[p]] != p) return 1;
    }
  // 2. cluster sizes, slot tables point back at the cells' dofs
  uint32_t max_c = 0, max_s = 0, n_add = 0;
  for (uint32_t k = 0; k < ncl; ++k)
    {
      const uint32_t nc = pl.cl_cell_ptr[k + 1] - pl.cl_cell_ptr[k];
      const uint32_t nsl = pl.cl_slot_ptr[k + 1] - pl.cl_slot_ptr[k];
      if (nc == 0 || nc > max_cells || nsl > W_max) return 2;
      max_c = std::max(max_c, nc);
      max_s = std::max(max_s, nsl);
      for (uint32_t p = pl.cl_cell_ptr[k]; p < pl.cl_cell_ptr[k + 1]; ++p)
        for (int j = 0; j < 4; ++j)
          {
            const uint32_t s = pl.cell_slots[4 * (size_t)p + j];
            if (s >= nsl) return 3;
            const uint32_t col = pl.slot_col[pl.cl_slot_ptr[k] + s] & WBEM_SLOT_COL_MASK;
            if (col >= N || pl.colperm[col] != cell_dofs[4 * (size_t)pl.cell_order[p] + j]) return 3;
          }
    }
  // 3. colperm / colpos are inverse permutations
  for (uint32_t d = 0; d < N; ++d)
    if (pl.colpos[d] >= N || pl.colperm[pl.colpos[d]] != d) return 4;
  // 4. launch order: within a colour no two clusters share a column; the first write of a
  //    column is a STORE, every later one an ADD
  std::vector<uint32_t> last_color(N, 0xffffffffu);
  std::vector<uint8_t> written(N, 0);
  for (uint32_t c = 0; c < pl.n_colors; ++c)
    for (uint32_t idx = pl.color_ptr[c]; idx < pl.color_ptr[c + 1]; ++idx)
      {
        const uint32_t k = pl.color_clusters[idx];
        for (uint32_t s = pl.cl_slot_ptr[k]; s < pl.cl_slot_ptr[k + 1]; ++s)
          {
            const uint32_t col = pl.slot_col[s] & WBEM_SLOT_COL_MASK;
            const bool add = pl.slot_col[s] >> 31;
            if (last_color[col] == c) return 5;
            last_color[col] = c;
            if (add != (written[col] != 0)) return 6;
            written[col] = 1;
            n_add += add;
          }
      }
  // 5. predecessor lists: exactly the earlier (lower-colour) clusters sharing a column
  {
    std::vector<std::vector<uint32_t>> col_users(N);
    for (uint32_t idx = 0; idx < ncl; ++idx)
      {
        const uint32_t k = pl.color_clusters[idx];
        std::vector<uint32_t> want;
        for (uint32_t s = pl.cl_slot_ptr[k]; s < pl.cl_slot_ptr[k + 1]; ++s)
          {
            const uint32_t col = pl.slot_col[s] & WBEM_SLOT_COL_MASK;
            want.insert(want.end(), col_users[col].begin(), col_users[col].end());
            col_users[col].push_back(idx);
          }
        std::sort(want.begin(), want.end());
        want.erase(std::unique(want.begin(), want.end()), want.end());
        if (pl.pred_ptr[idx + 1] - pl.pred_ptr[idx] != want.size()) return 8;
        if (!std::equal(want.begin(), want.end(), pl.pred.begin() + pl.pred_ptr[idx])) return 8;
        if (want.size() > pl.max_pred) return 8;
      }
  }
  uint32_t nw = 0;
  for (uint32_t col = 0; col < N; ++col)
    {
      if (written[col] && col >= pl.n_cols_written) return 7;
      nw += written[col];
    }
  if (nw != pl.n_cols_written) return 7;
  if (stats)
    {
      stats[0] = ncl;
      stats[1] = pl.n_colors;
      stats[2] = max_c;
      stats[3] = max_s;
      stats[4] = (double)pl.slot_col.size();
      stats[5] = n_add;
      stats[6] = pl.n_cols_written;
      // distinct 32-byte sectors (4 columns) the ADD slots of a cluster touch per row, summed over clusters
      size_t sectors = 0;
      for (uint32_t k = 0; k < ncl; ++k)
        {
          uint32_t last = 0xffffffffu;
          for (uint32_t s = pl.cl_slot_ptr[k]; s < pl.cl_slot_ptr[k + 1]; ++s)
            if (pl.slot_col[s] >> 31)
              {
                const uint32_t sec = (pl.slot_col[s] & WBEM_SLOT_COL_MASK) / 4;
                if (sec != last) ++sectors; // ADD slots are sorted by column
                last = sec;
              }
        }
      stats[7] = (double)sectors;
      stats[8] = pl.max_pred;
    }
  return 0;
}
