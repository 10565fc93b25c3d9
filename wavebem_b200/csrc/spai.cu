// spai.cu -- "local inverse" sparse approximate inverse preconditioner (precond_kind = 1).
//
// The reference preconditions GMRES with a sparse LU of the |i-j| < 50 band of the merged
// operator (source/bem_problem.cc:1107-1149); its quality depends on how well the dof
// NUMBERING happens to follow the geometry, and iteration counts grow with N (83 at 20 k
// nodes, restart stagnation beyond 80 k).  A dense BEM operator resident in HBM allows a
// better choice that needs no factorisation chain at all:
//
//   for every dof i take S_i = the K mesh-nearest dofs (breadth-first rings over the node
//   graph -- dofs sharing a cell or a double-node set -- cut to the K geometrically nearest),
//   solve the K x K system   A[S_i,S_i]^T m = e_i   and use m as row i of M ~ A^-1.
//
// M A has row i exact on S_i (a factorised-free sparse approximate inverse in the sense of
// Benzi & Tuma's survey).  Set-up = N independent K x K solves (one warp each, partial
// pivoting, shared memory); application = one sparse mat-vec with K entries per row -- both
// trivially parallel, no sequential sweep, no level-by-level barriers.  On the tank + Wigley
// meshes GMRES needs ~20 iterations (band: 55-175) nearly independent of N.
//
// Row-sharded runs: rank p holds matrix rows of its block only, but A[S_i,S_i] needs rows of
// the neighbours.  Every rank therefore extracts, for its rows r, the "near-field" entries
// A[r,c], c in E_r = union of the S_i that contain r (ELL rows of width EW ~ 110-180), the
// blocks are all-gathered (same volume as the band rows of the reference preconditioner),
// each rank solves the systems of its own rows and the K-entry rows of M are all-gathered.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <thread>
#include <vector>

#include "internal.h"

#define SPAI_K 32
#define SPAI_NONE 0xffffffffu

struct SpaiState
{
  uint32_t N = 0, npad = 0, EW = 0;
  uint32_t *d_nbr = nullptr;  // [npad][K] neighbour dofs, sorted ascending, padded with NONE
  uint32_t *d_ecol = nullptr; // [npad][EW] near-field columns, sorted ascending, padded with NONE
  double *d_nf = nullptr;     // [npad][EW] near-field values of the merged constrained operator
  double *d_val = nullptr;    // [npad][K] rows of M
  uint16_t *d_pos = nullptr;  // [nloc][K][K] position of S_b in the near-field row of S_a
  int *d_info = nullptr;      // number of singular local systems (those rows fall back to 1/a_ii)
  std::vector<uint32_t> h_nbr;
  bool pattern_ready = false;
};

// ---------------------------------------------------------------------------------------
// host: sparsity pattern, once per topology
// ---------------------------------------------------------------------------------------
// nbr[i][0..K): the K nearest dofs of i among its breadth-first rings (i itself included),
// sorted by dof id, padded with NONE when the mesh has fewer than K dofs in reach.
static void spai_build_neighbours(uint32_t N, uint32_t C, const uint32_t *cell_dofs, const uint32_t *dn_ptr,
                                  const uint32_t *dn_idx, const double *xyz, uint32_t K,
                                  std::vector<uint32_t> &nbr)
{
  std::vector<uint32_t> nptr(N + 1, 0), nadj(4 * (size_t)C);
  for (size_t k = 0; k < 4 * (size_t)C; ++k) nptr[cell_dofs[k] + 1]++;
  for (uint32_t i = 0; i < N; ++i) nptr[i + 1] += nptr[i];
  {
    std::vector<uint32_t> fill(nptr.begin(), nptr.end() - 1);
    for (uint32_t c = 0; c < C; ++c)
      for (int j = 0; j < 4; ++j) nadj[fill[cell_dofs[4 * (size_t)c + j]]++] = c;
  }
  nbr.assign((size_t)N * K, SPAI_NONE);
  // rows are independent: host threads over row ranges, each with its own scratch (reinit() runs on every remesh)
  auto rows = [&](uint32_t i0, uint32_t i1) {
    std::vector<uint32_t> mark(N, SPAI_NONE), order, frontier, next;
    std::vector<std::pair<double, uint32_t>> cand;
    for (uint32_t i = i0; i < i1; ++i)
      {
        order.clear();
        frontier.clear();
        mark[i] = i;
        order.push_back(i);
        frontier.push_back(i);
        // collect ~3K candidates (whole rings): on stretched cells the K nearest dofs are not the
        // first rings
        while (order.size() < 3 * (size_t)K && !frontier.empty())
          {
            next.clear();
            for (uint32_t u : frontier)
              {
                for (uint32_t a = nptr[u]; a < nptr[u + 1]; ++a)
                  for (int j = 0; j < 4; ++j)
                    {
                      const uint32_t v = cell_dofs[4 * (size_t)nadj[a] + j];
                      if (mark[v] != i)
                        {
                          mark[v] = i;
                          next.push_back(v);
                        }
                    }
                for (uint32_t a = dn_ptr[u]; a < dn_ptr[u + 1]; ++a)
                  {
                    const uint32_t v = dn_idx[a];
                    if (mark[v] != i)
                      {
                        mark[v] = i;
                        next.push_back(v);
                      }
                  }
              }
            order.insert(order.end(), next.begin(), next.end());
            frontier.swap(next);
          }
        cand.clear();
        for (uint32_t v : order)
          {
            const double dx = xyz[3 * (size_t)v] - xyz[3 * (size_t)i], dy = xyz[3 * (size_t)v + 1] - xyz[3 * (size_t)i + 1],
                         dz = xyz[3 * (size_t)v + 2] - xyz[3 * (size_t)i + 2];
            cand.emplace_back(v == i ? -1.0 : dx * dx + dy * dy + dz * dz, v);
          }
        const size_t keep = std::min<size_t>(K, cand.size());
        std::partial_sort(cand.begin(), cand.begin() + keep, cand.end());
        uint32_t *row = nbr.data() + (size_t)i * K;
        for (size_t k = 0; k < keep; ++k) row[k] = cand[k].second;
        std::sort(row, row + keep);
      }
  };
  unsigned nt = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  if (N < 4096) nt = 1;
  if (nt == 1)
    rows(0, N);
  else
    {
      std::vector<std::thread> pool;
      const uint32_t per = (N + nt - 1) / nt;
      for (unsigned t = 0; t < nt; ++t)
        {
          const uint32_t i0 = std::min(N, t * per), i1 = std::min(N, (t + 1) * per);
          if (i0 < i1) pool.emplace_back(rows, i0, i1);
        }
      for (auto &th : pool) th.join();
    }
}

// ecol[r] = sorted union of the S_i that contain r; returns the largest row length
static uint32_t spai_build_nearfield_pattern(uint32_t N, uint32_t K, const std::vector<uint32_t> &nbr,
                                             std::vector<uint32_t> &eptr, std::vector<uint32_t> &ecol)
{
  std::vector<uint32_t> rptr(N + 1, 0);
  for (size_t k = 0; k < nbr.size(); ++k)
    if (nbr[k] != SPAI_NONE) rptr[nbr[k] + 1]++;
  for (uint32_t i = 0; i < N; ++i) rptr[i + 1] += rptr[i];
  std::vector<uint32_t> rev(rptr[N]);
  {
    std::vector<uint32_t> fill(rptr.begin(), rptr.end() - 1);
    for (uint32_t i = 0; i < N; ++i)
      for (uint32_t a = 0; a < K; ++a)
        {
          const uint32_t r = nbr[(size_t)i * K + a];
          if (r != SPAI_NONE) rev[fill[r]++] = i;
        }
  }
  // rows are independent: host threads over row ranges, concatenated in row order afterwards
  unsigned nt = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  if (N < 4096) nt = 1;
  const uint32_t per = (N + nt - 1) / nt;
  std::vector<std::vector<uint32_t>> cols(nt), lens(nt);
  std::vector<uint32_t> widest_t(nt, 0);
  auto rows = [&](unsigned t) {
    const uint32_t r0 = std::min(N, t * per), r1 = std::min(N, (t + 1) * per);
    std::vector<uint32_t> mark(N, SPAI_NONE);
    std::vector<uint32_t> &ec = cols[t];
    ec.reserve((size_t)(r1 - r0) * 4 * K);
    lens[t].reserve(r1 - r0);
    for (uint32_t r = r0; r < r1; ++r)
      {
        const size_t b = ec.size();
        for (uint32_t a = rptr[r]; a < rptr[r + 1]; ++a)
          {
            const uint32_t *row = nbr.data() + (size_t)rev[a] * K;
            for (uint32_t k = 0; k < K; ++k)
              {
                const uint32_t c = row[k];
                if (c != SPAI_NONE && mark[c] != r)
                  {
                    mark[c] = r;
                    ec.push_back(c);
                  }
              }
          }
        std::sort(ec.begin() + b, ec.end());
        lens[t].push_back((uint32_t)(ec.size() - b));
        widest_t[t] = std::max(widest_t[t], (uint32_t)(ec.size() - b));
      }
  };
  if (nt == 1)
    rows(0);
  else
    {
      std::vector<std::thread> pool;
      for (unsigned t = 0; t < nt; ++t) pool.emplace_back(rows, t);
      for (auto &th : pool) th.join();
    }
  eptr.assign(N + 1, 0);
  ecol.clear();
  uint32_t widest = 0, r = 0;
  size_t total = 0;
  for (unsigned t = 0; t < nt; ++t) total += cols[t].size();
  ecol.reserve(total);
  for (unsigned t = 0; t < nt; ++t)
    {
      ecol.insert(ecol.end(), cols[t].begin(), cols[t].end());
      for (uint32_t l : lens[t])
        {
          eptr[r + 1] = eptr[r] + l;
          ++r;
        }
      widest = std::max(widest, widest_t[t]);
    }
  return widest;
}

// ---------------------------------------------------------------------------------------
// device kernels
// ---------------------------------------------------------------------------------------
// near-field entries of the merged, constrained operator for the local rows:
//   free row r:        A[r,c] = other(c) ? N[r,c] + alpha_r [r == c] : -D[r,c]     (bem_problem.cc:1126-1145)
//   constrained row r: A[r,c] = [r == c] - sum_k c_rk [col_k == c]                 (constrained_matrix.h:73-86)
__global__ void __launch_bounds__(256)
  k_spai_nearfield(uint32_t nloc, uint32_t row0, uint32_t ld, uint32_t EW, const uint32_t *__restrict__ ecol,
                   const double *__restrict__ Nm, const double *__restrict__ Dm,
                   const double *__restrict__ alpha, const double *__restrict__ surf,
                   const int32_t *__restrict__ line_of, const uint32_t *__restrict__ con_ptr,
                   const uint32_t *__restrict__ con_col, const double *__restrict__ con_val,
                   const uint32_t *__restrict__ colpos, double *__restrict__ nf)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t r = (uint32_t)(t / EW);
  if (r >= nloc) return;
  const uint32_t g = row0 + r;
  const uint32_t c = ecol[(size_t)g * EW + (t - (size_t)r * EW)];
  double v = 0.0;
  if (c != SPAI_NONE)
    {
      const int32_t line = line_of ? line_of[g] : -1;
      if (line >= 0)
        {
          v = (c == g) ? 1.0 : 0.0;
          for (uint32_t k = con_ptr[line]; k < con_ptr[line + 1]; ++k)
            if (con_col[k] == c) v -= con_val[k];
        }
      else if (surf[c] == 0)
        {
          v = Nm[(size_t)r * ld + colpos[c]];
          if (c == g) v += alpha[g];
        }
      else
        v = -Dm[(size_t)r * ld + colpos[c]];
    }
  nf[(size_t)g * EW + (t - (size_t)r * EW)] = v;
}

#define SPAI_LD (SPAI_K + 1)
#define SPAI_WARPS 4
#define SPAI_NOPOS 0xffffu

// Once per sparsity pattern, one warp per local row i: pos[i][a][b] = position of column S_b in
// the near-field row of S_a (binary search in the sorted ELL row), so that the per-solve gather
// of A[S_a, S_b] is two plain loads.
__global__ void __launch_bounds__(32 * SPAI_WARPS)
  k_spai_positions(uint32_t nloc, uint32_t row0, uint32_t EW, const uint32_t *__restrict__ nbr,
                   const uint32_t *__restrict__ ecol, uint16_t *__restrict__ pos)
{
  extern __shared__ __align__(16) unsigned char spai_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t *ec = reinterpret_cast<uint32_t *>(spai_smem) + (size_t)warp * EW;
  const uint32_t il = blockIdx.x * SPAI_WARPS + warp;
  if (il >= nloc) return;
  const uint32_t g = row0 + il;
  const uint32_t mine = nbr[(size_t)g * SPAI_K + lane]; // S_lane
  for (int a = 0; a < SPAI_K; ++a)
    {
      const uint32_t r = __shfl_sync(0xffffffffu, mine, a);
      uint16_t found = SPAI_NOPOS;
      if (r != SPAI_NONE)
        {
          for (uint32_t p = lane; p < EW; p += 32) ec[p] = ecol[(size_t)r * EW + p];
          __syncwarp();
          if (mine != SPAI_NONE)
            {
              uint32_t lo = 0, hi = EW; // first position with ec[pos] >= mine (NONE pads sort last)
              while (lo < hi)
                {
                  const uint32_t mid = (lo + hi) >> 1;
                  if (ec[mid] < mine)
                    lo = mid + 1;
                  else
                    hi = mid;
                }
              if (lo < EW && ec[lo] == mine) found = (uint16_t)lo;
            }
          __syncwarp();
        }
      pos[((size_t)il * SPAI_K + a) * SPAI_K + lane] = found;
    }
}

// One warp per local row i: gather T[b][a] = A[S_a, S_b] from the near-field rows, solve
// T m = e_i by Gaussian elimination with partial pivoting, store m.
__global__ void __launch_bounds__(32 * SPAI_WARPS)
  k_spai_solve(uint32_t nloc, uint32_t row0, uint32_t EW, const uint32_t *__restrict__ nbr,
               const uint32_t *__restrict__ ecol, const uint16_t *__restrict__ pos,
               const double *__restrict__ nf, double *__restrict__ val, int *__restrict__ info)
{
  extern __shared__ __align__(16) unsigned char spai_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *T = reinterpret_cast<double *>(spai_smem) + (size_t)warp * SPAI_K * SPAI_LD;
  const uint32_t il = blockIdx.x * SPAI_WARPS + warp;
  if (il >= nloc) return;
  const uint32_t g = row0 + il;
  const uint32_t mine = nbr[(size_t)g * SPAI_K + lane]; // S_lane
  const uint16_t *prow = pos + (size_t)il * SPAI_K * SPAI_K + lane;
#pragma unroll 8
  for (int a = 0; a < SPAI_K; ++a)
    {
      const uint32_t r = __shfl_sync(0xffffffffu, mine, a);
      const uint16_t p = prow[a * SPAI_K];
      double v = 0.0;
      if (r == SPAI_NONE)
        v = (a == lane) ? 1.0 : 0.0; // padding slot: identity row / column
      else if (p != SPAI_NOPOS)
        v = nf[(size_t)r * EW + p];
      T[lane * SPAI_LD + a] = v;
    }
  double e = (mine == g) ? 1.0 : 0.0;
  __syncwarp();
  bool singular = false;
  for (int k = 0; k < SPAI_K; ++k)
    {
      // pivot: largest |T[j][k]|, j >= k (ties: smallest j)
      double best = (lane >= k) ? fabs(T[lane * SPAI_LD + k]) : -1.0;
      int p = lane;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
        {
          const double ob = __shfl_xor_sync(0xffffffffu, best, off);
          const int op = __shfl_xor_sync(0xffffffffu, p, off);
          if (ob > best || (ob == best && op < p))
            {
              best = ob;
              p = op;
            }
        }
      if (!(best > 0.0))
        {
          singular = true;
          break;
        }
      if (p != k)
        { // swap rows k and p (lane = column), and the right-hand side entries
          const double tk = T[k * SPAI_LD + lane], tp = T[p * SPAI_LD + lane];
          T[k * SPAI_LD + lane] = tp;
          T[p * SPAI_LD + lane] = tk;
          const double ek = __shfl_sync(0xffffffffu, e, k), ep = __shfl_sync(0xffffffffu, e, p);
          if (lane == k) e = ep;
          if (lane == p) e = ek;
        }
      __syncwarp();
      const double ek = __shfl_sync(0xffffffffu, e, k);
      if (lane > k)
        {
          const double f = T[lane * SPAI_LD + k] / T[k * SPAI_LD + k];
          for (int c = k + 1; c < SPAI_K; ++c) T[lane * SPAI_LD + c] = fma(-f, T[k * SPAI_LD + c], T[lane * SPAI_LD + c]);
          e = fma(-f, ek, e);
        }
      __syncwarp();
    }
  double m = 0.0;
  if (!singular)
    {
      for (int k = SPAI_K - 1; k >= 0; --k)
        {
          const double xk = __shfl_sync(0xffffffffu, e, k) / T[k * SPAI_LD + k];
          if (lane == k) m = xk;
          if (lane < k) e = fma(-T[lane * SPAI_LD + k], xk, e);
        }
      // columns were never permuted: m_lane multiplies dof S_lane
    }
  else
    { // fall back to the diagonal (Jacobi) for this row
      if (lane == 0) atomicAdd(info, 1);
      double d = 0.0; // a_gg: exactly one lane finds column g in the near-field row of g
      for (uint32_t pp = lane; pp < EW; pp += 32)
        if (ecol[(size_t)g * EW + pp] == g) d = nf[(size_t)g * EW + pp];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
      m = (mine == g) ? (d != 0.0 ? 1.0 / d : 1.0) : 0.0;
    }
  if (mine == SPAI_NONE) m = 0.0;
  val[(size_t)g * SPAI_K + lane] = m;
}

// z_i = sum_a M[i][a] v[S_i[a]] : 8 lanes per row, 4 entries each, fixed summation order
__global__ void __launch_bounds__(256)
  k_spai_apply(uint32_t N, const uint32_t *__restrict__ nbr, const double *__restrict__ val,
               const double *__restrict__ v, double *__restrict__ z)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t >> 3, part = t & 7;
  double s = 0.0;
  if (i < N)
    {
      const uint4 c = reinterpret_cast<const uint4 *>(nbr + (size_t)i * SPAI_K)[part];
      const double2 m0 = reinterpret_cast<const double2 *>(val + (size_t)i * SPAI_K)[2 * part];
      const double2 m1 = reinterpret_cast<const double2 *>(val + (size_t)i * SPAI_K)[2 * part + 1];
      const double a0 = c.x != SPAI_NONE ? m0.x * v[c.x] : 0.0;
      const double a1 = c.y != SPAI_NONE ? m0.y * v[c.y] : 0.0;
      const double a2 = c.z != SPAI_NONE ? m1.x * v[c.z] : 0.0;
      const double a3 = c.w != SPAI_NONE ? m1.y * v[c.w] : 0.0;
      s = (a0 + a1) + (a2 + a3);
    }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (i < N && part == 0) z[i] = s;
}

// The same product with v = ConstrainedOperator::vmult's result taken straight from the gathered
// mat-vec rows: v_j = y_j, or src_j - sum_k c_jk src_k on a constrained row
// (include/constrained_matrix.h:73-86) -- the operator's epilogue and the preconditioner in ONE kernel.
__device__ __forceinline__ double epi_value(const EpilogueArgs &e, uint32_t j)
{
  if (e.line_of)
    {
      const int l = e.line_of[j];
      if (l >= 0)
        {
          double v = e.src[j];
          for (uint32_t k = e.cptr[l]; k < e.cptr[l + 1]; ++k) v -= e.cval[k] * e.src[e.ccol[k]];
          return v;
        }
    }
  return e.y[j];
}

__global__ void __launch_bounds__(256)
  k_spai_apply_fused(uint32_t N, const uint32_t *__restrict__ nbr, const double *__restrict__ val, const EpilogueArgs e,
                     double *__restrict__ z)
{
  if (e.flags)
    { // fused gather: wait (bounded) until every rank has delivered its rows of this epoch
      if ((int)threadIdx.x < e.n_peers)
        {
          unsigned long long v, t0 = 0, t1;
          unsigned int spins = 0;
          for (;;)
            {
              asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(e.flags + threadIdx.x) : "memory");
              if (v >= e.epoch) break;
              if ((++spins & 0x3ffu) == 0)
                {
                  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                  if (!t0) t0 = t1;
                  if (t1 - t0 > 20000000000ull)
                    {
                      atomicExch(e.timeout_flag, 1u);
                      break;
                    }
                }
            }
        }
      __syncthreads();
    }
  if (e.stop && *e.stop != 0) return; // enqueued ahead of a finished solve (uniform over the grid)
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t >> 3, part = t & 7;
  double s = 0.0;
  if (i < N)
    {
      const uint4 c = reinterpret_cast<const uint4 *>(nbr + (size_t)i * SPAI_K)[part];
      const double2 m0 = reinterpret_cast<const double2 *>(val + (size_t)i * SPAI_K)[2 * part];
      const double2 m1 = reinterpret_cast<const double2 *>(val + (size_t)i * SPAI_K)[2 * part + 1];
      const double a0 = c.x != SPAI_NONE ? m0.x * epi_value(e, c.x) : 0.0;
      const double a1 = c.y != SPAI_NONE ? m0.y * epi_value(e, c.y) : 0.0;
      const double a2 = c.z != SPAI_NONE ? m1.x * epi_value(e, c.z) : 0.0;
      const double a3 = c.w != SPAI_NONE ? m1.y * epi_value(e, c.w) : 0.0;
      s = (a0 + a1) + (a2 + a3);
    }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (i < N && part == 0) z[i] = s;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static SpaiState *spai_state(wbem_ctx *ctx)
{
  if (!ctx->spai) ctx->spai = new SpaiState();
  return reinterpret_cast<SpaiState *>(ctx->spai);
}

void wbem_spai_free(wbem_ctx *ctx)
{
  SpaiState *s = reinterpret_cast<SpaiState *>(ctx->spai);
  if (!s) return;
  cudaFree(s->d_nbr);
  cudaFree(s->d_ecol);
  cudaFree(s->d_nf);
  cudaFree(s->d_val);
  cudaFree(s->d_pos);
  cudaFree(s->d_info);
  delete s;
  ctx->spai = nullptr;
}

static int spai_build_pattern(wbem_ctx *ctx)
{
  SpaiState *s = spai_state(ctx);
  if (s->pattern_ready) return 0;
  if (!ctx->have_geometry) WBEM_FAIL(ctx, -3, "SPAI preconditioner needs the geometry (wbem_set_geometry)");
  const uint32_t N = ctx->N, K = SPAI_K;
  std::vector<double> xyz(3 * (size_t)N);
  CUDA_OK(ctx, cudaMemcpyAsync(xyz.data(), ctx->d_xyz, sizeof(double) * xyz.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  spai_build_neighbours(N, ctx->C, ctx->h_cell_dofs.data(), ctx->h_dn_ptr.data(), ctx->h_dn_idx.data(), xyz.data(), K,
                        s->h_nbr);
  std::vector<uint32_t> eptr, ecol;
  const uint32_t widest = spai_build_nearfield_pattern(N, K, s->h_nbr, eptr, ecol);
  s->N = N;
  s->npad = ctx->chunk * (uint32_t)ctx->p.world_size;
  s->EW = (widest + 7u) / 8u * 8u;
  std::vector<uint32_t> ell((size_t)s->npad * s->EW, SPAI_NONE), nbr_pad((size_t)s->npad * K, SPAI_NONE);
  for (uint32_t r = 0; r < N; ++r)
    std::copy(ecol.begin() + eptr[r], ecol.begin() + eptr[r + 1], ell.begin() + (size_t)r * s->EW);
  std::copy(s->h_nbr.begin(), s->h_nbr.end(), nbr_pad.begin());
  CUDA_OK(ctx, cudaMalloc((void **)&s->d_nbr, sizeof(uint32_t) * nbr_pad.size()));
  CUDA_OK(ctx, cudaMalloc((void **)&s->d_ecol, sizeof(uint32_t) * ell.size()));
  CUDA_OK(ctx, cudaMalloc((void **)&s->d_nf, sizeof(double) * ell.size()));
  CUDA_OK(ctx, cudaMalloc((void **)&s->d_val, sizeof(double) * nbr_pad.size()));
  CUDA_OK(ctx, cudaMalloc((void **)&s->d_info, sizeof(int)));
  CUDA_OK(ctx, cudaMemcpyAsync(s->d_nbr, nbr_pad.data(), sizeof(uint32_t) * nbr_pad.size(), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(ctx, cudaMemcpyAsync(s->d_ecol, ell.data(), sizeof(uint32_t) * ell.size(), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(ctx, cudaMemsetAsync(s->d_nf, 0, sizeof(double) * ell.size(), ctx->stream));
  CUDA_OK(ctx, cudaMemsetAsync(s->d_val, 0, sizeof(double) * nbr_pad.size(), ctx->stream));
  CUDA_OK(ctx, cudaMalloc((void **)&s->d_pos, sizeof(uint16_t) * std::max<size_t>(1, (size_t)ctx->nloc * K * K)));
  if (ctx->nloc)
    {
      k_spai_positions<<<(ctx->nloc + SPAI_WARPS - 1) / SPAI_WARPS, 32 * SPAI_WARPS, sizeof(uint32_t) * SPAI_WARPS * s->EW,
                         ctx->stream>>>(ctx->nloc, ctx->row0, s->EW, s->d_nbr, s->d_ecol, s->d_pos);
      ctx->launches++;
      CUDA_OK(ctx, cudaGetLastError());
    }
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  s->pattern_ready = true;
  return 0;
}


int wbem_spai_setup(wbem_ctx *ctx)
{
  int rc = spai_build_pattern(ctx);
  if (rc) return rc;
  SpaiState *s = spai_state(ctx);
  cudaStream_t st = ctx->stream;
  CUDA_OK(ctx, cudaMemsetAsync(s->d_info, 0, sizeof(int), st));
  if (ctx->nloc)
    {
      const size_t total = (size_t)ctx->nloc * s->EW;
      k_spai_nearfield<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        ctx->nloc, ctx->row0, ctx->ld, s->EW, s->d_ecol, ctx->d_Nm, ctx->d_Dm, ctx->d_alpha, ctx->d_surf,
        ctx->n_lines ? ctx->d_con_line_of : nullptr, ctx->d_con_ptr, ctx->d_con_col, ctx->d_con_val, ctx->d_colpos,
        s->d_nf);
      ctx->launches++;
      CUDA_OK(ctx, cudaGetLastError());
    }
  rc = wbem_allgather_bytes(ctx, s->d_nf, sizeof(double) * (size_t)ctx->chunk * s->EW);
  if (rc) return rc;
  if (ctx->nloc)
    {
      const size_t smem = sizeof(double) * SPAI_WARPS * SPAI_K * SPAI_LD; // 33 KB: below the default limit
      k_spai_solve<<<(ctx->nloc + SPAI_WARPS - 1) / SPAI_WARPS, 32 * SPAI_WARPS, smem, st>>>(
        ctx->nloc, ctx->row0, s->EW, s->d_nbr, s->d_ecol, s->d_pos, s->d_nf, s->d_val, s->d_info);
      ctx->launches++;
      CUDA_OK(ctx, cudaGetLastError());
    }
  return wbem_allgather_bytes(ctx, s->d_val, sizeof(double) * (size_t)ctx->chunk * SPAI_K);
}

int wbem_spai_apply(wbem_ctx *ctx, const double *d_in, double *d_out)
{
  SpaiState *s = reinterpret_cast<SpaiState *>(ctx->spai);
  if (!s || !s->pattern_ready) WBEM_FAIL(ctx, -3, "SPAI preconditioner applied before it was assembled");
  if (d_in == d_out) WBEM_FAIL(ctx, -1, "the sparse approximate inverse cannot be applied in place");
  k_spai_apply<<<(unsigned)(((size_t)ctx->N * 8 + 255) / 256), 256, 0, ctx->stream>>>(ctx->N, s->d_nbr, s->d_val, d_in, d_out);
  ctx->launches++;
  CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

int wbem_spai_apply_fused(wbem_ctx *ctx, const EpilogueArgs &ea, double *d_out)
{
  SpaiState *s = reinterpret_cast<SpaiState *>(ctx->spai);
  if (!s || !s->pattern_ready || !ctx->precond_ready) WBEM_FAIL(ctx, -3, "SPAI preconditioner applied before it was assembled");
  k_spai_apply_fused<<<(unsigned)(((size_t)ctx->N * 8 + 255) / 256), 256, 0, ctx->stream>>>(ctx->N, s->d_nbr, s->d_val, ea,
                                                                                           d_out);
  ctx->launches++;
  CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

void wbem_spai_invalidate_pattern(wbem_ctx *ctx) { wbem_spai_free(ctx); }

extern "C" {

int wbem_get_spai(wbem_ctx *ctx, uint32_t *k_out, uint32_t *nbr, double *val, int *n_singular)
{
  if (!ctx) return -1;
  if (k_out) *k_out = SPAI_K;
  SpaiState *s = reinterpret_cast<SpaiState *>(ctx->spai);
  if (!nbr && !val && !n_singular) return 0;
  if (!s || !s->pattern_ready) WBEM_FAIL(ctx, -3, "wbem_get_spai before the preconditioner was assembled");
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (nbr) CUDA_OK(ctx, cudaMemcpy(nbr, s->d_nbr, sizeof(uint32_t) * (size_t)ctx->N * SPAI_K, cudaMemcpyDeviceToHost));
  if (val) CUDA_OK(ctx, cudaMemcpy(val, s->d_val, sizeof(double) * (size_t)ctx->N * SPAI_K, cudaMemcpyDeviceToHost));
  if (n_singular) CUDA_OK(ctx, cudaMemcpy(n_singular, s->d_info, sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

// Host-only check of the sparsity pattern builder (no GPU): nbr_out[N][32] (0xffffffff pads),
// stats[0..3] = K, widest near-field row, mean near-field row, rows with fewer than K dofs.
// Returns 0 when every row is sorted, duplicate-free, in range and holds its own dof.
int wbem_spai_pattern_check(uint32_t N, uint32_t C, const uint32_t *cell_dofs, const uint32_t *dn_ptr,
                            const uint32_t *dn_idx, const double *xyz, uint32_t *nbr_out, double *stats)
{
  std::vector<uint32_t> nbr, eptr, ecol;
  spai_build_neighbours(N, C, cell_dofs, dn_ptr, dn_idx, xyz, SPAI_K, nbr);
  const uint32_t widest = spai_build_nearfield_pattern(N, SPAI_K, nbr, eptr, ecol);
  uint32_t n_short = 0;
  for (uint32_t i = 0; i < N; ++i)
    {
      const uint32_t *row = nbr.data() + (size_t)i * SPAI_K;
      bool self = false;
      uint32_t cnt = 0;
      for (uint32_t k = 0; k < SPAI_K; ++k)
        {
          if (row[k] == SPAI_NONE)
            {
              for (uint32_t kk = k; kk < SPAI_K; ++kk)
                if (row[kk] != SPAI_NONE) return 1; // pads must be trailing
              break;
            }
          if (row[k] >= N) return 2;
          if (k && row[k] <= row[k - 1]) return 3;
          self |= row[k] == i;
          ++cnt;
        }
      if (!self) return 4;
      n_short += cnt < SPAI_K;
      // near-field row of i must contain S_i
      for (uint32_t k = 0; k < cnt; ++k)
        if (!std::binary_search(ecol.begin() + eptr[i], ecol.begin() + eptr[i + 1], row[k])) return 5;
    }
  if (nbr_out) std::copy(nbr.begin(), nbr.end(), nbr_out);
  if (stats)
    {
      stats[0] = SPAI_K;
      stats[1] = widest;
      stats[2] = N ? (double)ecol.size() / N : 0.0;
      stats[3] = n_short;
    }
  return 0;
}

} // extern "C"
