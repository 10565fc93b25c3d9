// api.cu -- the C ABI of include/wbem.h: context, flattening of the mesh data, uploads,
// and thin wrappers over the kernels.  No CPU fallback anywhere: without a usable sm_100
// device wbem_create fails and nothing else can be called.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "internal.h"

static std::string g_create_error;

#define CHECK_CTX(ctx) \
  if (!(ctx)) return -1

template <typename T>
static int dev_alloc(wbem_ctx *ctx, T **p, size_t n)
{
  if (*p)
    {
      cudaFree(*p);
      *p = nullptr;
    }
  if (n == 0) n = 1;
  CUDA_OK(ctx, cudaMalloc((void **)p, n * sizeof(T)));
  return 0;
}
template <typename T>
static int dev_upload(wbem_ctx *ctx, T **p, const std::vector<T> &v)
{
  int rc = dev_alloc(ctx, p, v.size());
  if (rc) return rc;
  if (!v.empty())
    CUDA_OK(ctx, cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
#define FREE_DEV(p)     \
  do                    \
    {                   \
      if (p) cudaFree(p); \
      p = nullptr;      \
    }                   \
  while (0)

extern "C" {

int wbem_version(void) { return 100; }

void wbem_default_params(wbem_params *p)
{
  memset(p, 0, sizeof(*p));
  p->quad_order = 4;            // prm-files/default.prm:218-219
  p->sing_order = 5;            // prm-files/default.prm:220
  p->gmres_tol = 1e-16;         // prm-files/default.prm:229
  p->gmres_max_steps = 200;     // prm-files/default.prm:228
  p->gmres_n_tmp_vectors = 100; // source/bem_problem.cc:826-827
  p->preconditioner_band = 100; // source/bem_problem.cc:68
  p->device = 0;
  p->rank = 0;
  p->world_size = 1;
  p->assemble_variant = 0;
  p->precond_on_host = 0;
  p->precond_kind = 0;
  p->auto_constraints = 0;
  p->n_gpus = 0;
  p->fused_gather_on_shared_device = 0;
  for (int &d : p->devices) d = -1;
}

const char *wbem_last_error(const wbem_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int wbem_create(const wbem_params *p, wbem_ctx **out)
{
  if (!p || !out)
    {
      g_create_error = "null argument";
      return -1;
    }
  *out = nullptr;
  // n_gpus > 1: one handle, n_gpus row blocks driven from this process (group.cpp)
  if (p->n_gpus > 1) return wbem_group_create(p, out, &g_create_error);
  return wbem_create_single(p, out, &g_create_error);
}

} // extern "C"

int wbem_create_single(const wbem_params *p, wbem_ctx **out, std::string *errp)
{
  std::string &g_create_error = *errp;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    {
      g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                       " (libwbem has no CPU fallback)";
      cudaGetLastError();
      return -2;
    }
  if (p->device < 0 || p->device >= ndev)
    {
      g_create_error = "device ordinal out of range";
      return -1;
    }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, p->device);
  if (prop.major != 10)
    {
      char b[256];
      snprintf(b, sizeof(b), "device %d is sm_%d%d; libwbem is built for sm_100a only", p->device,
               prop.major, prop.minor);
      g_create_error = b;
      return -2;
    }
  if (p->world_size < 1 || p->rank < 0 || p->rank >= p->world_size || p->gmres_n_tmp_vectors < 3 ||
      p->gmres_n_tmp_vectors > WBEM_GMRES_KMAX || p->preconditioner_band < 0 || (p->preconditioner_band & 1) ||
      p->preconditioner_band > 128 || p->precond_kind < 0 || p->precond_kind > 1)
    {
      g_create_error = "bad wbem_params (world/rank, 3 <= n_tmp_vectors <= 1024, even band <= 128, precond_kind 0/1)";
      return -1;
    }
  wbem_ctx *ctx = new wbem_ctx();
  ctx->p = *p;
  ctx->dev = p->device;
  if (wbem_build_quadrature(p->quad_order, p->sing_order, &ctx->qt))
    {
      g_create_error = "unsupported quadrature order (1..8 regular, 1..12 singular)";
      delete ctx;
      return -1;
    }
  if (cudaSetDevice(ctx->dev) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess)
    {
      g_create_error = "cudaSetDevice/cudaStreamCreate failed";
      delete ctx;
      return -2;
    }
  for (auto &ev : ctx->ev) cudaEventCreate(&ev);
  if (wbem_upload_tables(ctx) || wbem_constraints_upload_tables(ctx))
    {
      g_create_error = ctx->err;
      delete ctx;
      return -2;
    }
  *out = ctx;
  return 0;
}

void wbem_p2p_close(wbem_ctx *ctx)
{
  for (int q = 0; q < WBEM_MAX_PEERS; ++q)
    {
      if (ctx->peer_opened[q] && ctx->peer_base[q]) cudaIpcCloseMemHandle(ctx->peer_base[q]);
      ctx->peer_opened[q] = false;
      ctx->peer_base[q] = nullptr;
    }
  ctx->p2p_ready = false;
}
extern "C" {

static void free_topology(wbem_ctx *ctx)
{
  wbem_p2p_close(ctx);
  FREE_DEV(ctx->d_p2p);
  FREE_DEV(ctx->d_ymulti);
  wbem_multi_free(ctx);
  FREE_DEV(ctx->d_done_counter);
  FREE_DEV(ctx->d_gather_timeout);
  FREE_DEV(ctx->d_cell_dofs);
  FREE_DEV(ctx->d_dir);
  FREE_DEV(ctx->d_cell_order);
  FREE_DEV(ctx->d_colpos);
  FREE_DEV(ctx->d_colperm);
  FREE_DEV(ctx->d_sing_ptr);
  FREE_DEV(ctx->d_sing_cellpos);
  FREE_DEV(ctx->d_sing_idx);
  FREE_DEV(ctx->d_tile_sing);
  FREE_DEV(ctx->d_slot_col);
  FREE_DEV(ctx->d_cta_desc);
  FREE_DEV(ctx->d_cell_slots);
  FREE_DEV(ctx->d_item_meta);
  FREE_DEV(ctx->d_item_desc);
  FREE_DEV(ctx->d_asm_sync);
  FREE_DEV(ctx->d_xyz);
  FREE_DEV(ctx->d_cellgeo);
  FREE_DEV(ctx->d_cellgeo2);
  FREE_DEV(ctx->d_Nm);
  FREE_DEV(ctx->d_Dm);
  FREE_DEV(ctx->d_alpha);
  FREE_DEV(ctx->d_alpha_part);
  FREE_DEV(ctx->d_surf);
  FREE_DEV(ctx->d_other);
  FREE_DEV(ctx->d_con_line_of);
  FREE_DEV(ctx->d_free_rows);
  FREE_DEV(ctx->d_con_lines);
  FREE_DEV(ctx->d_con_ptr);
  FREE_DEV(ctx->d_con_col);
  FREE_DEV(ctx->d_con_val);
  FREE_DEV(ctx->d_con_inhom);
  FREE_DEV(ctx->d_xn);
  FREE_DEV(ctx->d_xd);
  FREE_DEV(ctx->d_xdiag);
  FREE_DEV(ctx->d_list_o);
  FREE_DEV(ctx->d_list_s);
  FREE_DEV(ctx->d_yloc);
  for (auto &t : ctx->d_tmp) FREE_DEV(t);
  FREE_DEV(ctx->d_rhs);
  FREE_DEV(ctx->d_sol);
  FREE_DEV(ctx->d_V);
  FREE_DEV(ctx->d_h);
  FREE_DEV(ctx->d_gm);
  FREE_DEV(ctx->d_gm_ctl);
  ctx->gm_doubles = 0;
  FREE_DEV(ctx->d_band);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->h_asm_flag) cudaFreeHost(ctx->h_asm_flag);
  ctx->h_pinned = nullptr;
  ctx->h_asm_flag = nullptr;
  wbem_device_precond_free(ctx);
  wbem_spai_free(ctx);
  wbem_constraints_free(ctx);
  ctx->h_con_lines.clear();
  ctx->h_con_ptr.clear();
  ctx->h_con_col.clear();
  ctx->h_con_val.clear();
  ctx->have_geometry = ctx->assembled = ctx->have_alpha = ctx->have_masks = false;
  ctx->precond_ready = false;
  ctx->n_lines = 0;
}

int wbem_destroy(wbem_ctx *ctx)
{
  CHECK_CTX(ctx);
  if (ctx->group) return wbem_group_destroy(ctx);
  return wbem_destroy_single(ctx);
}

} // extern "C"

int wbem_destroy_single(wbem_ctx *ctx)
{
  cudaSetDevice(ctx->dev);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  free_topology(ctx);
  if (ctx->d_tables) cudaFree(ctx->d_tables);
  if (ctx->d_gauss) cudaFree(ctx->d_gauss);
  wbem_nccl_destroy(ctx);
  for (auto &ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  ctx->timer.release();
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

extern "C" {

int wbem_row_block(const wbem_ctx *ctx, uint32_t *row0, uint32_t *row1)
{
  CHECK_CTX(ctx);
  const bool all = wbem_group_forward(ctx); // the handle of a single-process group owns every row
  if (row0) *row0 = all ? 0 : ctx->row0;
  if (row1) *row1 = all ? ctx->N : ctx->row1;
  return 0;
}

int wbem_set_topology(wbem_ctx *ctx, uint32_t N, uint32_t C, const uint32_t *cell_dofs,
                      const uint8_t *cell_dir_flag, const uint32_t *dn_ptr, const uint32_t *dn_idx)
{
  CHECK_CTX(ctx);
  if (!cell_dofs || !cell_dir_flag || !dn_ptr || !dn_idx || N == 0)
    WBEM_FAIL(ctx, -1, "wbem_set_topology: null argument or N == 0");
  if ((uint64_t)N >= (1ull << 31)) WBEM_FAIL(ctx, -1, "N too large");
  GROUP_FORWARD(ctx, wbem_set_topology(s, N, C, cell_dofs, cell_dir_flag, dn_ptr, dn_idx));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->group)
    { // no peer may still be storing into this block's gather buffer when it is freed
      const int brc = wbem_group_barrier(ctx);
      if (brc) return brc;
    }
  free_topology(ctx);
  ctx->N = N;
  ctx->C = C;
  ctx->ld = (N + 63u) / 64u * 64u;
  const uint32_t P = (uint32_t)ctx->p.world_size, rank = (uint32_t)ctx->p.rank;
  ctx->chunk = (N + P - 1) / P;
  ctx->row0 = std::min(N, rank * ctx->chunk);
  ctx->row1 = std::min(N, ctx->row0 + ctx->chunk);
  ctx->nloc = ctx->row1 - ctx->row0;
  ctx->h_cell_dofs.assign(cell_dofs, cell_dofs + 4 * (size_t)C);
  ctx->h_dir.assign(cell_dir_flag, cell_dir_flag + C);
  // double_nodes_set is a vector<set<unsigned>> in the reference (computational_domain.h:213): every set
  // is sorted, duplicate-free and holds its own dof.  A flattened copy may not be: normalise it, so that
  // a missing i cannot send the self-cells of row i through the regular rule, an empty set cannot be
  // walked out of range and an unsorted one cannot change which dof compute_constraints picks first.
  if (dn_ptr[0] != 0) WBEM_FAIL(ctx, -1, "dn_ptr[0] must be 0");
  for (uint32_t i = 0; i < N; ++i)
    if (dn_ptr[i + 1] < dn_ptr[i]) WBEM_FAIL(ctx, -1, "dn_ptr is not monotone at dof %u", i);
  for (uint32_t k = 0; k < dn_ptr[N]; ++k)
    if (dn_idx[k] >= N) WBEM_FAIL(ctx, -1, "double_nodes_set entry out of range");
  ctx->h_dn_ptr.assign(N + 1, 0);
  ctx->h_dn_idx.clear();
  ctx->h_dn_idx.reserve((size_t)dn_ptr[N] + N);
  {
    std::vector<uint32_t> set;
    for (uint32_t i = 0; i < N; ++i)
      {
        set.assign(dn_idx + dn_ptr[i], dn_idx + dn_ptr[i + 1]);
        set.push_back(i);
        std::sort(set.begin(), set.end());
        set.erase(std::unique(set.begin(), set.end()), set.end());
        ctx->h_dn_idx.insert(ctx->h_dn_idx.end(), set.begin(), set.end());
        ctx->h_dn_ptr[i + 1] = (uint32_t)ctx->h_dn_idx.size();
      }
  }
  dn_ptr = ctx->h_dn_ptr.data(); // from here on: the normalised sets
  dn_idx = ctx->h_dn_idx.data();

  ctx->has_degenerate_cells = false;
  for (uint32_t c = 0; c < C && !ctx->has_degenerate_cells; ++c)
    for (int j = 1; j < 4; ++j)
      for (int i = 0; i < j; ++i)
        if (cell_dofs[4 * (size_t)c + i] == cell_dofs[4 * (size_t)c + j]) ctx->has_degenerate_cells = true;
  // tiling plan (plan.cpp); W and the cell cap match k_assemble_tiled's shared-memory tile
  int rc = wbem_build_plan(N, C, cell_dofs, wbem_tile_width(), 64, &ctx->plan);
  if (rc) WBEM_FAIL(ctx, -1, "wbem_build_plan failed (%d): cell dof out of range?", rc);
  const AssemblyPlan &pl = ctx->plan;

  // cell tables in processing order
  std::vector<uint32_t> dofs_po(4 * (size_t)C);
  std::vector<uint8_t> dir_po(C);
  for (uint32_t p = 0; p < C; ++p)
    {
      const uint32_t c = pl.cell_order[p];
      for (int j = 0; j < 4; ++j) dofs_po[4 * (size_t)p + j] = cell_dofs[4 * (size_t)c + j];
      dir_po[p] = cell_dir_flag[c] ? 1 : 0;
    }
  // singular (node, cell) pairs of the local rows (reference :223-230): the cell holds a dof
  // of double_nodes_set[i]; singular_index = first such local dof.
  std::vector<uint32_t> nptr(N + 1, 0), nadj(4 * (size_t)C);
  for (size_t k = 0; k < 4 * (size_t)C; ++k) nptr[cell_dofs[k] + 1]++;
  for (uint32_t i = 0; i < N; ++i) nptr[i + 1] += nptr[i];
  {
    std::vector<uint32_t> fill(nptr.begin(), nptr.end() - 1);
    for (uint32_t c = 0; c < C; ++c)
      for (int j = 0; j < 4; ++j) nadj[fill[cell_dofs[4 * (size_t)c + j]]++] = c;
  }
  std::vector<uint32_t> sing_ptr(ctx->nloc + 1, 0), sing_pos;
  std::vector<uint8_t> sing_idx;
  std::vector<std::pair<uint32_t, uint8_t>> tmp;
  for (uint32_t r = 0; r < ctx->nloc; ++r)
    {
      const uint32_t gi = ctx->row0 + r;
      tmp.clear();
      const uint32_t *sb = dn_idx + dn_ptr[gi], *se = dn_idx + dn_ptr[gi + 1];
      for (const uint32_t *m = sb; m != se; ++m)
        for (uint32_t a = nptr[*m]; a < nptr[*m + 1]; ++a)
          {
            const uint32_t c = nadj[a];
            int first = -1;
            for (int j = 0; j < 4 && first < 0; ++j)
              if (std::find(sb, se, cell_dofs[4 * (size_t)c + j]) != se) first = j;
            tmp.emplace_back(pl.cell_pos[c], (uint8_t)first);
          }
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      for (auto &pr : tmp)
        {
          sing_pos.push_back(pr.first);
          sing_idx.push_back(pr.second);
        }
      sing_ptr[r + 1] = (uint32_t)sing_pos.size();
    }
  ctx->n_sing = (uint32_t)sing_pos.size();
  // byte map: does (row tile, cluster) hold any singular pair?  (k_assemble_tiled)
  std::vector<uint32_t> cluster_of_pos(C);
  for (uint32_t k = 0; k < pl.n_clusters; ++k)
    for (uint32_t p = pl.cl_cell_ptr[k]; p < pl.cl_cell_ptr[k + 1]; ++p) cluster_of_pos[p] = k;
  const uint32_t tile_rows = wbem_tile_rows();
  const uint32_t row_tiles = (ctx->nloc + tile_rows - 1) / tile_rows;
  std::vector<uint8_t> tile_sing((size_t)std::max(1u, row_tiles) * std::max(1u, pl.n_clusters), 0);
  for (uint32_t r = 0; r < ctx->nloc; ++r)
    for (uint32_t k = sing_ptr[r]; k < sing_ptr[r + 1]; ++k)
      tile_sing[(size_t)(r / tile_rows) * pl.n_clusters + cluster_of_pos[sing_pos[k]]] = 1;
  // work items of the regular-pair kernel in launch (colour) order
  std::vector<uint32_t> cta_desc(4 * (size_t)std::max(1u, pl.n_clusters), 0);
  for (uint32_t k = 0; k < pl.n_clusters; ++k)
    {
      const uint32_t cl = pl.color_clusters[k];
      cta_desc[4 * (size_t)k + 0] = pl.cl_cell_ptr[cl];
      cta_desc[4 * (size_t)k + 1] = pl.cl_slot_ptr[cl];
      cta_desc[4 * (size_t)k + 2] = (pl.cl_cell_ptr[cl + 1] - pl.cl_cell_ptr[cl]) | ((pl.cl_slot_ptr[cl + 1] - pl.cl_slot_ptr[cl]) << 8);
      cta_desc[4 * (size_t)k + 3] = cl;
    }

  // uploads
  if ((rc = dev_upload(ctx, &ctx->d_cell_dofs, dofs_po))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_dir, dir_po))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_cell_order, pl.cell_order))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_colpos, pl.colpos))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_colperm, pl.colperm))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_sing_ptr, sing_ptr))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_sing_cellpos, sing_pos))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_sing_idx, sing_idx))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_tile_sing, tile_sing))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_slot_col, pl.slot_col))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_cta_desc, cta_desc))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_cell_slots, pl.cell_slots))) return rc;
  if ((rc = wbem_upload_stream_tables(ctx, sing_ptr, sing_pos, cluster_of_pos))) return rc;

  // storage (BEMProblem::reinit, :55-71)
  const size_t ld = ctx->ld, nloc = ctx->nloc;
  const int ntmp = ctx->p.gmres_n_tmp_vectors;
  const int band = std::max(ctx->p.preconditioner_band, 2);
  if ((rc = dev_alloc(ctx, &ctx->d_xyz, 3 * (size_t)N))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_cellgeo, (size_t)C * 8 * ctx->qt.nq + 16))) return rc;
  if (ctx->qt.n1 == 4)
    {
      if ((rc = dev_alloc(ctx, &ctx->d_cellgeo2, (size_t)C * wbem_line_record_doubles() + 64))) return rc;
    }
  else
    FREE_DEV(ctx->d_cellgeo2);
  if ((rc = dev_alloc(ctx, &ctx->d_Nm, nloc * ld))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_Dm, nloc * ld))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_alpha, ld))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_alpha_part, ((size_t)pl.n_clusters + 1) * nloc))) return rc;
  ctx->alpha_parts_valid = false;
  if ((rc = dev_alloc(ctx, &ctx->d_surf, ld))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_other, ld))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_xn, ld))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_xd, ld))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_xdiag, ld))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_list_o, ld / 64))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_list_s, ld / 64))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_yloc, (size_t)ctx->chunk * P + 64))) return rc;
  if (P > 1)
    {
      const size_t nd = 2 * (size_t)WBEM_MULTI_MAX * ctx->chunk * P + WBEM_MAX_PEERS;
      if ((rc = dev_alloc(ctx, &ctx->d_p2p, nd))) return rc;
      if ((rc = dev_alloc(ctx, &ctx->d_done_counter, 1))) return rc;
      if ((rc = dev_alloc(ctx, &ctx->d_gather_timeout, 1))) return rc;
      CUDA_OK(ctx, cudaMemsetAsync(ctx->d_gather_timeout, 0, sizeof(unsigned int), ctx->stream));
      CUDA_OK(ctx, cudaMemsetAsync(ctx->d_p2p, 0, sizeof(double) * nd, ctx->stream));
      CUDA_OK(ctx, cudaMemsetAsync(ctx->d_done_counter, 0, sizeof(unsigned long long), ctx->stream));
      ctx->p2p_epoch = 0;
      ctx->gemv_done_total = 0;
    }
  for (auto &t : ctx->d_tmp)
    if ((rc = dev_alloc(ctx, &t, ld))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_rhs, ld))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_sol, ld))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_V, (size_t)(ntmp - 1) * ld))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_h, 1024 + (size_t)(10 * WBEM_GMRES_KMAX) + (size_t)(N + 127) / 128 + 64))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_band, (size_t)ctx->chunk * P * band))) return rc;
  CUDA_OK(ctx, cudaMemsetAsync(ctx->d_xn, 0, sizeof(double) * ld, ctx->stream));
  CUDA_OK(ctx, cudaMemsetAsync(ctx->d_xd, 0, sizeof(double) * ld, ctx->stream));
  CUDA_OK(ctx, cudaMemsetAsync(ctx->d_alpha, 0, sizeof(double) * ld, ctx->stream));
  CUDA_OK(ctx, cudaMemsetAsync(ctx->d_yloc, 0, sizeof(double) * ((size_t)ctx->chunk * P + 64), ctx->stream));
  // the padding columns [N, ld) are read by the mat-vecs: keep them zero forever
  if (ld > N && nloc)
    {
      CUDA_OK(ctx, cudaMemset2DAsync(ctx->d_Nm + N, ld * sizeof(double), 0, (ld - N) * sizeof(double), nloc, ctx->stream));
      CUDA_OK(ctx, cudaMemset2DAsync(ctx->d_Dm + N, ld * sizeof(double), 0, (ld - N) * sizeof(double), nloc, ctx->stream));
    }
  ctx->pinned_doubles = 8 * (size_t)ld + 2048;
  CUDA_OK(ctx, cudaMallocHost((void **)&ctx->h_pinned, ctx->pinned_doubles * sizeof(double)));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->op_version++;
  if (ctx->group && (rc = wbem_group_p2p_setup(ctx))) return rc;
  return 0;
}

int wbem_set_geometry_dev(wbem_ctx *ctx, const double *d_support_points)
{
  CHECK_CTX(ctx);
  if (!ctx->N) WBEM_FAIL(ctx, -3, "wbem_set_geometry before wbem_set_topology");
  GROUP_FORWARD(ctx, wbem_set_geometry_dev(s, d_support_points));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_xyz, d_support_points, sizeof(double) * 3 * (size_t)ctx->N,
                               cudaMemcpyDeviceToDevice, ctx->stream));
  ctx->have_geometry = true;
  ctx->fevalues_given = false; // quadrature data follow the new support points again
  ctx->assembled = false;
  ctx->have_alpha = false;
  ctx->geom_version++;
  return 0;
}

int wbem_set_geometry(wbem_ctx *ctx, const double *support_points)
{
  CHECK_CTX(ctx);
  if (!ctx->N) WBEM_FAIL(ctx, -3, "wbem_set_geometry before wbem_set_topology");
  if (!support_points) WBEM_FAIL(ctx, -1, "null support_points");
  GROUP_FORWARD(ctx, wbem_set_geometry(s, support_points));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_xyz, support_points, sizeof(double) * 3 * (size_t)ctx->N,
                               cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream)); // caller may reuse its buffer
  ctx->have_geometry = true;
  ctx->fevalues_given = false; // quadrature data follow the new support points again
  ctx->assembled = false;
  ctx->have_alpha = false;
  ctx->geom_version++;
  return 0;
}

static int assemble_async(wbem_ctx *ctx)
{
  if (!ctx->have_geometry) WBEM_FAIL(ctx, -3, "wbem_assemble before wbem_set_geometry");
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  cudaStream_t st = ctx->stream;
  CUDA_OK(ctx, cudaEventRecord(ctx->ev[6], st));
  int rc = ctx->fevalues_given ? 0 : wbem_launch_geometry(ctx); // caller-supplied FEValues stay in place
  if (rc) return rc;
  rc = wbem_launch_assemble(ctx); // records ev[0..2]
  if (rc) return rc;
  ctx->assembled = true;
  ctx->have_alpha = false;
  ctx->op_version++;
  rc = wbem_launch_alpha(ctx);
  if (rc) return rc;
  CUDA_OK(ctx, cudaEventRecord(ctx->ev[3], st));
  return 0;
}

static int assemble_timings(wbem_ctx *ctx)
{
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_OK(ctx, cudaGetLastError());
  if (ctx->h_asm_flag && *ctx->h_asm_flag)
    WBEM_FAIL(ctx, -7, "assembly: a work item gave up waiting for the clusters it shares columns with (k_assemble_rows)");
  float a = 0, b = 0, c = 0, d = 0, e = 0;
  cudaEventElapsedTime(&a, ctx->ev[6], ctx->ev[0]);
  if (ctx->nloc && ctx->C)
    {
      cudaEventElapsedTime(&b, ctx->ev[0], ctx->ev[1]);
      cudaEventElapsedTime(&c, ctx->ev[1], ctx->ev[2]);
      cudaEventElapsedTime(&d, ctx->ev[2], ctx->ev[3]);
    }
  cudaEventElapsedTime(&e, ctx->ev[6], ctx->ev[3]);
  ctx->tm.geometry_ms = a;
  ctx->tm.assemble_regular_ms = b;
  ctx->tm.assemble_singular_ms = c;
  ctx->tm.alpha_ms = d;
  ctx->tm.assemble_total_ms = e;
  return 0;
}

int wbem_assemble(wbem_ctx *ctx)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_assemble(s));
  int rc = assemble_async(ctx);
  if (rc) return rc;
  return assemble_timings(ctx);
}

int wbem_compute_alpha(wbem_ctx *ctx)
{
  CHECK_CTX(ctx);
  if (!ctx->assembled) WBEM_FAIL(ctx, -3, "wbem_compute_alpha before wbem_assemble");
  GROUP_FORWARD(ctx, wbem_compute_alpha(s));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  int rc = wbem_launch_alpha(ctx, true); // the literal N * (-1) of the reference
  if (rc) return rc;
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int wbem_get_alpha(wbem_ctx *ctx, double *alpha)
{
  CHECK_CTX(ctx);
  if (!ctx->have_alpha) WBEM_FAIL(ctx, -3, "alpha not computed yet");
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  CUDA_OK(ctx, cudaMemcpyAsync(alpha, ctx->d_alpha, sizeof(double) * ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// out[r][colperm[c]] = M[r][c]
__global__ void k_unpermute_rows(uint32_t N, uint32_t ld, uint32_t nrows, const double *__restrict__ M,
                                 const uint32_t *__restrict__ colperm, double *__restrict__ out)
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t r = blockIdx.y;
  if (c < N && r < nrows) out[(size_t)r * N + colperm[c]] = M[(size_t)r * ld + c];
}

int wbem_get_rows(wbem_ctx *ctx, int which, uint32_t r0, uint32_t r1, double *out)
{
  CHECK_CTX(ctx);
  if (!ctx->assembled) WBEM_FAIL(ctx, -3, "wbem_get_rows before wbem_assemble");
  if (wbem_group_forward(ctx))
    { // every row block hands over its part of [r0, r1)
      if (r0 > r1 || r1 > ctx->N) WBEM_FAIL(ctx, -1, "rows [%u,%u) out of range", r0, r1);
      return wbem_group_run(ctx, [&](wbem_ctx *s) -> int {
        const uint32_t a = std::max(r0, s->row0), b = std::min(r1, s->row1);
        return a < b ? wbem_get_rows(s, which, a, b, out + (size_t)(a - r0) * s->N) : 0;
      });
    }
  if (r0 < ctx->row0 || r1 > ctx->row1 || r0 > r1)
    WBEM_FAIL(ctx, -1, "rows [%u,%u) outside this context's block [%u,%u)", r0, r1, ctx->row0, ctx->row1);
  if (r0 == r1) return 0;
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  const double *M = (which == 0 ? ctx->d_Nm : ctx->d_Dm) + (size_t)(r0 - ctx->row0) * ctx->ld;
  const uint32_t N = ctx->N;
  const uint32_t step = std::max<uint32_t>(1, std::min<uint32_t>(r1 - r0, (uint32_t)((64u << 20) / (8 * (size_t)N) + 1)));
  double *d_out = nullptr;
  CUDA_OK(ctx, cudaMalloc((void **)&d_out, sizeof(double) * (size_t)step * N));
  for (uint32_t r = r0; r < r1; r += step)
    {
      const uint32_t nr = std::min(step, r1 - r);
      dim3 grid((N + 255) / 256, nr);
      k_unpermute_rows<<<grid, 256, 0, ctx->stream>>>(N, ctx->ld, nr, M + (size_t)(r - r0) * ctx->ld,
                                                     ctx->d_colperm, d_out);
      ctx->launches++;
      cudaError_t e = cudaMemcpyAsync(out + (size_t)(r - r0) * N, d_out, sizeof(double) * (size_t)nr * N,
                                      cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess)
        {
          cudaFree(d_out);
          WBEM_FAIL(ctx, -2, "CUDA error %s in wbem_get_rows", cudaGetErrorString(e));
        }
    }
  cudaFree(d_out);
  return 0;
}

int wbem_set_masks(wbem_ctx *ctx, const double *surface_nodes, const double *other_nodes)
{
  CHECK_CTX(ctx);
  if (!ctx->N) WBEM_FAIL(ctx, -3, "wbem_set_masks before wbem_set_topology");
  GROUP_FORWARD(ctx, wbem_set_masks(s, surface_nodes, other_nodes));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  const uint32_t N = ctx->N;
  const bool same = ctx->have_masks && !memcmp(ctx->h_surf.data(), surface_nodes, sizeof(double) * N) &&
                    !memcmp(ctx->h_other.data(), other_nodes, sizeof(double) * N);
  if (same) return 0;
  ctx->h_surf.assign(surface_nodes, surface_nodes + N);
  ctx->h_other.assign(other_nodes, other_nodes + N);
  double linf = 0;
  for (uint32_t i = 0; i < N; ++i) linf = std::max(linf, std::fabs(surface_nodes[i]));
  ctx->pure_neumann = linf < 1e-10; // source/bem_problem.cc:667
  // 64-column chunks (storage order) that hold a non-zero mask entry
  std::vector<uint32_t> lo, ls;
  const uint32_t nchunks = ctx->ld / 64;
  for (uint32_t ch = 0; ch < nchunks; ++ch)
    {
      bool any_o = false, any_s = false;
      for (uint32_t c = ch * 64; c < std::min(N, (ch + 1) * 64); ++c)
        {
          const uint32_t dof = ctx->plan.colperm[c];
          any_o |= other_nodes[dof] != 0.0;
          any_s |= surface_nodes[dof] != 0.0;
        }
      if (any_o) lo.push_back(ch);
      if (any_s) ls.push_back(ch);
    }
  ctx->n_list_o = (int)lo.size();
  ctx->n_list_s = (int)ls.size();
  cudaStream_t st = ctx->stream;
  CUDA_OK(ctx, cudaStreamSynchronize(st));
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_surf, surface_nodes, sizeof(double) * N, cudaMemcpyHostToDevice, st));
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_other, other_nodes, sizeof(double) * N, cudaMemcpyHostToDevice, st));
  if (!lo.empty())
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_list_o, lo.data(), sizeof(uint32_t) * lo.size(), cudaMemcpyHostToDevice, st));
  if (!ls.empty())
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_list_s, ls.data(), sizeof(uint32_t) * ls.size(), cudaMemcpyHostToDevice, st));
  CUDA_OK(ctx, cudaStreamSynchronize(st));
  ctx->have_masks = true;
  ctx->op_version++;
  return 0;
}

int wbem_set_constraints(wbem_ctx *ctx, uint32_t n_lines, const uint32_t *lines, const uint32_t *ptr,
                         const uint32_t *col, const double *val, const double *inhom)
{
  CHECK_CTX(ctx);
  if (!ctx->N) WBEM_FAIL(ctx, -3, "wbem_set_constraints before wbem_set_topology");
  GROUP_FORWARD(ctx, wbem_set_constraints(s, n_lines, lines, ptr, col, val, inhom));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  const uint32_t N = ctx->N;
  std::vector<int32_t> line_of(N, -1);
  for (uint32_t k = 0; k < n_lines; ++k)
    {
      if (lines[k] >= N) WBEM_FAIL(ctx, -1, "constraint line dof out of range");
      line_of[lines[k]] = (int32_t)k;
    }
  const uint32_t nnz = n_lines ? ptr[n_lines] : 0;
  for (uint32_t k = 0; k < nnz; ++k)
    if (col[k] >= N) WBEM_FAIL(ctx, -1, "constraint entry out of range");
  std::vector<uint32_t> vl(lines, lines + n_lines), vp(ptr, ptr + (n_lines ? n_lines + 1 : 1)),
    vc(col, col + nnz);
  std::vector<double> vv(val, val + nnz), vi(inhom, inhom + n_lines);
  if (!n_lines) vp.assign(1, 0);
  // same lines and coefficients as last time (the per-solve pattern: only the inhomogeneities
  // follow tmp_rhs): one small copy, nothing else changes
  if (ctx->n_lines == n_lines && n_lines && ctx->d_con_inhom && vl == ctx->h_con_lines && vp == ctx->h_con_ptr &&
      vc == ctx->h_con_col && vv == ctx->h_con_val)
    {
      CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_con_inhom, vi.data(), sizeof(double) * n_lines, cudaMemcpyHostToDevice,
                                   ctx->stream));
      CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
      return 0;
    }
  // the preconditioners depend on which rows are constrained and on the line coefficients
  if (line_of != ctx->h_con_line_of || vc != ctx->h_con_col || vv != ctx->h_con_val) ctx->op_version++;
  ctx->h_con_line_of = line_of;
  ctx->h_con_lines = vl;
  ctx->h_con_ptr = vp;
  ctx->h_con_col = vc;
  ctx->h_con_val = vv;
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  int rc;
  std::vector<uint32_t> free_rows;
  for (uint32_t r = 0; r < ctx->nloc; ++r)
    if (line_of[ctx->row0 + r] < 0) free_rows.push_back(r);
  ctx->n_free_rows = (uint32_t)free_rows.size();
  if ((rc = dev_upload(ctx, &ctx->d_free_rows, free_rows))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_con_line_of, line_of))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_con_lines, vl))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_con_ptr, vp))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_con_col, vc))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_con_val, vv))) return rc;
  if ((rc = dev_upload(ctx, &ctx->d_con_inhom, vi))) return rc;
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->n_lines = n_lines;
  return 0;
}

static int ensure_alpha(wbem_ctx *ctx)
{
  if (!ctx->assembled) WBEM_FAIL(ctx, -3, "matrices not assembled");
  if (!ctx->have_alpha) return wbem_launch_alpha(ctx);
  return 0;
}

// host vector in -> operator -> host vector out
static int host_apply(wbem_ctx *ctx, int mode, bool constrained, double *dst, const double *src)
{
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  int rc = ensure_alpha(ctx);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_tmp[1], src, sizeof(double) * ctx->N, cudaMemcpyHostToDevice, st));
  rc = wbem_apply_operator(ctx, mode, ctx->d_tmp[1], ctx->d_tmp[2], constrained);
  if (rc) return rc;
  if (wbem_is_root(ctx)) // every row block holds the gathered result; one of them hands it over
    CUDA_OK(ctx, cudaMemcpyAsync(dst, ctx->d_tmp[2], sizeof(double) * ctx->N, cudaMemcpyDeviceToHost, st));
  CUDA_OK(ctx, cudaStreamSynchronize(st));
  return wbem_check_gather_timeout(ctx);
}

int wbem_vmult(wbem_ctx *ctx, double *dst, const double *src)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_vmult(s, dst, src));
  return host_apply(ctx, 0, false, dst, src);
}
int wbem_constrained_vmult(wbem_ctx *ctx, double *dst, const double *src)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_constrained_vmult(s, dst, src));
  return host_apply(ctx, 0, true, dst, src);
}
int wbem_compute_rhs(wbem_ctx *ctx, double *dst, const double *src)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_compute_rhs(s, dst, src));
  return host_apply(ctx, 1, false, dst, src);
}
int wbem_distribute_rhs(wbem_ctx *ctx, double *rhs)
{
  CHECK_CTX(ctx);
  // O(n_lines) host epilogue (include/constrained_matrix.h:89-94)
  if (!ctx->n_lines) return 0;
  std::vector<double> inhom(ctx->n_lines);
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  CUDA_OK(ctx, cudaMemcpy(inhom.data(), ctx->d_con_inhom, sizeof(double) * ctx->n_lines, cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < ctx->N; ++i)
    if (ctx->h_con_line_of[i] >= 0) rhs[i] = inhom[ctx->h_con_line_of[i]];
  return 0;
}

int wbem_assemble_preconditioner(wbem_ctx *ctx)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_assemble_preconditioner(s));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  if (!ctx->have_masks) WBEM_FAIL(ctx, -3, "assemble_preconditioner before wbem_set_masks");
  int rc = ensure_alpha(ctx);
  if (rc) return rc;
  rc = wbem_build_preconditioner(ctx);
  if (rc) return rc;
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int wbem_set_precond_kind(wbem_ctx *ctx, int kind)
{
  CHECK_CTX(ctx);
  if (kind < 0 || kind > 1) WBEM_FAIL(ctx, -1, "precond_kind must be 0 (band) or 1 (sparse approximate inverse)");
  GROUP_FORWARD(ctx, wbem_set_precond_kind(s, kind));
  if (kind != ctx->p.precond_kind)
    {
      ctx->p.precond_kind = kind;
      ctx->precond_ready = false; // the other kind's factors are not valid for this one
    }
  return 0;
}

int wbem_precond_vmult(wbem_ctx *ctx, double *dst, const double *src)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_precond_vmult(s, dst, src));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  cudaStream_t st = ctx->stream;
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_tmp[1], src, sizeof(double) * ctx->N, cudaMemcpyHostToDevice, st));
  int rc = wbem_apply_preconditioner(ctx, ctx->d_tmp[1], ctx->d_tmp[2]);
  if (rc) return rc;
  if (wbem_is_root(ctx))
    CUDA_OK(ctx, cudaMemcpyAsync(dst, ctx->d_tmp[2], sizeof(double) * ctx->N, cudaMemcpyDeviceToHost, st));
  CUDA_OK(ctx, cudaStreamSynchronize(st));
  return 0;
}

int wbem_get_band(wbem_ctx *ctx, double *out)
{
  CHECK_CTX(ctx);
  const int band = ctx->p.preconditioner_band;
  if (band <= 0 || !ctx->precond_ready) WBEM_FAIL(ctx, -3, "no band preconditioner assembled");
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  if (wbem_group_forward(ctx))
    { // the gathered band rows of all row blocks sit on every device: all N rows from this one
      CUDA_OK(ctx, cudaMemcpy(out, ctx->d_band, sizeof(double) * (size_t)ctx->N * band, cudaMemcpyDeviceToHost));
      return 0;
    }
  CUDA_OK(ctx, cudaMemcpy(out, ctx->d_band + (size_t)ctx->p.rank * ctx->chunk * band,
                          sizeof(double) * (size_t)ctx->nloc * band, cudaMemcpyDeviceToHost));
  return 0;
}

int wbem_solve_system_dev(wbem_ctx *ctx, double *d_phi, double *d_dphi_dn, const double *d_tmp_rhs,
                          int *iters, double *last_res)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_solve_system_dev(s, d_phi, d_dphi_dn, d_tmp_rhs, wbem_is_root(s) ? iters : nullptr,
                                           wbem_is_root(s) ? last_res : nullptr));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  if (!wbem_is_root(ctx))
    { // the caller's arrays live on the first row block's device: the others solve on private
      // copies (peer reads) and leave the in-out arrays to that block
      cudaStream_t st = ctx->stream;
      const size_t nb = sizeof(double) * ctx->N;
      CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_tmp[3], d_phi, nb, cudaMemcpyDefault, st));
      CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_tmp[4], d_dphi_dn, nb, cudaMemcpyDefault, st));
      CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_tmp[5], d_tmp_rhs, nb, cudaMemcpyDefault, st));
      return wbem_solve_system_device(ctx, ctx->d_tmp[3], ctx->d_tmp[4], ctx->d_tmp[5], iters, last_res);
    }
  return wbem_solve_system_device(ctx, d_phi, d_dphi_dn, d_tmp_rhs, iters, last_res);
}

int wbem_solve_system(wbem_ctx *ctx, double *phi, double *dphi_dn, const double *tmp_rhs, int *iters,
                      double *last_res)
{
  CHECK_CTX(ctx);
  if (!ctx->N) WBEM_FAIL(ctx, -3, "solve_system before wbem_set_topology");
  GROUP_FORWARD(ctx, wbem_solve_system(s, phi, dphi_dn, tmp_rhs, wbem_is_root(s) ? iters : nullptr,
                                       wbem_is_root(s) ? last_res : nullptr));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  cudaStream_t st = ctx->stream;
  const size_t nb = sizeof(double) * ctx->N;
  double *d_phi = ctx->d_tmp[3], *d_dphi = ctx->d_tmp[4], *d_bc = ctx->d_tmp[5];
  CUDA_OK(ctx, cudaMemcpyAsync(d_phi, phi, nb, cudaMemcpyHostToDevice, st));
  CUDA_OK(ctx, cudaMemcpyAsync(d_dphi, dphi_dn, nb, cudaMemcpyHostToDevice, st));
  CUDA_OK(ctx, cudaMemcpyAsync(d_bc, tmp_rhs, nb, cudaMemcpyHostToDevice, st));
  const int rc = wbem_solve_system_device(ctx, d_phi, d_dphi, d_bc, iters, last_res);
  if (rc < 0) return rc;
  if (wbem_is_root(ctx))
    {
      CUDA_OK(ctx, cudaMemcpyAsync(phi, d_phi, nb, cudaMemcpyDeviceToHost, st));
      CUDA_OK(ctx, cudaMemcpyAsync(dphi_dn, d_dphi, nb, cudaMemcpyDeviceToHost, st));
    }
  CUDA_OK(ctx, cudaStreamSynchronize(st));
  return rc;
}

static int block_apply(wbem_ctx *ctx, int mode, bool constrained, int nvec, double *dst, const double *src);
int wbem_constrained_vmult_multi(wbem_ctx *ctx, int nvec, double *dst, const double *src)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_constrained_vmult_multi(s, nvec, dst, src));
  return block_apply(ctx, 0, true, nvec, dst, src);
}
int wbem_compute_rhs_multi(wbem_ctx *ctx, int nvec, double *dst, const double *src)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_compute_rhs_multi(s, nvec, dst, src));
  return block_apply(ctx, 1, false, nvec, dst, src);
}
static int block_apply(wbem_ctx *ctx, int mode, bool constrained, int nvec, double *dst, const double *src)
{
  if (nvec < 1 || nvec > WBEM_MULTI_MAX || !dst || !src) WBEM_FAIL(ctx, -1, "block mat-vec takes 1..%d vectors", WBEM_MULTI_MAX);
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  int rc = ensure_alpha(ctx);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const size_t nd = (size_t)nvec * ctx->N;
  double *d = nullptr;
  CUDA_OK(ctx, cudaMalloc((void **)&d, 2 * sizeof(double) * nd));
  const double *ps[WBEM_MULTI_MAX];
  double *pd[WBEM_MULTI_MAX];
  for (int b = 0; b < nvec; ++b)
    {
      ps[b] = d + (size_t)b * ctx->N;
      pd[b] = d + nd + (size_t)b * ctx->N;
    }
  cudaError_t e = cudaMemcpyAsync(d, src, sizeof(double) * nd, cudaMemcpyHostToDevice, st);
  rc = e == cudaSuccess ? wbem_apply_operator_multi(ctx, mode, nvec, ps, pd, constrained) : -2;
  if (rc == 0 && wbem_is_root(ctx)) e = cudaMemcpyAsync(dst, d + nd, sizeof(double) * nd, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) WBEM_FAIL(ctx, -2, "CUDA error %s in wbem_constrained_vmult_multi", cudaGetErrorString(e));
  return rc ? rc : wbem_check_gather_timeout(ctx);
}

int wbem_solve_system_multi(wbem_ctx *ctx, int nrhs, double *phi, double *dphi_dn, const double *tmp_rhs, int *iters,
                            double *last_res)
{
  CHECK_CTX(ctx);
  if (!ctx->N) WBEM_FAIL(ctx, -3, "solve_system_multi before wbem_set_topology");
  if (nrhs < 0 || (nrhs && (!phi || !dphi_dn || !tmp_rhs))) WBEM_FAIL(ctx, -1, "bad argument");
  if (nrhs == 0) return 0;
  GROUP_FORWARD(ctx, wbem_solve_system_multi(s, nrhs, phi, dphi_dn, tmp_rhs, wbem_is_root(s) ? iters : nullptr,
                                             wbem_is_root(s) ? last_res : nullptr));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  cudaStream_t st = ctx->stream;
  const size_t nd = (size_t)nrhs * ctx->N, nbytes = sizeof(double) * nd;
  double *d = wbem_multi_io(ctx, 3 * nd);
  if (!d) return -2;
  double *d_phi = d, *d_dphi = d + nd, *d_bc = d + 2 * nd;
  cudaError_t e = cudaMemcpyAsync(d_phi, phi, nbytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_dphi, dphi_dn, nbytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_bc, tmp_rhs, nbytes, cudaMemcpyHostToDevice, st);
  int rc = e == cudaSuccess ? wbem_solve_system_multi_device(ctx, nrhs, d_phi, d_dphi, d_bc, iters, last_res) : -2;
  if (rc >= 0 && wbem_is_root(ctx))
    {
      e = cudaMemcpyAsync(phi, d_phi, nbytes, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(dphi_dn, d_dphi, nbytes, cudaMemcpyDeviceToHost, st);
    }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) WBEM_FAIL(ctx, -2, "CUDA error %s in wbem_solve_system_multi", cudaGetErrorString(e));
  return rc;
}

int wbem_gmres(wbem_ctx *ctx, const double *rhs, double *sol, int *iters, double *last_res)
{
  CHECK_CTX(ctx);
  if (!ctx->N) WBEM_FAIL(ctx, -3, "wbem_gmres before wbem_set_topology");
  if (!rhs || !sol) WBEM_FAIL(ctx, -1, "null argument");
  GROUP_FORWARD(ctx, wbem_gmres(s, rhs, sol, wbem_is_root(s) ? iters : nullptr, wbem_is_root(s) ? last_res : nullptr));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  cudaStream_t st = ctx->stream;
  const size_t nb = sizeof(double) * ctx->N;
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_tmp[5], rhs, nb, cudaMemcpyHostToDevice, st));
  const int rc = wbem_solve_system_device(ctx, nullptr, nullptr, ctx->d_tmp[5], iters, last_res);
  if (rc < 0) return rc;
  if (wbem_is_root(ctx)) CUDA_OK(ctx, cudaMemcpyAsync(sol, ctx->d_sol, nb, cudaMemcpyDeviceToHost, st));
  CUDA_OK(ctx, cudaStreamSynchronize(st));
  return rc;
}

int wbem_set_fevalues(wbem_ctx *ctx, const double *q_points, const double *normals, const double *JxW)
{
  CHECK_CTX(ctx);
  if (!ctx->have_geometry) WBEM_FAIL(ctx, -3, "wbem_set_fevalues needs the support points first (wbem_set_geometry)");
  if (!q_points || !normals || !JxW) WBEM_FAIL(ctx, -1, "null argument");
  GROUP_FORWARD(ctx, wbem_set_fevalues(s, q_points, normals, JxW));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  const int rc = wbem_upload_fevalues(ctx, q_points, normals, JxW);
  if (rc) return rc;
  ctx->fevalues_given = true;
  ctx->assembled = false;
  ctx->have_alpha = false;
  return 0;
}

int wbem_solve(wbem_ctx *ctx, const double *support_points, double *phi, double *dphi_dn,
               const double *tmp_rhs, int *iters, double *last_res)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_solve(s, support_points, phi, dphi_dn, tmp_rhs, wbem_is_root(s) ? iters : nullptr,
                                wbem_is_root(s) ? last_res : nullptr));
  int rc = wbem_set_geometry(ctx, support_points);
  if (rc) return rc;
  rc = assemble_async(ctx);
  if (rc) return rc;
  rc = wbem_solve_system(ctx, phi, dphi_dn, tmp_rhs, iters, last_res);
  if (rc < 0) return rc;
  const int rc2 = assemble_timings(ctx);
  return rc2 ? rc2 : rc;
}

int wbem_solve_dev(wbem_ctx *ctx, const double *d_support_points, double *d_phi, double *d_dphi_dn,
                   const double *d_tmp_rhs, int *iters, double *last_res)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_solve_dev(s, d_support_points, d_phi, d_dphi_dn, d_tmp_rhs, wbem_is_root(s) ? iters : nullptr,
                                    wbem_is_root(s) ? last_res : nullptr));
  int rc = wbem_set_geometry_dev(ctx, d_support_points);
  if (rc) return rc;
  rc = assemble_async(ctx);
  if (rc) return rc;
  rc = wbem_solve_system_dev(ctx, d_phi, d_dphi_dn, d_tmp_rhs, iters, last_res);
  if (rc < 0) return rc;
  const int rc2 = assemble_timings(ctx);
  return rc2 ? rc2 : rc;
}

// BEMProblem<3>::residual (source/bem_problem.cc:903-961):
//   rrhs = -distribute_rhs(compute_rhs(dphi_dn.o + phi.s)) ; sol = dphi_dn.s + phi.o ;
//   res = cc.vmult(sol) + rrhs
__global__ void k_residual_inputs(uint32_t N, const double *__restrict__ phi, const double *__restrict__ dphi,
                                  const double *__restrict__ s, const double *__restrict__ o,
                                  double *__restrict__ bc, double *__restrict__ sol)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  bc[i] = dphi[i] * o[i] + phi[i] * s[i];
  sol[i] = dphi[i] * s[i] + phi[i] * o[i];
}
__global__ void k_residual_combine(uint32_t N, const double *__restrict__ av, const double *__restrict__ rhs,
                                   const int32_t *__restrict__ line_of, const double *__restrict__ inhom,
                                   double *__restrict__ res)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double r = rhs[i];
  if (line_of && line_of[i] >= 0) r = inhom[line_of[i]];
  res[i] = av[i] + (-1.0 * r);
}

int wbem_residual(wbem_ctx *ctx, double *res, const double *phi, const double *dphi_dn)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_residual(s, res, phi, dphi_dn));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  if (!ctx->have_masks) WBEM_FAIL(ctx, -3, "residual before wbem_set_masks");
  int rc = ensure_alpha(ctx);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const uint32_t N = ctx->N;
  const size_t nb = sizeof(double) * N;
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_tmp[3], phi, nb, cudaMemcpyHostToDevice, st));
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_tmp[4], dphi_dn, nb, cudaMemcpyHostToDevice, st));
  k_residual_inputs<<<(N + 255) / 256, 256, 0, st>>>(N, ctx->d_tmp[3], ctx->d_tmp[4], ctx->d_surf,
                                                    ctx->d_other, ctx->d_tmp[5], ctx->d_tmp[6]);
  ctx->launches++;
  // compute_constraints(constraints, serv_tmp_rhs) (:929)
  if (ctx->p.auto_constraints && (rc = wbem_compute_constraints_device(ctx, ctx->d_tmp[5]))) return rc;
  if ((rc = wbem_apply_operator(ctx, 1, ctx->d_tmp[5], ctx->d_tmp[7], false))) return rc;
  if ((rc = wbem_apply_operator(ctx, 0, ctx->d_tmp[6], ctx->d_tmp[2], true))) return rc;
  k_residual_combine<<<(N + 255) / 256, 256, 0, st>>>(N, ctx->d_tmp[2], ctx->d_tmp[7],
                                                     ctx->n_lines ? ctx->d_con_line_of : nullptr,
                                                     ctx->d_con_inhom, ctx->d_tmp[1]);
  ctx->launches++;
  if (wbem_is_root(ctx)) CUDA_OK(ctx, cudaMemcpyAsync(res, ctx->d_tmp[1], nb, cudaMemcpyDeviceToHost, st));
  CUDA_OK(ctx, cudaStreamSynchronize(st));
  return 0;
}

int wbem_get_system_rhs(wbem_ctx *ctx, double *out)
{
  CHECK_CTX(ctx);
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  CUDA_OK(ctx, cudaMemcpy(out, ctx->d_rhs, sizeof(double) * ctx->N, cudaMemcpyDeviceToHost));
  return 0;
}
int wbem_get_sol(wbem_ctx *ctx, double *out)
{
  CHECK_CTX(ctx);
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  CUDA_OK(ctx, cudaMemcpy(out, ctx->d_sol, sizeof(double) * ctx->N, cudaMemcpyDeviceToHost));
  return 0;
}

int wbem_get_timings(wbem_ctx *ctx, wbem_timings *out)
{
  CHECK_CTX(ctx);
  ctx->tm.kernel_launches = 0;
  const bool all = wbem_group_forward(ctx); // timings of the first row block, launches of all of them
  for (int r = 0; r < (all ? wbem_group_size(ctx) : 1); ++r) ctx->tm.kernel_launches += wbem_group_shard(ctx, r)->launches;
  *out = ctx->tm;
  return 0;
}
int wbem_reset_counters(wbem_ctx *ctx)
{
  CHECK_CTX(ctx);
  const bool all = wbem_group_forward(ctx);
  for (int r = 0; r < (all ? wbem_group_size(ctx) : 1); ++r)
    {
      wbem_ctx *s = wbem_group_shard(ctx, r);
      s->launches = 0;
      memset(&s->tm, 0, sizeof(s->tm));
    }
  return 0;
}

// CUDA-event stopwatch on the library's stream (bench.py brackets its timed region with it)
int wbem_timer_start(wbem_ctx *ctx)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_timer_start(s));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_OK(ctx, cudaEventRecord(ctx->ev[10], ctx->stream));
  return 0;
}
int wbem_timer_stop(wbem_ctx *ctx, double *ms)
{
  CHECK_CTX(ctx);
  if (wbem_group_forward(ctx))
    { // device time of the slowest row block
      std::vector<double> v(wbem_group_size(ctx), 0.0);
      const int rc = wbem_group_run(ctx, [&](wbem_ctx *s) -> int { return wbem_timer_stop(s, &v[s->p.rank]); });
      *ms = *std::max_element(v.begin(), v.end());
      return rc;
    }
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  CUDA_OK(ctx, cudaEventRecord(ctx->ev[11], ctx->stream));
  CUDA_OK(ctx, cudaEventSynchronize(ctx->ev[11]));
  float f = 0;
  CUDA_OK(ctx, cudaEventElapsedTime(&f, ctx->ev[10], ctx->ev[11]));
  *ms = f;
  return 0;
}

int wbem_comm_unique_id(void *id128)
{
  std::string err;
  int rc = wbem_nccl_unique_id(id128, &err);
  if (rc) g_create_error = err;
  return rc;
}
// CUDA-IPC exchange of the gather buffers (after wbem_set_topology, on every rank):
// export -> the launcher all-gathers the 64-byte handles -> import.  Enables the fused
// GEMV + all-gather over NVLink peer memory; without it mat-vecs fall back to ncclAllGather.
int wbem_comm_ipc_export(wbem_ctx *ctx, void *handle64)
{
  CHECK_CTX(ctx);
  if (ctx->group) WBEM_FAIL(ctx, -3, "a single-process context (n_gpus > 1) needs no IPC exchange");
  if (ctx->p.world_size <= 1 || !ctx->d_p2p) WBEM_FAIL(ctx, -3, "ipc_export needs world_size > 1 and wbem_set_topology");
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  cudaIpcMemHandle_t h;
  CUDA_OK(ctx, cudaIpcGetMemHandle(&h, ctx->d_p2p));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}
int wbem_comm_ipc_import(wbem_ctx *ctx, const void *handles)
{
  CHECK_CTX(ctx);
  if (ctx->group) WBEM_FAIL(ctx, -3, "a single-process context (n_gpus > 1) needs no IPC exchange");
  const int P = ctx->p.world_size;
  if (P <= 1 || !ctx->d_p2p) WBEM_FAIL(ctx, -3, "ipc_import needs world_size > 1 and wbem_set_topology");
  if (P > WBEM_MAX_PEERS) WBEM_FAIL(ctx, -1, "at most %d ranks for the peer-to-peer gather", WBEM_MAX_PEERS);
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  wbem_p2p_close(ctx);
  for (int q = 0; q < P; ++q)
    {
      if (q == ctx->p.rank)
        {
          ctx->peer_base[q] = ctx->d_p2p;
          continue;
        }
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char *)handles + 64 * (size_t)q, 64);
      void *ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
        {
          cudaGetLastError();
          wbem_p2p_close(ctx);
          WBEM_FAIL(ctx, -5, "cudaIpcOpenMemHandle(rank %d): %s", q, cudaGetErrorString(e));
        }
      ctx->peer_base[q] = (double *)ptr;
      ctx->peer_opened[q] = true;
    }
  ctx->p2p_ready = true;
  return 0;
}

int wbem_comm_ipc_close(wbem_ctx *ctx)
{
  CHECK_CTX(ctx);
  GROUP_FORWARD(ctx, wbem_comm_ipc_close(s)); // a group falls back to the stream-ordered peer copies
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  wbem_p2p_close(ctx);
  return 0;
}

int wbem_comm_init(wbem_ctx *ctx, const void *id128)
{
  CHECK_CTX(ctx);
  if (ctx->group) WBEM_FAIL(ctx, -3, "a single-process context (n_gpus > 1) has no NCCL communicator");
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  return wbem_nccl_init(ctx, id128);
}

// ---------------------------------------------------------------------------------------
// diagnostics
// ---------------------------------------------------------------------------------------
// eight independent DFMA chains written in PTX (the compiler cannot re-associate them)
__global__ void k_dfma_chains(double *out, int iters)
{
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-9 + k;
  const double b = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i)
    {
#pragma unroll
      for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[k]) : "d"(b), "d"(c));
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}

__global__ void k_dfma_peak(double *out, int iters)
{
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i)
    {
      a0 = fma(a0, b, c);
      a1 = fma(a1, b, c);
      a2 = fma(a2, b, c);
      a3 = fma(a3, b, c);
      a4 = fma(a4, b, c);
      a5 = fma(a5, b, c);
      a6 = fma(a6, b, c);
      a7 = fma(a7, b, c);
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

int wbem_measure_fp64_peak(wbem_ctx *ctx, double *tflops)
{
  CHECK_CTX(ctx);
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  const int blocks = 148 * 8, threads = 256, iters = 1 << 15;
  double *d = nullptr;
  CUDA_OK(ctx, cudaMalloc((void **)&d, sizeof(double) * blocks * threads));
  cudaStream_t st = ctx->stream;
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep)
    {
      CUDA_OK(ctx, cudaEventRecord(ctx->ev[8], st));
      k_dfma_peak<<<blocks, threads, 0, st>>>(d, iters);
      ctx->launches++;
      CUDA_OK(ctx, cudaEventRecord(ctx->ev[9], st));
      CUDA_OK(ctx, cudaStreamSynchronize(st));
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]);
      if (rep >= 1 && ms < best) best = ms;
    }
  cudaFree(d);
  *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
  // the hand-scheduled chain loop (inline PTX, 16 warps per SM) runs a little closer to the
  // pipe's 64 FMA/clk/SM: report the better of the two as the roofline denominator
  {
    const int pb = 148 * 4, pt = 512, pit = 1 << 14;
    double *d2 = nullptr;
    CUDA_OK(ctx, cudaMalloc((void **)&d2, sizeof(double) * pb * pt));
    float best2 = 1e30f;
    for (int rep = 0; rep < 4; ++rep)
      {
        CUDA_OK(ctx, cudaEventRecord(ctx->ev[8], st));
        k_dfma_chains<<<pb, pt, 0, st>>>(d2, pit);
        ctx->launches++;
        CUDA_OK(ctx, cudaEventRecord(ctx->ev[9], st));
        CUDA_OK(ctx, cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]);
        if (rep >= 1 && ms < best2) best2 = ms;
      }
    cudaFree(d2);
    const double alt = 2.0 * 8.0 * (double)pit * pb * pt / (best2 * 1e-3) / 1e12;
    if (alt > *tflops) *tflops = alt;
  }
  return 0;
}

int wbem_measure_copy_bw(wbem_ctx *ctx, double *gbs)
{
  CHECK_CTX(ctx);
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  const size_t bytes = (size_t)2 << 30;
  char *a = nullptr, *b = nullptr;
  CUDA_OK(ctx, cudaMalloc((void **)&a, bytes));
  CUDA_OK(ctx, cudaMalloc((void **)&b, bytes));
  cudaStream_t st = ctx->stream;
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep)
    {
      CUDA_OK(ctx, cudaEventRecord(ctx->ev[8], st));
      CUDA_OK(ctx, cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, st));
      CUDA_OK(ctx, cudaEventRecord(ctx->ev[9], st));
      CUDA_OK(ctx, cudaStreamSynchronize(st));
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]);
      if (rep >= 1 && ms < best) best = ms;
    }
  cudaFree(a);
  cudaFree(b);
  *gbs = 2.0 * bytes / (best * 1e-3) / 1e9;
  return 0;
}

// one operator application (ConstrainedOperator::vmult) timed alone, `reps` times, with an
// optional L2 flush (a write larger than L2) between repetitions
int wbem_time_operator(wbem_ctx *ctx, int reps, int flush_l2, double *ms_avg, double *bytes)
{
  CHECK_CTX(ctx);
  if (wbem_group_forward(ctx))
    {
      std::vector<double> v(wbem_group_size(ctx), 0.0), b(wbem_group_size(ctx), 0.0);
      const int rc = wbem_group_run(
        ctx, [&](wbem_ctx *s) -> int { return wbem_time_operator(s, reps, flush_l2, &v[s->p.rank], &b[s->p.rank]); });
      *ms_avg = *std::max_element(v.begin(), v.end());
      if (bytes) *bytes = b[0];
      return rc;
    }
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  int rc = ensure_alpha(ctx);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  char *flush = nullptr;
  const size_t fb = (size_t)256 << 20;
  if (flush_l2) CUDA_OK(ctx, cudaMalloc((void **)&flush, fb));
  CUDA_OK(ctx, cudaMemsetAsync(ctx->d_tmp[1], 0, sizeof(double) * ctx->ld, st));
  double total = 0;
  for (int r = 0; r < reps + 1; ++r)
    {
      if (flush) CUDA_OK(ctx, cudaMemsetAsync(flush, r, fb, st));
      CUDA_OK(ctx, cudaEventRecord(ctx->ev[8], st));
      rc = wbem_apply_operator(ctx, 0, ctx->d_tmp[1], ctx->d_tmp[2], true);
      if (rc) return rc;
      CUDA_OK(ctx, cudaEventRecord(ctx->ev[9], st));
      CUDA_OK(ctx, cudaStreamSynchronize(st));
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]);
      if (r > 0) total += ms;
    }
  if (flush) cudaFree(flush);
  *ms_avg = total / reps;
  if (bytes) *bytes = ctx->tm.gemv_bytes_last;
  return 0;
}

int wbem_time_assemble(wbem_ctx *ctx, int reps, double *ms_avg)
{
  CHECK_CTX(ctx);
  if (wbem_group_forward(ctx))
    {
      std::vector<double> v(wbem_group_size(ctx), 0.0);
      const int rc = wbem_group_run(ctx, [&](wbem_ctx *s) -> int { return wbem_time_assemble(s, reps, &v[s->p.rank]); });
      *ms_avg = *std::max_element(v.begin(), v.end());
      return rc;
    }
  double total = 0;
  for (int r = 0; r < reps + 1; ++r)
    {
      int rc = wbem_assemble(ctx);
      if (rc) return rc;
      if (r > 0) total += ctx->tm.assemble_total_ms;
    }
  *ms_avg = total / reps;
  return 0;
}

} // extern "C"
