// internal.h -- shared declarations of libwbem (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/wbem.h"

#define WBEM_MAX_NQ 64   // regular rule: up to 8 x 8
#define WBEM_MAX_NS 288  // singular rule: up to 2 * 12^2
#define WBEM_MAX_PEERS 16
#define WBEM_GMRES_KMAX 1024 // largest gmres_n_tmp_vectors
#define WBEM_MULTI_MAX 8     // right-hand sides one block mat-vec streams the matrices for (wbem_solve_system_multi)

struct QuadTables
{ // host copies of the reference-cell tables (uploaded to __constant__ memory)
  int nq, ns;
  double g_u[WBEM_MAX_NQ], g_v[WBEM_MAX_NQ], g_w[WBEM_MAX_NQ];
  double g_shape[4][WBEM_MAX_NQ];
  double s_u[4][WBEM_MAX_NS], s_v[4][WBEM_MAX_NS], s_w[4][WBEM_MAX_NS];
  double g1_x[8], g1_w[8]; // 1-D rule (tensor structure of the regular rule)
  int n1;
};
int wbem_build_quadrature(int quad_order, int sing_order, QuadTables *qt);

struct NcclApi; // comm.cpp
struct WbemGroup; // group.cpp: one context driving several row blocks (GPUs) from one host thread

// slot_col entries: storage column | flags
#define WBEM_SLOT_ADD 0x80000000u  // an earlier cluster has written the column: add
#define WBEM_SLOT_COL_MASK 0x7fffffffu

// Tiling plan of the regular-pair kernel, built once per topology (plan.cpp).
struct AssemblyPlan
{
  uint32_t n_clusters = 0, n_colors = 0;
  uint32_t W = 0;                        // max column slots per cluster
  uint32_t n_cols_written = 0;           // storage columns [0,n) are written by the kernel
  uint32_t max_cells = 0;                // largest cluster (cells)
  std::vector<uint32_t> cl_cell_ptr;     // [ncl+1] range in cell processing order
  std::vector<uint32_t> cell_order;      // [C] processing position -> cell id
  std::vector<uint32_t> cell_pos;        // [C] cell id -> processing position
  std::vector<uint8_t> cell_slots;       // [C*4] in processing order: slot of local dof j
  std::vector<uint32_t> cl_slot_ptr;     // [ncl+1]
  std::vector<uint32_t> slot_col;        // storage column of each slot | (add flag << 31)
  std::vector<uint32_t> color_ptr;       // [ncolors+1] range in cluster launch order
  std::vector<uint32_t> color_clusters;  // cluster ids grouped by color
  std::vector<uint32_t> colpos;          // [N] global dof -> storage column
  std::vector<uint32_t> colperm;         // [N] storage column -> global dof
  // dependencies of the single-launch kernel: for the cluster at launch position k, the launch
  // positions (< k, lower colours) of the clusters it shares a dof with -- the ones whose flush
  // has to land before this cluster's (STORE before ADD, ADDs in colour order)
  std::vector<uint32_t> pred_ptr;        // [ncl+1] by launch position
  std::vector<uint32_t> pred;
  uint32_t max_pred = 0;
};
int wbem_build_plan(uint32_t N, uint32_t C, const uint32_t *cell_dofs, uint32_t W_max,
                    uint32_t max_cells_per_cluster, AssemblyPlan *plan);

struct EvTimer
{ // accumulates device time of (possibly nested) bracketed regions on one stream
  std::vector<cudaEvent_t> ev;
  struct Region { int tag; size_t e0, e1; };
  std::vector<Region> regions;
  std::vector<size_t> open; // stack of region ids
  size_t used = 0;
  cudaStream_t st = nullptr;
  bool on = false;
  size_t take()
  {
    if (used + 1 > ev.size())
      {
        const size_t old = ev.size();
        ev.resize(old + 256);
        for (size_t i = old; i < ev.size(); ++i) cudaEventCreate(&ev[i]);
      }
    return used++;
  }
  void begin(int t)
  {
    if (!on) return;
    const size_t e = take();
    regions.push_back({t, e, e});
    open.push_back(regions.size() - 1);
    cudaEventRecord(ev[e], st);
  }
  void end()
  {
    if (!on || open.empty()) return;
    const size_t e = take();
    regions[open.back()].e1 = e;
    open.pop_back();
    cudaEventRecord(ev[e], st);
  }
  void reset()
  {
    used = 0;
    regions.clear();
    open.clear();
  }
  void resolve(double *sums, int *counts, int ntags)
  {
    for (int i = 0; i < ntags; ++i) sums[i] = 0, counts[i] = 0;
    for (const Region &r : regions)
      {
        if (r.e1 == r.e0) continue;
        float ms = 0;
        cudaEventElapsedTime(&ms, ev[r.e0], ev[r.e1]);
        sums[r.tag] += ms;
        counts[r.tag]++;
      }
    reset();
  }
  void release()
  {
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear();
  }
};
enum { T_RHS = 0, T_PRECOND_SETUP, T_GMRES, T_ALPHA, T_GEMV, T_PRECOND_APPLY, T_ALLGATHER, T_CONSTRAINTS, T_NTAGS };

struct wbem_ctx
{
  wbem_params p;
  EvTimer timer;
  std::string err;
  int dev = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[16] = {};
  QuadTables qt;
  void *d_tables = nullptr; // DevTables (assemble.cu) in global memory: per context, not __constant__
  void *d_gauss = nullptr;  // GaussTable (constraints.cu)
  // per-device launch state (a context on another GPU needs its own opt-ins / occupancy)
  bool tiled_attr_set = false, bcr_attr_done = false;
  int gemv_ctas_per_sm = 0, gemv_multi_ctas_per_sm[3] = {0, 0, 0}, n_sm = 0;

  // sizes
  uint32_t N = 0, C = 0, ld = 0;
  uint32_t chunk = 0, row0 = 0, row1 = 0, nloc = 0;

  // host topology
  std::vector<uint32_t> h_cell_dofs, h_dn_ptr, h_dn_idx;
  std::vector<uint8_t> h_dir;
  AssemblyPlan plan;

  // device topology
  uint32_t *d_cell_dofs = nullptr; // [C][4] in processing order (storage columns NOT applied)
  uint8_t *d_dir = nullptr;        // processing order
  uint32_t *d_cell_order = nullptr;
  uint32_t *d_colpos = nullptr, *d_colperm = nullptr;
  // singular pairs of the local rows: CSR by local row
  uint32_t *d_sing_ptr = nullptr, *d_sing_cellpos = nullptr; // cell processing position, sorted
  uint8_t *d_sing_idx = nullptr;
  uint8_t *d_tile_sing = nullptr; // [row tiles][clusters]
  uint32_t n_sing = 0;
  // plan on device
  uint32_t *d_slot_col = nullptr;
  uint32_t *d_cta_desc = nullptr; // [clusters][4] in launch order: first cell, first slot, counts, cluster id
  uint8_t *d_cell_slots = nullptr;
  // single-launch (stream) kernel: per-cluster records, look-ahead descriptors, ticket + error flag + done flags
  void *d_item_meta = nullptr;
  void *d_item_desc = nullptr;
  unsigned int *d_asm_sync = nullptr; // [0] ticket, [1] error flag, [4...] done[row tiles][clusters]
  bool stream_ok = false;             // the plan fits the records (predecessor lists, cluster sizes)
  unsigned int asm_epoch = 0;
  unsigned int *h_asm_flag = nullptr; // pinned copy of the error flag, read after the next synchronisation
  int stream_ctas_per_sm = 0;
  int asm_group_tiles = 2; // G: row tiles interleaved colour by colour (environment WBEM_ASM_GROUP overrides)

  // geometry
  double *d_xyz = nullptr;     // [N][3]
  double *d_cellgeo = nullptr; // [C][8][nq] in processing order
  double *d_cellgeo2 = nullptr; // [C][GEO2_REC] line records of the stream kernel (4 x 4 rule only)
  bool have_geometry = false, assembled = false, have_alpha = false;
  bool fevalues_given = false; // cellgeo was filled by wbem_set_fevalues, not from the support points
  bool has_degenerate_cells = false; // a cell lists the same dof twice -> simple kernel

  // matrices: nloc x ld, row-major, columns in storage order (colperm)
  double *d_Nm = nullptr, *d_Dm = nullptr;
  double *d_alpha = nullptr; // [N] replicated
  double *d_alpha_part = nullptr; // [n_clusters + 1][nloc] partial row sums written by the assembly
  bool alpha_parts_valid = false;

  // masks / constraints (replicated)
  double *d_surf = nullptr, *d_other = nullptr;
  bool have_masks = false, pure_neumann = false;
  std::vector<double> h_surf, h_other;
  uint32_t n_lines = 0;
  int32_t *d_con_line_of = nullptr;
  uint32_t *d_con_lines = nullptr, *d_con_ptr = nullptr, *d_con_col = nullptr;
  double *d_con_val = nullptr, *d_con_inhom = nullptr;
  std::vector<int32_t> h_con_line_of;
  uint32_t *d_free_rows = nullptr; // local rows without a constraint line
  uint32_t n_free_rows = 0;
  uint64_t op_version = 0, precond_version = ~0ull; // preconditioner cache key

  // work vectors (device, length N or chunk*world)
  double *d_xn = nullptr, *d_xd = nullptr; // permuted multipliers of N and D
  double *d_yloc = nullptr;                // [chunk*world] gather buffer
  double *d_xdiag = nullptr;               // [N] multiplier of N in global order
  uint32_t *d_list_o = nullptr, *d_list_s = nullptr; // 64-column chunks with non-zero mask
  int n_list_o = 0, n_list_s = 0;
  double *d_tmp[8] = {};
  double *d_rhs = nullptr, *d_sol = nullptr;
  double *d_V = nullptr;   // Krylov basis [n_tmp-1][N]
  double *d_h = nullptr;   // small scalars
  double *h_pinned = nullptr; // pinned staging, >= 4N doubles
  size_t pinned_doubles = 0;

  // preconditioner
  double *d_band = nullptr;     // [N][band] band_system rows (all ranks' rows after gather)
  std::vector<double> h_lu;     // host LU (precond_on_host)
  std::vector<int> h_piv;
  int h_kl = 0, h_ku = 0, h_ldab = 0;
  bool precond_ready = false;
  bool precond_host_active = false; // this factorisation lives on the host (precond_on_host, or the device fallback)
  void *dev_precond = nullptr;  // precond.cu state
  void *spai = nullptr;         // spai.cu state (precond_kind = 1)
  void *con = nullptr;          // constraints.cu state (compute_constraints on the device)
  uint64_t geom_version = 0;    // bumped by every wbem_set_geometry
  std::vector<uint32_t> h_con_lines, h_con_ptr, h_con_col; // last installed constraint structure
  std::vector<double> h_con_val;

  // comm
  WbemGroup *group = nullptr; // wbem_params.n_gpus > 1: this shard belongs to a single-process group
  unsigned int *d_gather_timeout = nullptr; // set by k_epilogue when a peer's flag never arrived
  NcclApi *nccl = nullptr;
  void *nccl_comm = nullptr;
  // peer-to-peer gather fused into k_bem_gemv: every rank's gather buffer mapped through
  // CUDA IPC.  d_p2p = [2][WBEM_MULTI_MAX][chunk*P] doubles (double-buffered by epoch; one slab per
  // right-hand side of a block mat-vec) + WBEM_MAX_PEERS flags
  double *d_p2p = nullptr;
  double *peer_base[WBEM_MAX_PEERS] = {};
  bool peer_opened[WBEM_MAX_PEERS] = {};
  bool p2p_ready = false;
  unsigned long long p2p_epoch = 0, gemv_done_total = 0;
  unsigned long long *d_done_counter = nullptr;

  // GMRES iterations enqueued ahead of the host (gmres.cu): device-side Hessenberg / Givens state, the
  // normalisation factor of the newest Krylov vector and the stop flag the iteration kernels look at
  double *d_gm = nullptr;     // [0] 1/h[dim] (scale of the newest vector), [1] rho; then gamma, ci, si, Hs columns
  int *d_gm_ctl = nullptr;    // [0] stop state (0 run, 1 converged, 2 failed), [1] inner iterations done in this cycle
  size_t gm_doubles = 0;
  const double *op_scale_ptr = nullptr; // set around wbem_apply_operator_ex by the GMRES loop
  const int *op_stop_ptr = nullptr;
  void *multi = nullptr; // gmres.cu: work vectors of wbem_solve_system_multi (allocated at first use)
  double *d_ymulti = nullptr; // [WBEM_MULTI_MAX][chunk*world] gather buffers of the block mat-vec (no peer stores)
  wbem_timings tm = {};
  long long launches = 0;
};

#define WBEM_FAIL(ctx, code, ...)                                   \
  do                                                                \
    {                                                               \
      char _b[512];                                                 \
      snprintf(_b, sizeof(_b), __VA_ARGS__);                        \
      (ctx)->err = _b;                                              \
      return (code);                                                \
    }                                                               \
  while (0)

#define CUDA_OK(ctx, call)                                                            \
  do                                                                                  \
    {                                                                                 \
      cudaError_t _e = (call);                                                        \
      if (_e != cudaSuccess)                                                          \
        WBEM_FAIL(ctx, -2, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e),     \
                  __FILE__, __LINE__, #call);                                         \
    }                                                                                 \
  while (0)

// assemble.cu
int wbem_upload_tables(wbem_ctx *ctx);
int wbem_launch_geometry(wbem_ctx *ctx);
int wbem_upload_fevalues(wbem_ctx *ctx, const double *q_points, const double *normals, const double *JxW);
int wbem_launch_assemble(wbem_ctx *ctx);
int wbem_upload_stream_tables(wbem_ctx *ctx, const std::vector<uint32_t> &sing_ptr, const std::vector<uint32_t> &sing_pos,
                              const std::vector<uint32_t> &cluster_of_pos);
int wbem_launch_alpha(wbem_ctx *ctx, bool from_matrix = false);
uint32_t wbem_tile_rows(void);
uint32_t wbem_tile_width(void);
uint32_t wbem_line_record_doubles(void);
// operator.cu
int wbem_apply_operator(wbem_ctx *ctx, int mode /*0 vmult, 1 rhs*/, const double *d_src,
                        double *d_dst, bool constrained);
int wbem_apply_operator_ex(wbem_ctx *ctx, int mode, const double *d_src, double *d_dst, bool constrained, double src_scale,
                           double *src_scale_rw, bool with_precond);
struct EpilogueArgs
{ // what k_epilogue needs to turn the gathered mat-vec rows into ConstrainedOperator::vmult's result
  const double *y, *src;
  const int32_t *line_of;
  const uint32_t *cptr, *ccol;
  const double *cval;
  const unsigned long long *flags; // fused gather: one arrival flag per rank (null: nothing to wait for)
  int n_peers;
  unsigned long long epoch;
  unsigned int *timeout_flag;
  const int *stop; // device flag of the GMRES loop: non-zero = this iteration was enqueued ahead of a finished solve, do nothing
};
int wbem_spai_apply_fused(wbem_ctx *ctx, const EpilogueArgs &ea, double *d_out);
int wbem_apply_operator_multi(wbem_ctx *ctx, int mode, int nb, const double *const *d_src, double *const *d_dst,
                              bool constrained);
int wbem_allgather_rows(wbem_ctx *ctx, double *d_buf /* [chunk*world], own block filled */);
int wbem_allgather_bytes(wbem_ctx *ctx, void *d_buf, size_t bytes_per_rank);
int wbem_check_gather_timeout(wbem_ctx *ctx);
// gmres.cu
int wbem_build_preconditioner(wbem_ctx *ctx);
int wbem_apply_preconditioner(wbem_ctx *ctx, const double *d_in, double *d_out);
int wbem_solve_system_device(wbem_ctx *ctx, double *d_phi, double *d_dphi_dn,
                             const double *d_bc, int *iters, double *last_res);
int wbem_solve_system_multi_device(wbem_ctx *ctx, int nrhs, double *d_phi, double *d_dphi_dn, const double *d_bc,
                                   int *iters, double *last_res);
void wbem_multi_free(wbem_ctx *ctx);
struct MultiWork
{ // multiplier slabs of the block mat-vec, [WBEM_MULTI_MAX][ld] each (owned by gmres.cu's block state)
  double *d_xn, *d_xd, *d_xdiag;
};
MultiWork *wbem_multi_work(wbem_ctx *ctx);
double *wbem_multi_io(wbem_ctx *ctx, size_t doubles);
// precond.cu
int wbem_device_precond_factor(wbem_ctx *ctx);
int wbem_device_precond_solve(wbem_ctx *ctx, const double *d_in, double *d_out);
void wbem_device_precond_free(wbem_ctx *ctx);
// spai.cu
int wbem_spai_setup(wbem_ctx *ctx);
int wbem_spai_apply(wbem_ctx *ctx, const double *d_in, double *d_out);
void wbem_spai_free(wbem_ctx *ctx);
// constraints.cu
int wbem_constraints_upload_tables(wbem_ctx *ctx);
int wbem_compute_constraints_device(wbem_ctx *ctx, const double *d_tmp_rhs);
void wbem_constraints_free(wbem_ctx *ctx);
// api.cu
void wbem_p2p_close(wbem_ctx *ctx);
// group.cpp -- single-process multi-GPU: the caller's one host thread hands every collective call
// to one worker thread per row block; the workers run the same per-rank code as the
// one-process-per-GPU mode and meet in wbem_group_barrier / wbem_group_allgather
bool wbem_group_forward(const wbem_ctx *ctx);             // true: the call must be handed to the workers
bool wbem_is_root(const wbem_ctx *ctx);                   // writes the caller's host / device outputs
int wbem_group_create(const wbem_params *p, wbem_ctx **out, std::string *err);
int wbem_group_destroy(wbem_ctx *leader);
int wbem_group_size(const wbem_ctx *ctx);
wbem_ctx *wbem_group_shard(const wbem_ctx *ctx, int r);
int wbem_group_run_impl(wbem_ctx *leader, int (*fn)(wbem_ctx *, void *), void *arg);
int wbem_group_barrier(wbem_ctx *ctx);
int wbem_group_allgather(wbem_ctx *ctx, void *d_buf, size_t bytes_per_rank);
int wbem_group_p2p_setup(wbem_ctx *ctx);
template <typename F>
int wbem_group_run(wbem_ctx *leader, F f)
{ // f: int(wbem_ctx *shard), run on every shard's worker thread; first error wins
  return wbem_group_run_impl(
    leader, [](wbem_ctx *s, void *a) -> int { return (*reinterpret_cast<F *>(a))(s); }, &f);
}
#define GROUP_FORWARD(ctx, expr)                                   \
  if (wbem_group_forward(ctx))                                     \
  return wbem_group_run(ctx, [&](wbem_ctx *s) -> int { return (expr); })
// api.cu
int wbem_create_single(const wbem_params *p, wbem_ctx **out, std::string *err);
int wbem_destroy_single(wbem_ctx *ctx);
// comm.cpp
int wbem_nccl_unique_id(void *id128, std::string *err);
int wbem_nccl_init(wbem_ctx *ctx, const void *id128);
void wbem_nccl_destroy(wbem_ctx *ctx);
int wbem_nccl_allgather(wbem_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank);
