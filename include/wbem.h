/*
 * wbem.h -- C ABI of the B200-native collocation-BEM hot path (libwbem.so).
 *
 * Drop-in boundary for WaveBEM's BEMProblem<3> (reference include/bem_problem.h:87-180,
 * source/bem_problem.cc).  Every entry point is extern "C", takes plain pointers and sizes,
 * returns an int status: 0 = ok, <0 = fatal (CUDA / NCCL / allocation / bad argument),
 * >0 = solver did not converge (the adapter rethrows it as SolverControl::NoConvergence).
 * wbem_last_error() gives the message.  All host arrays are caller-owned, contiguous,
 * fp64 / uint32 / uint8; the library owns all device memory inside the opaque context.
 * One calling host thread per context (the reference is single-threaded, main.cc:26).
 *
 * There is NO CPU fallback: every compute entry point fails (<0) when no sm_100 device is
 * usable.
 *
 * Multi-GPU, two ways.  Row block r of P owns the contiguous matrix rows
 * [r*ceil(N/P), min(N,(r+1)*ceil(N/P))); vectors, geometry, masks and constraints are replicated.
 *  (1) single process (what the single-threaded reference needs, source/main.cc:26):
 *      wbem_params.n_gpus = P -- ONE context owns all P row blocks, every call below acts on all
 *      of them, results come back to the caller's host arrays once.  Peer access between the
 *      devices carries the fused mat-vec + gather; no NCCL, no IPC, no launcher.
 *  (2) one process per GPU (torchrun / MPI): wbem_params.rank / world_size; the mat-vecs
 *      all-gather over NCCL (wbem_comm_init) or, after the CUDA-IPC exchange, through the same
 *      fused peer stores.
 */
#ifndef WBEM_H
#define WBEM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wbem_ctx wbem_ctx;

/* Replaces the prm keys the path reads: "Quadrature rules/{Quadrature order, Singular
 * quadrature order}" (source/computational_domain.cc:84-91,123-131), "Solver/{Tolerance,
 * Max steps}" (source/bem_problem.cc:83-85,95-97), the GMRES AdditionalData(100)
 * (source/bem_problem.cc:826-827) and preconditioner_band = 100 (:68). */
typedef struct wbem_params
{
  int quad_order;          /* Gauss n x n, default 4 */
  int sing_order;          /* QGaussOneOverR n -> 2 n^2 points, default 5 */
  double gmres_tol;        /* absolute, on the preconditioned residual; default 1e-16 */
  int gmres_max_steps;     /* default 200 */
  int gmres_n_tmp_vectors; /* default 100 (Krylov basis 98) */
  int preconditioner_band; /* default 100; 0 disables the band preconditioner */
  int device;              /* CUDA device ordinal */
  int rank;                /* this context's row block */
  int world_size;          /* number of row blocks (GPUs) */
  int assemble_variant;    /* 0 = default (single-launch stream kernel, deterministic), 1 = simple atomic kernel,
                            * 2 = colour-per-launch tiled kernel (per-point arithmetic; also the path of wbem_set_fevalues) */
  int precond_on_host;     /* 1 = band LU + solves on the host (debug), 0 = on the device */
  int precond_kind;        /* 0 = the reference's band preconditioner (default); 1 = local-inverse
                              sparse approximate inverse (spai.cu): same solution within the solver
                              tolerance, ~4x fewer GMRES iterations, no dependence on the numbering */
  int auto_constraints;    /* 1 = solve_system / solve / residual call compute_constraints(tmp_rhs)
                              themselves, as the reference does (source/bem_problem.cc:845); 0 (default)
                              = the caller installs the lines with wbem_set_constraints */
  int n_gpus;              /* 0 / 1: this context is ONE row block (rank of world_size, one process per
                              GPU).  > 1: this ONE context owns n_gpus row blocks, block r on CUDA device
                              devices[r]; the caller stays single-threaded (source/main.cc:26) and the
                              library drives the devices with one internal worker thread + stream each.
                              rank / world_size / device are ignored then. */
  int fused_gather_on_shared_device; /* test switch: row blocks that share a device normally gather through
                              stream-ordered copies; 1 = use the fused peer-store mat-vec there as well */
  int reserved;
  int devices[16];         /* n_gpus > 1: device ordinal of every row block; -1 = block r on device r.  A
                              device may be listed more than once (several row blocks on one GPU: how the
                              single-GPU tests exercise the sharded code) */
} wbem_params;

typedef struct wbem_timings
{ /* device milliseconds measured with CUDA events on the library's stream (last call) */
  double geometry_ms;      /* cell quadrature data */
  double assemble_regular_ms;
  double assemble_singular_ms;
  double alpha_ms;
  double assemble_total_ms;
  double rhs_ms;
  double precond_setup_ms;
  double gmres_ms;         /* whole GMRES loop */
  double gemv_ms_sum;      /* sum over operator applications in the last solve */
  double precond_apply_ms_sum;
  double allgather_ms_sum;
  double solve_system_total_ms;
  int gemv_calls;
  int gmres_iters;
  long long kernel_launches; /* kernels of this library launched since create/reset */
  double gemv_bytes_last;  /* matrix bytes the last operator application had to stream */
  double constraints_ms;   /* compute_constraints inside solve_system (auto_constraints = 1) */
  double reserved[5];
} wbem_timings;

void wbem_default_params(wbem_params *p);
int wbem_create(const wbem_params *p, wbem_ctx **out);
int wbem_destroy(wbem_ctx *ctx);
const char *wbem_last_error(const wbem_ctx *ctx); /* ctx may be NULL: last create error */
int wbem_version(void);

/* BEMProblem<3>::reinit (source/bem_problem.cc:55-71) + the once-per-mesh flattening of
 * comp_dom.dh / comp_dom.double_nodes_set (source/bem_problem.cc:192-196, 223-230).
 * cell_dofs[C][4]: deal.II lexicographic vertex order; cell_dir_flag[C]:
 * cell->direction_flag(); dn_ptr[N+1], dn_idx[]: CSR of double_nodes_set (set i holds i). */
int wbem_set_topology(wbem_ctx *ctx, uint32_t n_dofs, uint32_t n_cells,
                      const uint32_t *cell_dofs, const uint8_t *cell_dir_flag,
                      const uint32_t *dn_ptr, const uint32_t *dn_idx);

/* Host-only helper for that flattening: ComputationalDomain<3>::generate_double_nodes_set
 * (source/computational_domain.cc:258-307, tol = 1e-8 at :271) with a uniform-grid search instead of
 * the reference's O(N_boundary N) loop.  boundary_dofs[N] (NULL = every dof is tested) is
 * DoFTools::extract_boundary_dofs; dn_ptr[N+1] is always filled; dn_idx receives the sorted sets when
 * it holds `capacity` entries.  Returns 0, or 1 when dn_idx is NULL / too small (*needed = entries). */
int wbem_generate_double_nodes_set(uint32_t n_dofs, const double *support_points,
                                   const uint8_t *boundary_dofs, double tol, uint32_t *dn_ptr,
                                   uint32_t *dn_idx, uint64_t capacity, uint64_t *needed);

/* Per assembly: support_points[N][3] = DoFTools::map_dofs_to_support_points
 * (source/bem_problem.cc:168-169).  The Q1 mapping's quadrature points, normals and JxW
 * (the FEValues of :133-137, 192-196, 495-501) are recomputed on the device from them. */
int wbem_set_geometry(wbem_ctx *ctx, const double *support_points);
int wbem_set_geometry_dev(wbem_ctx *ctx, const double *d_support_points);
/* Optional, after wbem_set_geometry: the literal FEValues output of the regular rule
 * (source/bem_problem.cc:192-196) -- q_points[C][nq][3], normals[C][nq][3], JxW[C][nq] in the
 * caller's cell order -- used for the regular pairs instead of the values recomputed from the
 * support points (bitwise the reference's inputs, e.g. for a mapping that is not Q1).  Singular
 * pairs keep using the Q1 mapping of the support points.  The next wbem_set_geometry drops them. */
int wbem_set_fevalues(wbem_ctx *ctx, const double *q_points, const double *normals, const double *JxW);

/* BEMProblem<3>::assemble_system (source/bem_problem.cc:106-590): both matrices for this
 * context's rows, plus alpha (compute_alpha, :594-618) fused behind it. */
int wbem_assemble(wbem_ctx *ctx);
/* BEMProblem<3>::compute_alpha alone (:594-618). */
int wbem_compute_alpha(wbem_ctx *ctx);
int wbem_get_alpha(wbem_ctx *ctx, double *alpha /* [N] */);

/* Test/preconditioner access to matrix entries (the reference exposes neumann_matrix and
 * dirichlet_matrix as public members, include/bem_problem.h:153-154).  which: 0 = Neumann
 * (double layer), 1 = Dirichlet (single layer).  Rows [r0,r1) must lie inside this
 * context's row block; out is (r1-r0) x N row-major in global dof numbering. */
int wbem_get_rows(wbem_ctx *ctx, int which, uint32_t r0, uint32_t r1, double *out);
int wbem_row_block(const wbem_ctx *ctx, uint32_t *row0, uint32_t *row1);

/* comp_dom.surface_nodes / other_nodes (source/numerical_towing_tank.cc:1799-1807); callers
 * mutate them between solves (source/free_surface.cc:642-675, 6048-6049). */
int wbem_set_masks(wbem_ctx *ctx, const double *surface_nodes, const double *other_nodes);

/* ConstraintMatrix built by compute_constraints (source/bem_problem.cc:990-1105), flattened:
 * line k constrains dof lines[k] = sum_j val[j] x[col[j]] (j in ptr[k]..ptr[k+1]) + inhom[k]. */
int wbem_set_constraints(wbem_ctx *ctx, uint32_t n_lines, const uint32_t *lines,
                         const uint32_t *ptr, const uint32_t *col, const double *val,
                         const double *inhom);

/* ComputationalDomain<3>::compute_normals (source/computational_domain.cc:1525-1620): L2
 * projection of the cell normals onto the Q1 space, normalised; normals[N][3] (may be NULL). */
int wbem_compute_normals(wbem_ctx *ctx, double *normals);
/* BEMProblem<3>::compute_surface_gradients (source/bem_problem.cc:1153-1293): L2 projection of
 * the surface gradient of phi = tmp_rhs o surface_nodes; gradients[N][3] (may be NULL). */
int wbem_compute_surface_gradients(wbem_ctx *ctx, const double *tmp_rhs, double *gradients);
/* BEMProblem<3>::compute_constraints (source/bem_problem.cc:990-1105) inside the library: runs the
 * two projections above when a double-node set needs them, walks the double-node sets and
 * installs the lines (like wbem_set_constraints).  Hanging-node lines
 * (DoFTools::make_hanging_node_constraints, :1000) stay the caller's: hand them over once with
 * wbem_set_hanging_constraints (weights only, no inhomogeneities).  wbem_get_constraints returns
 * the lines of the last call (sizes first, then the arrays; any pointer may be NULL). */
int wbem_set_hanging_constraints(wbem_ctx *ctx, uint32_t n_lines, const uint32_t *lines,
                                 const uint32_t *ptr, const uint32_t *col, const double *val);
int wbem_compute_constraints(wbem_ctx *ctx, const double *tmp_rhs);
int wbem_get_constraints(wbem_ctx *ctx, uint32_t *n_lines, uint32_t *nnz, uint32_t *lines,
                         uint32_t *ptr, uint32_t *col, double *val, double *inhom);
int wbem_mass_cg_iterations(wbem_ctx *ctx); /* CG iterations of the last projection solve */

/* BEMProblem<3>::vmult (source/bem_problem.cc:620-670). */
int wbem_vmult(wbem_ctx *ctx, double *dst, const double *src);
/* ConstrainedOperator::vmult / distribute_rhs (include/constrained_matrix.h:73-94). */
int wbem_constrained_vmult(wbem_ctx *ctx, double *dst, const double *src);
int wbem_distribute_rhs(wbem_ctx *ctx, double *rhs);
/* ConstrainedOperator::vmult on nvec <= 8 vectors (dst, src: [nvec][N]) with ONE pass over the matrices:
 * the block mat-vec wbem_solve_system_multi is built on. */
int wbem_constrained_vmult_multi(wbem_ctx *ctx, int nvec, double *dst, const double *src);
int wbem_compute_rhs_multi(wbem_ctx *ctx, int nvec, double *dst, const double *src); /* compute_rhs, same way */
/* BEMProblem<3>::compute_rhs (source/bem_problem.cc:673-707). */
int wbem_compute_rhs(wbem_ctx *ctx, double *dst, const double *src);
/* BEMProblem<3>::assemble_preconditioner (source/bem_problem.cc:1107-1149). */
int wbem_assemble_preconditioner(wbem_ctx *ctx);
int wbem_precond_vmult(wbem_ctx *ctx, double *dst, const double *src);
/* the band_system entries of this context's rows: out[(r-row0)*band + (i-(r-band/2+1))] */
int wbem_get_band(wbem_ctx *ctx, double *out);
/* switch wbem_params.precond_kind on a live context (the next solve rebuilds the preconditioner) */
int wbem_set_precond_kind(wbem_ctx *ctx, int kind);
/* precond_kind = 1: the assembled sparse approximate inverse M (row i = k entries at columns
 * nbr[i*k .. i*k+k), 0xffffffff = unused slot), and the number of local systems that were
 * singular (those rows fall back to 1/a_ii).  Any output pointer may be NULL. */
int wbem_get_spai(wbem_ctx *ctx, uint32_t *k, uint32_t *nbr, double *val, int *n_singular);
/* host-only check of its sparsity-pattern builder (no GPU): 0 = all invariants hold;
 * stats4 = k, widest / mean near-field row, rows with fewer than k dofs in reach */
int wbem_spai_pattern_check(uint32_t n_dofs, uint32_t n_cells, const uint32_t *cell_dofs,
                            const uint32_t *dn_ptr, const uint32_t *dn_idx, const double *xyz,
                            uint32_t *nbr_out, double *stats4);

/* BEMProblem<3>::solve_system (source/bem_problem.cc:821-895).  phi / dphi_dn are in-out:
 * only the unknown half is overwritten (:869-879).  iters / last_res may be NULL.
 * Returns 1 when GMRES hits max_steps (SolverControl::NoConvergence). */
int wbem_solve_system(wbem_ctx *ctx, double *phi, double *dphi_dn, const double *tmp_rhs,
                      int *iters, double *last_res);
/* nrhs calls of solve_system on the SAME matrices in one: FreeSurface::jacobian runs one inner GMRES per
 * Jacobian-vector product (source/free_surface.cc:4918-4993, source/dae_time_integrator.cc:375-377) with only
 * tmp_rhs changing.  phi / dphi_dn / tmp_rhs are [nrhs][N] row-major, iters / last_res [nrhs] (may be NULL).
 * Every system is the same left-preconditioned GMRES with its own Krylov space and stopping test; up to 8
 * of them share each pass over the matrices (block mat-vec), so the bytes streamed per solve drop by up to
 * 8x.  With auto_constraints = 1 the constraint inhomogeneities follow each system's tmp_rhs; with
 * caller-installed lines (wbem_set_constraints) all systems share them.  Returns 1 if any system hit max_steps. */
int wbem_solve_system_multi(wbem_ctx *ctx, int nrhs, double *phi, double *dphi_dn, const double *tmp_rhs,
                            int *iters, double *last_res);
/* solver.solve(cc, sol, system_rhs, preconditioner) alone (source/bem_problem.cc:853): GMRES on the
 * constrained operator for a right-hand side the caller prepared (compute_rhs + distribute_rhs),
 * x0 = 0, with the preconditioner wbem_params selects.  Same return convention as solve_system. */
int wbem_gmres(wbem_ctx *ctx, const double *rhs, double *sol, int *iters, double *last_res);
/* BEMProblem<3>::solve (source/bem_problem.cc:969-987) = assemble_system + solve_system,
 * with the geometry upload in front (the caller moved the mesh,
 * source/free_surface.cc:5306-5307). */
int wbem_solve(wbem_ctx *ctx, const double *support_points, double *phi, double *dphi_dn,
               const double *tmp_rhs, int *iters, double *last_res);
/* Same with every array already in device memory (bench `value` leg). */
int wbem_solve_dev(wbem_ctx *ctx, const double *d_support_points, double *d_phi,
                   double *d_dphi_dn, const double *d_tmp_rhs, int *iters, double *last_res);
int wbem_solve_system_dev(wbem_ctx *ctx, double *d_phi, double *d_dphi_dn,
                          const double *d_tmp_rhs, int *iters, double *last_res);
/* BEMProblem<3>::residual (source/bem_problem.cc:903-961). */
int wbem_residual(wbem_ctx *ctx, double *res, const double *phi, const double *dphi_dn);

/* Post-processing integrals that share the assembly's integrand (they read the geometry of the last
 * wbem_set_geometry, no matrix; quad_order = 4 only).
 * FreeSurface<3>::compute_internal_velocities (source/free_surface.cc:10426-10537): grad phi at
 * n_points field points x[n_points][3] from the boundary traces (the reference reads the points from
 * points.txt and writes velocities.txt): velocities[n_points][3]. */
int wbem_internal_velocities(wbem_ctx *ctx, const double *phi, const double *dphi_dn, uint32_t n_points,
                             const double *points, double *velocities);
/* The hull integrals of FreeSurface<3>::compute_pressure (source/free_surface.cc:9534-9598), steady
 * terms, over the cells with cell_marked[c] != 0 (the reference: material_id == wall_sur_ID1..3):
 * out11 = press_force_test_1[3], press_force_test_2[3], press_moment[3] about baricenter (NULL = origin),
 * marked area, integral of phi over the marked cells.  vinf[3] = the wind. */
int wbem_pressure_force(wbem_ctx *ctx, const double *phi, const double *dphi_dn, const uint8_t *cell_marked,
                        const double *vinf, double rho, double g, const double *baricenter, double *out11);

/* system_rhs / sol of the last solve_system (public members, include/bem_problem.h:155-158) */
int wbem_get_system_rhs(wbem_ctx *ctx, double *out);
int wbem_get_sol(wbem_ctx *ctx, double *out);

int wbem_get_timings(wbem_ctx *ctx, wbem_timings *out);
/* CUDA-event stopwatch on the library's own stream (torch.cuda.Event only sees torch's). */
int wbem_timer_start(wbem_ctx *ctx);
int wbem_timer_stop(wbem_ctx *ctx, double *ms);
int wbem_reset_counters(wbem_ctx *ctx);

/* Row-sharded runs: rank 0 makes an id (128 bytes), the launcher (torch.distributed, MPI,
 * ...) broadcasts it, every rank calls wbem_comm_init. */
int wbem_comm_unique_id(void *id128);
int wbem_comm_init(wbem_ctx *ctx, const void *id128);
/* Optional, after every wbem_set_topology: CUDA-IPC exchange of the gather buffers.  Each rank
 * exports 64 bytes, the launcher all-gathers them, each rank imports world_size*64 bytes.  With
 * it the mat-vec kernel stores its result rows straight into every rank's gather buffer over
 * NVLink (one fused kernel); without it each mat-vec is followed by ncclAllGather. */
int wbem_comm_ipc_export(wbem_ctx *ctx, void *handle64);
int wbem_comm_ipc_import(wbem_ctx *ctx, const void *handles);
int wbem_comm_ipc_close(wbem_ctx *ctx); /* back to the ncclAllGather path */

/* Diagnostics used by bench.py: measured FP64 FMA peak (TFLOP/s) and copy bandwidth (GB/s)
 * of this device; one operator application timed alone. */
int wbem_measure_fp64_peak(wbem_ctx *ctx, double *tflops);
int wbem_measure_copy_bw(wbem_ctx *ctx, double *gbs);
int wbem_time_operator(wbem_ctx *ctx, int reps, int flush_l2, double *ms_avg, double *bytes);
int wbem_time_assemble(wbem_ctx *ctx, int reps, double *ms_avg);
/* device self test of the fast 1/sqrt used by the regular-pair kernel: out[i] = rsqrt(in[i]) */
int wbem_selftest_rsqrt(wbem_ctx *ctx, const double *in, double *out, int n);
/* host-only check of the assembly tiling plan (no GPU): 0 = all invariants hold; stats9 = clusters, colours,
 * max cells, max slots, slots, ADD slots, columns written, ADD sectors, longest predecessor list */
int wbem_plan_check(uint32_t n_dofs, uint32_t n_cells, const uint32_t *cell_dofs, uint32_t w_max,
                    uint32_t max_cells, double *stats9);

/* host-only check of the item order of the single-launch assembly kernel (no GPU): every (row tile, cluster)
 * is drawn exactly once and after the clusters it has to wait for; stats4 = items, smallest / mean ticket
 * distance to a predecessor, row tiles.  0 = holds; 101 = this mesh takes the colour-per-launch kernel */
int wbem_stream_order_check(uint32_t n_dofs, uint32_t n_cells, const uint32_t *cell_dofs, uint32_t n_rows,
                            uint32_t group_tiles, double *stats4);

#ifdef __cplusplus
}
#endif
#endif /* WBEM_H */
