// bem_problem_b200.h -- header-only C++ host side above the C ABI (include/wbem.h).
//
// Mirrors the public interface of the reference's BEMProblem<3>
// (include/bem_problem.h:87-180): reinit, assemble_system, compute_alpha, vmult, compute_rhs,
// compute_constraints (taken as flattened lines), assemble_preconditioner, solve_system,
// solve, residual -- same names, argument meaning and error behaviour (a GMRES that reaches
// "Max steps" throws NoConvergence like deal.II's SolverControl; every other failure throws
// std::runtime_error with wbem_last_error()).
//
// deal.II is not available in this image, so the role of ComputationalDomain<3> is played by
// the plain struct FlatDomain below; INTEGRATION.md shows the 40-line adapter that fills it
// from comp_dom.dh / mapping / double_nodes_set inside the real WaveBEM.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/wbem.h"

namespace wbem
{
struct NoConvergence : public std::runtime_error
{ // SolverControl::NoConvergence
  unsigned int last_step;
  double last_residual;
  NoConvergence(unsigned int s, double r)
    : std::runtime_error("GMRES did not converge"), last_step(s), last_residual(r)
  {}
};

struct FlatDomain
{ // what BEMProblem reads from ComputationalDomain<3> (source/bem_problem.cc:58, 117-120, 133,
  // 169, 225, 643-644): dofs, cells, mapping (as support points), double nodes, masks
  std::vector<double> support_points;     // [N][3]   DoFTools::map_dofs_to_support_points
  std::vector<uint32_t> cell_dofs;        // [C][4]   cell->get_dof_indices, deal.II vertex order
  std::vector<uint8_t> cell_direction;    // [C]      cell->direction_flag()
  std::vector<uint32_t> dn_ptr, dn_idx;   // CSR      comp_dom.double_nodes_set
  std::vector<double> surface_nodes, other_nodes; // [N]
  unsigned int n_dofs() const { return (unsigned int)(support_points.size() / 3); }
  unsigned int n_cells() const { return (unsigned int)(cell_dofs.size() / 4); }
  // ComputationalDomain<3>::generate_double_nodes_set (source/computational_domain.cc:258-307) from the
  // support points; boundary_dofs = DoFTools::extract_boundary_dofs (empty: every dof is tested)
  void generate_double_nodes_set(const std::vector<uint8_t> &boundary_dofs = std::vector<uint8_t>(), double tol = 1e-8)
  {
    const uint32_t n = n_dofs();
    const uint8_t *b = boundary_dofs.empty() ? nullptr : boundary_dofs.data();
    dn_ptr.assign(n + 1, 0);
    uint64_t needed = 0;
    if (wbem_generate_double_nodes_set(n, support_points.data(), b, tol, dn_ptr.data(), nullptr, 0, &needed) < 0)
      throw std::runtime_error("wbem_generate_double_nodes_set: bad argument");
    dn_idx.assign(needed, 0);
    if (wbem_generate_double_nodes_set(n, support_points.data(), b, tol, dn_ptr.data(), dn_idx.data(), needed, &needed))
      throw std::runtime_error("wbem_generate_double_nodes_set failed");
  }
};

struct ConstraintLines
{ // flattened ConstraintMatrix (compute_constraints, source/bem_problem.cc:990-1105)
  std::vector<uint32_t> lines, ptr{0}, col;
  std::vector<double> val, inhom;
};

class BEMProblem
{
public:
  explicit BEMProblem(FlatDomain &comp_dom, const wbem_params *params = nullptr) : comp_dom(comp_dom)
  {
    wbem_params p;
    if (params)
      p = *params;
    else
      wbem_default_params(&p);
    if (wbem_create(&p, &ctx) != 0) throw std::runtime_error(std::string("wbem_create: ") + wbem_last_error(nullptr));
  }
  ~BEMProblem()
  {
    if (ctx) wbem_destroy(ctx);
  }
  BEMProblem(const BEMProblem &) = delete;
  BEMProblem &operator=(const BEMProblem &) = delete;

  // source/bem_problem.cc:55-71
  void reinit()
  {
    check(wbem_set_topology(ctx, comp_dom.n_dofs(), comp_dom.n_cells(), comp_dom.cell_dofs.data(),
                            comp_dom.cell_direction.data(), comp_dom.dn_ptr.data(), comp_dom.dn_idx.data()));
    const unsigned int n = comp_dom.n_dofs();
    system_rhs.assign(n, 0.0);
    sol.assign(n, 0.0);
    alpha.assign(n, 0.0);
  }
  // source/bem_problem.cc:106-590
  void assemble_system()
  {
    check(wbem_set_geometry(ctx, comp_dom.support_points.data()));
    check(wbem_assemble(ctx));
    check(wbem_get_alpha(ctx, alpha.data()));
  }
  // source/bem_problem.cc:594-618
  void compute_alpha()
  {
    check(wbem_compute_alpha(ctx));
    check(wbem_get_alpha(ctx, alpha.data()));
  }
  // source/bem_problem.cc:620-670
  void vmult(std::vector<double> &dst, const std::vector<double> &src)
  {
    masks();
    dst.resize(src.size());
    check(wbem_vmult(ctx, dst.data(), src.data()));
  }
  // source/bem_problem.cc:673-707
  void compute_rhs(std::vector<double> &dst, const std::vector<double> &src)
  {
    masks();
    dst.resize(src.size());
    check(wbem_compute_rhs(ctx, dst.data(), src.data()));
  }
  // source/bem_problem.cc:990-1105 inside the library (compute_normals + compute_surface_gradients
  // on the GPU, the walk over the double-node sets on the host); with wbem_params.auto_constraints
  // = 1 solve_system / solve / residual call it themselves, as the reference does (:845, :929)
  void compute_constraints(const std::vector<double> &tmp_rhs)
  {
    masks();
    check(wbem_compute_constraints(ctx, tmp_rhs.data()));
    uint32_t nl = 0, nnz = 0;
    check(wbem_get_constraints(ctx, &nl, &nnz, nullptr, nullptr, nullptr, nullptr, nullptr));
    constraints.lines.resize(nl);
    constraints.ptr.resize(nl + 1);
    constraints.col.resize(nnz);
    constraints.val.resize(nnz);
    constraints.inhom.resize(nl);
    check(wbem_get_constraints(ctx, nullptr, nullptr, constraints.lines.data(), constraints.ptr.data(),
                               constraints.col.data(), constraints.val.data(), constraints.inhom.data()));
  }
  // make_hanging_node_constraints lines stay the caller's (:1000): hand them over once per mesh
  void set_hanging_constraints(const ConstraintLines &c)
  {
    check(wbem_set_hanging_constraints(ctx, (uint32_t)c.lines.size(), c.lines.data(), c.ptr.data(), c.col.data(),
                                       c.val.data()));
  }
  // source/computational_domain.cc:1525-1620 and source/bem_problem.cc:1153-1293
  void compute_normals(std::vector<double> &nodes_normals)
  {
    nodes_normals.resize(3 * (size_t)comp_dom.n_dofs());
    check(wbem_compute_normals(ctx, nodes_normals.data()));
  }
  void compute_surface_gradients(const std::vector<double> &tmp_rhs, std::vector<double> &node_surface_gradients)
  {
    masks();
    node_surface_gradients.resize(3 * (size_t)comp_dom.n_dofs());
    check(wbem_compute_surface_gradients(ctx, tmp_rhs.data(), node_surface_gradients.data()));
  }
  // a ConstraintMatrix produced elsewhere (e.g. by the reference's own host compute_constraints)
  void set_constraints(const ConstraintLines &c)
  {
    check(wbem_set_constraints(ctx, (uint32_t)c.lines.size(), c.lines.data(), c.ptr.data(), c.col.data(),
                               c.val.data(), c.inhom.data()));
  }
  // source/bem_problem.cc:1107-1149
  void assemble_preconditioner()
  {
    masks();
    check(wbem_assemble_preconditioner(ctx));
  }
  // source/bem_problem.cc:821-895
  void solve_system(std::vector<double> &phi, std::vector<double> &dphi_dn, const std::vector<double> &tmp_rhs)
  {
    masks();
    int iters = 0;
    double res = 0;
    const int rc = wbem_solve_system(ctx, phi.data(), dphi_dn.data(), tmp_rhs.data(), &iters, &res);
    finish(rc, iters, res);
  }
  // the J.v pattern of FreeSurface::jacobian (source/free_surface.cc:4918-4993): nrhs solve_system calls on
  // unchanged matrices as one; phi / dphi_dn / tmp_rhs are [nrhs][n_dofs] row-major
  void solve_system_multi(unsigned int nrhs, std::vector<double> &phi, std::vector<double> &dphi_dn,
                          const std::vector<double> &tmp_rhs, std::vector<int> *iterations = nullptr)
  {
    masks();
    std::vector<int> it(nrhs, 0);
    std::vector<double> res(nrhs, 0.0);
    const int rc = wbem_solve_system_multi(ctx, (int)nrhs, phi.data(), dphi_dn.data(), tmp_rhs.data(), it.data(), res.data());
    check(rc);
    if (iterations) *iterations = it;
    if (rc > 0)
      {
        unsigned int worst = 0;
        for (unsigned int b = 1; b < nrhs; ++b)
          if (res[b] > res[worst]) worst = b;
        throw NoConvergence((unsigned int)it[worst], res[worst]);
      }
  }
  // FreeSurface<3>::compute_internal_velocities (source/free_surface.cc:10426-10537): points [n][3] -> velocities [n][3]
  void compute_internal_velocities(const std::vector<double> &phi, const std::vector<double> &dphi_dn,
                                   const std::vector<double> &points, std::vector<double> &velocities)
  {
    velocities.resize(points.size());
    check(wbem_internal_velocities(ctx, phi.data(), dphi_dn.data(), (uint32_t)(points.size() / 3), points.data(),
                                   velocities.data()));
  }
  // the hull integrals of FreeSurface<3>::compute_pressure (source/free_surface.cc:9534-9598): out[11], see wbem.h
  void pressure_force(const std::vector<double> &phi, const std::vector<double> &dphi_dn,
                      const std::vector<uint8_t> &cell_marked, const double vinf[3], double rho, double g,
                      const double baricenter[3], double out11[11])
  {
    check(wbem_pressure_force(ctx, phi.data(), dphi_dn.data(), cell_marked.data(), vinf, rho, g, baricenter, out11));
  }
  // source/bem_problem.cc:969-987
  void solve(std::vector<double> &phi, std::vector<double> &dphi_dn, const std::vector<double> &tmp_rhs)
  {
    masks();
    int iters = 0;
    double res = 0;
    const int rc = wbem_solve(ctx, comp_dom.support_points.data(), phi.data(), dphi_dn.data(), tmp_rhs.data(),
                              &iters, &res);
    finish(rc, iters, res);
  }
  // source/bem_problem.cc:903-961
  void residual(std::vector<double> &res, const std::vector<double> &phi, const std::vector<double> &dphi_dn)
  {
    masks();
    res.resize(phi.size());
    check(wbem_residual(ctx, res.data(), phi.data(), dphi_dn.data()));
  }
  // neumann_matrix(i, j) / dirichlet_matrix(i, j) rows (public members of the reference class)
  void matrix_rows(int which, uint32_t r0, uint32_t r1, std::vector<double> &out)
  {
    out.resize((size_t)(r1 - r0) * comp_dom.n_dofs());
    check(wbem_get_rows(ctx, which, r0, r1, out.data()));
  }

  FlatDomain &comp_dom;
  std::vector<double> system_rhs, sol, alpha; // include/bem_problem.h:155-158
  ConstraintLines constraints;                // include/bem_problem.h:163 (lines of the last compute_constraints)
  unsigned int last_step = 0;
  double last_residual = 0;
  wbem_ctx *ctx = nullptr;

private:
  void masks() { check(wbem_set_masks(ctx, comp_dom.surface_nodes.data(), comp_dom.other_nodes.data())); }
  void check(int rc)
  {
    if (rc < 0) throw std::runtime_error(std::string("libwbem: ") + wbem_last_error(ctx));
  }
  void finish(int rc, int iters, double res)
  {
    check(rc);
    last_step = (unsigned int)iters;
    last_residual = res;
    check(wbem_get_alpha(ctx, alpha.data()));
    check(wbem_get_system_rhs(ctx, system_rhs.data()));
    check(wbem_get_sol(ctx, sol.data()));
    if (rc > 0) throw NoConvergence(last_step, last_residual);
  }
};
} // namespace wbem
