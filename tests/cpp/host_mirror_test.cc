// Exercises the C++ host mirror (wavebem_b200/csrc/bem_problem_b200.h) of the reference's
// BEMProblem<3> through the C ABI: reinit -> solve -> solve_system -> vmult/residual, and the
// NoConvergence path.  Input/outputs are raw binary files written/read by
// tests/test_gpu_cpp_mirror.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../wavebem_b200/csrc/bem_problem_b200.h"

template <typename T>
static void rd(FILE *f, std::vector<T> &v, size_t n)
{
  v.resize(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n)
    {
      fprintf(stderr, "short read\n");
      exit(2);
    }
}

int main(int argc, char **argv)
{
  if (argc < 3)
    {
      fprintf(stderr, "usage: %s in.bin out.bin [n_gpus]\n", argv[0]);
      return 2;
    }
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  uint32_t hdr[5];
  if (fread(hdr, sizeof(uint32_t), 5, f) != 5) return 2;
  const uint32_t N = hdr[0], C = hdr[1], ndn = hdr[2], nl = hdr[3], nnz = hdr[4];
  wbem::FlatDomain dom;
  wbem::ConstraintLines con;
  std::vector<double> bc;
  rd(f, dom.support_points, 3 * (size_t)N);
  rd(f, dom.cell_dofs, 4 * (size_t)C);
  rd(f, dom.cell_direction, C);
  rd(f, dom.dn_ptr, N + 1);
  rd(f, dom.dn_idx, ndn);
  rd(f, dom.surface_nodes, N);
  rd(f, dom.other_nodes, N);
  rd(f, bc, N);
  rd(f, con.lines, nl);
  rd(f, con.ptr, nl + 1);
  rd(f, con.col, nnz);
  rd(f, con.val, nnz);
  rd(f, con.inhom, nl);
  fclose(f);

  wbem_params p;
  wbem_default_params(&p);
  p.gmres_tol = 1e-12;
  p.gmres_max_steps = 400;
  std::vector<double> phi(N, 0.0), dphi(N, 0.0), res, y;
  double checks[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  try
    {
      wbem::BEMProblem bem(dom, &p);
      bem.reinit();
      bem.set_constraints(con);
      bem.solve(phi, dphi, bc); // assemble_system + solve_system
      checks[0] = bem.last_step;
      // the J.v pattern: another solve_system on the same matrices
      std::vector<double> phi2(N, 0.0), dphi2(N, 0.0), bc2(bc);
      for (auto &v : bc2) v *= 2.0;
      wbem::ConstraintLines con2 = con;
      for (auto &v : con2.inhom) v *= 2.0;
      bem.set_constraints(con2);
      bem.solve_system(phi2, dphi2, bc2);
      double d = 0, s = 0;
      for (uint32_t i = 0; i < N; ++i)
        {
          d += (phi2[i] - 2 * phi[i]) * (phi2[i] - 2 * phi[i]) + (dphi2[i] - 2 * dphi[i]) * (dphi2[i] - 2 * dphi[i]);
          s += 4 * (phi[i] * phi[i] + dphi[i] * dphi[i]);
        }
      checks[1] = d / s; // linearity of the solve
      bem.set_constraints(con);
      std::vector<double> full_phi(N), full_dphi(N);
      for (uint32_t i = 0; i < N; ++i)
        {
          full_phi[i] = dom.surface_nodes[i] == 1 ? bc[i] : phi[i];
          full_dphi[i] = dom.surface_nodes[i] == 1 ? dphi[i] : bc[i];
        }
      bem.residual(res, full_phi, full_dphi);
      for (double v : res) checks[2] = std::max(checks[2], std::abs(v));
      // NoConvergence must surface as an exception
      wbem_params q = p;
      q.gmres_tol = 1e-30;
      q.gmres_max_steps = 5;
      wbem::BEMProblem bem2(dom, &q);
      bem2.reinit();
      bem2.set_constraints(con);
      std::vector<double> a(N, 0.0), b(N, 0.0);
      try
        {
          bem2.solve(a, b, bc);
        }
      catch (const wbem::NoConvergence &e)
        {
          checks[3] = e.last_step;
        }
      // the opt-in local-inverse preconditioner: same solution, fewer iterations
      wbem_params sp = p;
      sp.precond_kind = 1;
      wbem::BEMProblem bem3(dom, &sp);
      bem3.reinit();
      bem3.set_constraints(con);
      std::vector<double> phi3(N, 0.0), dphi3(N, 0.0);
      bem3.solve(phi3, dphi3, bc);
      double d3 = 0, s3 = 0;
      for (uint32_t i = 0; i < N; ++i)
        {
          d3 += (phi3[i] - phi[i]) * (phi3[i] - phi[i]) + (dphi3[i] - dphi[i]) * (dphi3[i] - dphi[i]);
          s3 += phi[i] * phi[i] + dphi[i] * dphi[i];
        }
      checks[4] = std::sqrt(d3 / s3);
      checks[5] = bem3.last_step;
      // compute_constraints inside the library (auto_constraints = 1): no lines handed over
      wbem_params ap = p;
      ap.auto_constraints = 1;
      wbem::BEMProblem bem4(dom, &ap);
      bem4.reinit();
      std::vector<double> phi4(N, 0.0), dphi4(N, 0.0);
      bem4.solve(phi4, dphi4, bc);
      double d4 = 0, s4 = 0;
      for (uint32_t i = 0; i < N; ++i)
        {
          d4 += (phi4[i] - phi[i]) * (phi4[i] - phi[i]) + (dphi4[i] - dphi[i]) * (dphi4[i] - dphi[i]);
          s4 += phi[i] * phi[i] + dphi[i] * dphi[i];
        }
      checks[6] = std::sqrt(d4 / s4);
      bem4.compute_constraints(bc);
      checks[7] = (double)bem4.constraints.lines.size();
      // the double-node sets regenerated from the support points equal the ones handed in
      wbem::FlatDomain dom2 = dom;
      dom2.generate_double_nodes_set();
      checks[8] = (dom2.dn_ptr == dom.dn_ptr && dom2.dn_idx == dom.dn_idx) ? 1.0 : 0.0;
      // ONE BEMProblem driving several row blocks from this single-threaded program (n_gpus > 1):
      // across the box's GPUs when it has several, else 3 row blocks on device 0.  Same kernels,
      // same rows, replicated Krylov vectors: bitwise the one-GPU result.
      {
        const int want = argc > 3 ? atoi(argv[3]) : 0; // visible GPUs to spread over (0 / 1: share device 0)
        wbem_params gp = p;
        gp.n_gpus = want >= 2 ? want : 3;
        for (int r = 0; r < gp.n_gpus; ++r) gp.devices[r] = want >= 2 ? r : 0;
        wbem::BEMProblem bemg(dom, &gp);
        bemg.reinit();
        bemg.set_constraints(con);
        std::vector<double> phig(N, 0.0), dphig(N, 0.0), rows1, rowsg;
        bemg.solve(phig, dphig, bc);
        bool same = bemg.last_step == (unsigned int)checks[0] && !memcmp(phig.data(), phi.data(), sizeof(double) * N) &&
                    !memcmp(dphig.data(), dphi.data(), sizeof(double) * N) &&
                    !memcmp(bemg.alpha.data(), bem.alpha.data(), sizeof(double) * N);
        bem.matrix_rows(0, 0, N, rows1);
        bemg.matrix_rows(0, 0, N, rowsg);
        same = same && !memcmp(rows1.data(), rowsg.data(), sizeof(double) * rows1.size());
        checks[9] = same ? 1.0 : 0.0;
        checks[10] = gp.n_gpus;
        checks[11] = want >= 2 ? 1.0 : 0.0;
      }
      FILE *o = fopen(argv[2], "wb");
      fwrite(phi.data(), sizeof(double), N, o);
      fwrite(dphi.data(), sizeof(double), N, o);
      fwrite(bem.alpha.data(), sizeof(double), N, o);
      fwrite(checks, sizeof(double), 12, o);
      fclose(o);
    }
  catch (const std::exception &e)
    {
      fprintf(stderr, "host_mirror_test: %s\n", e.what());
      return 1;
    }
  printf("host_mirror_test ok: N=%u iters=%g linearity=%.2e residual=%.2e noconv_step=%g\n", N, checks[0], checks[1],
         checks[2], checks[3]);
  return 0;
}
