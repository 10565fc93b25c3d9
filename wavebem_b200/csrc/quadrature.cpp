// quadrature.cpp -- reference-cell tables of the two rules the path integrates with.
//
//  * regular pairs:  deal.II QGauss<2>(n)  (reference source/computational_domain.cc:125-128,
//    prm "Quadrature order" = 4) -- tensor Gauss-Legendre on [0,1]^2, first index fastest;
//  * singular pairs: deal.II QGaussOneOverR<2>(n, unit_support_point[j], true)
//    (reference source/bem_problem.cc:116-120, prm "Singular quadrature order" = 5) -- a
//    Duffy/Lachat-Watson polar rule with 2 n^2 points centred on vertex j, weights already
//    multiplied by the distance to the vertex.
// deal.II is not vendored in the reference tree; the constructions below follow its published
// algorithms (Legendre roots by Newton iteration; polar map (u, u tan(pi v / 4)) mirrored on
// the diagonal and rotated by 0, +pi/2, -pi/2, pi for vertices 0..3).
#include <cmath>
#include <vector>

#include "internal.h"

namespace
{
const long double kPi = 3.14159265358979323846264338327950288L;

// Gauss-Legendre nodes/weights on [0,1]; symmetric pairs are filled together.
void legendre_rule_unit(int n, double *x, double *w)
{
  for (int k = 0; k < (n + 1) / 2; ++k)
    {
      long double z = std::cos(kPi * (k + 0.75L) / (n + 0.5L)); // k-th root estimate
      long double dp = 1, step = 1;
      for (int it = 0; it < 100 && std::fabs(step) > 4e-19L; ++it)
        {
          long double pm1 = 0, p = 1; // Bonnet recursion up to degree n
          for (int j = 0; j < n; ++j)
            {
              long double pn = ((2 * j + 1) * z * p - j * pm1) / (j + 1);
              pm1 = p;
              p = pn;
            }
          dp = n * (z * p - pm1) / (z * z - 1);
          step = p / dp;
          z -= step;
        }
      const double half = (double)(0.5L * z);
      const double wt = (double)(1.0L / ((1 - z * z) * dp * dp));
      x[k] = 0.5 - half;
      x[n - 1 - k] = 0.5 + half;
      w[k] = w[n - 1 - k] = wt;
    }
}
} // namespace

int wbem_build_quadrature(int quad_order, int sing_order, QuadTables *qt)
{
  if (quad_order < 1 || quad_order * quad_order > WBEM_MAX_NQ) return -1;
  if (sing_order < 1 || 2 * sing_order * sing_order > WBEM_MAX_NS) return -1;
  const int n = quad_order;
  qt->n1 = n;
  qt->nq = n * n;
  legendre_rule_unit(n, qt->g1_x, qt->g1_w);
  for (int iy = 0; iy < n; ++iy)
    for (int ix = 0; ix < n; ++ix)
      {
        const int q = iy * n + ix;
        const double u = qt->g1_x[ix], v = qt->g1_x[iy];
        qt->g_u[q] = u;
        qt->g_v[q] = v;
        qt->g_w[q] = qt->g1_w[ix] * qt->g1_w[iy];
        qt->g_shape[0][q] = (1 - u) * (1 - v);
        qt->g_shape[1][q] = u * (1 - v);
        qt->g_shape[2][q] = (1 - u) * v;
        qt->g_shape[3][q] = u * v;
      }

  const int m = sing_order, m2 = m * m;
  qt->ns = 2 * m2;
  std::vector<double> x1(m), w1(m);
  legendre_rule_unit(m, x1.data(), w1.data());
  const double quarter_pi = (double)(kPi / 4);
  // rule for vertex 0, then rotations about the cell centre
  std::vector<double> bu(2 * m2), bv(2 * m2), bw(2 * m2);
  for (int iy = 0; iy < m; ++iy)
    for (int ix = 0; ix < m; ++ix)
      {
        const int q = iy * m + ix;
        const double rad = x1[ix], ang = quarter_pi * x1[iy];
        const double pu = rad, pv = rad * std::tan(ang);
        double wt = w1[ix] * w1[iy] * quarter_pi / std::cos(ang);
        wt *= std::sqrt(pu * pu + pv * pv); // factor_out_singularity = true
        bu[q] = pu;
        bv[q] = pv;
        bw[q] = wt;
        bu[m2 + q] = pv; // mirror image across the diagonal
        bv[m2 + q] = pu;
        bw[m2 + q] = wt;
      }
  const double angle[4] = {0.0, (double)(kPi / 2), -(double)(kPi / 2), (double)kPi};
  for (int vtx = 0; vtx < 4; ++vtx)
    {
      const double cs = std::cos(angle[vtx]), sn = std::sin(angle[vtx]);
      for (int q = 0; q < 2 * m2; ++q)
        {
          double u = bu[q], v = bv[q];
          if (vtx != 0)
            {
              const double a = u - 0.5, b = v - 0.5;
              u = cs * a - sn * b + 0.5;
              v = sn * a + cs * b + 0.5;
            }
          qt->s_u[vtx][q] = u;
          qt->s_v[vtx][q] = v;
          qt->s_w[vtx][q] = bw[q];
        }
    }
  return 0;
}
