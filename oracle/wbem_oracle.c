/*
 * wbem_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic of WaveBEM's collocation-BEM hot path
 * (reference: mathLab/WaveBEM, source/bem_problem.cc, include/laplace_kernel.h,
 * include/constrained_matrix.h) and of the deal.II 8.4 primitives it calls
 * (QGauss, QGaussOneOverR, FE_Q(1), MappingQ1 for codimension one, FullMatrix::vmult,
 * SolverGMRES), which are NOT vendored in the reference tree.
 *
 * PARITY UNPINNED: the reference cannot be built in this image (deal.II, SUNDIALS,
 * OpenCASCADE, Trilinos and UMFPACK are absent) and its tests/ directory holds no
 * known-answer test for this path (tests/template.cc prints "0").  The oracle is
 * therefore pinned only by first-principles known-answer tests (tests/test_oracle_kat.py):
 * quadrature identities, solid angles, manufactured harmonic solutions.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libwbem.so) never does.
 *
 * Compile twice: default (double) and -DORC_LONG_DOUBLE (x87 80-bit accumulation,
 * function suffix _ld) -- the latter is the "more exact" arbiter for tolerance tests.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORC_LONG_DOUBLE
typedef long double real;
#define FN(name) name##_ld
#define R_SQRT sqrtl
#define R_FABS fabsl
#define R_TAN tanl
#define R_COS cosl
#define R_SIN sinl
#else
typedef double real;
#define FN(name) name
#define R_SQRT sqrt
#define R_FABS fabs
#define R_TAN tan
#define R_COS cos
#define R_SIN sin
#endif

#define ORC_PI 3.14159265358979323846264338327950288L /* numbers::PI */

/* ------------------------------------------------------------------------------------
 * deal.II QGauss<1>(n) on [0,1]  (deal.II 8.4 base/quadrature_lib.cc, QGauss<1>::QGauss;
 * called through QuadratureSelector<2>("gauss",4), reference
 * source/computational_domain.cc:125-128).  Newton iteration on the Legendre polynomial
 * in long double, tolerance max(eps_double/100, 5 eps_long_double).
 * ---------------------------------------------------------------------------------- */
int FN(orc_gauss01)(int n, double *x, double *w)
{
  if (n < 1) return -1;
  const int m = (n + 1) / 2;
  const long double ld_eps = LDBL_EPSILON, d_eps = DBL_EPSILON;
  long double tol = d_eps / 100;
  if (ld_eps * 5 > tol) tol = ld_eps * 5;
  for (int i = 1; i <= m; ++i)
    {
      long double z = cosl(ORC_PI * (i - .25L) / (n + .5L));
      long double pp, p1, p2, p3;
      do
        {
          p1 = 1.;
          p2 = 0.;
          for (int j = 0; j < n; ++j)
            {
              p3 = p2;
              p2 = p1;
              p1 = ((2. * j + 1.) * z * p2 - j * p3) / (j + 1);
            }
          pp = n * (z * p1 - p2) / (z * z - 1);
          z = z - p1 / pp;
        }
      while (fabsl(p1 / pp) > tol);
      double xx = .5 * z;
      x[i - 1] = .5 - xx;
      x[n - i] = .5 + xx;
      double ww = 1. / ((1. - z * z) * pp * pp);
      w[i - 1] = ww;
      w[n - i] = ww;
    }
  return 0;
}

/* QGauss<2>(n): tensor product, first coordinate fastest (deal.II Quadrature<dim>
 * tensor constructor).  uv is [n*n][2]. */
int FN(orc_qgauss2)(int n, double *uv, double *w)
{
  double *x1 = (double *)malloc(sizeof(double) * 2 * n);
  double *w1 = x1 + n;
  if (FN(orc_gauss01)(n, x1, w1)) { free(x1); return -1; }
  for (int iy = 0; iy < n; ++iy)
    for (int ix = 0; ix < n; ++ix)
      {
        int q = ix + n * iy;
        uv[2 * q] = x1[ix];
        uv[2 * q + 1] = x1[iy];
        w[q] = w1[ix] * w1[iy];
      }
  free(x1);
  return 0;
}

/* QGaussOneOverR<2>(n, vertex_index, factor_out_singularity)
 * (deal.II 8.4 base/quadrature_lib.cc).  The reference constructs it from the unit
 * support point of local dof j (source/bem_problem.cc:116-120); for a vertex that
 * Point<2> constructor reduces to this vertex-index rule with 2 n^2 points.
 * uv is [2*n*n][2]. */
int FN(orc_qgauss_one_over_r)(int n, int vertex, int factor_out, double *uv, double *w)
{
  if (vertex < 0 || vertex > 3) return -1;
  const int n2 = n * n;
  double *guv = (double *)malloc(sizeof(double) * 3 * n2);
  double *gw = guv + 2 * n2;
  if (FN(orc_qgauss2)(n, guv, gw)) { free(guv); return -1; }
  const double pi4 = (double)(ORC_PI) / 4;
  for (int q = 0; q < n2; ++q)
    {
      const double g0 = guv[2 * q], g1 = guv[2 * q + 1];
      double px = g0;
      double py = g0 * tan(pi4 * g1);
      double ww = gw[q] * pi4 / cos(pi4 * g1);
      if (factor_out) ww *= sqrt(px * px + py * py);
      uv[2 * q] = px;
      uv[2 * q + 1] = py;
      w[q] = ww;
      w[n2 + q] = ww;
      uv[2 * (n2 + q)] = py;
      uv[2 * (n2 + q) + 1] = px;
    }
  double theta = 0;
  switch (vertex)
    {
    case 0: theta = 0; break;
    case 1: theta = (double)(ORC_PI) / 2; break;
    case 2: theta = -(double)(ORC_PI) / 2; break;
    case 3: theta = (double)(ORC_PI); break;
    }
  const double R00 = cos(theta), R01 = -sin(theta);
  const double R10 = sin(theta), R11 = cos(theta);
  if (vertex != 0)
    for (int q = 0; q < 2 * n2; ++q)
      {
        double xx = uv[2 * q] - .5, yy = uv[2 * q + 1] - .5;
        uv[2 * q] = R00 * xx + R01 * yy + .5;
        uv[2 * q + 1] = R10 * xx + R11 * yy + .5;
      }
  free(guv);
  return 0;
}

/* FE_Q<2,3>(1) shape functions, deal.II lexicographic vertex order
 * (reference source/computational_domain.cc:57). */
static void shape_q1(real u, real v, real phi[4], real du[4], real dv[4])
{
  phi[0] = (1 - u) * (1 - v);
  phi[1] = u * (1 - v);
  phi[2] = (1 - u) * v;
  phi[3] = u * v;
  du[0] = -(1 - v);
  du[1] = (1 - v);
  du[2] = -v;
  du[3] = v;
  dv[0] = -(1 - u);
  dv[1] = -u;
  dv[2] = (1 - u);
  dv[3] = u;
}

/* FEValues<2,3> with a Q1 mapping on one cell: quadrature points, cell normals, JxW,
 * shape values (deal.II MappingQ1<2,3>: contravariant DX_t, G = DX_t DX_t^T,
 * JxW = sqrt(det G) w, normal = DX_t[0] x DX_t[1] normalised, flipped when
 * cell->direction_flag() is false).  X is the 4 vertex positions [4][3]; vertex k is
 * support_points[dof k] (source/computational_domain.cc:1516-1517).
 * Outputs: qp[nq][3], nrm[nq][3], jxw[nq], shp[4][nq] (any may be NULL). */
static void fe_values_cell(const double *X, int dir_flag, int nq, const double *uv,
                           const double *w, real *qp, real *nrm, real *jxw, real *shp)
{
  for (int q = 0; q < nq; ++q)
    {
      real phi[4], du[4], dv[4];
      shape_q1((real)uv[2 * q], (real)uv[2 * q + 1], phi, du, dv);
      real p[3] = {0, 0, 0}, t0[3] = {0, 0, 0}, t1[3] = {0, 0, 0};
      for (int k = 0; k < 4; ++k)
        for (int d = 0; d < 3; ++d)
          {
            p[d] += phi[k] * (real)X[3 * k + d];
            t0[d] += du[k] * (real)X[3 * k + d];
            t1[d] += dv[k] * (real)X[3 * k + d];
          }
      real g00 = t0[0] * t0[0] + t0[1] * t0[1] + t0[2] * t0[2];
      real g01 = t0[0] * t1[0] + t0[1] * t1[1] + t0[2] * t1[2];
      real g11 = t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2];
      real detG = g00 * g11 - g01 * g01;
      real c[3] = {t0[1] * t1[2] - t0[2] * t1[1], t0[2] * t1[0] - t0[0] * t1[2],
                   t0[0] * t1[1] - t0[1] * t1[0]};
      real cn = R_SQRT(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
      real sgn = dir_flag ? 1 : -1;
      if (qp) for (int d = 0; d < 3; ++d) qp[3 * q + d] = p[d];
      if (nrm) for (int d = 0; d < 3; ++d) nrm[3 * q + d] = sgn * (c[d] / cn);
      if (jxw) jxw[q] = R_SQRT(detG) * (real)w[q];
      if (shp) for (int k = 0; k < 4; ++k) shp[k * nq + q] = phi[k];
    }
}

/* exported double-precision view of fe_values_cell for unit tests */
void FN(orc_fe_values)(const double *X, int dir_flag, int nq, const double *uv,
                       const double *w, double *qp, double *nrm, double *jxw, double *shp)
{
  real *b = (real *)malloc(sizeof(real) * (size_t)nq * 11);
  real *rq = b, *rn = b + 3 * nq, *rj = b + 6 * nq, *rs = b + 7 * nq;
  fe_values_cell(X, dir_flag, nq, uv, w, rq, rn, rj, rs);
  for (int i = 0; i < 3 * nq; ++i) { qp[i] = (double)rq[i]; nrm[i] = (double)rn[i]; }
  for (int i = 0; i < nq; ++i) jxw[i] = (double)rj[i];
  for (int i = 0; i < 4 * nq; ++i) shp[i] = (double)rs[i];
  free(b);
}

/* LaplaceKernel::kernels<3> (reference include/laplace_kernel.h:45-62, 3-D branch
 * :55-58): r = |R|, r2 = r*r, d = 1/(r 4 pi), D = R / (-4 pi r2 r). */
static inline void laplace_kernels(const real R[3], real D[3], real *d)
{
  real r = R_SQRT(R[0] * R[0] + R[1] * R[1] + R[2] * R[2]);
  real r2 = r * r;
  *d = (1. / (r * 4 * (real)ORC_PI));
  real den = (-4 * (real)ORC_PI * r2 * r);
  D[0] = R[0] / den;
  D[1] = R[1] / den;
  D[2] = R[2] / den;
}

/* BEMProblem<3>::assemble_system (reference source/bem_problem.cc:106-590) restricted
 * to rows [row0,row1).  Loop order: cell outer, node inner, q, j (:190,:214,:243,:248).
 * dn_ptr/dn_idx: CSR of comp_dom.double_nodes_set (each set contains i itself,
 * source/computational_domain.cc:279-303).  A (node,cell) pair whose cell holds a dof
 * of double_nodes_set[i] is integrated ONLY with QGaussOneOverR centred on the first
 * such local dof (:223-230, :261-525).  Nm, Dm: row-major (row1-row0) x N, zeroed here
 * (:111-112).  With OpenMP each thread runs the same cell-outer loop on its own block of
 * rows, so every entry sees its cell contributions in the reference order. */
int FN(orc_assemble_rows)(int N, int C, const double *xyz, const uint32_t *cell_dofs,
                          const uint8_t *dir_flag, const uint32_t *dn_ptr,
                          const uint32_t *dn_idx, int quad_order, int sing_order, int row0,
                          int row1, double *Nm, double *Dm, int nthreads)
{
  const int nq = quad_order * quad_order;
  const int ns = 2 * sing_order * sing_order;
  double *uv = (double *)malloc(sizeof(double) * 3 * nq);
  double *wq = uv + 2 * nq;
  if (FN(orc_qgauss2)(quad_order, uv, wq)) return -1;
  double *suv = (double *)malloc(sizeof(double) * 3 * ns * 4);
  double *sw = suv + 2 * ns * 4;
  for (int j = 0; j < 4; ++j)
    if (FN(orc_qgauss_one_over_r)(sing_order, j, 1, suv + 2 * ns * j, sw + ns * j)) return -1;

  const long nrows = row1 - row0;
  memset(Nm, 0, sizeof(double) * nrows * (size_t)N);
  memset(Dm, 0, sizeof(double) * nrows * (size_t)N);

  /* FEValues per cell with the regular rule (fe_v.reinit(cell), :192-196), cached. */
  real *cq = (real *)malloc(sizeof(real) * (size_t)C * nq * 7);
  real *shp = (real *)malloc(sizeof(real) * 4 * nq);
  for (int c = 0; c < C; ++c)
    {
      double X[12];
      for (int k = 0; k < 4; ++k)
        for (int d = 0; d < 3; ++d) X[3 * k + d] = xyz[3 * (size_t)cell_dofs[4 * c + k] + d];
      real *b = cq + (size_t)c * nq * 7;
      fe_values_cell(X, dir_flag[c], nq, uv, wq, b, b + 3 * nq, b + 6 * nq, c == 0 ? shp : NULL);
    }
  if (C == 0) { free(cq); free(shp); free(uv); free(suv); return 0; }

#ifdef _OPENMP
  if (nthreads < 1) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  const long chunk = (nrows + nthreads - 1) / nthreads;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
#endif
  for (int t = 0; t < nthreads; ++t)
    {
      const long a = row0 + t * chunk;
      const long b = (a + chunk < row1) ? a + chunk : row1;
      real *sqp = (real *)malloc(sizeof(real) * ns * 11);
      real *snr = sqp + 3 * ns, *sjw = sqp + 6 * ns, *ssh = sqp + 7 * ns;
      for (int c = 0; c < C; ++c)
        {
          const uint32_t *dofs = cell_dofs + 4 * c;
          const real *qp = cq + (size_t)c * nq * 7;
          const real *nr = qp + 3 * nq, *jw = qp + 6 * nq;
          for (long i = a; i < b; ++i)
            {
              real locN[4] = {0, 0, 0, 0}, locD[4] = {0, 0, 0, 0};
              int is_singular = 0, singular_index = -1;
              for (int j = 0; j < 4 && !is_singular; ++j)
                for (uint32_t k = dn_ptr[i]; k < dn_ptr[i + 1]; ++k)
                  if (dn_idx[k] == dofs[j])
                    {
                      singular_index = j;
                      is_singular = 1;
                      break;
                    }
              const real xi[3] = {(real)xyz[3 * i], (real)xyz[3 * i + 1], (real)xyz[3 * i + 2]};
              if (!is_singular)
                {
                  for (int q = 0; q < nq; ++q)
                    {
                      real R[3] = {qp[3 * q] + -1.0 * xi[0], qp[3 * q + 1] + -1.0 * xi[1],
                                   qp[3 * q + 2] + -1.0 * xi[2]};
                      real D[3], s;
                      laplace_kernels(R, D, &s);
                      real Dn = D[0] * nr[3 * q] + D[1] * nr[3 * q + 1] + D[2] * nr[3 * q + 2];
                      for (int j = 0; j < 4; ++j)
                        {
                          locN[j] += (Dn * shp[j * nq + q] * jw[q]);
                          locD[j] += (s * shp[j * nq + q] * jw[q]);
                        }
                    }
                }
              else
                {
                  double X[12];
                  for (int k = 0; k < 4; ++k)
                    for (int d = 0; d < 3; ++d) X[3 * k + d] = xyz[3 * (size_t)dofs[k] + d];
                  fe_values_cell(X, dir_flag[c], ns, suv + 2 * ns * singular_index,
                                 sw + ns * singular_index, sqp, snr, sjw, ssh);
                  for (int q = 0; q < ns; ++q)
                    {
                      real R[3] = {sqp[3 * q] + -1.0 * xi[0], sqp[3 * q + 1] + -1.0 * xi[1],
                                   sqp[3 * q + 2] + -1.0 * xi[2]};
                      real D[3], s;
                      laplace_kernels(R, D, &s);
                      real Dn = D[0] * snr[3 * q] + D[1] * snr[3 * q + 1] + D[2] * snr[3 * q + 2];
                      for (int j = 0; j < 4; ++j)
                        {
                          locN[j] += (Dn * ssh[j * ns + q] * sjw[q]);
                          locD[j] += (s * ssh[j * ns + q] * sjw[q]);
                        }
                    }
                }
              double *Ni = Nm + (size_t)(i - row0) * N, *Di = Dm + (size_t)(i - row0) * N;
#ifdef ORC_LONG_DOUBLE
              /* keep the long-double variant's extra precision only inside the pair */
#endif
              for (int j = 0; j < 4; ++j)
                {
                  Ni[dofs[j]] += (double)locN[j];
                  Di[dofs[j]] += (double)locD[j];
                }
            }
        }
      free(sqp);
    }
  free(cq);
  free(shp);
  free(uv);
  free(suv);
  return 0;
}

#ifndef ORC_LONG_DOUBLE
/* ====================================================================================
 * Operator algebra, constraints, preconditioner, GMRES: double only.
 * ==================================================================================== */

/* deal.II FullMatrix<double>::vmult(w, v, adding): serial row-major dot products
 * (deal.II 8.4 lac/full_matrix.templates.h).  A is nrows x N (a row slab). */
static void fullmatrix_vmult(int nrows, int N, const double *A, double *w, const double *v,
                             int adding, int nthreads)
{
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : omp_get_max_threads()) schedule(static)
#endif
  for (int i = 0; i < nrows; ++i)
    {
      double s = adding ? w[i] : 0.;
      const double *e = A + (size_t)i * N;
      for (int j = 0; j < N; ++j) s += v[j] * e[j];
      w[i] = s;
    }
}

void orc_fullmatrix_vmult(int nrows, int N, const double *A, double *w, const double *v,
                          int adding, int nthreads)
{
  fullmatrix_vmult(nrows, N, A, w, v, adding, nthreads);
}

/* BEMProblem<3>::compute_alpha, Direct branch (reference source/bem_problem.cc:594-618):
 * alpha = neumann_matrix * (-1,...,-1). */
void orc_compute_alpha(int nrows, int N, const double *Nm, double *alpha, int nthreads)
{
  double *ones = (double *)malloc(sizeof(double) * N);
  for (int j = 0; j < N; ++j) ones[j] = -1.;
  fullmatrix_vmult(nrows, N, Nm, alpha, ones, 0, nthreads);
  free(ones);
}

typedef struct
{
  int N;
  const double *Nm, *Dm; /* N x N row-major */
  const double *alpha;
  const double *surface_nodes, *other_nodes;
  /* ConstraintMatrix lines (may be 0 lines): is_con[i] = line index or -1 */
  const int32_t *con_line_of; /* [N] */
  const uint32_t *con_ptr;    /* [n_lines+1] */
  const uint32_t *con_col;
  const double *con_val;
  const double *con_inhom; /* [n_lines] */
  int nthreads;
  long n_vmult;
} orc_op;

/* BEMProblem<3>::vmult, Direct branch (reference source/bem_problem.cc:620-670). */
void orc_vmult(int N, const double *Nm, const double *Dm, const double *alpha,
               const double *surface_nodes, const double *other_nodes, double *dst,
               const double *src, int nthreads)
{
  double *phi = (double *)malloc(sizeof(double) * 2 * N);
  double *dphi_dn = phi + N;
  for (int i = 0; i < N; ++i)
    {
      dst[i] = 0;
      phi[i] = src[i] * other_nodes[i];
      dphi_dn[i] = src[i] * surface_nodes[i];
    }
  fullmatrix_vmult(N, N, Dm, dst, dphi_dn, 0, nthreads); /* dirichlet_matrix.vmult(dst,dphi_dn) */
  for (int i = 0; i < N; ++i) dst[i] *= -1;              /* dst *= -1 */
  fullmatrix_vmult(N, N, Nm, dst, phi, 1, nthreads);     /* neumann_matrix.vmult_add(dst,phi) */
  for (int i = 0; i < N; ++i)
    {
      phi[i] *= alpha[i]; /* phi.scale(alpha) */
      dst[i] += phi[i];   /* dst += phi */
    }
  double linf = 0;
  for (int i = 0; i < N; ++i)
    if (fabs(surface_nodes[i]) > linf) linf = fabs(surface_nodes[i]);
  if (linf < 1e-10) /* pure Neumann problem: dst.add(-dst.l2_norm()) (:667-668) */
    {
      double s = 0;
      for (int i = 0; i < N; ++i) s += dst[i] * dst[i];
      s = sqrt(s);
      for (int i = 0; i < N; ++i) dst[i] += -s;
    }
  free(phi);
}

/* BEMProblem<3>::compute_rhs, Direct branch (reference source/bem_problem.cc:673-707). */
void orc_compute_rhs(int N, const double *Nm, const double *Dm, const double *alpha,
                     const double *surface_nodes, const double *other_nodes, double *dst,
                     const double *src, int nthreads)
{
  double *phi = (double *)malloc(sizeof(double) * 2 * N);
  double *dphi_dn = phi + N;
  for (int i = 0; i < N; ++i)
    {
      phi[i] = src[i] * surface_nodes[i];
      dphi_dn[i] = src[i] * other_nodes[i];
    }
  fullmatrix_vmult(N, N, Nm, dst, phi, 0, nthreads);
  for (int i = 0; i < N; ++i)
    {
      phi[i] *= alpha[i];
      dst[i] += phi[i];
      dst[i] *= -1;
    }
  fullmatrix_vmult(N, N, Dm, dst, dphi_dn, 1, nthreads);
  free(phi);
}

/* ConstrainedOperator::vmult (reference include/constrained_matrix.h:73-86). */
static void constrained_vmult(orc_op *op, double *dst, const double *src)
{
  orc_vmult(op->N, op->Nm, op->Dm, op->alpha, op->surface_nodes, op->other_nodes, dst, src,
            op->nthreads);
  op->n_vmult++;
  if (op->con_line_of)
    for (int i = 0; i < op->N; ++i)
      {
        int l = op->con_line_of[i];
        if (l >= 0)
          {
            dst[i] = src[i];
            for (uint32_t k = op->con_ptr[l]; k < op->con_ptr[l + 1]; ++k)
              dst[i] -= op->con_val[k] * src[op->con_col[k]];
          }
      }
}

void orc_constrained_vmult(int N, const double *Nm, const double *Dm, const double *alpha,
                           const double *surface_nodes, const double *other_nodes,
                           const int32_t *con_line_of, const uint32_t *con_ptr,
                           const uint32_t *con_col, const double *con_val, double *dst,
                           const double *src, int nthreads)
{
  orc_op op = {N, Nm, Dm, alpha, surface_nodes, other_nodes, con_line_of, con_ptr, con_col,
               con_val, NULL, nthreads, 0};
  constrained_vmult(&op, dst, src);
}

/* ConstrainedOperator::distribute_rhs (reference include/constrained_matrix.h:89-94). */
void orc_distribute_rhs(int N, const int32_t *con_line_of, const double *con_inhom, double *rhs)
{
  if (!con_line_of) return;
  for (int i = 0; i < N; ++i)
    if (con_line_of[i] >= 0) rhs[i] = con_inhom[con_line_of[i]];
}

/* ------------------------------------------------------------------------------------
 * BEMProblem<3>::assemble_preconditioner (reference source/bem_problem.cc:1107-1149):
 * band_system(j,i) for j in [max(i-50,0), min(i+50,N)) -- preconditioner_band = 100
 * (:68).  Column i comes from neumann(j,i) (+alpha_i on the diagonal) when
 * surface_nodes(i)==0, else -dirichlet(j,i); constrained rows hold only a unit diagonal.
 * The reference factorises it with UMFPACK (SparseDirectUMFPACK::initialize, :1147) --
 * an exact sparse LU; restated here as a banded LU with partial pivoting (the same
 * inverse up to rounding).  Storage: LAPACK-style band, kl = hb, ku = hb-1.
 * ---------------------------------------------------------------------------------- */
typedef struct
{
  int n, kl, ku, ldab; /* ldab = 2*kl+ku+1 */
  double *ab;          /* column-major band storage, [n][ldab] */
  int *ipiv;
} orc_band;

#define AB(B, i, j) ((B)->ab[(size_t)(j) * (B)->ldab + ((B)->kl + (B)->ku + (i) - (j))])

static orc_band *band_alloc(int n, int kl, int ku)
{
  orc_band *B = (orc_band *)malloc(sizeof(orc_band));
  B->n = n;
  B->kl = kl;
  B->ku = ku;
  B->ldab = 2 * kl + ku + 1;
  B->ab = (double *)calloc((size_t)n * B->ldab, sizeof(double));
  B->ipiv = (int *)malloc(sizeof(int) * n);
  return B;
}
static void band_free(orc_band *B)
{
  if (!B) return;
  free(B->ab);
  free(B->ipiv);
  free(B);
}

/* unblocked banded LU with partial pivoting (LAPACK dgbtf2 algorithm) */
static int band_factor(orc_band *B)
{
  const int n = B->n, kl = B->kl, ku = B->ku, kv = kl + ku;
  int ju = 0;
  for (int j = 0; j < n; ++j)
    {
      int km = (kl < n - 1 - j) ? kl : n - 1 - j;
      int jp = 0;
      double amax = fabs(AB(B, j, j));
      for (int k = 1; k <= km; ++k)
        if (fabs(AB(B, j + k, j)) > amax)
          {
            amax = fabs(AB(B, j + k, j));
            jp = k;
          }
      B->ipiv[j] = j + jp;
      if (amax == 0.) return j + 1;
      int t = j + ku + jp;
      if (t > n - 1) t = n - 1;
      if (t > ju) ju = t;
      if (jp != 0)
        for (int c = j; c <= ju; ++c)
          {
            double tmp = AB(B, j + jp, c);
            AB(B, j + jp, c) = AB(B, j, c);
            AB(B, j, c) = tmp;
          }
      double piv = 1. / AB(B, j, j);
      for (int k = 1; k <= km; ++k) AB(B, j + k, j) *= piv;
      for (int c = j + 1; c <= ju; ++c)
        {
          double ujc = AB(B, j, c);
          if (ujc != 0.)
            for (int k = 1; k <= km; ++k) AB(B, j + k, c) -= AB(B, j + k, j) * ujc;
        }
      (void)kv;
    }
  return 0;
}

static void band_solve(const orc_band *B, double *x)
{
  const int n = B->n, kl = B->kl, kv = B->kl + B->ku;
  for (int j = 0; j < n; ++j)
    {
      int km = (kl < n - 1 - j) ? kl : n - 1 - j;
      int p = B->ipiv[j];
      if (p != j)
        {
          double t = x[p];
          x[p] = x[j];
          x[j] = t;
        }
      double xj = x[j];
      for (int k = 1; k <= km; ++k) x[j + k] -= AB(B, j + k, j) * xj;
    }
  for (int j = n - 1; j >= 0; --j)
    {
      x[j] /= AB(B, j, j);
      double xj = x[j];
      int lo = j - kv;
      if (lo < 0) lo = 0;
      for (int i = lo; i < j; ++i) x[i] -= AB(B, i, j) * xj;
    }
}

static orc_band *build_preconditioner(const orc_op *op, int band)
{
  const int N = op->N, hb = band / 2;
  orc_band *B = band_alloc(N, hb, hb); /* rows j-i in [-hb, hb-1]: ku = hb, kl = hb-1 (stored as hb) */
  for (int i = 0; i < N; ++i)
    {
      if (op->con_line_of && op->con_line_of[i] >= 0) AB(B, i, i) = 1;
      int j0 = i - hb > 0 ? i - hb : 0;
      int j1 = i + hb < N ? i + hb : N;
      for (int j = j0; j < j1; ++j)
        if (!(op->con_line_of && op->con_line_of[j] >= 0))
          {
            if (op->surface_nodes[i] == 0)
              {
                double v = op->Nm[(size_t)j * N + i];
                if (i == j) v += op->alpha[i];
                AB(B, j, i) = v;
              }
            else
              AB(B, j, i) = -op->Dm[(size_t)j * N + i];
          }
    }
  return B;
}

/* exported: dense copy of band_system (N x N row-major, zeros outside), for tests */
void orc_band_system_dense(int N, const double *Nm, const double *Dm, const double *alpha,
                           const double *surface_nodes, const int32_t *con_line_of, int band,
                           double *out)
{
  orc_op op = {N, Nm, Dm, alpha, surface_nodes, NULL, con_line_of, NULL, NULL, NULL, NULL, 1, 0};
  orc_band *B = build_preconditioner(&op, band);
  memset(out, 0, sizeof(double) * (size_t)N * N);
  const int hb = band / 2;
  for (int i = 0; i < N; ++i)
    {
      int j0 = i - hb > 0 ? i - hb : 0;
      int j1 = i + hb < N ? i + hb : N;
      for (int j = j0; j < j1; ++j) out[(size_t)j * N + i] = AB(B, j, i);
      out[(size_t)i * N + i] = AB(B, i, i);
    }
  band_free(B);
}

/* ------------------------------------------------------------------------------------
 * deal.II SolverGMRES<Vector<double>>::solve with AdditionalData(max_n_tmp_vectors),
 * left preconditioning, default residual (deal.II 8.4 lac/solver_gmres.h), as called at
 * reference source/bem_problem.cc:826-827, :853.  Stopping: SolverControl::check --
 * success when the preconditioned residual estimate <= tol (absolute), failure when
 * step >= max_steps.  Returns 0 on success, 1 on no convergence.
 * ---------------------------------------------------------------------------------- */
static double vdot(int n, const double *a, const double *b)
{
  double s = 0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

static void givens_rotation(double *h, double *b, double *ci, double *si, int col)
{
  for (int i = 0; i < col; i++)
    {
      const double s = si[i];
      const double c = ci[i];
      const double dummy = h[i];
      h[i] = c * dummy + s * h[i + 1];
      h[i + 1] = -s * dummy + c * h[i + 1];
    }
  const double r = 1. / sqrt(h[col] * h[col] + h[col + 1] * h[col + 1]);
  si[col] = h[col + 1] * r;
  ci[col] = h[col] * r;
  h[col] = ci[col] * h[col] + si[col] * h[col + 1];
  b[col + 1] = -si[col] * b[col];
  b[col] *= ci[col];
}

static int gmres_left(orc_op *op, const orc_band *P, int n_tmp_vectors, double tol,
                      int max_steps, double *x, const double *b, int *iters, double *last_res,
                      double *res_hist, int res_hist_len)
{
  const int n = op->N;
  const int m = n_tmp_vectors - 2; /* Krylov basis length per cycle */
  double *V = (double *)malloc(sizeof(double) * (size_t)n * (n_tmp_vectors - 1));
  double *p = (double *)malloc(sizeof(double) * n);
  double *H = (double *)calloc((size_t)(n_tmp_vectors) * (n_tmp_vectors - 1), sizeof(double));
  double *gamma = (double *)calloc(n_tmp_vectors, sizeof(double));
  double *ci = (double *)calloc(n_tmp_vectors, sizeof(double));
  double *si = (double *)calloc(n_tmp_vectors, sizeof(double));
  double *h = (double *)calloc(n_tmp_vectors, sizeof(double));
  const int ldh = n_tmp_vectors - 1;
  int accumulated = 0, state = 0; /* 0 iterate, 1 success, 2 failure */
  int re_orth = 0;
  double rho = 0;
  do
    {
      double *v = V; /* tmp_vectors[0] */
      /* p = b - A x ; v = P^-1 p */
      constrained_vmult(op, p, x);
      for (int i = 0; i < n; ++i) p[i] = b[i] - p[i];
      memcpy(v, p, sizeof(double) * n);
      if (P) band_solve(P, v);
      rho = sqrt(vdot(n, v, v));
      if (res_hist && accumulated < res_hist_len) res_hist[accumulated] = rho;
      state = (rho <= tol) ? 1 : ((accumulated >= max_steps) ? 2 : 0);
      if (!(rho == rho)) state = 2;
      if (state != 0) break;
      gamma[0] = rho;
      for (int i = 0; i < n; ++i) v[i] *= 1. / rho;
      int dim = 0;
      for (int inner = 0; inner < m && state == 0; ++inner)
        {
          ++accumulated;
          double *vv = V + (size_t)(inner + 1) * n;
          constrained_vmult(op, p, V + (size_t)inner * n);
          memcpy(vv, p, sizeof(double) * n);
          if (P) band_solve(P, vv);
          dim = inner + 1;
          /* modified Gram-Schmidt with the Kelley loss-of-orthogonality test every 5 its */
          const int consider = (re_orth == 0) && (inner % 5 == 4);
          double norm_vv_start = 0;
          if (consider) norm_vv_start = sqrt(vdot(n, vv, vv));
          for (int i = 0; i < dim; ++i)
            {
              h[i] = vdot(n, vv, V + (size_t)i * n);
              const double hi = h[i];
              const double *vi = V + (size_t)i * n;
              for (int k = 0; k < n; ++k) vv[k] += -hi * vi[k];
            }
          double s = sqrt(vdot(n, vv, vv));
          if (consider && norm_vv_start > 0)
            if (s < norm_vv_start * 10. * sqrt(DBL_EPSILON) * 1.) re_orth = 1;
          if (re_orth)
            {
              for (int i = 0; i < dim; ++i)
                {
                  double htmp = vdot(n, vv, V + (size_t)i * n);
                  h[i] += htmp;
                  const double *vi = V + (size_t)i * n;
                  for (int k = 0; k < n; ++k) vv[k] += -htmp * vi[k];
                }
              s = sqrt(vdot(n, vv, vv));
            }
          h[inner + 1] = s;
          for (int k = 0; k < n; ++k) vv[k] *= 1. / s;
          givens_rotation(h, gamma, ci, si, inner);
          for (int i = 0; i < dim; ++i) H[(size_t)i * ldh + inner] = h[i];
          rho = fabs(gamma[dim]);
          if (res_hist && accumulated < res_hist_len) res_hist[accumulated] = rho;
          state = (rho <= tol) ? 1 : ((accumulated >= max_steps) ? 2 : 0);
          if (!(rho == rho)) state = 2;
        }
      /* back substitution H1 y = gamma ; x += V y */
      for (int i = dim - 1; i >= 0; --i)
        {
          double s = gamma[i];
          for (int k = i + 1; k < dim; ++k) s -= H[(size_t)i * ldh + k] * h[k];
          h[i] = s / H[(size_t)i * ldh + i];
        }
      for (int i = 0; i < dim; ++i)
        {
          const double hi = h[i];
          const double *vi = V + (size_t)i * n;
          for (int k = 0; k < n; ++k) x[k] += hi * vi[k];
        }
    }
  while (state == 0);
  if (iters) *iters = accumulated;
  if (last_res) *last_res = rho;
  free(V); free(p); free(H); free(gamma); free(ci); free(si); free(h);
  return state == 1 ? 0 : 1;
}

/* BEMProblem<3>::solve_system, Direct branch (reference source/bem_problem.cc:821-895).
 * Inputs: assembled matrices, masks, boundary data tmp_rhs, the ConstraintMatrix lines
 * produced by compute_constraints (host code, :990-1105).  In/out: phi, dphi_dn (only
 * the unknown half is overwritten, :869-879).  Also returns alpha, system_rhs, sol. */
int orc_solve_system(int N, const double *Nm, const double *Dm, const double *surface_nodes,
                     const double *other_nodes, const double *tmp_rhs,
                     const int32_t *con_line_of, const uint32_t *con_ptr,
                     const uint32_t *con_col, const double *con_val, const double *con_inhom,
                     double tol, int max_steps, int n_tmp_vectors, int band, int use_precond,
                     double *phi, double *dphi_dn, double *alpha_out, double *rhs_out,
                     double *sol_out, int *iters, double *last_res, double *res_hist,
                     int res_hist_len, int nthreads)
{
  double *alpha = (double *)calloc(N, sizeof(double));
  double *rhs = (double *)calloc(N, sizeof(double));
  double *sol = (double *)calloc(N, sizeof(double));
  orc_compute_alpha(N, N, Nm, alpha, nthreads);
  orc_compute_rhs(N, Nm, Dm, alpha, surface_nodes, other_nodes, rhs, tmp_rhs, nthreads);
  orc_distribute_rhs(N, con_line_of, con_inhom, rhs);
  orc_op op = {N, Nm, Dm, alpha, surface_nodes, other_nodes, con_line_of, con_ptr, con_col,
               con_val, con_inhom, nthreads, 0};
  orc_band *P = NULL;
  if (use_precond)
    {
      P = build_preconditioner(&op, band);
      if (band_factor(P)) { band_free(P); free(alpha); free(rhs); free(sol); return -2; }
    }
  int rc = gmres_left(&op, P, n_tmp_vectors, tol, max_steps, sol, rhs, iters, last_res,
                      res_hist, res_hist_len);
  for (int i = 0; i < N; ++i)
    {
      if (surface_nodes[i] == 0)
        phi[i] = sol[i];
      else
        dphi_dn[i] = sol[i];
    }
  if (alpha_out) memcpy(alpha_out, alpha, sizeof(double) * N);
  if (rhs_out) memcpy(rhs_out, rhs, sizeof(double) * N);
  if (sol_out) memcpy(sol_out, sol, sizeof(double) * N);
  band_free(P);
  free(alpha); free(rhs); free(sol);
  return rc;
}

/* apply the band preconditioner alone (tests) */
int orc_precond_apply(int N, const double *Nm, const double *Dm, const double *alpha,
                      const double *surface_nodes, const int32_t *con_line_of, int band,
                      const double *in, double *out)
{
  orc_op op = {N, Nm, Dm, alpha, surface_nodes, NULL, con_line_of, NULL, NULL, NULL, NULL, 1, 0};
  orc_band *P = build_preconditioner(&op, band);
  if (band_factor(P)) { band_free(P); return -2; }
  memcpy(out, in, sizeof(double) * N);
  band_solve(P, out);
  band_free(P);
  return 0;
}

/* BEMProblem<3>::residual, Direct branch (reference source/bem_problem.cc:903-961). */
void orc_residual(int N, const double *Nm, const double *Dm, const double *surface_nodes,
                  const double *other_nodes, const int32_t *con_line_of,
                  const uint32_t *con_ptr, const uint32_t *con_col, const double *con_val,
                  const double *con_inhom, const double *phi, const double *dphi_dn,
                  double *res, int nthreads)
{
  double *alpha = (double *)calloc(N, sizeof(double));
  double *tmp = (double *)calloc(N, sizeof(double));
  double *rrhs = (double *)calloc(N, sizeof(double));
  double *sol = (double *)calloc(N, sizeof(double));
  orc_compute_alpha(N, N, Nm, alpha, nthreads);
  for (int i = 0; i < N; ++i) tmp[i] = dphi_dn[i] * other_nodes[i] + phi[i] * surface_nodes[i];
  orc_compute_rhs(N, Nm, Dm, alpha, surface_nodes, other_nodes, rrhs, tmp, nthreads);
  orc_distribute_rhs(N, con_line_of, con_inhom, rrhs);
  for (int i = 0; i < N; ++i) rrhs[i] *= -1;
  for (int i = 0; i < N; ++i) sol[i] = dphi_dn[i] * surface_nodes[i] + phi[i] * other_nodes[i];
  orc_op op = {N, Nm, Dm, alpha, surface_nodes, other_nodes, con_line_of, con_ptr, con_col,
               con_val, con_inhom, nthreads, 0};
  constrained_vmult(&op, res, sol);
  for (int i = 0; i < N; ++i) res[i] += rrhs[i];
  free(alpha); free(tmp); free(rrhs); free(sol);
}

/* ------------------------------------------------------------------------------------
 * L2 projections run by compute_constraints on every solve_system (reference
 * source/bem_problem.cc:996-997):
 *   ComputationalDomain<3>::compute_normals        source/computational_domain.cc:1525-1620
 *   BEMProblem<3>::compute_surface_gradients       source/bem_problem.cc:1153-1293
 * Both assemble, cell by cell, the mass matrix of FESystem(FE_Q(1),3) -- component-wise the
 * scalar Q1 mass matrix, (:1576-1588 / :1246-1258) -- and a right-hand side
 *   int phi_i n_d dS        (:1589-1590)          resp.   int phi_i (grad_s phi)_d dS  (:1259-1260)
 * and solve with SparseDirectUMFPACK.  Restated with a dense Cholesky factorisation per
 * connected mesh component (patches do not couple: their edge dofs are double nodes), which
 * is the same exact solve up to rounding.  which = 0: normals (normalised, :1613),
 * which = 1: surface gradients of the nodal field phi[N] (= tmp_rhs o surface_nodes, :1157-1158).
 * The surface gradient in a quadrature point is deal.II's covariant transformation of the
 * reference gradients: grad_s f = DX_t^T G^-1 [d_u f, d_v f].
 * ---------------------------------------------------------------------------------- */
static int uf_find(int *parent, int i)
{
  while (parent[i] != i)
    {
      parent[i] = parent[parent[i]];
      i = parent[i];
    }
  return i;
}

int orc_l2_projection(int which, int N, int C, const double *xyz, const uint32_t *cell_dofs,
                      const uint8_t *cell_dir_flag, int quad_order, const double *phi, double *out)
{
  const int nq = quad_order * quad_order;
  double *uv = (double *)malloc(sizeof(double) * 2 * nq), *w = (double *)malloc(sizeof(double) * nq);
  if (orc_qgauss2(quad_order, uv, w)) return -1;
  /* connected components */
  int *parent = (int *)malloc(sizeof(int) * N);
  for (int i = 0; i < N; ++i) parent[i] = i;
  for (int c = 0; c < C; ++c)
    for (int j = 1; j < 4; ++j)
      {
        int a = uf_find(parent, (int)cell_dofs[4 * c]), b = uf_find(parent, (int)cell_dofs[4 * c + j]);
        if (a != b) parent[b] = a;
      }
  int *comp = (int *)malloc(sizeof(int) * N), *local = (int *)malloc(sizeof(int) * N);
  int ncomp = 0;
  int *root_id = (int *)malloc(sizeof(int) * N);
  for (int i = 0; i < N; ++i) root_id[i] = -1;
  int *csize = (int *)calloc((size_t)N, sizeof(int));
  for (int i = 0; i < N; ++i)
    {
      int r = uf_find(parent, i);
      if (root_id[r] < 0) root_id[r] = ncomp++;
      comp[i] = root_id[r];
      local[i] = csize[comp[i]]++;
    }
  double **M = (double **)calloc((size_t)ncomp, sizeof(double *));
  double **B = (double **)calloc((size_t)ncomp, sizeof(double *));
  for (int k = 0; k < ncomp; ++k)
    {
      M[k] = (double *)calloc((size_t)csize[k] * csize[k], sizeof(double));
      B[k] = (double *)calloc((size_t)csize[k] * 3, sizeof(double));
    }
  /* cell loop (:1566-1600 / :1229-1273) */
  for (int c = 0; c < C; ++c)
    {
      const uint32_t *dofs = cell_dofs + 4 * c;
      double X[12];
      for (int k = 0; k < 4; ++k)
        for (int d = 0; d < 3; ++d) X[3 * k + d] = xyz[3 * (size_t)dofs[k] + d];
      double lm[4][4] = {{0}}, lr[4][3] = {{0}};
      for (int q = 0; q < nq; ++q)
        {
          double ph[4], du[4], dv[4];
          shape_q1(uv[2 * q], uv[2 * q + 1], ph, du, dv);
          double t0[3] = {0, 0, 0}, t1[3] = {0, 0, 0};
          for (int k = 0; k < 4; ++k)
            for (int d = 0; d < 3; ++d)
              {
                t0[d] += du[k] * X[3 * k + d];
                t1[d] += dv[k] * X[3 * k + d];
              }
          const double g00 = t0[0] * t0[0] + t0[1] * t0[1] + t0[2] * t0[2];
          const double g01 = t0[0] * t1[0] + t0[1] * t1[1] + t0[2] * t1[2];
          const double g11 = t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2];
          const double detG = g00 * g11 - g01 * g01;
          const double jxw = sqrt(detG) * w[q];
          double vec[3];
          if (which == 0)
            {
              const double cr[3] = {t0[1] * t1[2] - t0[2] * t1[1], t0[2] * t1[0] - t0[0] * t1[2],
                                    t0[0] * t1[1] - t0[1] * t1[0]};
              const double cn = sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
              const double sgn = cell_dir_flag[c] ? 1.0 : -1.0;
              for (int d = 0; d < 3; ++d) vec[d] = sgn * (cr[d] / cn);
            }
          else
            { /* fe_v.get_function_gradients(phi, phi_surf_grads) (:1236): sum_k phi_k grad_s(shape_k) */
              for (int d = 0; d < 3; ++d) vec[d] = 0.0;
              for (int k = 0; k < 4; ++k)
                {
                  const double a = (g11 * du[k] - g01 * dv[k]) / detG, b = (g00 * dv[k] - g01 * du[k]) / detG;
                  for (int d = 0; d < 3; ++d) vec[d] += phi[dofs[k]] * (t0[d] * a + t1[d] * b);
                }
            }
          for (int i = 0; i < 4; ++i)
            {
              for (int j = 0; j < 4; ++j) lm[i][j] += ph[i] * ph[j] * jxw;
              for (int d = 0; d < 3; ++d) lr[i][d] += ph[i] * vec[d] * jxw;
            }
        }
      const int k = comp[dofs[0]], n = csize[k];
      for (int i = 0; i < 4; ++i)
        {
          for (int j = 0; j < 4; ++j) M[k][(size_t)local[dofs[i]] * n + local[dofs[j]]] += lm[i][j];
          for (int d = 0; d < 3; ++d) B[k][(size_t)local[dofs[i]] * 3 + d] += lr[i][d];
        }
    }
  /* Cholesky M = L L^T per component, three right-hand sides */
  int rc = 0;
  for (int k = 0; k < ncomp && !rc; ++k)
    {
      const int n = csize[k];
      double *A = M[k], *b = B[k];
      if (n == 1 && A[0] == 0.0)
        { /* a dof no cell touches */
          b[0] = b[1] = b[2] = 0.0;
          continue;
        }
      for (int j = 0; j < n && !rc; ++j)
        {
          double s = A[(size_t)j * n + j];
          for (int t = 0; t < j; ++t) s -= A[(size_t)j * n + t] * A[(size_t)j * n + t];
          if (!(s > 0.0)) { rc = -2; break; }
          const double ljj = sqrt(s);
          A[(size_t)j * n + j] = ljj;
          for (int i = j + 1; i < n; ++i)
            {
              double v = A[(size_t)i * n + j];
              for (int t = 0; t < j; ++t) v -= A[(size_t)i * n + t] * A[(size_t)j * n + t];
              A[(size_t)i * n + j] = v / ljj;
            }
        }
      for (int d = 0; d < 3 && !rc; ++d)
        {
          for (int i = 0; i < n; ++i)
            {
              double v = b[(size_t)i * 3 + d];
              for (int t = 0; t < i; ++t) v -= A[(size_t)i * n + t] * b[(size_t)t * 3 + d];
              b[(size_t)i * 3 + d] = v / A[(size_t)i * n + i];
            }
          for (int i = n - 1; i >= 0; --i)
            {
              double v = b[(size_t)i * 3 + d];
              for (int t = i + 1; t < n; ++t) v -= A[(size_t)t * n + i] * b[(size_t)t * 3 + d];
              b[(size_t)i * 3 + d] = v / A[(size_t)i * n + i];
            }
        }
    }
  for (int i = 0; i < N && !rc; ++i)
    {
      double v[3];
      for (int d = 0; d < 3; ++d) v[d] = B[comp[i]][(size_t)local[i] * 3 + d];
      if (which == 0)
        { /* nodes_normals[i] /= |nodes_normals[i]| (:1613) */
          const double nn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
          for (int d = 0; d < 3; ++d) v[d] /= nn;
        }
      for (int d = 0; d < 3; ++d) out[3 * (size_t)i + d] = v[d];
    }
  for (int k = 0; k < ncomp; ++k) { free(M[k]); free(B[k]); }
  free(M); free(B); free(uv); free(w); free(parent); free(comp); free(local); free(root_id); free(csize);
  return rc;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
#endif /* !ORC_LONG_DOUBLE */
