// postproc.cu -- the post-processing integrals that share the assembly's integrand.
//
//   FreeSurface<3>::compute_internal_velocities   (reference source/free_surface.cc:10426-10537)
//       v(x_i) = sum_cells sum_q [ dphi_dn(q) grad_x G - phi(q) grad_x dG/dn ] JxW(q)
//       G = 1/(4 pi |r|), dG/dn = -(r.n)/(4 pi |r|^3), r = y_q - x_i; the reference differentiates
//       with Sacado (fad_double, :10495-10521), here the gradients are written out:
//         grad_x G     =  r / (4 pi |r|^3)
//         grad_x dG/dn = [ n / |r|^3 - 3 (r.n) r / |r|^5 ] / (4 pi)
//       Same panel records as the assembly (y_q, n JxW/(-4 pi), JxW/(4 pi)); points x cells x 16.
//
//   the hull integrals of FreeSurface<3>::compute_pressure   (source/free_surface.cc:9534-9598)
//       gradient = n dphi_dn + grad_s phi ;  press = rho Vinf^2/2 - rho |gradient + Vinf|^2/2 - rho g z
//       press2   = -rho Vinf.gradient - rho |gradient|^2/2 - rho g z
//       press_force_test_1/2 += press(2) n JxW ,  press_moment += press (y - baricenter) x n JxW
//       over the cells the caller marks (material_id == wall_sur_ID1..3, :9580-9582); steady terms:
//       the DphiDt / node-velocity terms of :9556-9557 belong to the time integrator and are zero in
//       the steady problem this library solves.
// Both are O(points x cells) / O(cells) and read no matrix: they run on the first row block only.
#include <algorithm>
#include <cstdio>
#include <vector>

#include "internal.h"
#include "q1map.cuh"

#define PP_POINTS 128 // points per CTA
#define PP_CELLS 8    // cells staged per pass
#define GEO_REC 8

// values of a nodal field at the 16 Gauss points of every cell (processing order): q_f[p][q]
__global__ void k_pp_qvalues(uint32_t C, const uint32_t *__restrict__ cell_dofs, const double *__restrict__ g1,
                             const double *__restrict__ phi, const double *__restrict__ dphi,
                             double *__restrict__ qphi, double *__restrict__ qdphi)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t p = t >> 4;
  const int q = t & 15;
  if (p >= C) return;
  const double u = g1[q & 3], v = g1[q >> 2];
  const double s0 = (1 - u) * (1 - v), s1 = u * (1 - v), s2 = (1 - u) * v, s3 = u * v;
  const uint32_t *d = cell_dofs + 4 * (size_t)p;
  qphi[t] = s0 * phi[d[0]] + s1 * phi[d[1]] + s2 * phi[d[2]] + s3 * phi[d[3]];
  qdphi[t] = s0 * dphi[d[0]] + s1 * dphi[d[1]] + s2 * dphi[d[2]] + s3 * dphi[d[3]];
}

// partial[chunk][point][3]: one thread per point, the chunk's cells staged through shared memory
__global__ void __launch_bounds__(PP_POINTS)
  k_pp_velocities(uint32_t n_points, uint32_t C, uint32_t cells_per_chunk, const double *__restrict__ pts,
                  const double *__restrict__ geo, const double *__restrict__ qphi, const double *__restrict__ qdphi,
                  double *__restrict__ partial)
{
  __shared__ __align__(16) double sg[PP_CELLS][GEO_REC * 16];
  __shared__ double sp[PP_CELLS][16], sd[PP_CELLS][16];
  const uint32_t i = blockIdx.x * PP_POINTS + threadIdx.x;
  const uint32_t ic = i < n_points ? i : n_points - 1;
  const double x0 = pts[3 * (size_t)ic], x1 = pts[3 * (size_t)ic + 1], x2 = pts[3 * (size_t)ic + 2];
  const uint32_t cb = blockIdx.y * cells_per_chunk, ce = min(C, cb + cells_per_chunk);
  double v0 = 0, v1 = 0, v2 = 0;
  for (uint32_t c0 = cb; c0 < ce; c0 += PP_CELLS)
    {
      const int nc = (int)min((uint32_t)PP_CELLS, ce - c0);
      __syncthreads();
      for (int k = threadIdx.x; k < nc * GEO_REC * 16; k += PP_POINTS) (&sg[0][0])[k] = geo[(size_t)c0 * GEO_REC * 16 + k];
      for (int k = threadIdx.x; k < nc * 16; k += PP_POINTS)
        {
          (&sp[0][0])[k] = qphi[(size_t)c0 * 16 + k];
          (&sd[0][0])[k] = qdphi[(size_t)c0 * 16 + k];
        }
      __syncthreads();
      for (int k = 0; k < nc; ++k)
        {
          const double *g = sg[k];
#pragma unroll 4
          for (int q = 0; q < 16; ++q)
            {
              const double r0 = g[q] - x0, r1 = g[16 + q] - x1, r2 = g[32 + q] - x2;
              const double rr = r0 * r0 + r1 * r1 + r2 * r2;
              const double ri = rsqrt(rr);
              const double ri3 = ri * ri * ri;
              // nJ = n JxW / (-4 pi), wJ = JxW / (4 pi)
              const double nJ0 = g[48 + q], nJ1 = g[64 + q], nJ2 = g[80 + q], wJ = g[96 + q];
              const double rn = r0 * nJ0 + r1 * nJ1 + r2 * nJ2;
              const double ph = sp[k][q], dp = sd[k][q];
              // dphi grad G JxW - phi grad dG/dn JxW = r (dphi wJ / r^3 - 3 phi (r.nJ) / r^5) + phi nJ / r^3
              const double a = dp * wJ * ri3 - 3.0 * ph * rn * ri3 * (ri * ri);
              const double b = ph * ri3;
              v0 += a * r0 + b * nJ0;
              v1 += a * r1 + b * nJ1;
              v2 += a * r2 + b * nJ2;
            }
        }
    }
  if (i < n_points)
    {
      double *o = partial + ((size_t)blockIdx.y * n_points + i) * 3;
      o[0] = v0;
      o[1] = v1;
      o[2] = v2;
    }
}

__global__ void k_pp_reduce(uint32_t n, uint32_t n_chunks, const double *__restrict__ partial, double *__restrict__ out)
{ // out[j] = sum over chunks, fixed order
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s = 0;
  for (uint32_t c = 0; c < n_chunks; ++c) s += partial[(size_t)c * n + j];
  out[j] = s;
}

// one thread per cell (processing order): the 11 hull integrals of that cell (zero when unmarked)
__global__ void __launch_bounds__(128)
  k_pp_pressure_cells(uint32_t C, const double *__restrict__ xyz, const uint32_t *__restrict__ cell_dofs,
                      const uint8_t *__restrict__ dir, const uint32_t *__restrict__ cell_order,
                      const uint8_t *__restrict__ marked, const double *__restrict__ g1, const double *__restrict__ w1,
                      const double *__restrict__ phi, const double *__restrict__ dphi, double vx, double vy, double vz,
                      double rho, double grav, double bx, double by, double bz, double *__restrict__ out /* [11][C] */)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= C) return;
  double acc[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) acc[k] = 0.0;
  if (marked[cell_order[p]])
    {
      const uint32_t *d = cell_dofs + 4 * (size_t)p;
      QuadVerts X;
      load_verts(xyz, d, X);
      const double sgn = dir[p] ? 1.0 : -1.0;
      const double f0 = phi[d[0]], f1 = phi[d[1]], f2 = phi[d[2]], f3 = phi[d[3]];
      const double h0 = dphi[d[0]], h1 = dphi[d[1]], h2 = dphi[d[2]], h3 = dphi[d[3]];
      const double vv = vx * vx + vy * vy + vz * vz;
      for (int q = 0; q < 16; ++q)
        {
          const double u = g1[q & 3], v = g1[q >> 2], w = w1[q & 3] * w1[q >> 2];
          double y[3], cr[3], sh[4], tu[3], tv[3];
          map_q1(X, u, v, y, cr, sh);
          q1_tangents(X, u, v, tu, tv);
          const double cn = sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
          const double n0 = sgn * cr[0] / cn, n1 = sgn * cr[1] / cn, n2 = sgn * cr[2] / cn;
          const double jxw = cn * w;
          // surface gradient: [t_u t_v] G^-1 [d_u phi, d_v phi]^T
          const double E = tu[0] * tu[0] + tu[1] * tu[1] + tu[2] * tu[2];
          const double F = tu[0] * tv[0] + tu[1] * tv[1] + tu[2] * tv[2];
          const double G = tv[0] * tv[0] + tv[1] * tv[1] + tv[2] * tv[2];
          const double det = E * G - F * F;
          const double pu = (1 - v) * (f1 - f0) + v * (f3 - f2), pv = (1 - u) * (f2 - f0) + u * (f3 - f1);
          const double ca = (G * pu - F * pv) / det, cb = (E * pv - F * pu) / det;
          const double dn = sh[0] * h0 + sh[1] * h1 + sh[2] * h2 + sh[3] * h3;
          const double g0 = n0 * dn + ca * tu[0] + cb * tv[0];
          const double g1v = n1 * dn + ca * tu[1] + cb * tv[1];
          const double g2 = n2 * dn + ca * tu[2] + cb * tv[2];
          const double t0 = g0 + vx, t1 = g1v + vy, t2 = g2 + vz;
          const double press = rho * vv / 2 - rho * (t0 * t0 + t1 * t1 + t2 * t2) / 2 - rho * grav * y[2];
          const double press2 = -rho * (vx * g0 + vy * g1v + vz * g2) - rho * (g0 * g0 + g1v * g1v + g2 * g2) / 2 -
                                rho * grav * y[2];
          acc[0] += press * n0 * jxw;
          acc[1] += press * n1 * jxw;
          acc[2] += press * n2 * jxw;
          acc[3] += press2 * n0 * jxw;
          acc[4] += press2 * n1 * jxw;
          acc[5] += press2 * n2 * jxw;
          const double a0 = y[0] - bx, a1 = y[1] - by, a2 = y[2] - bz;
          acc[6] += press * (a1 * n2 - a2 * n1) * jxw;
          acc[7] += press * (a2 * n0 - a0 * n2) * jxw;
          acc[8] += press * (a0 * n1 - a1 * n0) * jxw;
          acc[9] += jxw;
          acc[10] += (sh[0] * f0 + sh[1] * f1 + sh[2] * f2 + sh[3] * f3) * jxw;
        }
    }
#pragma unroll
  for (int k = 0; k < 11; ++k) out[(size_t)k * C + p] = acc[k];
}

// out[k] = sum over cells of in[k][.], one CTA per k, fixed order (deterministic)
__global__ void __launch_bounds__(256) k_pp_sum_rows(uint32_t C, const double *__restrict__ in, double *__restrict__ out)
{
  __shared__ double red[256];
  const double *r = in + (size_t)blockIdx.x * C;
  double s = 0;
  for (uint32_t i = threadIdx.x; i < C; i += 256) s += r[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1)
    {
      if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
      __syncthreads();
    }
  if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

struct PpTables
{
  double g1[4], w1[4];
};

static int pp_prepare(wbem_ctx *ctx, double **d_tab)
{
  if (!ctx->N || !ctx->have_geometry) WBEM_FAIL(ctx, -3, "post-processing needs wbem_set_topology and wbem_set_geometry");
  if (ctx->qt.n1 != 4) WBEM_FAIL(ctx, -1, "post-processing kernels are written for the Gauss 4x4 rule (quad_order = 4)");
  PpTables t;
  // the 1-D rule: nodes from the tables, weights recovered from the 2-D ones (w_q = w1[qx] w1[qy], sum w1 = 1)
  for (int k = 0; k < 4; ++k) t.g1[k] = ctx->qt.g1_x[k];
  for (int k = 0; k < 4; ++k)
    {
      double s = 0;
      for (int j = 0; j < 4; ++j) s += ctx->qt.g_w[4 * j + k];
      t.w1[k] = s;
    }
  CUDA_OK(ctx, cudaMalloc((void **)d_tab, sizeof(t)));
  CUDA_OK(ctx, cudaMemcpyAsync(*d_tab, &t, sizeof(t), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

extern "C" {

int wbem_internal_velocities(wbem_ctx *ctx, const double *phi, const double *dphi_dn, uint32_t n_points,
                             const double *points, double *velocities)
{
  if (!ctx || !phi || !dphi_dn || (n_points && (!points || !velocities))) return -1;
  if (wbem_group_forward(ctx)) ctx = wbem_group_shard(ctx, 0); // no matrix involved: the first row block does it
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  if (n_points == 0) return 0;
  double *d_tab = nullptr;
  int rc = pp_prepare(ctx, &d_tab);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const uint32_t N = ctx->N, C = ctx->C;
  // the panel records follow the support points (they are rebuilt by every assembly anyway)
  if (!ctx->fevalues_given && (rc = wbem_launch_geometry(ctx))) return rc;
  const uint32_t pblocks = (n_points + PP_POINTS - 1) / PP_POINTS;
  uint32_t chunks = std::max(1u, (4u * 148u + pblocks - 1) / pblocks);
  uint32_t per = (C + chunks - 1) / chunks;
  per = std::max<uint32_t>(PP_CELLS, (per + PP_CELLS - 1) / PP_CELLS * PP_CELLS);
  chunks = (C + per - 1) / per;
  double *d_pts = nullptr, *d_q = nullptr, *d_part = nullptr, *d_out = nullptr;
  cudaError_t e = cudaMalloc((void **)&d_pts, sizeof(double) * 3 * (size_t)n_points);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_q, sizeof(double) * 32 * (size_t)C);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_part, sizeof(double) * 3 * (size_t)n_points * chunks);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_out, sizeof(double) * 3 * (size_t)n_points);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_pts, points, sizeof(double) * 3 * (size_t)n_points, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_tmp[3], phi, sizeof(double) * N, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_tmp[4], dphi_dn, sizeof(double) * N, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    {
      k_pp_qvalues<<<(C * 16 + 255) / 256, 256, 0, st>>>(C, ctx->d_cell_dofs, d_tab, ctx->d_tmp[3], ctx->d_tmp[4], d_q,
                                                        d_q + 16 * (size_t)C);
      k_pp_velocities<<<dim3(pblocks, chunks), PP_POINTS, 0, st>>>(n_points, C, per, d_pts, ctx->d_cellgeo, d_q,
                                                                  d_q + 16 * (size_t)C, d_part);
      k_pp_reduce<<<(3 * n_points + 255) / 256, 256, 0, st>>>(3 * n_points, chunks, d_part, d_out);
      ctx->launches += 3;
      e = cudaGetLastError();
    }
  if (e == cudaSuccess) e = cudaMemcpyAsync(velocities, d_out, sizeof(double) * 3 * (size_t)n_points, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_pts);
  cudaFree(d_q);
  cudaFree(d_part);
  cudaFree(d_out);
  cudaFree(d_tab);
  if (e != cudaSuccess) WBEM_FAIL(ctx, -2, "CUDA error %s in wbem_internal_velocities", cudaGetErrorString(e));
  return 0;
}

int wbem_pressure_force(wbem_ctx *ctx, const double *phi, const double *dphi_dn, const uint8_t *cell_marked,
                        const double *vinf, double rho, double g, const double *baricenter, double *out11)
{
  if (!ctx || !phi || !dphi_dn || !cell_marked || !vinf || !out11) return -1;
  if (wbem_group_forward(ctx)) ctx = wbem_group_shard(ctx, 0);
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  double *d_tab = nullptr;
  int rc = pp_prepare(ctx, &d_tab);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const uint32_t N = ctx->N, C = ctx->C;
  const double b0 = baricenter ? baricenter[0] : 0.0, b1 = baricenter ? baricenter[1] : 0.0, b2 = baricenter ? baricenter[2] : 0.0;
  uint8_t *d_mark = nullptr;
  double *d_cells = nullptr, *d_sum = nullptr;
  cudaError_t e = cudaMalloc((void **)&d_mark, std::max(1u, C));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_cells, sizeof(double) * 11 * (size_t)std::max(1u, C));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_sum, sizeof(double) * 11);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_mark, cell_marked, C, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_tmp[3], phi, sizeof(double) * N, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_tmp[4], dphi_dn, sizeof(double) * N, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && C)
    {
      const PpTables *t = reinterpret_cast<const PpTables *>(d_tab);
      k_pp_pressure_cells<<<(C + 127) / 128, 128, 0, st>>>(C, ctx->d_xyz, ctx->d_cell_dofs, ctx->d_dir, ctx->d_cell_order,
                                                          d_mark, t->g1, t->w1, ctx->d_tmp[3], ctx->d_tmp[4], vinf[0],
                                                          vinf[1], vinf[2], rho, g, b0, b1, b2, d_cells);
      k_pp_sum_rows<<<11, 256, 0, st>>>(C, d_cells, d_sum);
      ctx->launches += 2;
      e = cudaGetLastError();
    }
  else if (e == cudaSuccess)
    e = cudaMemsetAsync(d_sum, 0, sizeof(double) * 11, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out11, d_sum, sizeof(double) * 11, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_mark);
  cudaFree(d_cells);
  cudaFree(d_sum);
  cudaFree(d_tab);
  if (e != cudaSuccess) WBEM_FAIL(ctx, -2, "CUDA error %s in wbem_pressure_force", cudaGetErrorString(e));
  return 0;
}

} // extern "C"
