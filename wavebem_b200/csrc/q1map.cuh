// q1map.cuh -- Q1 codimension-one mapping of a surface quad (deal.II FE_Q<2,3>(1) with a
// MappingQ1-type mapping): position, d_u x d_v and shape values at a
// reference point.  Shared by assemble.cu and constraints.cu.
#pragma once
#include <stdint.h>

struct QuadVerts
{
  double x[4][3];
};

__device__ __forceinline__ void load_verts(const double *__restrict__ xyz,
                                           const uint32_t *__restrict__ dofs, QuadVerts &X)
{
#pragma unroll
  for (int k = 0; k < 4; ++k)
    {
      const double *p = xyz + 3 * (size_t)dofs[k];
      X.x[k][0] = p[0];
      X.x[k][1] = p[1];
      X.x[k][2] = p[2];
    }
}

__device__ __forceinline__ void q1_tangents(const QuadVerts &X, double u, double v, double tu[3], double tv[3])
{
#pragma unroll
  for (int d = 0; d < 3; ++d)
    {
      tu[d] = (1 - v) * (X.x[1][d] - X.x[0][d]) + v * (X.x[3][d] - X.x[2][d]);
      tv[d] = (1 - u) * (X.x[2][d] - X.x[0][d]) + u * (X.x[3][d] - X.x[1][d]);
    }
}

__device__ __forceinline__ void map_q1(const QuadVerts &X, double u, double v, double y[3],
                                       double cr[3], double phi[4])
{
  phi[0] = (1 - u) * (1 - v);
  phi[1] = u * (1 - v);
  phi[2] = (1 - u) * v;
  phi[3] = u * v;
  double tu[3], tv[3];
  q1_tangents(X, u, v, tu, tv);
#pragma unroll
  for (int d = 0; d < 3; ++d)
    y[d] = phi[0] * X.x[0][d] + phi[1] * X.x[1][d] + phi[2] * X.x[2][d] + phi[3] * X.x[3][d];
  cr[0] = tu[1] * tv[2] - tu[2] * tv[1];
  cr[1] = tu[2] * tv[0] - tu[0] * tv[2];
  cr[2] = tu[0] * tv[1] - tu[1] * tv[0];
}
