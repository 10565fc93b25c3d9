// probes.cu -- development micro-benchmarks of the FP64 issue path (NOT part of libwbem.so: built as
// lib/libwbem_probes.so by `python -m wavebem_b200.build --probes`, used by scripts/probe.py; DESIGN 4.1
// quotes their results).  Standalone: plain CUDA runtime, no library context.
#include <cuda_runtime.h>

#include <cstdio>

// issue-port probe: 8 independent DFMA chains + NI integer/LDS-free ALU ops per iteration
template <int NI>
__global__ void k_issue_probe(double *out, int *iout, int iters)
{
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-7;
  int x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
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
#pragma unroll
      for (int k = 0; k < NI; ++k)
        {
          if ((k & 3) == 0) x0 = (x0 ^ i) + 0x9e37;
          if ((k & 3) == 1) x1 = (x1 ^ i) + 0x79b9;
          if ((k & 3) == 2) x2 = (x2 ^ i) + 0x7f4a;
          if ((k & 3) == 3) x3 = (x3 ^ i) + 0x7c15;
        }
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  iout[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}


// opcode probe: 8 independent chains of one FP64 opcode (0 DFMA, 1 DADD, 2 DMUL, 3 = 4 DFMA + 4 DADD)
template <int OP>
__global__ void k_opcode_probe(double *out, int iters)
{
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-9 + k;
  const double b = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i)
    {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        {
          if (OP == 0 || (OP == 3 && (k & 1) == 0))
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[k]) : "d"(b), "d"(c));
          else if (OP == 1 || OP == 3)
            asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[k]) : "d"(c));
          else
            asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[k]) : "d"(b));
        }
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}


// DMMA probe: NF independent DFMA chains + NM independent m8n8k4 FP64 tensor-core MMAs per
// iteration -- does the FP64 tensor pipe run beside the FP64 FMA pipe on this chip?
template <int NF, int NM>
__global__ void k_dmma_probe(double *out, int iters)
{
  double a[8], c0[4], c1[4];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-9 + k;
#pragma unroll
  for (int k = 0; k < 4; ++k) c0[k] = c1[k] = 0.0;
  const double b = 1.0000001, c = 1e-7, ma = 1e-3 * (threadIdx.x & 3), mb = 1e-3 * (threadIdx.x >> 2);
  for (int i = 0; i < iters; ++i)
    {
#pragma unroll
      for (int k = 0; k < NF; ++k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[k]) : "d"(b), "d"(c));
#pragma unroll
      for (int k = 0; k < NM; ++k)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[k]), "+d"(c1[k])
                     : "d"(ma), "d"(mb));
    }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) s += c0[k] + c1[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}



extern "C" int wbem_probe(int device, int n_int, double *tflops)
{
  if (cudaSetDevice(device) != cudaSuccess) return -2;
  const int blocks = 148 * 4, threads = 512, iters = 1 << 14;
  double *d = nullptr;
  int *di = nullptr;
  if (cudaMalloc((void **)&d, sizeof(double) * blocks * threads) != cudaSuccess) return -2;
  if (cudaMalloc((void **)&di, sizeof(int) * blocks * threads) != cudaSuccess) return -2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaStream_t st = nullptr;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep)
    {
      cudaEventRecord(e0, st);
      switch (n_int)
        {
        case 0: k_issue_probe<0><<<blocks, threads, 0, st>>>(d, di, iters); break;
        case 2: k_issue_probe<2><<<blocks, threads, 0, st>>>(d, di, iters); break;
        case 4: k_issue_probe<4><<<blocks, threads, 0, st>>>(d, di, iters); break;
        case 8: k_issue_probe<8><<<blocks, threads, 0, st>>>(d, di, iters); break;
        case 16: k_issue_probe<16><<<blocks, threads, 0, st>>>(d, di, iters); break;
        case 100: k_opcode_probe<0><<<blocks, threads, 0, st>>>(d, iters); break; // DFMA only
        case 101: k_opcode_probe<1><<<blocks, threads, 0, st>>>(d, iters); break; // DADD only
        case 102: k_opcode_probe<2><<<blocks, threads, 0, st>>>(d, iters); break; // DMUL only
        case 103: k_opcode_probe<3><<<blocks, threads, 0, st>>>(d, iters); break; // DFMA/DADD alternating
        case 104: k_dmma_probe<0, 4><<<blocks, threads, 0, st>>>(d, iters); break; // 4 DMMA, no DFMA
        case 105: k_dmma_probe<8, 1><<<blocks, threads, 0, st>>>(d, iters); break; // 8 DFMA + 1 DMMA
        case 106: k_dmma_probe<8, 2><<<blocks, threads, 0, st>>>(d, iters); break; // 8 DFMA + 2 DMMA
        case 107: k_dmma_probe<8, 0><<<blocks, threads, 0, st>>>(d, iters); break; // 8 DFMA (same loop)
        default: cudaFree(d); cudaFree(di); return -1;
        }
      cudaEventRecord(e1, st);
      cudaStreamSynchronize(st);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 1 && ms < best) best = ms;
    }
  cudaFree(d);
  cudaFree(di);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
  // 104: report the tensor-pipe rate itself (4 MMAs of 8x8x4 FMAs per warp and iteration)
  if (n_int == 104) *tflops = 2.0 * 4.0 * 256.0 * (double)iters * blocks * (threads / 32) / (best * 1e-3) / 1e12;
  return 0;
}
