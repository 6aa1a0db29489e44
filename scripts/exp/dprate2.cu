// dprate2.cu -- does the operand-reuse cache survive a warp switch?  x_c = fma(x_c, a, b) with NC chains per warp and W warps
// per SM sub-partition: with few chains per warp the scheduler has to rotate between warps after every NC instructions.
// Also: dependent-issue latency of DFMA (1 warp, 1 chain).
#include <cstdio>
#include <cuda_runtime.h>

template <int NC, int P>
__global__ void k(double *out, int iters, double a, double b) {
  double x[NC], d[NC], m[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) { x[c] = threadIdx.x + c; d[c] = a + 1e-9 * c * (threadIdx.x + 1); m[c] = b * (c + 1) + 1e-12 * threadIdx.x; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 48 / NC; u++) {
#pragma unroll
      for (int c = 0; c < NC; c++) {
        if (P == 0) x[c] = fma(x[c], d[0], m[0]);
        else if (P == 1) x[c] = fma(x[c], 0.99951171875, m[0]);
        else x[c] = fma(x[c], d[c], m[c]);
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int c = 0; c < NC; c++) r += x[c];
  if (r == 123.456) out[0] = r;
}

template <int NC, int P>
static void run(const char *name, double *out, int wps) {
  const int tps = 128 * wps;  // threads per SM
  const int threads = tps > 1024 ? 1024 : tps, blocks = 148 * (tps / threads);
  const int iters = 40000 / wps;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NC, P><<<blocks, threads>>>(out, 100, 0.999999, 1e-9);
  float best = 1e30f;
  for (int t = 0; t < 3; t++) {
    cudaEventRecord(e0); k<NC, P><<<blocks, threads>>>(out, iters, 0.999999, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double inst = (double)blocks * threads * (48 / NC * NC) * (double)iters;
  const double frac = inst / (best * 1e-3) / (148 * 64 * 1.965e9);
  // cycles between two instructions of the same warp = 2 / frac * wps (pipe cycles per warp-instruction x warps sharing it)
  printf("%-22s chains/warp %2d warps/SMSP %2d  %.4f of peak lane-inst/s; %.1f cycles per instruction per warp (%s)\n", name, NC, wps, frac,
         2.0 / frac * wps, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  double *out; cudaMalloc(&out, 1024);
  run<1, 1>("x=fma(x,imm,m0)", out, 1);   // latency: 1 chain, 1 warp per SMSP
  run<1, 0>("x=fma(x,d0,m0)", out, 1);
  run<1, 5>("x=fma(x,d_c,m_c)", out, 1);
  run<2, 1>("x=fma(x,imm,m0)", out, 16);
  run<2, 0>("x=fma(x,d0,m0)", out, 16);
  run<2, 5>("x=fma(x,d_c,m_c)", out, 16);
  run<1, 1>("x=fma(x,imm,m0)", out, 16);
  run<1, 0>("x=fma(x,d0,m0)", out, 16);
  run<4, 1>("x=fma(x,imm,m0)", out, 8);
  run<4, 0>("x=fma(x,d0,m0)", out, 8);
  run<4, 0>("x=fma(x,d0,m0)", out, 16);
  run<8, 0>("x=fma(x,d0,m0)", out, 4);
  run<8, 0>("x=fma(x,d0,m0)", out, 8);
  run<16, 0>("x=fma(x,d0,m0)", out, 2);
  run<16, 0>("x=fma(x,d0,m0)", out, 4);
  return 0;
}
