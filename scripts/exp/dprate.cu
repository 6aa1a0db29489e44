// dprate.cu -- FP64 issue rate by operand pattern AND warps per SM sub-partition (is the operand-reuse cache kept when
// the scheduler alternates between warps?).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o dprate dprate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int P>
__global__ void k(double *out, int iters, double a, double b) {
  constexpr int NC = 15;
  double x[NC], d[NC], m[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) { x[c] = threadIdx.x + c; d[c] = a + 1e-9 * c * (threadIdx.x + 1); m[c] = b * (c + 1) + 1e-12 * threadIdx.x; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int c = 0; c < NC; c++) {
        if (P == 0) x[c] = fma(x[c], d[0], m[0]);            // two shared operands
        else if (P == 1) x[c] = fma(x[c], 0.99951171875, m[0]);  // immediate + shared
        else if (P == 5) x[c] = fma(x[c], d[c], m[c]);       // three distinct
        else if (P == 6) x[c] = fma(x[c], d[0], m[c]);       // one shared multiplier for all
        else if (P == 7) x[c] = fma(d[c], m[c / 3], x[c]);   // acc += m_g * d_c, groups of 3 share m_g
        else if (P == 8) x[c] = fma(d[c], m[c / 5], x[c]);   // groups of 5
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int c = 0; c < NC; c++) r += x[c];
  if (r == 123.456) out[0] = r;
}

template <int P>
static void run(const char *name, double *out) {
  for (int wps = 1; wps <= 8; wps *= 2) {  // warps per sub-partition
    const int threads = 128 * wps > 1024 ? 1024 : 128 * wps, blocks = 148 * (128 * wps / threads);
    const int iters = 20000 / wps;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<P><<<blocks, threads>>>(out, 100, 0.999999, 1e-9);
    float best = 1e30f;
    for (int t = 0; t < 3; t++) {
      cudaEventRecord(e0); k<P><<<blocks, threads>>>(out, iters, 0.999999, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double inst = (double)blocks * threads * 60.0 * iters;
    printf("%-40s warps/SMSP %d  %.4f of 148x64x1.965e9 lane-inst/s (%s)\n", name, wps, inst / (best * 1e-3) / (148 * 64 * 1.965e9),
           cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  double *out; cudaMalloc(&out, 1024);
  run<1>("P1 x=fma(x,imm,m0)", out);
  run<0>("P0 x=fma(x,d0,m0)", out);
  run<6>("P6 x=fma(x,d0,m_c)", out);
  run<5>("P5 x=fma(x,d_c,m_c)", out);
  run<7>("P7 x_c=fma(d_c,m_g,x_c) groups of 3", out);
  run<8>("P8 x_c=fma(d_c,m_g,x_c) groups of 5", out);
  return 0;
}
