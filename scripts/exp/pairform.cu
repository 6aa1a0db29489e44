// pairform.cu -- stand-alone probe of source forms of the Hermite pair interaction (FP64 operand-fetch cost).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -lineinfo -o pairform pairform.cu
// Each variant: 296 CTAs x THREADS, lanes hold IPT i-particles, every warp streams the j tile from shared memory
// (broadcast LDS.128), as k_force does.  Prints pairs/s and the fraction of 148 x 64 x clk / 32 DP per pair.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

struct Acc7 { double ax, ay, az, jx, jy, jz, pot, bx, by, bz; };

__device__ __forceinline__ double rsq_seed(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  int hi = __double2hiint(x);
  int yhi = (hi == 0) ? 0 : __double2hiint(y0);
  return __hiloint2double(yhi, 0);
}

template <int F>
__device__ __forceinline__ void pair(const double4 pj, const double4 vj, const double eps2, const double xi,
                                     const double yi, const double zi, const double vxi, const double vyi,
                                     const double vzi, Acc7 &s) {
  const double dx = pj.x - xi, dy = pj.y - yi, dz = pj.z - zi;
  const double dvx = vj.x - vxi, dvy = vj.y - vyi, dvz = vj.z - vzi;
  if (F == 0) {  // the shipped form
    const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
    const double rv = fma(dx, dvx, fma(dy, dvy, dz * dvz));
    const double y0 = rsq_seed(r2);
    const double y2 = y0 * y0;
    const double e = fma(-r2, y2, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double ye = y0 * e;
    const double rinv = fma(ye, p, y0);
    const double rinv2 = rinv * rinv;
    const double mrinv = pj.w * rinv;
    const double mrinv3 = mrinv * rinv2;
    const double al = (-3.0 * rv) * rinv2;
    s.pot -= mrinv;
    s.ax = fma(mrinv3, dx, s.ax);
    s.ay = fma(mrinv3, dy, s.ay);
    s.az = fma(mrinv3, dz, s.az);
    s.jx = fma(mrinv3, fma(al, dx, dvx), s.jx);
    s.jy = fma(mrinv3, fma(al, dy, dvy), s.jy);
    s.jz = fma(mrinv3, fma(al, dz, dvz), s.jz);
  } else if (F == 1) {  // rsqrt tail as a product: no three-register FMA in the refinement
    const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
    const double rv = fma(dx, dvx, fma(dy, dvy, dz * dvz));
    const double y0 = rsq_seed(r2);
    const double y2 = y0 * y0;
    const double e = fma(-r2, y2, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double q = fma(e, p, 1.0);
    const double rinv = y0 * q;
    const double rinv2 = rinv * rinv;
    const double mrinv = pj.w * rinv;
    const double mrinv3 = mrinv * rinv2;
    const double al = (-3.0 * rv) * rinv2;
    s.pot -= mrinv;
    s.ax = fma(mrinv3, dx, s.ax);
    s.ay = fma(mrinv3, dy, s.ay);
    s.az = fma(mrinv3, dz, s.az);
    s.jx = fma(mrinv3, fma(al, dx, dvx), s.jx);
    s.jy = fma(mrinv3, fma(al, dy, dvy), s.jy);
    s.jz = fma(mrinv3, fma(al, dz, dvz), s.jz);
  } else if (F == 2) {  // as 1, r.v interleaved with r2 so that dy, dz can stay in the reuse cache
    const double r2a = fma(dz, dz, eps2);
    const double rva = dz * dvz;
    const double r2b = fma(dy, dy, r2a);
    const double rvb = fma(dy, dvy, rva);
    const double r2 = fma(dx, dx, r2b);
    const double rv = fma(dx, dvx, rvb);
    const double y0 = rsq_seed(r2);
    const double y2 = y0 * y0;
    const double e = fma(-r2, y2, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double q = fma(e, p, 1.0);
    const double rinv = y0 * q;
    const double rinv2 = rinv * rinv;
    const double mrinv = pj.w * rinv;
    const double mrinv3 = mrinv * rinv2;
    const double al = (-3.0 * rv) * rinv2;
    s.pot -= mrinv;
    const double tx = fma(al, dx, dvx), ty = fma(al, dy, dvy), tz = fma(al, dz, dvz);
    s.ax = fma(mrinv3, dx, s.ax);
    s.ay = fma(mrinv3, dy, s.ay);
    s.az = fma(mrinv3, dz, s.az);
    s.jx = fma(mrinv3, tx, s.jx);
    s.jy = fma(mrinv3, ty, s.jy);
    s.jz = fma(mrinv3, tz, s.jz);
  } else if (F == 3 || F == 4 || F == 5 || F == 6) {
    // jerk split into two sums: jA += mr3 dv, jB += (mr3 rv / r^2) dx, jerk = jA - 3 jB at the end: no tmp -> jerk dependency,
    // six FMAs in a row share mr3, three share c.  4: + rsqrt tail as a product.  5: as 3 without pot.  6: as 4 without pot.
    const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
    const double rv = fma(dx, dvx, fma(dy, dvy, dz * dvz));
    const double y0 = rsq_seed(r2);
    const double y2 = y0 * y0;
    const double e = fma(-r2, y2, 1.0);
    const double p = fma(0.375, e, 0.5);
    double rinv;
    if (F == 3 || F == 5) { const double ye = y0 * e; rinv = fma(ye, p, y0); }
    else { const double q = fma(e, p, 1.0); rinv = y0 * q; }
    const double rinv2 = rinv * rinv;
    const double mrinv = pj.w * rinv;
    const double mrinv3 = mrinv * rinv2;
    const double c = (rv * rinv2) * mrinv3;
    if (F == 3 || F == 4) s.pot -= mrinv;
    s.ax = fma(mrinv3, dx, s.ax);
    s.ay = fma(mrinv3, dy, s.ay);
    s.az = fma(mrinv3, dz, s.az);
    s.jx = fma(mrinv3, dvx, s.jx);
    s.jy = fma(mrinv3, dvy, s.jy);
    s.jz = fma(mrinv3, dvz, s.jz);
    s.bx = fma(c, dx, s.bx);
    s.by = fma(c, dy, s.by);
    s.bz = fma(c, dz, s.bz);
  } else if (F == 7) {  // the shipped form without pot
    const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
    const double rv = fma(dx, dvx, fma(dy, dvy, dz * dvz));
    const double y0 = rsq_seed(r2);
    const double y2 = y0 * y0;
    const double e = fma(-r2, y2, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double ye = y0 * e;
    const double rinv = fma(ye, p, y0);
    const double rinv2 = rinv * rinv;
    const double mrinv3 = (pj.w * rinv) * rinv2;
    const double al = (-3.0 * rv) * rinv2;
    s.ax = fma(mrinv3, dx, s.ax);
    s.ay = fma(mrinv3, dy, s.ay);
    s.az = fma(mrinv3, dz, s.az);
    s.jx = fma(mrinv3, fma(al, dx, dvx), s.jx);
    s.jy = fma(mrinv3, fma(al, dy, dvy), s.jy);
    s.jz = fma(mrinv3, fma(al, dz, dvz), s.jz);
  }
}

template <int F, int THREADS, int MINB, int IPT, int UNR, int TJ>
__global__ void __launch_bounds__(THREADS, MINB) k_probe(const double4 *__restrict__ jp, const double4 *__restrict__ jv,
                                                         double *out, int reps, double eps2) {
  __shared__ double4 sp[TJ], sv[TJ];
  for (int k = threadIdx.x; k < TJ; k += THREADS) { sp[k] = jp[k]; sv[k] = jv[k]; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int WARPS = THREADS / 32;
  double xi[IPT], yi[IPT], zi[IPT], vxi[IPT], vyi[IPT], vzi[IPT];
  Acc7 s[IPT];
#pragma unroll
  for (int q = 0; q < IPT; q++) {
    const int id = (blockIdx.x * 32 + lane) * IPT + q;
    const double4 pi = jp[256 + id % 1024], vi = jv[256 + id % 1024];
    xi[q] = pi.x + 0.001 * id; yi[q] = pi.y; zi[q] = pi.z; vxi[q] = vi.x; vyi[q] = vi.y; vzi[q] = vi.z;
    s[q].ax = s[q].ay = s[q].az = s[q].jx = s[q].jy = s[q].jz = s[q].pot = s[q].bx = s[q].by = s[q].bz = 0.0;
  }
  for (int r = 0; r < reps; r++) {
#pragma unroll UNR
    for (int jj = warp; jj < TJ; jj += WARPS) {
      const double4 pj = sp[jj];
      const double4 vj = sv[jj];
#pragma unroll
      for (int q = 0; q < IPT; q++) pair<F>(pj, vj, eps2, xi[q], yi[q], zi[q], vxi[q], vyi[q], vzi[q], s[q]);
    }
  }
  double a = 0;
#pragma unroll
  for (int q = 0; q < IPT; q++) a += s[q].ax + s[q].ay + s[q].az + s[q].jx + s[q].jy + s[q].jz + s[q].pot + s[q].bx + s[q].by + s[q].bz;
  out[blockIdx.x * THREADS + threadIdx.x] = a;
}

template <int F, int THREADS, int MINB, int IPT, int UNR>
static void run(const char *name, const double4 *jp, const double4 *jv, double *out, double clk_ghz) {
  constexpr int TJ = 256;
  const int grid = 148 * MINB, reps = 400;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_probe<F, THREADS, MINB, IPT, UNR, TJ><<<grid, THREADS>>>(jp, jv, out, 20, 0.0);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int t = 0; t < 5; t++) {
    cudaEventRecord(e0);
    k_probe<F, THREADS, MINB, IPT, UNR, TJ><<<grid, THREADS>>>(jp, jv, out, reps, 0.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double pairs = (double)grid * 32.0 * IPT * TJ * reps;
  const double rate = pairs / (best * 1e-3);
  const double peak = 148.0 * 64.0 * clk_ghz * 1e9 / 32.0;
  printf("%-34s %8.3f ms  %.4e pairs/s  frac of DP issue peak %.4f  (%s)\n", name, best, rate, rate / peak,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int TJ = 256 + 1024;
  double4 *hp = (double4 *)malloc(TJ * sizeof(double4)), *hv = (double4 *)malloc(TJ * sizeof(double4));
  srand(1);
  for (int k = 0; k < TJ; k++) {
    hp[k] = make_double4(rand() / (double)RAND_MAX, rand() / (double)RAND_MAX, rand() / (double)RAND_MAX, 1.0 / TJ);
    hv[k] = make_double4(rand() / (double)RAND_MAX, rand() / (double)RAND_MAX, rand() / (double)RAND_MAX, 0.0);
  }
  double4 *jp, *jv; double *out;
  cudaMalloc(&jp, TJ * sizeof(double4)); cudaMalloc(&jv, TJ * sizeof(double4)); cudaMalloc(&out, 148 * 4 * 1024 * 8);
  cudaMemcpy(jp, hp, TJ * sizeof(double4), cudaMemcpyHostToDevice);
  cudaMemcpy(jv, hv, TJ * sizeof(double4), cudaMemcpyHostToDevice);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double clk = 1.965;
  printf("device clock attr %d kHz, using %.3f GHz\n", clk_khz, clk);
#define RUN(F, T, M, I, U) run<F, T, M, I, U>("F" #F " thr" #T " minb" #M " ipt" #I " unr" #U, jp, jv, out, clk);
#include "variants.inc"
  return 0;
}
