// analysis.cu -- post-processing kernels next to the hot path.
//
// k_local_density: the 10-nearest-neighbour local density of every star, the reference's
// `local_densities_numba` (/root/reference/plotting/al26_plot.py:324-359): for star i, d_ij for all j, the ten
// nearest (the sorted list's entries 1..10; entry 0 is the star itself), rho = (sum of their masses, added in
// ascending-distance order) / (ftp * d10^3) with the script's literal ftp = 4.18879020479.
// The reference materialises an N x N distance matrix (8 N^2 bytes); here each thread keeps the 11 smallest
// squared distances (and the masses that go with them) sorted in registers while the j-set streams through shared
// memory, so memory is O(N).  Same evaluation order (--fmad=false) => bit-identical to the reference.
// Ties in distance are broken towards the lower index (numba's argsort leaves them unspecified).
#include "al26_internal.cuh"

namespace al26 {

constexpr int LD_T = 256;
constexpr int LD_TJ = 512;
constexpr int LD_K = 11;  // self + 10 neighbours

__global__ void __launch_bounds__(LD_T) k_local_density(int n, const double *__restrict__ x, const double *__restrict__ y,
                                                        const double *__restrict__ z, const double *__restrict__ m,
                                                        double *__restrict__ rho) {
  __shared__ double4 sj[LD_TJ];
  const int i = blockIdx.x * LD_T + threadIdx.x;
  const bool valid = i < n;
  const double xi = valid ? x[i] : 0.0, yi = valid ? y[i] : 0.0, zi = valid ? z[i] : 0.0;
  double bd[LD_K], bm[LD_K];
#pragma unroll
  for (int k = 0; k < LD_K; k++) {
    bd[k] = __longlong_as_double(0x7FF0000000000000ll);
    bm[k] = 0.0;
  }
  for (int b = 0; b < n; b += LD_TJ) {
    const int cnt = min(LD_TJ, n - b);
    __syncthreads();
    for (int k = threadIdx.x; k < cnt; k += LD_T) sj[k] = make_double4(x[b + k], y[b + k], z[b + k], m[b + k]);
    __syncthreads();
    for (int k = 0; k < cnt; k++) {
      const double4 pj = sj[k];
      const double dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
      const double d2 = dx * dx + dy * dy + dz * dz;  // (xi-xj)**2 + (yi-yj)**2 + (zi-zj)**2
      if (d2 < bd[LD_K - 1]) {                        // strict: an equal distance keeps the earlier (lower) index ahead
        bd[LD_K - 1] = d2;
        bm[LD_K - 1] = pj.w;
#pragma unroll
        for (int q = LD_K - 1; q > 0; q--) {
          if (bd[q] < bd[q - 1]) {
            const double td = bd[q]; bd[q] = bd[q - 1]; bd[q - 1] = td;
            const double tm = bm[q]; bm[q] = bm[q - 1]; bm[q - 1] = tm;
          }
        }
      }
    }
  }
  if (!valid) return;
  double mass = 0.0;
#pragma unroll
  for (int q = 1; q < LD_K; q++) mass += bm[q];       // for j in idx_nr: mass += masses[j]
  const double d10 = sqrt(bd[LD_K - 1]);              // d[i, j] = sqrt(d2); the 10th neighbour's distance
  const double vol = 4.18879020479 * d10 * d10 * d10; // ftp * d_10 * d_10 * d_10
  rho[i] = mass / vol;
}

int launch_local_density(int n, const double *x, const double *y, const double *z, const double *m, double *rho,
                         cudaStream_t s) {
  k_local_density<<<(n + LD_T - 1) / LD_T, LD_T, 0, s>>>(n, x, y, z, m, rho);
  return 1;
}

}  // namespace al26
