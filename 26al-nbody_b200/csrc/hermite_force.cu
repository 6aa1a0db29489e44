// hermite_force.cu -- K1: softened fp64 pairwise acceleration + jerk + potential on the active
// i-particles from the predicted j-set (stands in for ph4's
// idata::get_acc_and_jerk behind gravity.evolve_model, al26_nbody.py:833), and K4's pair sums.
//
// Mapping (B200, sm_100a): FP64-pipe bound, 32 DP instructions per pair.
//   * persistent grid (2 CTAs per SM), 256 threads = 8 warps per CTA;
//   * a work item = (i-tile of 32*IPT active particles) x (one contiguous j-chunk); items are
//     handed out by an atomic counter, so SMs stay busy for any n_act, and every item writes its
//     own partial slot -> results do not depend on which CTA ran which item;
//   * the 32 lanes of EVERY warp hold the item's i-particles (IPT per lane, in registers); the
//     8 warps split the j's of each tile, and are summed in fixed warp order through shared
//     memory at the end of the item -> bitwise reproducible run to run;
//   * j-tiles ({x,y,z,m},{vx,vy,vz,-}: 64 B per j) are staged global -> shared by 1-D TMA bulk
//     copies (cp.async.bulk + mbarrier complete_tx), 3 stages deep, issued by one thread;
//     every lane reads the same j (shared-memory broadcast, LDS.128);
//   * 1/sqrt: MUFU.RSQ64H seed + one 3rd-order step (5 DP ops), error ~1e-18 relative;
//   * pairs with r^2 + eps2 == 0 (self, coincident) are masked on the integer pipe.
// Tensor cores are not used: this is not a dense contraction.
#include "hermite_force.cuh"

namespace al26 {

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB) k_force(const GravDev g, const int phase) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ForceSmemT<C> &sm = *reinterpret_cast<ForceSmemT<C> *>(smem_raw);
  StepCtrl *ctl = &g.ctrl[phase];
  const int n_act = ctl->n_act;
  if (n_act <= 0) return;
  force_smem_init<C>(sm);
  uint32_t it = 0;
  force_items<C>(g, sm, ctl, n_act, gridDim.x, it);
}

int force_variant_count() { return FORCE_VARIANT_COUNT; }

int force_variant_info(int v, int *ctas_per_sm, int *ipt) {
  switch (v) {
#define X(id, cfg) case id: *ctas_per_sm = cfg::MINB; *ipt = cfg::IPT; return 0;
    FOR_EACH_FORCE_VARIANT(X)
#undef X
  }
  return -1;
}

cudaError_t force_kernel_setup() {
  cudaError_t e = cudaSuccess;
#define X(id, cfg) if (e == cudaSuccess) e = cudaFuncSetAttribute(k_force<cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ForceSmemT<cfg>));
  FOR_EACH_FORCE_VARIANT(X)
#undef X
  return e;
}

int launch_force(const GravDev &g, int phase, cudaStream_t s) {
  switch (g.variant) {
#define X(id, cfg) case id: k_force<cfg><<<g.grid_force, cfg::THREADS, sizeof(ForceSmemT<cfg>), s>>>(g, phase); break;
    FOR_EACH_FORCE_VARIANT(X)
#undef X
  }
  return 1;
}

int force_smem_bytes() { return (int)sizeof(ForceSmemT<FV0>); }

// DFMA-only microkernel: the measured FP64 roofline denominator.  8 independent
// chains per thread, 512 threads x 2 CTAs per SM.
__global__ void __launch_bounds__(512, 2) k_dfma_peak(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (r == 123.456) out[0] = r;  // never true; keeps the chains alive
}

// the same 64 DFMAs per iteration plus two MUFU.RSQ64H (one per 32 DFMAs: the force kernel's ratio), whose
// results feed nothing the DFMA chains wait for: does the 64-bit MUFU take FP64 issue slots?
__global__ void __launch_bounds__(512, 2) k_dfma_mufu_mix(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  double m0 = 1.0 + threadIdx.x, m1 = 2.0 + threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
      if (u == 3) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(m0));
      if (u == 7) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(m1));
    }
  }
  const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7)) + (m0 + m1);
  if (r == 123.456) out[0] = r;  // never true; keeps the chains alive
}

// Operand-pattern variants of the FP64 issue-rate microbenchmark (8 independent chains per thread, 512 threads x 2 CTAs
// per SM): what limits a DFMA with three distinct register operands -- the pipe or the register file?
//   1: x = fma(x, imm, b)   two register reads per instruction    2: x = x + b (DADD)    3: x = x * imm (DMUL)
//   4: x = fma(x, x, b)     5: x = fma(x, a_k, b_k) with per-chain a_k, b_k (three distinct registers, no sharing)
template <int V>
__global__ void __launch_bounds__(512, 2) k_fp64_rate(double *out, int iters, double a, double b) {
  double x[8], ak[8], bk[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    x[k] = (double)threadIdx.x + k;
    ak[k] = a + 1e-9 * k;
    bk[k] = b * (k + 1);
  }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (V == 1) x[k] = fma(x[k], 0.99951171875, b);
        else if (V == 2) x[k] = x[k] + b;
        else if (V == 3) x[k] = x[k] * 0.99951171875;
        else if (V == 4) x[k] = fma(x[k], x[k], b);
        else x[k] = fma(x[k], ak[k], bk[k]);
      }
    }
  }
  double r = 0.0;
#pragma unroll
  for (int k = 0; k < 8; k++) r += x[k];
  if (r == 123.456) out[0] = r;  // never true; keeps the chains alive
}

// returns DP lane-instructions per launch
double launch_fp64_rate(int variant, int sm_count, int iters, double *scratch, cudaStream_t s) {
  const int blocks = sm_count * 2, threads = 512;
  switch (variant) {
    case 1: k_fp64_rate<1><<<blocks, threads, 0, s>>>(scratch, iters, 0.999999, 1e-9); break;
    case 2: k_fp64_rate<2><<<blocks, threads, 0, s>>>(scratch, iters, 0.999999, 1e-9); break;
    case 3: k_fp64_rate<3><<<blocks, threads, 0, s>>>(scratch, iters, 0.999999, 1e-9); break;
    case 4: k_fp64_rate<4><<<blocks, threads, 0, s>>>(scratch, iters, 0.999999, 1e-9); break;
    case 5: k_fp64_rate<5><<<blocks, threads, 0, s>>>(scratch, iters, 0.999999, 1e-9); break;
    default: k_dfma_peak<<<blocks, threads, 0, s>>>(scratch, iters, 0.999999, 1e-9); break;
  }
  return (double)blocks * threads * 64.0 * (double)iters;
}

double launch_dfma_mufu_mix(int sm_count, int iters, double *scratch, cudaStream_t s) {
  const int blocks = sm_count * 2, threads = 512;
  k_dfma_mufu_mix<<<blocks, threads, 0, s>>>(scratch, iters, 0.999999, 1e-9);
  return 2.0 * (double)blocks * threads * 64.0 * (double)iters;  // DFMA flops of one launch
}

double launch_dfma_peak(int sm_count, int iters, double *scratch, cudaStream_t s) {
  const int blocks = sm_count * 2, threads = 512;
  k_dfma_peak<<<blocks, threads, 0, s>>>(scratch, iters, 0.999999, 1e-9);
  return 2.0 * (double)blocks * threads * 64.0 * (double)iters;  // flops of one launch
}


// ------------------------------------------------------------------------------------------
// K4: pair sums for the energies: for local i over all j,
//   u_i = sum_j m_j / sqrt(r^2 + eps2),  s_i = sum_j m_j / r   (self / coincident pairs masked)
// then per-block partials of  K = 1/2 m v^2,  U = -1/2 m_i u_i,  S = 1/2 m_i s_i.
// 2-D grid: x = i-tiles of 256, y = j-chunks; partials reduced in fixed order by k_energy_final.
// ------------------------------------------------------------------------------------------
constexpr int EN_THREADS = 256;
constexpr int EN_TJ = 512;
constexpr int EN_JSPLIT_TARGET = 1184;  // aim for ~8 CTAs per SM in total

__global__ void __launch_bounds__(EN_THREADS) k_energy_pairs(const EnergyDev e, double *us_part, int n_jsplit,
                                                             int jchunk) {
  __shared__ double4 sj[EN_TJ];
  const int i = blockIdx.x * EN_THREADS + threadIdx.x;
  const bool valid = i < e.n_loc;
  const double4 pi = e.jpos[e.i0 + (valid ? i : 0)];
  const int j0 = blockIdx.y * jchunk, j1 = min(e.n_tot, j0 + jchunk);
  double u = 0.0, s = 0.0;
  const double eps2 = e.eps2;
  const bool soft = eps2 != 0.0;
  const int self = valid ? e.i0 + i : -1;
  for (int b = j0; b < j1; b += EN_TJ) {
    const int cnt = min(EN_TJ, j1 - b);
    __syncthreads();
    for (int k = threadIdx.x; k < cnt; k += EN_THREADS) sj[k] = e.jpos[b + k];
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < cnt; k++) {
      const double4 pj = sj[k];
      const double dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
      const double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
      const double ri = rsqrt_masked(r2);
      s = fma(pj.w, ri, s);
      // softened sum: the self pair (not dropped by the masked rsqrt: r^2 + eps2 > 0) is taken out by index
      if (soft) u = fma((b + k == self) ? 0.0 : pj.w, rsqrt_masked(r2 + eps2), u);
    }
  }
  if (!soft) u = s;
  if (valid) {
    us_part[((long long)blockIdx.y * e.n_loc + i) * 2 + 0] = u;
    us_part[((long long)blockIdx.y * e.n_loc + i) * 2 + 1] = s;
  }
  (void)n_jsplit;
}

__device__ __forceinline__ double block_sum_fixed(double v, double *sh) {
  // fixed-order: butterfly in the warp, then warp 0 sums the warp totals in order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) r += sh[w];
  return r;  // valid in thread 0
}

__global__ void __launch_bounds__(EN_THREADS) k_energy_reduce(const EnergyDev e, const double *us_part, int n_jsplit) {
  __shared__ double sh[EN_THREADS / 32];
  const int i = blockIdx.x * EN_THREADS + threadIdx.x;
  double K = 0.0, U = 0.0, S = 0.0;
  if (i < e.n_loc) {
    double u = 0.0, s = 0.0;
    for (int js = 0; js < n_jsplit; js++) {
      u += us_part[((long long)js * e.n_loc + i) * 2 + 0];
      s += us_part[((long long)js * e.n_loc + i) * 2 + 1];
    }
    const double4 p = e.pos[i], v = e.vel[i];
    K = 0.5 * p.w * fma(v.x, v.x, fma(v.y, v.y, v.z * v.z));
    U = -0.5 * p.w * u;
    S = 0.5 * p.w * s;
  }
  const double k = block_sum_fixed(K, sh);
  const double u = block_sum_fixed(U, sh);
  const double s = block_sum_fixed(S, sh);
  if (threadIdx.x == 0) {
    e.block_part[blockIdx.x * 3 + 0] = k;
    e.block_part[blockIdx.x * 3 + 1] = u;
    e.block_part[blockIdx.x * 3 + 2] = s;
  }
}

__global__ void __launch_bounds__(EN_THREADS) k_energy_final(const EnergyDev e, int nblocks) {
  __shared__ double sh[EN_THREADS / 32];
  double acc[3] = {0.0, 0.0, 0.0};
  for (int b = threadIdx.x; b < nblocks; b += EN_THREADS)
    for (int c = 0; c < 3; c++) acc[c] += e.block_part[b * 3 + c];
  for (int c = 0; c < 3; c++) {
    const double r = block_sum_fixed(acc[c], sh);
    if (threadIdx.x == 0) e.out[c] = r;
  }
}

int energy_grid(int n_loc) { return (n_loc + EN_THREADS - 1) / EN_THREADS; }

// us_part scratch is carved from block_part's tail by the caller: see api.cu (energy_scratch_doubles)
int launch_energies(const EnergyDev &e, cudaStream_t s) {
  const int gx = energy_grid(e.n_loc);
  int n_jsplit = (EN_JSPLIT_TARGET + gx - 1) / gx;
  const int max_by_j = (e.n_tot + EN_TJ - 1) / EN_TJ;
  if (n_jsplit > max_by_j) n_jsplit = max_by_j;
  if (n_jsplit < 1) n_jsplit = 1;
  if (n_jsplit > 65535) n_jsplit = 65535;
  int jchunk = (e.n_tot + n_jsplit - 1) / n_jsplit;
  n_jsplit = (e.n_tot + jchunk - 1) / jchunk;
  double *us_part = e.block_part + (size_t)gx * 3;
  k_energy_pairs<<<dim3(gx, n_jsplit), EN_THREADS, 0, s>>>(e, us_part, n_jsplit, jchunk);
  k_energy_reduce<<<gx, EN_THREADS, 0, s>>>(e, us_part, n_jsplit);
  k_energy_final<<<1, EN_THREADS, 0, s>>>(e, gx);
  return 3;
}

}  // namespace al26
