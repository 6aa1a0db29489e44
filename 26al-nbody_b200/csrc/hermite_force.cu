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
#include "al26_internal.cuh"

namespace al26 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

// 1/sqrt(x) for x >= 0; returns 0 for x == 0 (and sub-2^-1042 denormals): the self-pair mask.
__device__ __forceinline__ double rsqrt_masked(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));  // MUFU.RSQ64H, ~2^-20 relative
  // low word of y0 is zero; select on the high word only (integer pipe)
  int hi = __double2hiint(x);
  int yhi = (hi == 0) ? 0 : __double2hiint(y0);
  y0 = __hiloint2double(yhi, 0);
  const double y2 = y0 * y0;
  const double e = fma(-x, y2, 1.0);
  const double p = fma(0.375, e, 0.5);
  const double ye = y0 * e;
  return fma(ye, p, y0);  // y0 (1 + e/2 + 3e^2/8): residual 5e^3/16
}

struct Acc7 {
  double ax, ay, az, jx, jy, jz, pot;
};

__device__ __forceinline__ void pair_interaction(const double4 pj, const double4 vj, const double eps2,
                                                 const double xi, const double yi, const double zi,
                                                 const double vxi, const double vyi, const double vzi, Acc7 &s) {
  const double dx = pj.x - xi, dy = pj.y - yi, dz = pj.z - zi;
  const double dvx = vj.x - vxi, dvy = vj.y - vyi, dvz = vj.z - vzi;
  const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
  const double rv = fma(dx, dvx, fma(dy, dvy, dz * dvz));
  const double rinv = rsqrt_masked(r2);
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
}

// Kernel configuration: threads per CTA, CTAs per SM, i-particles per lane for big blocks, unroll of the
// j loop.  Several configurations are instantiated; GravDev::variant selects one (al26_set_force_variant).
template <int THREADS_, int MINB_, int IPT_, int UNR_>
struct FCfg {
  static constexpr int THREADS = THREADS_, MINB = MINB_, IPT = IPT_, UNR = UNR_, WARPS = THREADS_ / 32;
};

template <class C>
struct ForceSmemT {
  double4 pos[FORCE_STAGES][FORCE_TJ];
  double4 vel[FORCE_STAGES][FORCE_TJ];
  double red[C::WARPS][7][32 * C::IPT];
  unsigned long long full[FORCE_STAGES];
  int item;
};

__device__ __forceinline__ void issue_tile(const GravDev &g, double4 *spos, double4 *svel, unsigned long long *bar,
                                           int b, int cnt) {
  mbar_expect_tx(bar, (uint32_t)cnt * 64u);
  tma_load_1d(spos, g.jpos + b, (uint32_t)cnt * 32u, bar);
  tma_load_1d(svel, g.jvel + b, (uint32_t)cnt * 32u, bar);
}

// One work item.  IPT i-particles per lane.  SPLIT (IPT == 1 only, tiny blocks of n_act <= 16): the 32 lanes
// are divided into 32/iw groups that take different j's for the same iw i-particles, and are summed with a
// fixed xor-butterfly at the end -> the FP64 pipe time of a tiny block drops by the same factor.
template <class C, int IPT, bool SPLIT>
__device__ __forceinline__ void run_item(const GravDev &g, ForceSmemT<C> &sm, const Decomp &d, const int n_act,
                                         const int item, uint32_t &it) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int itile = item / d.n_jsplit, js = item - itile * d.n_jsplit;
  const int j0 = js * d.jchunk;
  const int j1 = min(g.n_tot, j0 + d.jchunk);
  const int ntiles = (j1 - j0 + FORCE_TJ - 1) / FORCE_TJ;

  // producer prologue: all stages are free here (previous item ended with __syncthreads)
  if (tid == 0) {
    const int pre = ntiles < FORCE_STAGES ? ntiles : FORCE_STAGES;
    for (int k = 0; k < pre; k++) {
      const int st = (it + k) % FORCE_STAGES;
      const int b = j0 + k * FORCE_TJ;
      issue_tile(g, &sm.pos[st][0], &sm.vel[st][0], &sm.full[st], b, min(FORCE_TJ, j1 - b));
    }
  }

  int iw = 32;  // lanes that hold distinct i-particles
  if (SPLIT) {
    iw = 1;
    while (iw < n_act) iw <<= 1;
  }
  const int isub = SPLIT ? (lane & (iw - 1)) : lane;
  const int jsub = SPLIT ? (lane / iw) : 0;
  const int jgroups = SPLIT ? (32 / iw) : 1;

  double xi[IPT], yi[IPT], zi[IPT], vxi[IPT], vyi[IPT], vzi[IPT];
  Acc7 s[IPT];
#pragma unroll
  for (int q = 0; q < IPT; q++) {
    const int slot = itile * d.ti + q * 32 + isub;
    const int li = (slot < n_act) ? g.list[slot] : g.list[0];
    const double4 p = g.jpos[g.i0 + li];
    const double4 v = g.jvel[g.i0 + li];
    xi[q] = p.x; yi[q] = p.y; zi[q] = p.z;
    vxi[q] = v.x; vyi[q] = v.y; vzi[q] = v.z;
    s[q].ax = s[q].ay = s[q].az = s[q].jx = s[q].jy = s[q].jz = s[q].pot = 0.0;
  }
  const double eps2 = g.eps2;
  const int jfirst = warp * jgroups + jsub, jstride = C::WARPS * jgroups;

  for (int k = 0; k < ntiles; k++, it++) {
    const int st = it % FORCE_STAGES;
    const uint32_t parity = (it / FORCE_STAGES) & 1u;
    const int cnt = min(FORCE_TJ, j1 - (j0 + k * FORCE_TJ));
    mbar_wait(&sm.full[st], parity);
    const double4 *__restrict__ sp = sm.pos[st];
    const double4 *__restrict__ sv = sm.vel[st];
#pragma unroll C::UNR
    for (int jj = jfirst; jj < cnt; jj += jstride) {
      const double4 pj = sp[jj];
      const double4 vj = sv[jj];
#pragma unroll
      for (int q = 0; q < IPT; q++) pair_interaction(pj, vj, eps2, xi[q], yi[q], zi[q], vxi[q], vyi[q], vzi[q], s[q]);
    }
    __syncthreads();  // every warp is done with stage st
    if (tid == 0 && k + FORCE_STAGES < ntiles) {
      const int b = j0 + (k + FORCE_STAGES) * FORCE_TJ;
      issue_tile(g, &sm.pos[st][0], &sm.vel[st][0], &sm.full[st], b, min(FORCE_TJ, j1 - b));
    }
  }

  if (SPLIT) {  // sum the j-groups of a warp: fixed butterfly over the lane bits above log2(iw)
    for (int o = iw; o < 32; o <<= 1) {
      s[0].ax += __shfl_xor_sync(0xffffffffu, s[0].ax, o); s[0].ay += __shfl_xor_sync(0xffffffffu, s[0].ay, o);
      s[0].az += __shfl_xor_sync(0xffffffffu, s[0].az, o); s[0].jx += __shfl_xor_sync(0xffffffffu, s[0].jx, o);
      s[0].jy += __shfl_xor_sync(0xffffffffu, s[0].jy, o); s[0].jz += __shfl_xor_sync(0xffffffffu, s[0].jz, o);
      s[0].pot += __shfl_xor_sync(0xffffffffu, s[0].pot, o);
    }
  }
  // fixed-order reduction over the warps
  if (!SPLIT || lane < iw) {
#pragma unroll
    for (int q = 0; q < IPT; q++) {
      const int c = q * 32 + lane;
      sm.red[warp][0][c] = s[q].ax; sm.red[warp][1][c] = s[q].ay; sm.red[warp][2][c] = s[q].az;
      sm.red[warp][3][c] = s[q].jx; sm.red[warp][4][c] = s[q].jy; sm.red[warp][5][c] = s[q].jz;
      sm.red[warp][6][c] = s[q].pot;
    }
  }
  __syncthreads();
  if (tid < 32 * IPT && (!SPLIT || tid < iw)) {
    double r[7];
#pragma unroll
    for (int c = 0; c < 7; c++) {
      double a = sm.red[0][c][tid];
#pragma unroll
      for (int w = 1; w < C::WARPS; w++) a += sm.red[w][c][tid];
      r[c] = a;
    }
    const long long o = (long long)js * d.slot_stride + (long long)itile * d.ti + tid;
    g.part_a[o] = make_double4(r[0], r[1], r[2], r[6]);
    g.part_j[o] = make_double4(r[3], r[4], r[5], 0.0);
  }
  // red[] is next written only after the item-fetch barrier of the next item -> no hazard.
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB) k_force(const GravDev g, const int phase) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ForceSmemT<C> &sm = *reinterpret_cast<ForceSmemT<C> *>(smem_raw);
  StepCtrl *ctl = &g.ctrl[phase];
  const int n_act = ctl->n_act;
  if (n_act <= 0) return;
  const Decomp d = make_decomp(n_act, g.n_tot, gridDim.x, C::IPT);
  const int n_items = d.n_itiles * d.n_jsplit;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < FORCE_STAGES; s++) mbar_init(&sm.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t it = 0;
  while (true) {
    if (tid == 0) sm.item = atomicAdd(&ctl->work_counter, 1);
    __syncthreads();
    const int item = sm.item;
    if (item >= n_items) break;
    if (d.ipt > 1) run_item<C, C::IPT, false>(g, sm, d, n_act, item, it);
    else if (n_act <= FORCE_SPLIT_MAX_NACT) run_item<C, 1, true>(g, sm, d, n_act, item, it);
    else run_item<C, 1, false>(g, sm, d, n_act, item, it);
    __syncthreads();
  }
}

// ---- the instantiated configurations ----
using FV0 = FCfg<256, 2, 2, 2>;  // default
using FV1 = FCfg<256, 2, 2, 4>;
using FV2 = FCfg<256, 2, 2, 1>;
using FV3 = FCfg<256, 1, 4, 1>;
using FV4 = FCfg<256, 1, 4, 2>;
using FV5 = FCfg<512, 1, 2, 2>;
using FV6 = FCfg<256, 3, 1, 2>;
using FV7 = FCfg<128, 4, 2, 2>;
using FV8 = FCfg<256, 1, 3, 2>;
using FV9 = FCfg<128, 3, 3, 1>;

#define FOR_EACH_FORCE_VARIANT(X) X(0, FV0) X(1, FV1) X(2, FV2) X(3, FV3) X(4, FV4) X(5, FV5) X(6, FV6) X(7, FV7) X(8, FV8) X(9, FV9)

int force_variant_count() { return 10; }

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
      if (soft) u = fma(pj.w, rsqrt_masked(r2 + eps2), u);
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
