// hermite_force.cuh -- device code of K1 (see hermite_force.cu for the description): TMA / mbarrier
// helpers, the pair interaction, one work item (run_item) and the kernel configurations.  Shared by
// k_force (hermite_force.cu) and the persistent loop kernel (hermite_loop.cu).
#pragma once
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
                                         const int item, uint32_t &it, const int *__restrict__ list) {
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
    const int li = (slot < n_act) ? list[slot] : list[0];
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

// the item loop of one force evaluation: items handed out by the atomic work counter of `ctl`
template <class C>
__device__ __forceinline__ void force_items(const GravDev &g, ForceSmemT<C> &sm, StepCtrl *ctl, const int n_act,
                                            const int n_ctas, uint32_t &it, const int *__restrict__ list = nullptr) {
  if (!list) list = g.list;  // the peer-memory loop passes the list of the particles this rank owns instead
  const Decomp d = make_decomp(n_act, g.n_tot, g.decomp_tab, C::IPT, g.big_nact);
  const int n_items = d.n_itiles * d.n_jsplit;
  const int tid = threadIdx.x;
  // first round: CTA b takes item b with no atomic; later rounds (only when there are more items than CTAs)
  // draw from the shared counter, which therefore counts items n_ctas, n_ctas + 1, ...  Small blocks -- one
  // round -- pay no atomic round trip at all.
  int item = blockIdx.x;
  while (item < n_items) {
    if (d.ipt > 1) run_item<C, C::IPT, false>(g, sm, d, n_act, item, it, list);
    else if (n_act <= FORCE_SPLIT_MAX_NACT) run_item<C, 1, true>(g, sm, d, n_act, item, it, list);
    else run_item<C, 1, false>(g, sm, d, n_act, item, it, list);
    if (n_items <= n_ctas) break;
    __syncthreads();
    if (tid == 0) sm.item = n_ctas + atomicAdd(&ctl->work_counter, 1);
    __syncthreads();
    item = sm.item;
  }
  __syncthreads();  // red[] / stage buffers are free again for whoever runs next in this CTA
}

template <class C>
__device__ __forceinline__ void force_smem_init(ForceSmemT<C> &sm) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < FORCE_STAGES; s++) mbar_init(&sm.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

// ---- the instantiated configurations ----
using FV0 = FCfg<256, 2, 2, 2>;  // default: 2 CTAs/SM x 8 warps, 2 i per lane, j loop unrolled x2
using FV1 = FCfg<256, 2, 2, 4>;
using FV2 = FCfg<256, 1, 4, 2>;  // 1 CTA/SM, 4 i per lane
using FV3 = FCfg<128, 3, 3, 1>;  // 3 CTAs/SM x 4 warps, 3 i per lane

#define FOR_EACH_FORCE_VARIANT(X) X(0, FV0) X(1, FV1) X(2, FV2) X(3, FV3)
constexpr int FORCE_VARIANT_COUNT = 4;


}  // namespace al26
