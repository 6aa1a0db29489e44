// enrich.cu -- K5: the fused short-lived-radionuclide enrichment pass, one outer step of
// /root/reference/al26_nbody.py:878-1086 (interloper block excluded):
//   classify (:1194-1216) -> wind deposit, local + global bubble, 26Al + 60Fe (:642-702, :897-938)
//   -> supernova events + deposit (:945-967, :1326-1334) -> decay (:1048-1064) -> condense (:1071-1086).
// Compiled with --fmad=false and written in the reference's evaluation order, so the per-disc
// wind sums are BIT-IDENTICAL to the reference's numba kernel: per disc the massive stars are
// visited in ascending index order (the sorted source table), each term is
// ((wind_ratio*mdot)*eta_bub)*dt with eta_bub = ((0.75*(r*r))*d_trav)/(R*(R*R)).
// The reference evaluates the geometry four times per step (once per calc_wind_abs call); here
// it is evaluated once per (disc, source) pair and feeds all four accumulators.
//
// Kernels (3 launches per outer step):
//   k_enrich_classify  all stars: m >= 13 Msun -> atomic append to the source list
//   k_enrich_sources   one CTA: bitonic sort of the list (ascending index), build the source
//                      table {x,y,z,c26},{c60,sn26,sn60,event}, detect SN events (mdot == 0 and
//                      not kicked), set kicked, emit the ordered event list
//   k_enrich_discs     one thread per star of this rank's slice: source loop from shared memory
//                      (broadcast reads), decay of all rows, condense flags; FP64-issue bound
//                      for many sources, HBM bound (~260 B per disc-update) for few.
#include "al26_internal.cuh"

namespace al26 {

constexpr int EN_T = 256;
constexpr int SRC_TILE = 512;   // sources staged per shared-memory tile (2 x 16 KB)

__global__ void __launch_bounds__(EN_T) k_enrich_classify(const EnrichDev e) {
  const int i = blockIdx.x * EN_T + threadIdx.x;
  if (i == 0) {
    e.counters[1] = 0;
  }
  if (i >= e.n_tot) return;
  if (e.mass_msun[i] >= 13.0) {
    const int p = atomicAdd(&e.counters[0], 1);
    if (p < ENR_MAX_SOURCES) e.hm_list[p] = i;
    else e.counters[2] = 1;
  }
}

__device__ __forceinline__ void star_pos(const EnrichDev &e, int i, double &x, double &y, double &z) {
  if (e.px) {
    x = e.px[i]; y = e.py[i]; z = e.pz[i];
  } else {
    const double4 p = e.gpos[i];
    x = p.x * e.km_per_length; y = p.y * e.km_per_length; z = p.z * e.km_per_length;
  }
}
__device__ __forceinline__ void star_vel(const EnrichDev &e, int i, double &x, double &y, double &z) {
  if (e.px) {
    x = e.pvx[i]; y = e.pvy[i]; z = e.pvz[i];
  } else {
    const double4 p = e.gvel[i];
    x = p.x * e.kms_per_speed; y = p.y * e.kms_per_speed; z = p.z * e.kms_per_speed;
  }
}

__global__ void __launch_bounds__(1024) k_enrich_sources(const EnrichDev e) {
  __shared__ int keys[ENR_MAX_SOURCES];
  int n_hm = e.counters[0];
  if (n_hm > ENR_MAX_SOURCES) n_hm = ENR_MAX_SOURCES;
  int np2 = 1;
  while (np2 < n_hm) np2 <<= 1;
  for (int k = threadIdx.x; k < np2; k += blockDim.x) keys[k] = (k < n_hm) ? e.hm_list[k] : 0x7fffffff;
  __syncthreads();
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int k = threadIdx.x; k < np2; k += blockDim.x) {
        const int partner = k ^ stride;
        if (partner > k) {
          const bool up = ((k & size) == 0);
          const int a = keys[k], b = keys[partner];
          if ((a > b) == up) {
            keys[k] = b;
            keys[partner] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int k = threadIdx.x; k < n_hm; k += blockDim.x) {
    const int i = keys[k];
    e.hm_list[k] = i;
    double x, y, z;
    star_pos(e, i, x, y, z);
    const double mdot = e.mdot[i];
    const bool ev = (mdot == 0.0) && (e.kicked[i] == 0);
    e.src_a[k] = make_double4(x, y, z, e.wr26[i] * mdot);
    e.src_b[k] = make_double4(e.wr60[i] * mdot, ev ? e.sn26[i] : 0.0, ev ? e.sn60[i] : 0.0, ev ? 1.0 : 0.0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int ne = 0;
    for (int k = 0; k < n_hm; k++) {
      if (e.src_b[k].w != 0.0) {
        const int i = keys[k];
        e.sn_events[ne++] = i;
        e.kicked[i] = 1;
      }
    }
    e.counters[1] = ne;
    e.counters[3] = n_hm;
  }
}

__global__ void __launch_bounds__(EN_T) k_enrich_discs(const EnrichDev e, const EnrichParams p) {
  __shared__ double4 sa[SRC_TILE];
  __shared__ double4 sb[SRC_TILE];
  const int li = blockIdx.x * EN_T + threadIdx.x;
  const bool valid = li < e.n_loc;
  const int gi = e.d0 + (valid ? li : 0);
  int n_hm = e.counters[0];
  if (n_hm > ENR_MAX_SOURCES) n_hm = ENR_MAX_SOURCES;
  const int n_ev = e.counters[1];

  const double m = valid ? e.mass_msun[gi] : 0.0;
  const bool is_lm = valid && (m >= 0.1) && (m <= 3.0);
  const size_t n = (size_t)e.n_loc;

  double inv[ENR_NINV];
#pragma unroll
  for (int r = 0; r < ENR_NINV; r++) inv[r] = valid ? e.inv[r * n + li] : 0.0;

  double x = 0, y = 0, z = 0, rd = 0, eta_l = 0, eta_g = 0;
  if (is_lm && n_hm > 0) {
    double vx, vy, vz;
    star_pos(e, gi, x, y, z);
    star_vel(e, gi, vx, vy, vz);
    rd = e.r_disk[li];
    const double spd = sqrt(vx * vx + vy * vy + vz * vz);   // (lm_vx**2 + lm_vy**2 + lm_vz**2)**0.5
    const double trav = spd * p.dt_s;                       // d_disk_trav = disk_spd * dt
    const double k0 = 0.75 * (rd * rd) * trav;              // 0.75 * (r_disk**2) * d_disk_trav
    eta_l = k0 / p.r_local3;                                // / (bubble_radius ** 3)
    eta_g = k0 / p.r_global3;
  }
  double l26 = 0.0, l60 = 0.0, g26 = 0.0, g60 = 0.0;
  double s26 = inv[2], s60 = inv[6];  // SN deposits add straight onto the inventories (:964-965)

  for (int b = 0; b < n_hm; b += SRC_TILE) {
    const int cnt = min(SRC_TILE, n_hm - b);
    __syncthreads();
    for (int k = threadIdx.x; k < cnt; k += EN_T) {
      sa[k] = e.src_a[b + k];
      sb[k] = e.src_b[b + k];
    }
    __syncthreads();
    if (is_lm) {
      // Sources in groups of 64: the hot loop only MARKS the sources whose local bubble holds this disc (a predicated
      // integer OR) -- their deposit, 6 DP instructions that predication would otherwise issue for every pair, is
      // added afterwards for the marked ones only (rare: R = 0.1 pc in a ~1 pc cluster), still in ascending source
      // order, so the sums keep the reference's rounding.
      for (int k0 = 0; k0 < cnt; k0 += 64) {
        const int gn = min(64, cnt - k0);
        unsigned long long marked = 0ull;
#pragma unroll 4
        for (int kk = 0; kk < gn; kk++) {
          const double4 A = sa[k0 + kk];
          const double4 B = sb[k0 + kk];
          // global model: distance_limit == 0 -> no test (:688)
          g26 += (A.w * eta_g) * p.dt_s;
          g60 += (B.x * eta_g) * p.dt_s;
          // local model: skip when bubble_radius <= d_sep (:689-691); q_local is the exact
          // d^2 threshold of that test, so no sqrt is needed here
          const double dx = x - A.x, dy = y - A.y, dz = z - A.z;
          const double d2 = dx * dx + dy * dy + dz * dz;
          if (!(d2 >= p.q_local)) marked |= 1ull << kk;
          if (n_ev > 0 && B.w != 0.0) {
            // calc_star_distance + calc_eta_disk_sne (:1331-1333)
            const double d = sqrt(d2);
            const double eta = (0.5 * 0.7) * ((0.5 * (rd * rd)) / (4.0 * (d * d)));
            s26 += B.y * eta;
            s60 += B.z * eta;
          }
        }
        while (marked) {
          const int kk = __ffsll((long long)marked) - 1;
          marked &= marked - 1ull;
          l26 += (sa[k0 + kk].w * eta_l) * p.dt_s;
          l60 += (sb[k0 + kk].x * eta_l) * p.dt_s;
        }
      }
    }
  }
  if (!valid) return;
  // accumulate (:935-938), decay (:1054-1064)
  inv[0] = (inv[0] + l26) * p.decay26;
  inv[1] = (inv[1] + g26) * p.decay26;
  inv[2] = s26 * p.decay26;
  inv[4] = (inv[4] + l60) * p.decay60;
  inv[5] = (inv[5] + g60) * p.decay60;
  inv[6] = s60 * p.decay60;
  if (p.with_agb) {
    inv[3] *= p.decay26;
    inv[7] *= p.decay60;
  }
#pragma unroll
  for (int r = 0; r < ENR_NINV; r++) e.inv[r * n + li] = inv[r];
  // condense (:1071-1086)
  if (is_lm && e.alive[li]) {
    const double tau = e.tau_disk[li];
    if (tau >= p.t_new_myr) {
#pragma unroll
      for (int r = 0; r < ENR_NINV; r++)
        if (p.with_agb || (r != 3 && r != 7)) e.fin[r * n + li] = inv[r];
    }
    if (tau < p.t_new_myr) e.alive[li] = 0;
  }
}

// AGB interloper deposit (al26_nbody.py:985-1028): per disc, the fraction of the step spent within r_test of
// the interloper, by the reference's own recipe -- 1024 np.linspace samples of both straight-line paths
// (calc_intersection, :1156-1190; linspace = k*step + start with step = (stop-start)/1023, last = stop) --
// then the sweep-up deposit onto the agb rows.  The interloper's samples are shared by the CTA.
constexpr int ISECT_N = 1024;

__global__ void __launch_bounds__(EN_T) k_enrich_interloper(const EnrichDev e, const InterloperParams p) {
  __shared__ double s1[3][ISECT_N];
  const size_t nt = (size_t)e.n_tot;
  for (int c = 0; c < 3; c++) {
    const double a = p.old_pc[c * nt + p.k_int], b = p.new_pc[c * nt + p.k_int];
    const double step = (b - a) / (double)(ISECT_N - 1);
    for (int k = threadIdx.x; k < ISECT_N; k += EN_T) s1[c][k] = (k == ISECT_N - 1) ? b : ((double)k * step + a);
  }
  __syncthreads();
  const int li = blockIdx.x * EN_T + threadIdx.x;
  if (li >= e.n_loc) return;
  const int gi = e.d0 + li;
  const double m = e.mass_msun[gi];
  if (!((m >= 0.1) && (m <= 3.0)) || gi == p.k_int) return;  // for i in lm_id: if not is_interloper (:990-991)
  const double xo = p.old_pc[gi], yo = p.old_pc[nt + gi], zo = p.old_pc[2 * nt + gi];
  const double xn = p.new_pc[gi], yn = p.new_pc[nt + gi], zn = p.new_pc[2 * nt + gi];
  const double sx = (xn - xo) / (double)(ISECT_N - 1), sy = (yn - yo) / (double)(ISECT_N - 1),
               sz = (zn - zo) / (double)(ISECT_N - 1);
  int cnt = 0;
#pragma unroll 4
  for (int k = 0; k < ISECT_N - 1; k++) {
    const double kk = (double)k;
    const double dx = s1[0][k] - (kk * sx + xo), dy = s1[1][k] - (kk * sy + yo), dz = s1[2][k] - (kk * sz + zo);
    cnt += (dx * dx + dy * dy + dz * dz <= p.q_test) ? 1 : 0;
  }
  {
    const double dx = s1[0][ISECT_N - 1] - xn, dy = s1[1][ISECT_N - 1] - yn, dz = s1[2][ISECT_N - 1] - zn;
    cnt += (dx * dx + dy * dy + dz * dz <= p.q_test) ? 1 : 0;
  }
  if (cnt == 0) return;                                                   // if intersection_frac != 0.0 (:1015)
  const double frac = (double)cnt / (double)ISECT_N;
  const double ex = (xn - xo) * p.km_per_pc, ey = (yn - yo) * p.km_per_pc, ez = (zn - zo) * p.km_per_pc;
  double trav = sqrt(ex * ex + ey * ey + ez * ez);                        // :1020
  trav *= frac;                                                           // :1021
  const double rd = e.r_disk[li];
  const double eta = 0.75 * (rd * rd) * trav / p.r_bub3;                  // :1022
  const double a26 = p.rate26 * eta * p.dt_s, a60 = p.rate60 * eta * p.dt_s;  // :1023-1024
  const size_t n = (size_t)e.n_loc;
  e.inv[3 * n + li] += a26;                                               // mass_26al_agb (:1025)
  e.inv[7 * n + li] += a60;                                               // mass_60fe_agb (:1026)
  p.raw[li] += a26;                                                       // *_agb_raw (:1027-1028)
  p.raw[n + li] += a60;
}

int launch_interloper(const EnrichDev &e, const InterloperParams &p, cudaStream_t s) {
  k_enrich_interloper<<<(e.n_loc + EN_T - 1) / EN_T, EN_T, 0, s>>>(e, p);
  return 1;
}

int launch_enrich(const EnrichDev &e, const EnrichParams &p, cudaStream_t s) {
  k_enrich_classify<<<(e.n_tot + EN_T - 1) / EN_T, EN_T, 0, s>>>(e);
  k_enrich_sources<<<1, 1024, 0, s>>>(e);
  k_enrich_discs<<<(e.n_loc + EN_T - 1) / EN_T, EN_T, 0, s>>>(e, p);
  return 3;
}

}  // namespace al26
