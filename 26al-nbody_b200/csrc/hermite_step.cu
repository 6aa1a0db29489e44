// hermite_step.cu -- the block-timestep machinery around the force kernel (compiled with
// --fmad=false so that, given identical inputs, every timestep / scheduling decision is
// bit-identical to the CPU oracle's):
//   K2  k_predict_list   predictor for every local j fused with the scheduler
//                        (row G5): ballot-compaction of the active list for THIS step and a
//                        warp-shuffle + atomicMin reduction of min(t+dt) over the non-active
//                        particles for the NEXT step;
//   K3  k_correct        fixed-order reduction of the force partials, Hermite corrector, Aarseth
//                        criterion and the dyadic ladder (row G4); folds the active particles'
//                        new t+dt into the same minimum;
//   k_begin              start of an evolve call: tau = 0, clamp dt to the call's ladder.
// Stand-ins for ph4's jdata::predict_all, idata::correct and scheduler behind
// gravity.evolve_model (al26_nbody.py:833).  Three StepCtrl records rotate (this step / next step /
// being reset) so a whole sequence of block steps runs from one CUDA graph with no host in the loop.
#include "al26_internal.cuh"

namespace al26 {

constexpr int ST_THREADS = 256;

__device__ __forceinline__ unsigned long long dbits(double x) { return (unsigned long long)__double_as_longlong(x); }
__device__ __forceinline__ double bitsd(unsigned long long b) { return __longlong_as_double((long long)b); }

__device__ __forceinline__ double pow2floor(double x) {  // x > 0, normal
  return __longlong_as_double(__double_as_longlong(x) & 0x7FF0000000000000ll);
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}

// block-wide min -> one atomicMin per block
__device__ __forceinline__ void block_min_to(unsigned long long v, unsigned long long *dst, unsigned long long *sh) {
  v = warp_min_u64(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : INF_BITS;
    w = warp_min_u64(w);
    if (lane == 0 && w != INF_BITS) atomicMin(dst, w);
  }
}

__global__ void __launch_bounds__(ST_THREADS) k_begin(const GravDev g, const double span, const double D) {
  __shared__ unsigned long long sh[ST_THREADS / 32];
  const int i = blockIdx.x * ST_THREADS + threadIdx.x;
  unsigned long long c = INF_BITS;
  if (i < g.n_loc) {
    double dt = g.dt[i];
    if (dt > D) dt = D;
    g.dt[i] = dt;
    g.t[i] = 0.0;
    c = dbits(dt);
  }
  if (i == 0) {
    g.hdr->span = span;
    g.hdr->D = D;
    g.hdr->done = 0;
  }
  block_min_to(c, &g.ctrl[0].t_next_bits, sh);
}

template <int MODE>
__global__ void __launch_bounds__(ST_THREADS) k_predict_list(const GravDev g, const int phase) {
  __shared__ unsigned long long sh[ST_THREADS / 32];
  StepCtrl *cur = &g.ctrl[phase];
  StepCtrl *nxt = &g.ctrl[(phase + 1) % 3];
  StepCtrl *old = &g.ctrl[(phase + 2) % 3];
  const int i = blockIdx.x * ST_THREADS + threadIdx.x;
  double tn;
  if (MODE == MODE_STEP) {
    const unsigned long long tb = cur->t_next_bits;
    tn = bitsd(tb);
    if (tn > g.hdr->span) {  // finished: carry the time forward so later steps in the graph are no-ops too
      if (i == 0) {
        atomicMin(&nxt->t_next_bits, tb);
        g.hdr->done = 1;
        old->t_next_bits = INF_BITS;
        old->n_act = 0;
        old->work_counter = 0;
      }
      return;
    }
  } else if (MODE == MODE_INIT) {
    tn = 0.0;
  } else {
    tn = g.hdr->span;
  }
  bool active = false;
  unsigned long long c_bits = INF_BITS;
  if (i < g.n_loc) {
    const double4 p = g.pos[i], v = g.vel[i], a = g.acc[i], j = g.jrk[i];
    const double ti = g.t[i], dti = g.dt[i];
    const double s = (MODE == MODE_INIT) ? 0.0 : (tn - ti);  // init: predicted == current, whatever t holds
    const double s2 = s * s * 0.5, s3 = s * s * s * (1.0 / 6.0);
    double4 pp, pv;
    pp.x = p.x + v.x * s + a.x * s2 + j.x * s3;
    pp.y = p.y + v.y * s + a.y * s2 + j.y * s3;
    pp.z = p.z + v.z * s + a.z * s2 + j.z * s3;
    pp.w = p.w;
    pv.x = v.x + a.x * s + j.x * s2;
    pv.y = v.y + a.y * s + j.y * s2;
    pv.z = v.z + a.z * s + j.z * s2;
    pv.w = 0.0;
    g.jpos[g.i0 + i] = pp;
    g.jvel[g.i0 + i] = pv;
    const double c = ti + dti;
    if (MODE == MODE_STEP) active = (c == tn);
    else if (MODE == MODE_INIT) active = true;
    else active = (ti < tn);
    if (!active) c_bits = dbits(c);
  }
  // ballot compaction of the active list (order within the list is irrelevant to the results:
  // every slot's force sum runs over j in a fixed order)
  const unsigned m = __ballot_sync(0xffffffffu, active);
  if (m) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == (__ffs(m) - 1)) base = atomicAdd(&cur->n_act, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (active) g.list[base + __popc(m & ((1u << lane) - 1u))] = i;
  }
  if (MODE == MODE_STEP) block_min_to(c_bits, &nxt->t_next_bits, sh);
  if (i == 0) {
    old->t_next_bits = INF_BITS;
    old->n_act = 0;
    old->work_counter = 0;
  }
}

__global__ void __launch_bounds__(ST_THREADS) k_snapshot_j(const GravDev g) {
  const int i = blockIdx.x * ST_THREADS + threadIdx.x;
  if (i < g.n_loc) {
    g.jpos[g.i0 + i] = g.pos[i];
    double4 v = g.vel[i];
    v.w = 0.0;
    g.jvel[g.i0 + i] = v;
  }
}

// Aarseth estimate; mirrors oracle/hermite_oracle.c: aarseth()
__device__ __forceinline__ double aarseth(const double eta, const double a1[3], const double j1[3],
                                          const double a2[3], const double a3[3]) {
  const double sa = a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2];
  const double sj = j1[0] * j1[0] + j1[1] * j1[1] + j1[2] * j1[2];
  const double s2 = a2[0] * a2[0] + a2[1] * a2[1] + a2[2] * a2[2];
  const double s3 = a3[0] * a3[0] + a3[1] * a3[1] + a3[2] * a3[2];
  const double num = sqrt(sa * s2) + sj;
  const double den = sqrt(sj * s3) + s2;
  if (!(den > 0.0) || !(num > 0.0)) return 1.0e300;
  return eta * sqrt(num / den);
}

// corrector + ladder for one active slot, given its reduced force r[7] = {ax,ay,az,jx,jy,jz,pot}
template <int MODE>
__device__ __forceinline__ void apply_slot(const GravDev &g, const StepCtrl *cur, const int slot, const double r[7],
                                           unsigned long long &c_bits) {
  if (MODE == MODE_RAW) {
    g.raw_a[slot] = make_double4(r[0], r[1], r[2], r[6]);
    g.raw_j[slot] = make_double4(r[3], r[4], r[5], 0.0);
    return;
  }
  const int i = g.list[slot];
  const double a1[3] = {r[0], r[1], r[2]};
  const double j1[3] = {r[3], r[4], r[5]};
  if (MODE == MODE_INIT) {
    g.acc[i] = make_double4(a1[0], a1[1], a1[2], r[6]);
    g.jrk[i] = make_double4(j1[0], j1[1], j1[2], 0.0);
    const double sa = a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2];
    const double sj = j1[0] * j1[0] + j1[1] * j1[1] + j1[2] * j1[2];
    double dt0 = g.dt_max;
    if (sa > 0.0 && sj > 0.0) dt0 = g.eta * 0.0625 * sqrt(sa / sj);
    if (dt0 > 0.03125) dt0 = 0.03125;
    if (dt0 > g.dt_max) dt0 = g.dt_max;
    double dd = pow2floor(dt0);
    if (dd < g.dt_min) dd = g.dt_min;
    g.dt[i] = dd;
    g.t[i] = 0.0;
    return;
  }
  const double4 a0v = g.acc[i], j0v = g.jrk[i];
  const double4 xpv = g.jpos[g.i0 + i], vpv = g.jvel[g.i0 + i];
  const double ti = g.t[i], dti = g.dt[i];
  const double tn = (MODE == MODE_STEP) ? bitsd(cur->t_next_bits) : g.hdr->span;
  const double s = (MODE == MODE_STEP) ? dti : (tn - ti);
  const double a0[3] = {a0v.x, a0v.y, a0v.z}, j0[3] = {j0v.x, j0v.y, j0v.z};
  const double xp[3] = {xpv.x, xpv.y, xpv.z}, vp[3] = {vpv.x, vpv.y, vpv.z};
  double x1[3], v1[3], a2[3], a3[3];
  const double s2 = s * s;
  const double is2 = 1.0 / s2, is3 = 1.0 / (s2 * s);
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const double da = a0[c] - a1[c];
    const double alpha = -3.0 * da - s * (2.0 * j0[c] + j1[c]);
    const double beta = 2.0 * da + s * (j0[c] + j1[c]);
    x1[c] = xp[c] + s2 * (alpha * (1.0 / 12.0) + beta * (1.0 / 20.0));
    v1[c] = vp[c] + s * (alpha * (1.0 / 3.0) + beta * 0.25);
    a2[c] = (2.0 * alpha + 6.0 * beta) * is2;
    a3[c] = (6.0 * beta) * is3;
  }
  const double m = g.pos[i].w;
  g.pos[i] = make_double4(x1[0], x1[1], x1[2], m);
  g.vel[i] = make_double4(v1[0], v1[1], v1[2], 0.0);
  g.acc[i] = make_double4(a1[0], a1[1], a1[2], r[6]);
  g.jrk[i] = make_double4(j1[0], j1[1], j1[2], 0.0);
  double dtA = aarseth(g.eta, a1, j1, a2, a3);
  double nd;
  if (MODE == MODE_STEP) {
    nd = dti;
    if (dtA < dti) {
      if (0.5 * dti >= g.dt_min) nd = 0.5 * dti;
    } else if (dtA >= 2.0 * dti && 2.0 * dti <= g.hdr->D) {
      const double q = tn / (2.0 * dti);
      if (q == floor(q)) nd = 2.0 * dti;
    }
    g.t[i] = tn;
    g.dt[i] = nd;
    const unsigned long long cb = dbits(tn + nd);
    c_bits = cb < c_bits ? cb : c_bits;
  } else {  // MODE_SYNC
    if (dtA > g.dt_max) dtA = g.dt_max;
    nd = pow2floor(dtA);
    if (nd < g.dt_min) nd = g.dt_min;
    g.t[i] = tn;
    g.dt[i] = nd;
  }
}

// Reduction of the j-chunk partials in a FIXED order, then the corrector.
//   few chunks  (n_jsplit <= 32): one warp per active slot, lanes stride the chunks, xor-butterfly;
//   many chunks (tiny blocks cut into up to `grid` chunks): one CTA per slot, all 256 threads load in
//   parallel (one round trip instead of n_jsplit/32 dependent ones), butterfly + ordered warp sum.
template <int MODE>
__global__ void __launch_bounds__(ST_THREADS) k_correct(const GravDev g, const int phase) {
  __shared__ unsigned long long sh[ST_THREADS / 32];
  __shared__ double shr[ST_THREADS / 32][7];
  StepCtrl *cur = &g.ctrl[phase];
  StepCtrl *nxt = &g.ctrl[(phase + 1) % 3];
  const int n_act = cur->n_act;
  if (n_act <= 0) return;
  const Decomp d = make_decomp(n_act, g.n_tot, g.grid_force, g.force_ipt);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = ST_THREADS / 32;
  unsigned long long c_bits = INF_BITS;
  if (d.n_jsplit > 32) {
    for (int slot = blockIdx.x; slot < n_act; slot += gridDim.x) {
      double r[7] = {0, 0, 0, 0, 0, 0, 0};
      for (int js = threadIdx.x; js < d.n_jsplit; js += ST_THREADS) {
        const long long o = (long long)js * d.slot_stride + slot;
        const double4 pa = g.part_a[o], pj = g.part_j[o];
        r[0] += pa.x; r[1] += pa.y; r[2] += pa.z; r[6] += pa.w;
        r[3] += pj.x; r[4] += pj.y; r[5] += pj.z;
      }
#pragma unroll
      for (int c = 0; c < 7; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r[c] += __shfl_xor_sync(0xffffffffu, r[c], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 7; c++) shr[warp][c] = r[c];
      }
      __syncthreads();
      if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < 7; c++) {
          double a = shr[0][c];
          for (int w = 1; w < wpb; w++) a += shr[w][c];
          r[c] = a;
        }
        apply_slot<MODE>(g, cur, slot, r, c_bits);
      }
      __syncthreads();
    }
  } else {
    for (int slot = blockIdx.x * wpb + warp; slot < n_act; slot += gridDim.x * wpb) {
      double r[7] = {0, 0, 0, 0, 0, 0, 0};
      if (lane < d.n_jsplit) {
        const long long o = (long long)lane * d.slot_stride + slot;
        const double4 pa = g.part_a[o], pj = g.part_j[o];
        r[0] = pa.x; r[1] = pa.y; r[2] = pa.z; r[6] = pa.w;
        r[3] = pj.x; r[4] = pj.y; r[5] = pj.z;
      }
#pragma unroll
      for (int c = 0; c < 7; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r[c] += __shfl_xor_sync(0xffffffffu, r[c], o);
      }
      if (lane == 0) apply_slot<MODE>(g, cur, slot, r, c_bits);
    }
  }
  if (MODE == MODE_STEP) block_min_to(c_bits, &nxt->t_next_bits, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0 && MODE != MODE_RAW) {
    if (MODE != MODE_INIT) g.hdr->n_steps += 1;
    g.hdr->n_pairs += (long long)n_act * (long long)g.n_tot;
    int b = 0;
    while ((1 << (b + 1)) <= n_act && b < 31) b++;
    g.hdr->nact_hist[b] += 1;  // diagnostic: log2 histogram of the block sizes
  }
}

static inline int blocks_for(int n) { return (n + ST_THREADS - 1) / ST_THREADS; }

int launch_begin(const GravDev &g, double span, double D, cudaStream_t s) {
  k_begin<<<blocks_for(g.n_loc > 0 ? g.n_loc : 1), ST_THREADS, 0, s>>>(g, span, D);
  return 1;
}

int launch_predict_list(const GravDev &g, int mode, int phase, cudaStream_t s) {
  const int nb = blocks_for(g.n_loc > 0 ? g.n_loc : 1);
  if (mode == MODE_STEP) k_predict_list<MODE_STEP><<<nb, ST_THREADS, 0, s>>>(g, phase);
  else if (mode == MODE_INIT) k_predict_list<MODE_INIT><<<nb, ST_THREADS, 0, s>>>(g, phase);
  else k_predict_list<MODE_SYNC><<<nb, ST_THREADS, 0, s>>>(g, phase);
  return 1;
}

int launch_snapshot_j(const GravDev &g, cudaStream_t s) {
  k_snapshot_j<<<blocks_for(g.n_loc > 0 ? g.n_loc : 1), ST_THREADS, 0, s>>>(g);
  return 1;
}

int launch_correct(const GravDev &g, int mode, int phase, cudaStream_t s) {
  // fixed grid (n_act lives on the device): enough warps for a full block, grid-stride beyond
  int nb = (g.n_loc + (ST_THREADS / 32) - 1) / (ST_THREADS / 32);
  const int cap = 148 * 8;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  if (mode == MODE_STEP) k_correct<MODE_STEP><<<nb, ST_THREADS, 0, s>>>(g, phase);
  else if (mode == MODE_INIT) k_correct<MODE_INIT><<<nb, ST_THREADS, 0, s>>>(g, phase);
  else if (mode == MODE_SYNC) k_correct<MODE_SYNC><<<nb, ST_THREADS, 0, s>>>(g, phase);
  else k_correct<MODE_RAW><<<nb, ST_THREADS, 0, s>>>(g, phase);
  return 1;
}

}  // namespace al26
