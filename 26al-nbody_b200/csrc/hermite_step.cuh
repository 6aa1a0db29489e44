// hermite_step.cuh -- device-side pieces of the block-timestep machinery, shared by the stand-alone
// kernels (hermite_step.cu) and the persistent loop kernel (hermite_loop.cu).  Every translation unit
// of the library is compiled with --fmad=false: the arithmetic below follows the CPU oracle's
// evaluation order, so given identical inputs every timestep / scheduling decision is bit-identical.
//   phase_predict_list  predictor for every local j fused with the scheduler (row G5):
//                       ballot compaction of THIS step's active list, warp-shuffle + one atomicMin per
//                       block of min(t+dt) over the non-active particles for the NEXT step
//   phase_correct       fixed-order reduction of the force partials, Hermite corrector, Aarseth criterion
//                       and the dyadic ladder (row G4); folds the active particles' new t+dt into the
//                       same minimum
// Stand-ins for ph4's jdata::predict_all, idata::correct and scheduler behind gravity.evolve_model
// (al26_nbody.py:833).
#pragma once
#include "al26_internal.cuh"

namespace al26 {

constexpr int ST_THREADS = 256;

__device__ __forceinline__ unsigned long long dbits(double x) { return (unsigned long long)__double_as_longlong(x); }
__device__ __forceinline__ double bitsd(unsigned long long b) { return __longlong_as_double((long long)b); }

__device__ __forceinline__ double pow2floor(double x) {  // x > 0, normal
  return __longlong_as_double(__double_as_longlong(x) & 0x7FF0000000000000ll);
}

// L2 (cache-global) load of a double4: bypasses the non-coherent L1
__device__ __forceinline__ double4 ldcg_d4(const double4 *p) {
  const double2 a = __ldcg(reinterpret_cast<const double2 *>(p));
  const double2 b = __ldcg(reinterpret_cast<const double2 *>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}

// block-wide min -> one atomicMin per block.  sh: >= blockDim.x/32 entries of shared memory.
__device__ __forceinline__ void block_min_to(unsigned long long v, unsigned long long *dst, unsigned long long *sh) {
  v = warp_min_u64(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : INF_BITS;
    w = warp_min_u64(w);
    if (lane == 0 && w != INF_BITS) atomicMin(dst, w);
  }
}

// Predict every local particle to `tn` into the global j-set, build the active list of this step in
// cur->n_act / g.list, and fold min(t+dt) of the non-active particles into nxt->t_next_bits.
// (block_id, n_blocks) describe the caller's grid; every block runs the same number of iterations.
// DIST (peer-memory mode): the state is replicated; first pull the records peers staged during exchange
// `pull_tag` (parity pull_tag & 1; 0 = nothing to pull) into the local state; build BOTH the list of all active
// particles (g.list, cur->n_act) and the list of those this rank owns (g.list_own, cur->pad[0]) -- the loop
// kernel decides after the barrier which one this step uses.
template <int MODE, bool DIST>
__device__ __forceinline__ void phase_predict_list(const GravDev &g, StepCtrl *cur, StepCtrl *nxt, const double tn,
                                                   const int block_id, const int n_blocks, unsigned long long *sh,
                                                   const unsigned long long pull_tag = 0) {
  const int nthr = blockDim.x;
  unsigned long long c_min = INF_BITS;
  StagingView sv;
  if (DIST) sv = staging_view(g.slab[g.rank], g.n_tot, (int)(pull_tag & 1ull));
  for (int base = block_id * nthr; base < g.n_loc; base += n_blocks * nthr) {
    const int i = base + threadIdx.x;
    bool active = false;
    if (i < g.n_loc) {
      double4 p, v, a, j;
      double ti, dti;
      if (DIST && pull_tag != 0ull && sv.tag[i] == (unsigned int)pull_tag) {  // staged by its owner in the previous block step
        p = sv.pos[i]; v = sv.vel[i]; a = sv.acc[i]; j = sv.jrk[i];
        ti = sv.t[i]; dti = sv.dt[i];
        g.pos[i] = p; g.vel[i] = v; g.acc[i] = a; g.jrk[i] = j;
        g.t[i] = ti; g.dt[i] = dti;
      } else {
        p = g.pos[i]; v = g.vel[i]; a = g.acc[i]; j = g.jrk[i];
        ti = g.t[i]; dti = g.dt[i];
      }
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
      if (!active) {
        const unsigned long long cb = dbits(c);
        c_min = cb < c_min ? cb : c_min;
      }
    }
    // ballot compaction (order within the list is irrelevant to the results: every slot's force sum
    // runs over j in a fixed order)
    const unsigned m = __ballot_sync(0xffffffffu, active);
    if (m) {
      const int lane = threadIdx.x & 31;
      int b0 = 0;
      if (lane == (__ffs(m) - 1)) b0 = atomicAdd(&cur->n_act, __popc(m));
      b0 = __shfl_sync(0xffffffffu, b0, __ffs(m) - 1);
      if (active) g.list[b0 + __popc(m & ((1u << lane) - 1u))] = i;
    }
    if (DIST) {
      const bool mine = active && (i % g.world) == g.rank;
      const unsigned mo = __ballot_sync(0xffffffffu, mine);
      if (mo) {
        const int lane = threadIdx.x & 31;
        int b0 = 0;
        if (lane == (__ffs(mo) - 1)) b0 = atomicAdd(&cur->pad[0], __popc(mo));
        b0 = __shfl_sync(0xffffffffu, b0, __ffs(mo) - 1);
        if (mine) g.list_own[b0 + __popc(mo & ((1u << lane) - 1u))] = i;
      }
    }
  }
  if (MODE == MODE_STEP) block_min_to(c_min, &nxt->t_next_bits, sh);
}

// The scheduler pass of the persistent loop kernels when the fused small-step path is on (MODE_STEP only, state
// not sliced: n_loc == n_tot).  Same work as phase_predict_list, but every CTA owns a CONTIGUOUS chunk
// [j0, j0 + cnt) whose predicted particles it also keeps in its own shared memory (sp / sv, cnt <= the stage
// buffers' capacity), and every active particle's predicted state + old force go to the compact record g.act, so
// a small block step needs neither the TMA pipeline nor dependent loads through the list.
#ifdef AL26_NOINLINE_SCAN
#define AL26_SCAN_INLINE __noinline__
#else
#define AL26_SCAN_INLINE __forceinline__
#endif
template <bool DIST>
__device__ AL26_SCAN_INLINE void phase_scan_chunk(const GravDev &g, StepCtrl *cur, StepCtrl *nxt, const double tn,
                                                 const int j0, const int cnt, double4 *sp, double4 *sv,
                                                 unsigned long long *sh, const unsigned long long pull_tag = 0) {
  unsigned long long c_min = INF_BITS;
  StagingView sview;
  if (DIST) sview = staging_view(g.slab[g.rank], g.n_tot, (int)(pull_tag & 1ull));
  const int lane = threadIdx.x & 31;
  for (int base = 0; base < cnt; base += blockDim.x) {
    const int k = base + threadIdx.x;
    const int i = j0 + k;
    bool active = false;
    double4 pp, pv, a, j;
    double ti = 0.0, dti = 0.0;
    if (k < cnt) {
      double4 p, v;
      if (DIST && pull_tag != 0ull && sview.tag[i] == (unsigned int)pull_tag) {  // staged by its owner in the previous exchange
        p = sview.pos[i]; v = sview.vel[i]; a = sview.acc[i]; j = sview.jrk[i];
        ti = sview.t[i]; dti = sview.dt[i];
        g.pos[i] = p; g.vel[i] = v; g.acc[i] = a; g.jrk[i] = j;
        g.t[i] = ti; g.dt[i] = dti;
      } else {
        p = g.pos[i]; v = g.vel[i]; a = g.acc[i]; j = g.jrk[i];
        ti = g.t[i]; dti = g.dt[i];
      }
      const double s = tn - ti;
      const double s2 = s * s * 0.5, s3 = s * s * s * (1.0 / 6.0);
      pp.x = p.x + v.x * s + a.x * s2 + j.x * s3;
      pp.y = p.y + v.y * s + a.y * s2 + j.y * s3;
      pp.z = p.z + v.z * s + a.z * s2 + j.z * s3;
      pp.w = p.w;
      pv.x = v.x + a.x * s + j.x * s2;
      pv.y = v.y + a.y * s + j.y * s2;
      pv.z = v.z + a.z * s + j.z * s2;
      pv.w = 0.0;
      sp[k] = pp;
      sv[k] = pv;
      g.jpos[i] = pp;  // the big-block path (TMA) and the parity hooks read the global copy
      g.jvel[i] = pv;
      const double c = ti + dti;
      active = (c == tn);
      if (!active) {
        const unsigned long long cb = dbits(c);
        c_min = cb < c_min ? cb : c_min;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, active);
    if (m) {
      int b0 = 0;
      if (lane == (__ffs(m) - 1)) b0 = atomicAdd(&cur->n_act, __popc(m));
      b0 = __shfl_sync(0xffffffffu, b0, __ffs(m) - 1);
      if (active) {
        const int slot = b0 + __popc(m & ((1u << lane) - 1u));
        g.list[slot] = i;
        if (slot < FUSE_CAP) {
          ActBuf *ab = g.act;
          ab->pos[slot] = pp; ab->vel[slot] = pv; ab->acc[slot] = a; ab->jrk[slot] = j;
          ab->tdt[slot] = make_double2(ti, dti);
          ab->idx[slot] = i;
        }
      }
    }
    if (DIST) {
      const bool mine = active && (i % g.world) == g.rank;
      const unsigned mo = __ballot_sync(0xffffffffu, mine);
      if (mo) {
        int b0 = 0;
        if (lane == (__ffs(mo) - 1)) b0 = atomicAdd(&cur->pad[0], __popc(mo));
        b0 = __shfl_sync(0xffffffffu, b0, __ffs(mo) - 1);
        if (mine) g.list_own[b0 + __popc(mo & ((1u << lane) - 1u))] = i;
      }
    }
  }
  // the stage buffers were written through the generic proxy; a later TMA refill is an async-proxy write
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  block_min_to(c_min, &nxt->t_next_bits, sh);
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

struct NewState {
  double4 pos, vel, acc, jrk;
  double t, dt;
};

// With softening the force kernel's masked rsqrt does not drop the self pair (r^2 + eps2 = eps2 > 0): dx = 0 nulls its
// acceleration and jerk, but -m_i/eps lands in the potential.  Taken out again here, so `pot` is the potential of the
// OTHER particles, as in every direct N-body code (and the oracle).
__device__ __forceinline__ double pot_without_self(const double pot, const double m_i, const double eps2) {
  return (eps2 > 0.0) ? pot + m_i * (1.0 / sqrt(eps2)) : pot;
}

// where a corrected particle goes: the local state, or (peer-memory mode) every rank's staging slab
template <bool DIST>
__device__ __forceinline__ void store_state(const GravDev &g, const int i, const NewState &n, const unsigned long long step_id) {
  if (!DIST) {
    g.pos[i] = n.pos; g.vel[i] = n.vel; g.acc[i] = n.acc; g.jrk[i] = n.jrk;
    g.t[i] = n.t; g.dt[i] = n.dt;
  } else {
    for (int q = 0; q < g.world; q++) {  // NVLink peer stores (q == rank: local)
      const StagingView v = staging_view(g.slab[q], g.n_tot, (int)(step_id & 1ull));
      v.pos[i] = n.pos; v.vel[i] = n.vel; v.acc[i] = n.acc; v.jrk[i] = n.jrk;
      v.t[i] = n.t; v.dt[i] = n.dt;
      v.tag[i] = (unsigned int)step_id;
    }
  }
}

// what the corrector needs to know about one active particle besides its new force
struct SlotIn {
  int i;
  double4 a0, j0;  // force at the start of the step
  double4 xp, vp;  // predicted position (w = mass) and velocity
  double t, dt;
};

// corrector + ladder for one active particle, given its reduced force r[7] = {ax,ay,az,jx,jy,jz,pot}; tn = the
// block time (MODE_STEP) or the span (MODE_SYNC); Dmax = the largest step of the call's ladder (hdr->D)
template <int MODE, bool DIST>
__device__ __forceinline__ void correct_slot(const GravDev &g, const double tn, const SlotIn &in, const double r[7],
                                             unsigned long long &c_bits, const unsigned long long step_id,
                                             const double Dmax, NewState *out = nullptr) {
  const int i = in.i;
  const double a1[3] = {r[0], r[1], r[2]};
  const double j1[3] = {r[3], r[4], r[5]};
  NewState n;
  n.acc = make_double4(a1[0], a1[1], a1[2], pot_without_self(r[6], in.xp.w, g.eps2));
  n.jrk = make_double4(j1[0], j1[1], j1[2], 0.0);
  const double4 a0v = in.a0, j0v = in.j0;
  const double4 xpv = in.xp, vpv = in.vp;
  const double ti = in.t, dti = in.dt;
  const double s = (MODE == MODE_STEP) ? dti : (tn - ti);
  const double a0[3] = {a0v.x, a0v.y, a0v.z}, j0[3] = {j0v.x, j0v.y, j0v.z};
  const double xp[3] = {xpv.x, xpv.y, xpv.z}, vp[3] = {vpv.x, vpv.y, vpv.z};
  double x1[3], v1[3], a2[3], a3[3];
  const double s2 = s * s;
  // 1/s^2, 1/s^3: in a block step s is a power of two (>= dt_min = 2^-40 by default), so the reciprocal is an exponent
  // flip and the powers are exact products -- bit-identical to the oracle's divisions (a quotient by an exact power of
  // two is exact), without three divisions in the step's dependent chain
  double is1 = 0.0, is2, is3;
  if (MODE == MODE_STEP) {
    is1 = __longlong_as_double(0x7FE0000000000000ll - __double_as_longlong(s));
    is2 = is1 * is1;
    is3 = is2 * is1;
  } else {
    is2 = 1.0 / s2;
    is3 = 1.0 / (s2 * s);
  }
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
  n.pos = make_double4(x1[0], x1[1], x1[2], xpv.w);  // jpos.w carries the mass
  n.vel = make_double4(v1[0], v1[1], v1[2], 0.0);
  double dtA = aarseth(g.eta, a1, j1, a2, a3);
  double nd;
  if (MODE == MODE_STEP) {
    nd = dti;
    if (dtA < dti) {
      if (0.5 * dti >= g.dt_min) nd = 0.5 * dti;
    } else if (dtA >= 2.0 * dti && 2.0 * dti <= Dmax) {
      const double q = tn * (0.5 * is1);  // == tn / (2 dt), exactly
      if (q == floor(q)) nd = 2.0 * dti;
    }
    const unsigned long long cb = dbits(tn + nd);
    c_bits = cb < c_bits ? cb : c_bits;
  } else {  // MODE_SYNC
    if (dtA > g.dt_max) dtA = g.dt_max;
    nd = pow2floor(dtA);
    if (nd < g.dt_min) nd = g.dt_min;
  }
  n.t = tn;
  n.dt = nd;
  store_state<DIST>(g, i, n, step_id);
  if (out) *out = n;
}

// one active slot of the list: raw output, initial timestep, or the corrector on the particle's global records
template <int MODE, bool DIST>
__device__ __forceinline__ void apply_slot(const GravDev &g, const double tn, const int slot, const double r[7],
                                           unsigned long long &c_bits, const unsigned long long step_id,
                                           const int *__restrict__ list) {
  if (MODE == MODE_RAW) {
    g.raw_a[slot] = make_double4(r[0], r[1], r[2], pot_without_self(r[6], g.jpos[g.i0 + list[slot]].w, g.eps2));
    g.raw_j[slot] = make_double4(r[3], r[4], r[5], 0.0);
    return;
  }
  const int i = list[slot];
  if (MODE == MODE_INIT) {
    const double a1[3] = {r[0], r[1], r[2]};
    const double j1[3] = {r[3], r[4], r[5]};
    NewState n;
    n.acc = make_double4(a1[0], a1[1], a1[2], pot_without_self(r[6], g.jpos[g.i0 + i].w, g.eps2));
    n.jrk = make_double4(j1[0], j1[1], j1[2], 0.0);
    const double sa = a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2];
    const double sj = j1[0] * j1[0] + j1[1] * j1[1] + j1[2] * j1[2];
    double dt0 = g.dt_max;
    if (sa > 0.0 && sj > 0.0) dt0 = g.eta * 0.0625 * sqrt(sa / sj);
    if (dt0 > 0.03125) dt0 = 0.03125;
    if (dt0 > g.dt_max) dt0 = g.dt_max;
    double dd = pow2floor(dt0);
    if (dd < g.dt_min) dd = g.dt_min;
    n.dt = g.keep_dt ? g.dt[i] : dd;
    n.t = 0.0;
    if (!DIST) {  // positions and velocities are untouched
      g.acc[i] = n.acc; g.jrk[i] = n.jrk; g.dt[i] = n.dt; g.t[i] = n.t;
    } else {
      n.pos = g.pos[i];
      n.vel = g.vel[i];
      store_state<true>(g, i, n, step_id);
    }
    return;
  }
  SlotIn in;
  in.i = i;
  in.a0 = g.acc[i]; in.j0 = g.jrk[i];
  in.xp = g.jpos[g.i0 + i]; in.vp = g.jvel[g.i0 + i];
  in.t = g.t[i]; in.dt = g.dt[i];
  correct_slot<MODE, DIST>(g, tn, in, r, c_bits, step_id, g.hdr->D);
}

// Reduction of the j-chunk partials in a FIXED order, then the corrector.
//   few chunks  (n_jsplit <= 32): one warp per active slot, lanes take one chunk each, xor-butterfly;
//   many chunks (tiny blocks cut into up to `grid` chunks): one CTA per slot, all threads load in
//   parallel (one round trip instead of n_jsplit/32 dependent ones), butterfly + ordered warp sum.
// sh: blockDim/32 u64, shr: blockDim/32 x 7 doubles of shared memory.
template <int MODE, bool DIST>
__device__ __forceinline__ void phase_correct(const GravDev &g, StepCtrl *nxt, const int n_act, const double tn,
                                              const int block_id, const int n_blocks, unsigned long long *sh,
                                              double (*shr)[7], const unsigned long long step_id = 0,
                                              const int count_n = -1, const int *__restrict__ list = nullptr) {
  if (!list) list = g.list;
  const Decomp d = make_decomp(n_act, g.n_tot, g.decomp_tab, g.force_ipt, g.big_nact);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  unsigned long long c_bits = INF_BITS;
  if (d.n_jsplit > 32) {
    for (int slot = block_id; slot < n_act; slot += n_blocks) {
      double r[7] = {0, 0, 0, 0, 0, 0, 0};
      for (int js = threadIdx.x; js < d.n_jsplit; js += blockDim.x) {
        const long long o = (long long)js * d.slot_stride + slot;
        const double4 pa = ldcg_d4(&g.part_a[o]), pj = ldcg_d4(&g.part_j[o]);
        r[0] += pa.x; r[1] += pa.y; r[2] += pa.z; r[6] += pa.w;
        r[3] += pj.x; r[4] += pj.y; r[5] += pj.z;
      }
#pragma unroll
      for (int c = 0; c < 7; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r[c] += __shfl_xor_sync(0xffffffffu, r[c], o);
      }
      __syncthreads();
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
        apply_slot<MODE, DIST>(g, tn, slot, r, c_bits, step_id, list);
      }
    }
  } else {
    for (int slot = block_id * wpb + warp; slot < n_act; slot += n_blocks * wpb) {
      double r[7] = {0, 0, 0, 0, 0, 0, 0};
      if (lane < d.n_jsplit) {
        const long long o = (long long)lane * d.slot_stride + slot;
        const double4 pa = ldcg_d4(&g.part_a[o]), pj = ldcg_d4(&g.part_j[o]);
        r[0] = pa.x; r[1] = pa.y; r[2] = pa.z; r[6] = pa.w;
        r[3] = pj.x; r[4] = pj.y; r[5] = pj.z;
      }
#pragma unroll
      for (int c = 0; c < 7; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r[c] += __shfl_xor_sync(0xffffffffu, r[c], o);
      }
      if (lane == 0) apply_slot<MODE, DIST>(g, tn, slot, r, c_bits, step_id, list);
    }
  }
  if (MODE == MODE_STEP) block_min_to(c_bits, &nxt->t_next_bits, sh);
  if (block_id == 0 && threadIdx.x == 0 && MODE != MODE_RAW) {
    if (MODE != MODE_INIT) g.hdr->n_steps += 1;
    // count_n: the particles whose pairs this rank accounts for (its own share when a step is computed redundantly)
    g.hdr->n_pairs += (long long)(count_n >= 0 ? count_n : n_act) * (long long)g.n_tot;
    int b = 0;
    while ((1 << (b + 1)) <= n_act && b < 31) b++;
    g.hdr->nact_hist[b] += 1;  // diagnostic: log2 histogram of the block sizes
  }
}

}  // namespace al26
