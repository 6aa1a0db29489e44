// hermite_loop.cu -- the persistent block-step kernel: predict/schedule -> force -> correct for up to
// `max_steps` consecutive block steps inside ONE cooperative launch, the three phases separated by
// grid-wide barriers instead of kernel boundaries.
//
// Why: with individual block timesteps ~3/4 of all block steps advance fewer than 32 particles
// (profiles/: block-size histogram); at N = 1e5 such a step is a few microseconds of work, so kernel
// launch latency, graph-node dependencies and cold TMA prologues set the price.  Here the CTAs stay
// resident (2 per SM, the force kernel's own configuration), the mbarrier pipeline and the L2-resident
// state stay warm, and the host is only consulted when the whole evolve call is done.
//
// Grid barrier: one monotonic counter in global memory; thread 0 of every CTA arrives with an atomicAdd
// after a __threadfence() and spins (bounded -- a logic error raises hdr->loop_error instead of hanging the
// GPU) until all CTAs of the launch have arrived; the trailing __threadfence() is a gpu-scope acquire, which
// also invalidates the SM's L1 so data written by other SMs in the previous phase is re-read from L2.
// Co-residency of all CTAs is guaranteed by the cooperative launch.
#include <cooperative_groups.h>
#include <cstdlib>

#include "hermite_force.cuh"
#include "hermite_step.cuh"

namespace al26 {

constexpr unsigned LOOP_SPIN_LIMIT = 1u << 28;

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// returns false when the launch must be abandoned (error flag raised by some CTA).
// Arrival is a release-RED at gpu scope on ONE counter (orders the CTA's earlier writes, made visible to thread 0 by the
// bar.sync), the spin is a relaxed volatile load, and one acquire fence after the spin (ptxas: CCTL.IVALL + MEMBAR)
// makes the other CTAs' writes visible to every thread released by the trailing bar.sync.
// Measured and rejected (round 2): a two-level arrival (16 sub-counters, the last arriver of a group -- found from the
// value its atomic returns -- arrives at the top word).  The CTAs do not arrive at once, so the ~300 same-address REDs
// are absorbed as they come and what the step waits for is the latency behind the LAST arrival; the extra atomic round
// trip put 0.7 us on every barrier (N = 1e5, one GPU: 134.4 -> 136.2 us per block step).
__device__ __forceinline__ bool grid_barrier(GravHeader *hdr, unsigned &target, const unsigned n_ctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n_ctas;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&hdr->bar_counter) : "memory");
    unsigned spins = 0;
    while (ld_volatile_u32(&hdr->bar_counter) < target) {
      if (++spins > LOOP_SPIN_LIMIT) {
        atomicExch(&hdr->loop_error, 1);
        break;
      }
      if ((spins & 0xfff) == 0 && ld_volatile_u32((const unsigned *)&hdr->loop_error)) break;
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");  // later TMA reads must see the generic-proxy writes
  }
  __syncthreads();
  return ld_volatile_u32((const unsigned *)&hdr->loop_error) == 0;
}

// ------------------------------------------------------------------------------------------------
// Fused small block steps.  Three quarters of all block steps advance fewer than 32 particles; such a step is
// pure latency (dependent L2 round trips, fences and grid barriers: ~0.4 us per hop), so it gets its own path
// with ONE full grid barrier:
//   scan     every CTA predicts its own contiguous chunk of particles into its shared memory (and into the
//            global j-set), appends the active ones -- predicted state + old force -- to the compact record
//   barrier  (the active set is complete)
//   force    every CTA: the <= 32 active particles (one round trip: count + compact record) against the chunk
//            it still holds in shared memory -- no TMA prologue, no j re-read; lanes split over i and j as in
//            run_item<SPLIT>; one partial per (slot, CTA); release-RED on the step's arrival counter
//   correct  CTA s (s < n_act) is the corrector of slot s: it polls the arrival counter, sums the slot's
//            partials (one per thread -> one round trip; fixed order), runs the corrector from the compact
//            record, folds the new t + dt into the next block time and release-REDs a second counter
//   release  everybody polls that second counter (== n_act), then reads the next block time
// Bitwise reproducible run to run (fixed orders everywhere); the j summation order differs from the big-block
// path's, i.e. the two paths agree to rounding, not bit for bit.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(int *p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// thread 0 only: spin until *p >= want (bounded); acquire.  Returns false when the launch must be abandoned.
__device__ __forceinline__ bool spin_until_ge(GravHeader *hdr, const int *p, const int want, const int err_code) {
  unsigned spins = 0;
  bool ok = true;
  while ((int)ld_volatile_u32((const unsigned *)p) < want) {
    if (++spins > LOOP_SPIN_LIMIT) {
      atomicExch(&hdr->loop_error, err_code);
      ok = false;
      break;
    }
    if ((spins & 0xfff) == 0 && ld_volatile_u32((const unsigned *)&hdr->loop_error)) {
      ok = false;
      break;
    }
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  return ok;
}

#ifdef AL26_FUSE_TIMING
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FT_STAMP(var) const long long var = gtime()
#define FT_ADD(k, dt) g.hdr->fuse_ns[k] += (dt)
#else
#define FT_STAMP(var)
#define FT_ADD(k, dt)
#endif

template <class C>
__device__ __forceinline__ void fused_force(const GravDev &g, ForceSmemT<C> &sm, const int n_act, const int cnt) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int iw = 1;  // lanes that hold distinct i-particles
  while (iw < n_act) iw <<= 1;
  const int isub = lane & (iw - 1), jsub = lane / iw, jgroups = 32 / iw;
  const int li = isub < n_act ? isub : 0;
  const double4 p = ldcg_d4(&g.act->pos[li]);
  const double4 v = ldcg_d4(&g.act->vel[li]);
  Acc7 s;
  s.ax = s.ay = s.az = s.jx = s.jy = s.jz = s.pot = 0.0;
  const double eps2 = g.eps2;
  const double4 *__restrict__ sp = &sm.pos[0][0];
  const double4 *__restrict__ sv = &sm.vel[0][0];
#pragma unroll 2
  for (int jj = warp * jgroups + jsub; jj < cnt; jj += C::WARPS * jgroups)
    pair_interaction(sp[jj], sv[jj], eps2, p.x, p.y, p.z, v.x, v.y, v.z, s);
  for (int o = iw; o < 32; o <<= 1) {  // sum the j-groups of a warp: fixed butterfly over the lane bits above log2(iw)
    s.ax += __shfl_xor_sync(0xffffffffu, s.ax, o); s.ay += __shfl_xor_sync(0xffffffffu, s.ay, o);
    s.az += __shfl_xor_sync(0xffffffffu, s.az, o); s.jx += __shfl_xor_sync(0xffffffffu, s.jx, o);
    s.jy += __shfl_xor_sync(0xffffffffu, s.jy, o); s.jz += __shfl_xor_sync(0xffffffffu, s.jz, o);
    s.pot += __shfl_xor_sync(0xffffffffu, s.pot, o);
  }
  if (lane < iw) {
    sm.red[warp][0][lane] = s.ax; sm.red[warp][1][lane] = s.ay; sm.red[warp][2][lane] = s.az;
    sm.red[warp][3][lane] = s.jx; sm.red[warp][4][lane] = s.jy; sm.red[warp][5][lane] = s.jz;
    sm.red[warp][6][lane] = s.pot;
  }
  __syncthreads();
  if (tid < n_act) {  // fixed-order sum over the warps; partials laid out [slot][CTA]
    double r[7];
#pragma unroll
    for (int c = 0; c < 7; c++) {
      double a = sm.red[0][c][tid];
#pragma unroll
      for (int w = 1; w < C::WARPS; w++) a += sm.red[w][c][tid];
      r[c] = a;
    }
    const long long o = (long long)tid * gridDim.x + blockIdx.x;
    g.part_a[o] = make_double4(r[0], r[1], r[2], r[6]);
    g.part_j[o] = make_double4(r[3], r[4], r[5], 0.0);
  }
}

// the corrector CTA of `slot`: sum the slot's n_parts partials (one per thread and pass; fixed order), corrector,
// fold the new t + dt into the next block time.  count_n as in phase_correct.
template <class C>
__device__ __forceinline__ void fused_correct(const GravDev &g, StepCtrl *nxt, const int slot, const int n_act,
                                              const double tn, const int n_parts, const double Dmax,
                                              double (*shr)[7], const SlotIn &in, const int count_n) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double r[7] = {0, 0, 0, 0, 0, 0, 0};
  const long long row = (long long)slot * gridDim.x;
  for (int c = threadIdx.x; c < n_parts; c += C::THREADS) {
    const double4 pa = ldcg_d4(&g.part_a[row + c]), pj = ldcg_d4(&g.part_j[row + c]);
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
      for (int w = 1; w < C::WARPS; w++) a += shr[w][c];
      r[c] = a;
    }
    unsigned long long c_bits = INF_BITS;
    correct_slot<MODE_STEP, false>(g, tn, in, r, c_bits, 0ull, Dmax);
    atomicMin(&nxt->t_next_bits, c_bits);
    if (slot == 0) {  // accounting (RED: off the critical path)
      atomicAdd((unsigned long long *)&g.hdr->n_steps, 1ull);
      atomicAdd((unsigned long long *)&g.hdr->n_fused, 1ull);
      atomicAdd((unsigned long long *)&g.hdr->n_pairs,
                (unsigned long long)((long long)(count_n >= 0 ? count_n : n_act) * (long long)g.n_tot));
      int b = 0;
      while ((1 << (b + 1)) <= n_act && b < 31) b++;
      atomicAdd((unsigned long long *)&g.hdr->nact_hist[b], 1ull);
    }
  }
}

// force + correctors of a fused step, then wait for the release counter.  Returns false on error.
#ifdef AL26_NOINLINE_FUSED
#define AL26_FUSED_INLINE __noinline__
#else
#define AL26_FUSED_INLINE __forceinline__
#endif
template <class C>
__device__ AL26_FUSED_INLINE bool fused_step(const GravDev &g, ForceSmemT<C> &sm, StepCtrl *cur, StepCtrl *nxt,
                                           const int n_act, const int cnt, const int n_parts, const double tn,
                                           const double Dmax, double (*shr)[7], unsigned long long *sh_word,
                                           const int count_n, unsigned long long &tnext_out) {
  const bool first = (blockIdx.x == 0 && threadIdx.x == 0);
  const bool corrector = (int)blockIdx.x < n_act;
  FT_STAMP(ft0);  // after the barrier + n_act load
  SlotIn in;
  if (corrector && threadIdx.x == 0) {  // the compact record of this CTA's slot: complete since the barrier
    const int slot = blockIdx.x;
    in.i = __ldcg(&g.act->idx[slot]);
    in.a0 = ldcg_d4(&g.act->acc[slot]); in.j0 = ldcg_d4(&g.act->jrk[slot]);
    in.xp = ldcg_d4(&g.act->pos[slot]); in.vp = ldcg_d4(&g.act->vel[slot]);
    const double2 td = __ldcg(&g.act->tdt[slot]);
    in.t = td.x; in.dt = td.y;
  }
  if (cnt > 0) {
    fused_force<C>(g, sm, n_act, cnt);
    __syncthreads();
    if (threadIdx.x == 0) red_release_add(&cur->work_counter, 1);
  }
  FT_STAMP(ft1);
  if (first) FT_ADD(0, ft1 - ft0);  // force + partial store + arrival, CTA 0
  bool ok = true;
  if (corrector) {
    if (threadIdx.x == 0) ok = spin_until_ge(g.hdr, &cur->work_counter, n_parts, 3);
    __syncthreads();
    FT_STAMP(ft2);
    fused_correct<C>(g, nxt, blockIdx.x, n_act, tn, n_parts, Dmax, shr, in, count_n);
    if (threadIdx.x == 0) red_release_add(&cur->pad[1], 1);
    FT_STAMP(ft3);
    if (first) {
      FT_ADD(1, ft2 - ft1);  // CTA 0 (corrector of slot 0): wait for all partials
      FT_ADD(2, ft3 - ft2);  // reduce + correct + arrival
    }
  }
  __syncthreads();
  FT_STAMP(ft4);
  if (threadIdx.x == 0) {
    if (ok) ok = spin_until_ge(g.hdr, &cur->pad[1], n_act, 4);
    *sh_word = __ldcg(&nxt->t_next_bits);
  }
  __syncthreads();
  FT_STAMP(ft5);
  if (first) {
    FT_ADD(4, ft5 - ft4);  // CTA 0: wait for the release counter + next block time
    FT_ADD(5, ft5 - ft0);  // CTA 0: whole fused part
  }
  tnext_out = *sh_word;
  return ld_volatile_u32((const unsigned *)&g.hdr->loop_error) == 0;
}

// FUSE: compile the fused small-step path in.  It is a separate instantiation because the extra live state costs
// the big-block force loop registers (measured: -3 % on N = 1e5 block steps when compiled in but unused).
template <class C, bool FUSE>
__global__ void __launch_bounds__(C::THREADS, C::MINB) k_loop(const __grid_constant__ GravDev g, const int phase0, const int max_steps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ForceSmemT<C> &sm = *reinterpret_cast<ForceSmemT<C> *>(smem_raw);
  __shared__ unsigned long long sh[C::THREADS / 32];
  __shared__ double shr[C::THREADS / 32][7];
  __shared__ unsigned long long sh_word;
  const unsigned n_ctas = gridDim.x;
  const bool first = (blockIdx.x == 0 && threadIdx.x == 0);
  force_smem_init<C>(sm);
  uint32_t it = 0;
  unsigned target = 0;
  int ph = phase0;
  const double span = g.hdr->span;
  const bool fuse = FUSE && g.fuse_max > 0;
  unsigned long long tnext_bits = __ldcg(&g.ctrl[phase0].t_next_bits);
  long long prof[6] = {0, 0, 0, 0, 0, 0}, tk = clock64();  // CTA 0's cycles in A, barrier, B, barrier, C, barrier
#define PROF(k)                      \
  if (!FUSE && first) {              \
    const long long now = clock64(); \
    prof[k] += now - tk;             \
    tk = now;                        \
  }
// this CTA's chunk of the particle set (fused path); recomputed where needed rather than kept in registers
#define CHUNK_J0 min((int)blockIdx.x * g.fuse_jc, g.n_tot)
#define CHUNK_CNT min(g.fuse_jc, g.n_tot - CHUNK_J0)
#define CHUNK_PARTS ((g.n_tot + g.fuse_jc - 1) / g.fuse_jc)
  for (int step = 0; step < max_steps; step++) {
    StepCtrl *cur = &g.ctrl[ph];
    StepCtrl *nxt = &g.ctrl[(ph + 1) % 3];
    StepCtrl *old = &g.ctrl[(ph + 2) % 3];
    // ---- phase A: predictor + scheduler ------------------------------------------------------
    const double tn = bitsd(tnext_bits);
    if (tn > span) {  // uniform: every CTA holds the same value
      if (first) g.hdr->done = 1;
      break;
    }
    FT_STAMP(fa0);
    if (fuse) phase_scan_chunk<false>(g, cur, nxt, tn, CHUNK_J0, CHUNK_CNT, &sm.pos[0][0], &sm.vel[0][0], sh);
    else phase_predict_list<MODE_STEP, false>(g, cur, nxt, tn, blockIdx.x, n_ctas, sh);
    PROF(0)
    FT_STAMP(fa1);
    if (!grid_barrier(g.hdr, target, n_ctas)) break;
    // Recycle the record of the step before last only now: the previous step's fused path ends with every CTA polling
    // old->pad[1] with no barrier behind it, so a reset before this barrier could be seen by a CTA that is still
    // polling.  Past the barrier every CTA has left those polls, and the reset is ordered before the record's next
    // use (as `nxt` of the following step) by this step's later barriers / release counters, all of which CTA 0 takes
    // part in after these stores.
    if (first) {
      old->t_next_bits = INF_BITS;
      old->n_act = 0;
      old->work_counter = 0;
      old->pad[1] = 0;
    }
    PROF(1)
    const int n_act = __ldcg(&cur->n_act);
    FT_STAMP(fa2);
    if (first) {
      FT_ADD(6, fa1 - fa0);  // scan, CTA 0 (all steps)
      FT_ADD(7, fa2 - fa1);  // barrier + n_act load
    }
    if (fuse && n_act > 0 && n_act <= g.fuse_max) {
      // ---- fused small step: force from shared memory, per-slot corrector CTAs, release counter ------
      if (!fused_step<C>(g, sm, cur, nxt, n_act, CHUNK_CNT, CHUNK_PARTS, tn, g.Dmax, shr, &sh_word, -1, tnext_bits)) break;
      PROF(2)
    } else {
      // ---- phase B: force on the active particles ---------------------------------------------
      if (n_act > 0) force_items<C>(g, sm, cur, n_act, n_ctas, it);
      PROF(2)
      if (!grid_barrier(g.hdr, target, n_ctas)) break;
      PROF(3)
      // ---- phase C: reduce partials, corrector, ladder, next block time -------------------------
      if (n_act > 0) phase_correct<MODE_STEP, false>(g, nxt, n_act, tn, blockIdx.x, n_ctas, sh, shr);
      PROF(4)
      if (!grid_barrier(g.hdr, target, n_ctas)) break;
      PROF(5)
      tnext_bits = __ldcg(&nxt->t_next_bits);
    }
    ph = (ph + 1) % 3;
  }
#undef PROF
  if (first) {
    g.hdr->phase = ph;
    if (!FUSE)
      for (int k = 0; k < 6; k++) g.hdr->loop_cycles[k] += prof[k];
  }
}

// ------------------------------------------------------------------------------------------------
// Peer-memory multi-GPU loop.  Every rank runs this kernel on its own GPU at the same
// time.  Per block step: predict ALL particles locally (pulling the records peers staged in the previous
// step), force on the active particles this rank owns, corrector, and the corrected particles are stored
// straight into every rank's staging slab over NVLink.  The third barrier of the step is a cross-GPU one:
// the last CTA of a rank to arrive (so everything the rank wrote is fenced at system scope) posts the rank's
// candidate for the next block time plus the step id into every rank's mailbox; every CTA then waits until
// all `world` mailbox entries carry this step id.  No NCCL, no host, one NVLink round per block step.
// ------------------------------------------------------------------------------------------------
// returns false on error; tmin_out = global minimum of the ranks' candidates (the next block time).
// Mailbox entry of (parity, source rank): two self-validating 64-bit words, each carrying the low 32 bits of the
// step id next to one half of the candidate time, so the two stores need no fence between them and a reader can
// never pair a new tag with an old value.  did_store: this CTA issued peer stores in the corrector phase (only
// those CTAs pay a system-scope fence before arriving; the others fence at gpu scope).
__device__ __forceinline__ bool dist_barrier(const GravDev &g, unsigned &target, const unsigned n_ctas,
                                             const unsigned long long step_id, const StepCtrl *nxt, const bool did_store,
                                             unsigned long long *sh_tmin, unsigned long long &tmin_out) {
  GravHeader *hdr = g.hdr;
  const int par = (int)(step_id & 1ull);
  const unsigned long long tag = (step_id & 0xffffffffull) << 32;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (did_store) __threadfence_system();
    else __threadfence();
    target += n_ctas;
    const unsigned old = atomicAdd(&hdr->bar_counter, 1u);
    if (old == target - 1u) {  // last CTA of this rank: all of the rank's stores (local and peer) are fenced
      __threadfence_system();
      const unsigned long long lm = ld_volatile_u64(&nxt->t_next_bits);
      const unsigned long long w0 = tag | (lm >> 32), w1 = tag | (lm & 0xffffffffull);
      for (int q = 0; q < g.world; q++) {
        MboxEntry *mb = mbox_of(g.slab[q], g.n_tot) + par * MAX_PEERS + g.rank;
        st_volatile_u64(&mb->tmin_bits, w0);
        st_volatile_u64(&mb->tag, w1);
      }
    }
    const MboxEntry *mine = mbox_of(g.slab[g.rank], g.n_tot) + par * MAX_PEERS;
    unsigned spins = 0;
    bool ok = true;
    unsigned long long tm = INF_BITS;
    for (int q = 0; q < g.world && ok; q++) {
      while (true) {
        const unsigned long long w0 = ld_volatile_u64(&mine[q].tmin_bits), w1 = ld_volatile_u64(&mine[q].tag);
        if ((w0 & 0xffffffff00000000ull) == tag && (w1 & 0xffffffff00000000ull) == tag) {
          const unsigned long long v = (w0 << 32) | (w1 & 0xffffffffull);
          tm = v < tm ? v : tm;
          break;
        }
        if (++spins > LOOP_SPIN_LIMIT) {
          atomicExch(&hdr->loop_error, 2);
          ok = false;
          break;
        }
        if ((spins & 0xfff) == 0 && ld_volatile_u32((const unsigned *)&hdr->loop_error)) {
          ok = false;
          break;
        }
      }
    }
    __threadfence_system();
    asm volatile("fence.proxy.async;" ::: "memory");
    *sh_tmin = tm;
  }
  __syncthreads();
  tmin_out = *sh_tmin;
  return ld_volatile_u32((const unsigned *)&hdr->loop_error) == 0;
}

// phase_arg < 0 (chained with the chip engine, hermite_chip.cu): the StepCtrl phase and the exchange id are taken from
// the header, where the previous launch left them -- the host queues [k_chip, k_loop_dist] pairs without reading anything
// back -- and a block step the chip engine can take is handed over to it: the scheduler pass (which has pulled the
// previous exchange) is abandoned, its list / minimum are cleaned up by k_chip, and the launch ends.
template <class C, int MODE, bool FUSE>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
    k_loop_dist(const __grid_constant__ GravDev g, const int phase_arg, const int max_steps, const unsigned long long xid_arg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ForceSmemT<C> &sm = *reinterpret_cast<ForceSmemT<C> *>(smem_raw);
  __shared__ unsigned long long sh[C::THREADS / 32];
  __shared__ double shr[C::THREADS / 32][7];
  __shared__ unsigned long long sh_tmin;
  // time split of the call as CTA 0 sees it (al26_dist_profile): stamps and accumulators live in shared memory, so
  // the instrumentation holds no registers across the force loop; flushed to the header once, at the end
  __shared__ long long tp_stamp[5];
  __shared__ long long tp_acc[DIST_PROF_N];
  const unsigned n_ctas = gridDim.x;
  const bool first = (blockIdx.x == 0 && threadIdx.x == 0);
  if (first)
    for (int k = 0; k < DIST_PROF_N; k++) tp_acc[k] = 0;
#define TP_STAMP(k) if (first) tp_stamp[k] = clock64()
#define TP_ADD(k, a, b) tp_acc[k] += tp_stamp[b] - tp_stamp[a]
  force_smem_init<C>(sm);
  uint32_t it = 0;
  unsigned target = 0;
  const bool chained = (MODE == MODE_STEP) && phase_arg < 0;
  const int phase0 = chained ? g.hdr->phase : phase_arg;
  const unsigned long long xid0 = chained ? g.hdr->dist_step : xid_arg;
  int ph = phase0;
  const double span = g.hdr->span;
  const bool fuse = FUSE && (MODE == MODE_STEP) && g.fuse_max > 0;  // fused small steps: redundant, non-exchanged ones only
  // Exchanges are numbered (xid); a block step with few active particles is NOT exchanged: the state is
  // replicated, so every rank computes all of its active particles itself -- bit-identical on every rank (same
  // code, same inputs, fixed reduction order) -- with no NVLink traffic and no cross-GPU barrier.  The decision is
  // a function of the global active count, hence the same on every rank; ranks drift apart during such steps
  // and meet again at the next exchange.
  unsigned long long xid = xid0;
  // the last block step of the call may itself have been an exchanged one (a dyadic span ends on a coarse level of the
  // block-time hierarchy): the synchronisation step pulls what it staged, like any block step would
  bool prev_exch = (MODE != MODE_INIT) && g.hdr->dist_prev_exch != 0;
  // the next block time: written by k_begin (identical on every rank) or by the previous launch
  unsigned long long tnext_bits = __ldcg(&g.ctrl[phase0].t_next_bits);
  for (int step = 0; step < max_steps; step++) {
    StepCtrl *cur = &g.ctrl[ph];
    StepCtrl *nxt = &g.ctrl[(ph + 1) % 3];
    StepCtrl *old = &g.ctrl[(ph + 2) % 3];
    double tn;
    if (MODE == MODE_STEP) {
      tn = bitsd(tnext_bits);
      if (tn > span) {
        if (first) g.hdr->done = 1;
        break;
      }
    } else if (MODE == MODE_INIT) {
      tn = 0.0;
    } else {
      tn = span;
    }
    TP_STAMP(0);
    if (fuse) phase_scan_chunk<true>(g, cur, nxt, tn, CHUNK_J0, CHUNK_CNT, &sm.pos[0][0], &sm.vel[0][0], sh, prev_exch ? xid : 0ull);
    else phase_predict_list<MODE, true>(g, cur, nxt, tn, blockIdx.x, n_ctas, sh, prev_exch ? xid : 0ull);
    if (!grid_barrier(g.hdr, target, n_ctas)) break;
    TP_STAMP(1);
    if (first) {  // after the barrier: nobody polls the previous step's counters any more (see k_loop)
      old->t_next_bits = INF_BITS;
      old->n_act = 0;
      old->work_counter = 0;
      old->pad[0] = 0;
      old->pad[1] = 0;
    }
    const int n_all = __ldcg(&cur->n_act);
    const int n_own = __ldcg(&cur->pad[0]);
    const bool exchange = (MODE != MODE_STEP) || n_all >= g.split_min;
    if (chained && g.chip_max > 0 && n_all > 0 && n_all <= g.chip_max) {  // uniform: a run of small steps begins
      prev_exch = false;  // this pass has pulled what the last exchange staged
      break;
    }
    if (!exchange && fuse && n_all > 0 && n_all <= g.fuse_max) {
      // ---- redundant AND small: the fused path (one barrier, see above) ----
      if (!fused_step<C>(g, sm, cur, nxt, n_all, CHUNK_CNT, CHUNK_PARTS, tn, g.Dmax, shr, &sh_tmin, n_own, tnext_bits)) break;
      prev_exch = false;
      TP_STAMP(2);
      if (first) {
        TP_ADD(0, 0, 2);
        tp_acc[1] += 1;
      }
    } else if (!exchange) {
      // ---- redundant step: all active particles, local stores, local barrier ----
      if (n_all > 0) force_items<C>(g, sm, cur, n_all, n_ctas, it);
      if (!grid_barrier(g.hdr, target, n_ctas)) break;
      if (n_all > 0) phase_correct<MODE_STEP, false>(g, nxt, n_all, tn, blockIdx.x, n_ctas, sh, shr, 0ull, n_own);
      if (!grid_barrier(g.hdr, target, n_ctas)) break;
      tnext_bits = __ldcg(&nxt->t_next_bits);
      prev_exch = false;
      TP_STAMP(2);
      if (first) {
        TP_ADD(2, 0, 2);
        tp_acc[3] += 1;
        tp_acc[4] += n_all;
      }
    } else {
      // ---- exchanged step: own share, peer stores, cross-GPU barrier ----
      const unsigned long long this_id = xid + 1;
      if (n_own > 0) force_items<C>(g, sm, cur, n_own, n_ctas, it, g.list_own);
      if (!grid_barrier(g.hdr, target, n_ctas)) break;
      TP_STAMP(2);
      if (n_own > 0) phase_correct<MODE, true>(g, nxt, n_own, tn, blockIdx.x, n_ctas, sh, shr, this_id, -1, g.list_own);
      const bool did_store = n_own > 0 && (int)blockIdx.x < n_own;  // superset of the CTAs that corrected a slot
      TP_STAMP(3);
      if (!dist_barrier(g, target, n_ctas, this_id, nxt, did_store, &sh_tmin, tnext_bits)) break;
      xid = this_id;
      prev_exch = true;
      TP_STAMP(4);
      if (first && MODE == MODE_STEP) {
        TP_ADD(5, 0, 1);   // predictor + scheduler + barrier
        TP_ADD(6, 1, 2);   // force on the own share + barrier
        TP_ADD(7, 2, 3);   // corrector + peer stores
        TP_ADD(8, 3, 4);   // cross-GPU barrier
        tp_acc[9] += 1;
        tp_acc[10] += n_all;
      }
    }
    ph = (ph + 1) % 3;
    if (first) g.ctrl[ph].t_next_bits = tnext_bits;  // the next block time, for the host / the next launch
  }
  if (first) {
    g.hdr->phase = ph;
    g.hdr->dist_step = xid;
    g.hdr->dist_tnext_bits = tnext_bits;
    g.hdr->dist_prev_exch = (MODE == MODE_STEP && prev_exch) ? 1 : 0;  // init / sync steps are pulled by k_pull
    for (int k = 0; k < DIST_PROF_N; k++) g.hdr->dist_prof[k] += tp_acc[k];
  }
#undef TP_STAMP
#undef TP_ADD
}

// pull the records staged during block step `step_id` into the local state (after an init / sync step)
__global__ void __launch_bounds__(256) k_pull(const GravDev g, const unsigned long long step_id) {
  const StagingView sv = staging_view(g.slab[g.rank], g.n_tot, (int)(step_id & 1ull));
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.n_loc && sv.tag[i] == (unsigned int)step_id) {
    g.pos[i] = sv.pos[i]; g.vel[i] = sv.vel[i]; g.acc[i] = sv.acc[i]; g.jrk[i] = sv.jrk[i];
    g.t[i] = sv.t[i]; g.dt[i] = sv.dt[i];
    sv.tag[i] = 0u;  // consumed: the next predictor pass must not pull it again (k_begin edits t / dt in between)
  }
}

int launch_pull(const GravDev &g, unsigned long long step_id, cudaStream_t s) {
  k_pull<<<(g.n_loc + 255) / 256, 256, 0, s>>>(g, step_id);
  return 1;
}

template <class C>
static cudaError_t launch_loop_dist_t(const GravDev &g, int mode, void **args, cudaStream_t s) {
  const void *fn = mode == MODE_STEP ? (g.fuse_max > 0 ? (const void *)k_loop_dist<C, MODE_STEP, true>
                                                       : (const void *)k_loop_dist<C, MODE_STEP, false>)
                   : mode == MODE_INIT ? (const void *)k_loop_dist<C, MODE_INIT, false>
                                       : (const void *)k_loop_dist<C, MODE_SYNC, false>;
  return cudaLaunchCooperativeKernel(fn, dim3(g.grid_force), dim3(C::THREADS), args, sizeof(ForceSmemT<C>), s);
}

int launch_loop_dist(const GravDev &g, int mode, int phase, int max_steps, unsigned long long step_id0,
                     cudaStream_t s, cudaError_t *err) {
  GravDev gg = g;
  void *args[] = {(void *)&gg, (void *)&phase, (void *)&max_steps, (void *)&step_id0};
  cudaError_t e = cudaErrorInvalidValue;
  switch (g.variant) {
#define X(id, cfg) case id: e = launch_loop_dist_t<cfg>(g, mode, args, s); break;
    FOR_EACH_FORCE_VARIANT(X)
#undef X
  }
  if (err) *err = e;
  return 1;
}

template <class C>
static cudaError_t launch_loop_t(const GravDev &g, void **args, cudaStream_t s) {
  const void *fn = g.fuse_max > 0 ? (const void *)k_loop<C, true> : (const void *)k_loop<C, false>;
  return cudaLaunchCooperativeKernel(fn, dim3(g.grid_force), dim3(C::THREADS), args, sizeof(ForceSmemT<C>), s);
}

int launch_loop(const GravDev &g, int phase, int max_steps, cudaStream_t s, cudaError_t *err) {
  GravDev gg = g;
  void *args[] = {(void *)&gg, (void *)&phase, (void *)&max_steps};
  cudaError_t e = cudaErrorInvalidValue;
  switch (g.variant) {
#define X(id, cfg) case id: e = launch_loop_t<cfg>(g, args, s); break;
    FOR_EACH_FORCE_VARIANT(X)
#undef X
  }
  if (err) *err = e;
  return 1;
}

template <class K>
static cudaError_t set_smem(K *k, int bytes) {
  return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

cudaError_t loop_kernel_setup() {
  cudaError_t e = cudaSuccess;
#define X(id, cfg)                                                                                       \
  if (e == cudaSuccess) e = set_smem(k_loop<cfg, false>, (int)sizeof(ForceSmemT<cfg>));                   \
  if (e == cudaSuccess) e = set_smem(k_loop<cfg, true>, (int)sizeof(ForceSmemT<cfg>));                    \
  if (e == cudaSuccess) e = set_smem(k_loop_dist<cfg, MODE_STEP, false>, (int)sizeof(ForceSmemT<cfg>));   \
  if (e == cudaSuccess) e = set_smem(k_loop_dist<cfg, MODE_STEP, true>, (int)sizeof(ForceSmemT<cfg>));    \
  if (e == cudaSuccess) e = set_smem(k_loop_dist<cfg, MODE_INIT, false>, (int)sizeof(ForceSmemT<cfg>));   \
  if (e == cudaSuccess) e = set_smem(k_loop_dist<cfg, MODE_SYNC, false>, (int)sizeof(ForceSmemT<cfg>));
  FOR_EACH_FORCE_VARIANT(X)
#undef X
  return e;
}

// largest cooperative grid (CTAs per SM) the loop kernels of this variant can run with
int loop_max_ctas_per_sm(int variant) {
  int n = 0, m = 0;
  switch (variant) {
#define X(id, cfg)                                                                                                    \
  case id:                                                                                                            \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_loop<cfg, false>, cfg::THREADS, sizeof(ForceSmemT<cfg>));       \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, k_loop<cfg, true>, cfg::THREADS, sizeof(ForceSmemT<cfg>));        \
    break;
    FOR_EACH_FORCE_VARIANT(X)
#undef X
  }
  return n < m ? n : m;
}

}  // namespace al26
