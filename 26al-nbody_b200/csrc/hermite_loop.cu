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
// Arrival is a release-RED at gpu scope (orders the CTA's earlier writes, made visible to thread 0 by the
// bar.sync), the spin is a relaxed volatile load, and one acquire fence after the spin (ptxas: CCTL.IVALL +
// MEMBAR) makes the other CTAs' writes visible to every thread released by the trailing bar.sync.
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

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB) k_loop(const GravDev g, const int phase0, const int max_steps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ForceSmemT<C> &sm = *reinterpret_cast<ForceSmemT<C> *>(smem_raw);
  __shared__ unsigned long long sh[C::THREADS / 32];
  __shared__ double shr[C::THREADS / 32][7];
  const unsigned n_ctas = gridDim.x;
  const bool first = (blockIdx.x == 0 && threadIdx.x == 0);
  force_smem_init<C>(sm);
  uint32_t it = 0;
  unsigned target = 0;
  int ph = phase0;
  const double span = g.hdr->span;
  long long prof[6] = {0, 0, 0, 0, 0, 0}, tk = clock64();  // CTA 0's cycles in A, barrier, B, barrier, C, barrier
#define PROF(k)                    \
  if (first) {                     \
    const long long now = clock64(); \
    prof[k] += now - tk;           \
    tk = now;                      \
  }
  for (int step = 0; step < max_steps; step++) {
    StepCtrl *cur = &g.ctrl[ph];
    StepCtrl *nxt = &g.ctrl[(ph + 1) % 3];
    StepCtrl *old = &g.ctrl[(ph + 2) % 3];
    // ---- phase A: predictor + scheduler ------------------------------------------------------
    const unsigned long long tb = __ldcg(&cur->t_next_bits);
    const double tn = bitsd(tb);
    if (tn > span) {  // uniform: every CTA reads the same record after the previous barrier
      if (first) g.hdr->done = 1;
      break;
    }
    phase_predict_list<MODE_STEP, false>(g, cur, nxt, tn, blockIdx.x, n_ctas, sh);
    if (first) {
      old->t_next_bits = INF_BITS;
      old->n_act = 0;
      old->work_counter = 0;
    }
    PROF(0)
    if (!grid_barrier(g.hdr, target, n_ctas)) break;
    PROF(1)
    // ---- phase B: force on the active particles ---------------------------------------------
    const int n_act = __ldcg(&cur->n_act);
    if (n_act > 0) force_items<C>(g, sm, cur, n_act, n_ctas, it);
    PROF(2)
    if (!grid_barrier(g.hdr, target, n_ctas)) break;
    PROF(3)
    // ---- phase C: reduce partials, corrector, ladder, next block time -------------------------
    if (n_act > 0) phase_correct<MODE_STEP, false>(g, nxt, n_act, tn, blockIdx.x, n_ctas, sh, shr);
    PROF(4)
    if (!grid_barrier(g.hdr, target, n_ctas)) break;
    PROF(5)
    ph = (ph + 1) % 3;
  }
#undef PROF
  if (first) {
    g.hdr->phase = ph;
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
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

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

template <class C, int MODE>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
    k_loop_dist(const GravDev g, const int phase0, const int max_steps, const unsigned long long xid0) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ForceSmemT<C> &sm = *reinterpret_cast<ForceSmemT<C> *>(smem_raw);
  __shared__ unsigned long long sh[C::THREADS / 32];
  __shared__ double shr[C::THREADS / 32][7];
  __shared__ unsigned long long sh_tmin;
  const unsigned n_ctas = gridDim.x;
  const bool first = (blockIdx.x == 0 && threadIdx.x == 0);
  force_smem_init<C>(sm);
  uint32_t it = 0;
  unsigned target = 0;
  int ph = phase0;
  const double span = g.hdr->span;
  // Exchanges are numbered (xid); a block step with few active particles is NOT exchanged: the state is
  // replicated, so every rank computes all of its active particles itself -- bit-identical on every rank (same
  // code, same inputs, fixed reduction order) -- with no NVLink traffic and no cross-GPU barrier.  The decision is
  // a function of the global active count, hence the same on every rank; ranks drift apart during such steps
  // and meet again at the next exchange.
  unsigned long long xid = xid0;
  bool prev_exch = (MODE == MODE_STEP) && g.hdr->dist_prev_exch != 0;
  // the next block time: written by k_begin (identical on every rank) or by the previous launch
  unsigned long long tnext_bits = __ldcg(&g.ctrl[phase0].t_next_bits);
  GravDev gown = g;
  gown.list = g.list_own;
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
    phase_predict_list<MODE, true>(g, cur, nxt, tn, blockIdx.x, n_ctas, sh, prev_exch ? xid : 0ull);
    if (first) {
      old->t_next_bits = INF_BITS;
      old->n_act = 0;
      old->work_counter = 0;
      old->pad[0] = 0;
    }
    if (!grid_barrier(g.hdr, target, n_ctas)) break;
    const int n_all = __ldcg(&cur->n_act);
    const int n_own = __ldcg(&cur->pad[0]);
    const bool exchange = (MODE != MODE_STEP) || n_all >= g.split_min;
    if (!exchange) {
      // ---- redundant step: all active particles, local stores, local barrier ----
      if (n_all > 0) force_items<C>(g, sm, cur, n_all, n_ctas, it);
      if (!grid_barrier(g.hdr, target, n_ctas)) break;
      if (n_all > 0) phase_correct<MODE_STEP, false>(g, nxt, n_all, tn, blockIdx.x, n_ctas, sh, shr, 0ull, n_own);
      if (!grid_barrier(g.hdr, target, n_ctas)) break;
      tnext_bits = __ldcg(&nxt->t_next_bits);
      prev_exch = false;
    } else {
      // ---- exchanged step: own share, peer stores, cross-GPU barrier ----
      const unsigned long long this_id = xid + 1;
      if (n_own > 0) force_items<C>(gown, sm, cur, n_own, n_ctas, it);
      if (!grid_barrier(g.hdr, target, n_ctas)) break;
      if (n_own > 0) phase_correct<MODE, true>(gown, nxt, n_own, tn, blockIdx.x, n_ctas, sh, shr, this_id);
      const bool did_store = n_own > 0 && (int)blockIdx.x < n_own;  // superset of the CTAs that corrected a slot
      if (!dist_barrier(g, target, n_ctas, this_id, nxt, did_store, &sh_tmin, tnext_bits)) break;
      xid = this_id;
      prev_exch = true;
    }
    ph = (ph + 1) % 3;
    if (first) g.ctrl[ph].t_next_bits = tnext_bits;  // the next block time, for the host / the next launch
  }
  if (first) {
    g.hdr->phase = ph;
    g.hdr->dist_step = xid;
    g.hdr->dist_tnext_bits = tnext_bits;
    g.hdr->dist_prev_exch = (MODE == MODE_STEP && prev_exch) ? 1 : 0;  // init / sync steps are pulled by k_pull
  }
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
  const void *fn = mode == MODE_STEP ? (const void *)k_loop_dist<C, MODE_STEP>
                   : mode == MODE_INIT ? (const void *)k_loop_dist<C, MODE_INIT> : (const void *)k_loop_dist<C, MODE_SYNC>;
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

int launch_loop(const GravDev &g, int phase, int max_steps, cudaStream_t s, cudaError_t *err) {
  GravDev gg = g;
  void *args[] = {(void *)&gg, (void *)&phase, (void *)&max_steps};
  cudaError_t e = cudaErrorInvalidValue;
  switch (g.variant) {
#define X(id, cfg)                                                                                                   \
  case id:                                                                                                           \
    e = cudaLaunchCooperativeKernel((const void *)k_loop<cfg>, dim3(g.grid_force), dim3(cfg::THREADS), args,          \
                                    sizeof(ForceSmemT<cfg>), s);                                                      \
    break;
    FOR_EACH_FORCE_VARIANT(X)
#undef X
  }
  if (err) *err = e;
  return 1;
}

cudaError_t loop_kernel_setup() {
  cudaError_t e = cudaSuccess;
#define X(id, cfg)                                                                                                  \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_loop<cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ForceSmemT<cfg>)); \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_loop_dist<cfg, MODE_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ForceSmemT<cfg>)); \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_loop_dist<cfg, MODE_INIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ForceSmemT<cfg>)); \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_loop_dist<cfg, MODE_SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ForceSmemT<cfg>));
  FOR_EACH_FORCE_VARIANT(X)
#undef X
  return e;
}

// largest cooperative grid (CTAs per SM) the loop kernel of this variant can run with
int loop_max_ctas_per_sm(int variant) {
  int n = 0;
  switch (variant) {
#define X(id, cfg) case id: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_loop<cfg>, cfg::THREADS, sizeof(ForceSmemT<cfg>)); break;
    FOR_EACH_FORCE_VARIANT(X)
#undef X
  }
  return n;
}

}  // namespace al26
