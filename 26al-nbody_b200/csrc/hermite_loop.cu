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

// returns false when the launch must be abandoned (error flag raised by some CTA)
__device__ __forceinline__ bool grid_barrier(GravHeader *hdr, unsigned &target, const unsigned n_ctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    target += n_ctas;
    atomicAdd(&hdr->bar_counter, 1u);
    unsigned spins = 0;
    while (ld_volatile_u32(&hdr->bar_counter) < target) {
      if (++spins > LOOP_SPIN_LIMIT) {
        atomicExch(&hdr->loop_error, 1);
        break;
      }
      if ((spins & 0xfff) == 0 && ld_volatile_u32((const unsigned *)&hdr->loop_error)) break;
    }
    __threadfence();
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
    phase_predict_list<MODE_STEP>(g, cur, nxt, tn, blockIdx.x, n_ctas, sh);
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
    if (n_act > 0) phase_correct<MODE_STEP>(g, cur, nxt, n_act, blockIdx.x, n_ctas, sh, shr);
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
#define X(id, cfg) if (e == cudaSuccess) e = cudaFuncSetAttribute(k_loop<cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ForceSmemT<cfg>));
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
