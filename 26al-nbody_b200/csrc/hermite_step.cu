// hermite_step.cu -- stand-alone kernels of the block-timestep machinery (used for the init / sync
// steps, the multi-GPU graph path and the parity hooks); the device code lives in hermite_step.cuh.
//   k_begin            start of an evolve call: tau = 0, clamp dt to the call's ladder, first t_next
//   k_predict_list     K2: predictor + scheduler
//   k_correct          K3: partial reduction + corrector + ladder
// Three StepCtrl records rotate (this step / next step / being reset), so a sequence of block steps runs
// with no host in the loop.
#include "hermite_step.cuh"

namespace al26 {

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
    g.hdr->phase = 0;  // the chained peer-memory launches (chip engine + loop kernel) pick the StepCtrl phase up from here
  }
  block_min_to(c, &g.ctrl[0].t_next_bits, sh);
}

template <int MODE>
__global__ void __launch_bounds__(ST_THREADS) k_predict_list(const GravDev g, const int phase) {
  __shared__ unsigned long long sh[ST_THREADS / 32];
  StepCtrl *cur = &g.ctrl[phase];
  StepCtrl *nxt = &g.ctrl[(phase + 1) % 3];
  StepCtrl *old = &g.ctrl[(phase + 2) % 3];
  const bool first = (blockIdx.x == 0 && threadIdx.x == 0);
  double tn;
  if (MODE == MODE_STEP) {
    const unsigned long long tb = cur->t_next_bits;
    tn = bitsd(tb);
    if (tn > g.hdr->span) {  // finished: carry the time forward so later steps in the graph are no-ops too
      if (first) {
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
  phase_predict_list<MODE, false>(g, cur, nxt, tn, blockIdx.x, gridDim.x, sh);
  if (first) {
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

template <int MODE>
__global__ void __launch_bounds__(ST_THREADS) k_correct(const GravDev g, const int phase) {
  __shared__ unsigned long long sh[ST_THREADS / 32];
  __shared__ double shr[ST_THREADS / 32][7];
  StepCtrl *cur = &g.ctrl[phase];
  StepCtrl *nxt = &g.ctrl[(phase + 1) % 3];
  const int n_act = cur->n_act;
  if (n_act <= 0) return;
  const double tn = (MODE == MODE_STEP) ? bitsd(cur->t_next_bits) : g.hdr->span;
  phase_correct<MODE, false>(g, nxt, n_act, tn, blockIdx.x, gridDim.x, sh, shr);
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
