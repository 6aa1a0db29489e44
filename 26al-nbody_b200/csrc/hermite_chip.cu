// hermite_chip.cu -- the chip engine: runs of SMALL block steps with the whole particle set resident in the shared
// memory of the chip (one CTA per SM), for the N that no longer fits one thread-block cluster (hermite_engine.cu).
//
// Why: at N = 1e5 (BASELINE config 3) nine block steps out of ten advance fewer than 256 particles and three quarters
// fewer than 32; through the grid-wide kernels such a step costs 14 us on one GPU (three kernel boundaries, a full
// predictor pass over 14 MB of state, a TMA prologue, a partial-sum pass) and 21 us inside the multi-GPU loop kernel,
// where every rank computes it redundantly -- half of an 8-GPU outer step.  Here:
//   * CTA c owns the contiguous chunk c of the particles and keeps their full state (pos, vel, acc, jerk, t, dt: 144 B
//     per particle, + 64 B of predicted state) in its shared memory for the whole launch; corrected particles are written
//     through to the global records, which stay authoritative;
//   * there is NO grid barrier and NO atomic.  Every word that crosses CTAs has exactly one writer and carries the step
//     number it belongs to:
//       mail[c]   header {step, min(t+dt) over chunk c, number of particles that attain it} + the predicted state of
//                 those particles (published speculatively: if chunk c's minimum turns out to be the global one, these
//                 ARE the step's active particles, already predicted to the block time);
//       flag[c]   "CTA c's force partials of step s are stored";
//       part[slot][c]  the partial force of chunk c on active particle `slot`.
//   * per block step:  every CTA derives the block time, the owners and the slot numbering from its copy of the 148
//     headers (one warp), fetches the owners' records (one L2 round trip), predicts its chunk, computes the <= CHIP_CAP
//     active particles against it (lanes split over i and j for tiny blocks, tiles of 32 i for bigger ones; fixed-order
//     sums), stores one partial per slot and raises its flag.  An OWNER waits for all flags, sums its slots' partials in
//     CTA order (one warp per slot), runs the same correct_slot as every other path on the resident state, rescans its
//     chunk and publishes the new header + records.  Everybody waits for the owners' headers only: the other chunks'
//     minima cannot have changed.
//   Critical path of a step: flag -> partials -> corrector -> header -> records -> force, i.e. 5-6 dependent L2 hops
//   instead of 10-18, and no pass over the state in L2.
// The launch ends when the next block has more than `chip_max` active particles (the grid-wide kernels that follow
// take it) or the span is exhausted.  Same arithmetic as the other paths except the order of the j sum (CTA chunks):
// identical integer work, positions to rounding.  Stands in, like the rest, for ph4's evolve loop behind
// gravity.evolve_model (al26_nbody.py:833).
#include <cooperative_groups.h>

#include "hermite_force.cuh"
#include "hermite_step.cuh"

namespace al26 {

constexpr int CHIP_T = 512;
constexpr int CHIP_WARPS = CHIP_T / 32;
static_assert(CHIP_WARPS == 16, "the cross-warp sum below is a 16-lane butterfly");
constexpr int CHIP_ENT = CHIP_MAX_CTAS / 32;  // header entries per lane of the scheduling warp
constexpr unsigned CHIP_SPIN_LIMIT = 1u << 22;

struct alignas(32) ChipShared {
  double red[CHIP_WARPS][7][32];
  double4 a_pp[CHIP_CAP], a_pv[CHIP_CAP];  // the step's active set (predicted), gathered from the owners' mail
  int a_idx[CHIP_CAP];
  unsigned long long cmin[CHIP_MAX_CTAS];  // every chunk's min(t + dt) ...
  int ccount[CHIP_MAX_CTAS];               // ... and how many of its particles attain it
  int own_cta[CHIP_MAX_CTAS], own_base[CHIP_MAX_CTAS], own_cnt[CHIP_MAX_CTAS];  // this step's owners, in CTA order
  unsigned long long wmin[CHIP_WARPS];
  unsigned long long tn_bits;
  int n_act, n_own, my_base, my_cnt, cand, pad[3];
};
static_assert(sizeof(ChipShared) % 32 == 0, "ChipShared must keep the double4 arrays behind it aligned");

constexpr int CHIP_BYTES_PER_PARTICLE = 6 * 32 + 2 * 8;  // pos, vel, acc, jrk, ppos, pvel, t, dt

int chip_smem_bytes(int p_cap) { return (int)sizeof(ChipShared) + p_cap * CHIP_BYTES_PER_PARTICLE; }
size_t chip_mail_bytes(int n_ctas) { return (size_t)n_ctas * sizeof(ChipMail) + (size_t)n_ctas * sizeof(unsigned); }

// n particles over n_ctas CTAs: per-CTA capacity, or false when the chunks do not fit `max_smem` bytes per block
bool chip_plan(int n, int n_ctas, int max_smem, int *p_cap_out) {
  if (n < 1 || n_ctas < 1 || n_ctas > CHIP_MAX_CTAS) return false;
  const int p_cap = (((n + n_ctas - 1) / n_ctas) + 7) & ~7;
  if (chip_smem_bytes(p_cap) > max_smem) return false;
  *p_cap_out = p_cap;
  return true;
}

namespace {

__device__ __forceinline__ unsigned long long ldv_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stv_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ldv_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stv_u32(unsigned *p, unsigned v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Block-wide poll: every thread evaluates `ready()` (threads with nothing to wait for return true) until all do.
// Bounded: a protocol error raises hdr->loop_error (code 5) instead of hanging the GPU; returns false when the launch
// must be abandoned (uniform over the block).  Ends with a gpu-scope acquire by the polling threads; the block barrier
// inside __syncthreads_and hands the visibility on to the other threads, which read remote data with ld.cg only.
template <class F>
__device__ __forceinline__ bool chip_poll(GravHeader *hdr, F ready) {
  unsigned spins = 0;
  while (true) {
    const bool ok = ready();
    if (__syncthreads_and(ok)) break;
    if ((++spins & 1023u) == 0u) {
      int bad = 0;
      if (threadIdx.x == 0) {
        if (spins > CHIP_SPIN_LIMIT) atomicExch(&hdr->loop_error, 5);
        bad = ldv_u32((const unsigned *)&hdr->loop_error) != 0u;
      }
      if (__syncthreads_or(bad)) return false;
    }
  }
  fence_gpu();
  return true;
}

// header of one chunk: three self-validating words (step number in the high half), so a reader can never pair a new
// step number with an old value and the writer needs no fence between them
__device__ __forceinline__ void mail_put_header(ChipMail *m, unsigned seq, unsigned long long min_bits, int count) {
  const unsigned long long tag = (unsigned long long)seq << 32;
  stv_u64(&m->w[0], tag | (min_bits >> 32));
  stv_u64(&m->w[1], tag | (min_bits & 0xffffffffull));
  stv_u64(&m->w[2], tag | (unsigned long long)(unsigned)count);
}
__device__ __forceinline__ bool mail_get_header(const ChipMail *m, unsigned seq, unsigned long long &min_bits, int &count) {
  const unsigned long long tag = (unsigned long long)seq << 32, hi = 0xffffffff00000000ull;
  const unsigned long long w0 = ldv_u64(&m->w[0]), w1 = ldv_u64(&m->w[1]), w2 = ldv_u64(&m->w[2]);
  if ((w0 & hi) != tag || (w1 & hi) != tag || (w2 & hi) != tag) return false;
  min_bits = (w0 << 32) | (w1 & 0xffffffffull);
  count = (int)(unsigned)(w2 & 0xffffffffull);
  return true;
}

__device__ __forceinline__ void predict_to(const double tn, const double ti, const double4 p, const double4 v, const double4 a,
                                           const double4 j, double4 &pp, double4 &pv) {
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
}

}  // namespace

// phase_arg < 0: chained with the peer-memory loop kernel (hermite_loop.cu) -- the StepCtrl phase is where that kernel
// left it in the header
__global__ void __launch_bounds__(CHIP_T, 1) k_chip(const GravDev g, const int phase_arg, const int p_cap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ChipShared &S = *reinterpret_cast<ChipShared *>(smem_raw);
  double4 *pos = reinterpret_cast<double4 *>(smem_raw + sizeof(ChipShared));
  double4 *vel = pos + p_cap, *acc = vel + p_cap, *jrk = acc + p_cap, *ppos = jrk + p_cap, *pvel = ppos + p_cap;
  double *tt = reinterpret_cast<double *>(pvel + p_cap), *dtt = tt + p_cap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int me = blockIdx.x, nc = gridDim.x;

  const int phase = phase_arg < 0 ? g.hdr->phase : phase_arg;
  StepCtrl *ctl = &g.ctrl[phase];
  const double span = g.hdr->span;
  const double Dmax = g.hdr->D;  // not g.Dmax: a graph captures g by value at commit
  if (bitsd(ctl->t_next_bits) > span) return;  // uniform over the grid; the predict kernel that follows raises `done`
  const int max_act = g.chip_max < CHIP_CAP ? g.chip_max : CHIP_CAP;
  if (me == 0 && tid == 0) {
    // the loop kernel hands a run of small steps over after its scheduler pass has started on the step: drop that pass's
    // list and minimum (no-ops behind the grid-wide kernels of the graph, which leave these records clean)
    StepCtrl *nxt = &g.ctrl[(phase + 1) % 3];
    ctl->n_act = 0;
    ctl->work_counter = 0;
    ctl->pad[0] = 0;
    ctl->pad[1] = 0;
    nxt->t_next_bits = INF_BITS;
  }

  ChipMail *mail = reinterpret_cast<ChipMail *>(g.chip_mail);
  unsigned *flag = reinterpret_cast<unsigned *>(mail + nc);
  ChipMail *mine = mail + me;
  unsigned seq = g.hdr->chip_seq + 1u;  // step numbers go on from launch to launch (the last one is left in the header)

  const int per = (g.n_tot + nc - 1) / nc;
  const int j0 = min(me * per, g.n_tot), cnt = min(per, g.n_tot - j0);
  for (int k = tid; k < cnt; k += CHIP_T) {
    const int i = j0 + k;
    pos[k] = g.pos[i]; vel[k] = g.vel[i]; acc[k] = g.acc[i]; jrk[k] = g.jrk[i];
    tt[k] = g.t[i]; dtt[k] = g.dt[i];
  }
  __syncthreads();

  // this chunk's min(t + dt), the particles that attain it predicted to that time, then the header
  auto publish = [&](const unsigned s) {
    unsigned long long v = INF_BITS;
    for (int k = tid; k < cnt; k += CHIP_T) {
      const unsigned long long cb = dbits(tt[k] + dtt[k]);
      v = cb < v ? cb : v;
    }
    v = warp_min_u64(v);
    if (lane == 0) S.wmin[warp] = v;
    if (tid == 0) S.cand = 0;
    __syncthreads();
    unsigned long long m = S.wmin[0];
#pragma unroll
    for (int w = 1; w < CHIP_WARPS; w++) m = S.wmin[w] < m ? S.wmin[w] : m;
    const double tm = bitsd(m);
    for (int k = tid; k < cnt; k += CHIP_T) {
      if (dbits(tt[k] + dtt[k]) == m) {
        const int slot = atomicAdd(&S.cand, 1);  // the order of the slots is irrelevant to the results
        if (slot < CHIP_CAP) {
          double4 pp, pv;
          predict_to(tm, tt[k], pos[k], vel[k], acc[k], jrk[k], pp, pv);
          mine->idx[slot] = j0 + k;
          mine->pp[slot] = pp;
          mine->pv[slot] = pv;
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      fence_gpu();  // release: the records (made visible to this thread by the barrier) before the header
      mail_put_header(mine, s, m, S.cand);
    }
  };

  publish(seq);
  // the first step needs everybody's header
  bool alive = chip_poll(g.hdr, [&]() {
    if (tid >= nc) return true;
    unsigned long long mb;
    int c;
    if (!mail_get_header(mail + tid, seq, mb, c)) return false;
    S.cmin[tid] = mb;
    S.ccount[tid] = c;
    return true;
  });

  long long prof[6] = {0, 0, 0, 0, 0, 0}, tk = clock64();  // CTA 0 / thread 0: cycles per segment (al26_grav_loop_profile)
#define PROF(k)                      \
  if (me == 0 && tid == 0) {         \
    const long long now = clock64(); \
    prof[k] += now - tk;             \
    tk = now;                        \
  }
  long long n_steps = 0, n_pairs = 0;
  int hist[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const double eps2 = g.eps2;
  const int ent = (nc + 31) / 32;
  unsigned long long tnb = INF_BITS;

  while (alive) {
    // ---- 1: block time, owners, slot numbering (warp 0, from the copies of the headers) -----------
    __syncthreads();  // cmin / ccount complete (the polling threads wrote them)
    if (warp == 0) {
      unsigned long long mymin = INF_BITS;
      for (int q = 0; q < ent; q++) {
        const int e = lane * ent + q;
        if (e < nc) mymin = S.cmin[e] < mymin ? S.cmin[e] : mymin;
      }
      const unsigned long long tn_b = warp_min_u64(mymin);
      int c_l = 0, o_l = 0;
      for (int q = 0; q < ent; q++) {
        const int e = lane * ent + q;
        if (e < nc && S.cmin[e] == tn_b) {
          c_l += S.ccount[e];
          o_l += 1;
        }
      }
      int c_inc = c_l, o_inc = o_l;  // inclusive scans over the lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, c_inc, o), b = __shfl_up_sync(0xffffffffu, o_inc, o);
        if (lane >= o) {
          c_inc += a;
          o_inc += b;
        }
      }
      int base = c_inc - c_l, opos = o_inc - o_l;
      if (lane == 0) {
        S.my_cnt = 0;
        S.my_base = 0;
      }
      __syncwarp();
      for (int q = 0; q < ent; q++) {
        const int e = lane * ent + q;
        if (e < nc && S.cmin[e] == tn_b) {
          S.own_cta[opos] = e;
          S.own_base[opos] = base;
          S.own_cnt[opos] = S.ccount[e];
          if (e == me) {
            S.my_base = base;
            S.my_cnt = S.ccount[e];
          }
          base += S.ccount[e];
          opos++;
        }
      }
      if (lane == 31) {
        S.n_act = c_inc;
        S.n_own = o_inc;
        S.tn_bits = tn_b;
      }
    }
    __syncthreads();
    tnb = S.tn_bits;
    const int n_act = S.n_act, n_own = S.n_own;
    const double tn = bitsd(tnb);
    if (tn > span || n_act > max_act) break;  // uniform: a block for the whole chip, or the call is over
    // ---- 2: the owners' records (one L2 round trip), meanwhile the predictor on the own chunk --------------------
    if (tid < n_act) {
      int k = 0;
      while (k + 1 < n_own && S.own_base[k + 1] <= tid) k++;
      const ChipMail *m = mail + S.own_cta[k];
      const int r = tid - S.own_base[k];
      S.a_idx[tid] = __ldcg(&m->idx[r]);
      S.a_pp[tid] = ldcg_d4(&m->pp[r]);
      S.a_pv[tid] = ldcg_d4(&m->pv[r]);
    }
    for (int k = tid; k < cnt; k += CHIP_T) {
      double4 pp, pv;
      predict_to(tn, tt[k], pos[k], vel[k], acc[k], jrk[k], pp, pv);
      ppos[k] = pp;
      pvel[k] = pv;
    }
    __syncthreads();
    PROF(0)
    // ---- 3: force of the active set against this chunk, one partial per slot ---------------------------------------
    for (int t0 = 0; t0 < n_act; t0 += 32) {
      const int nt = min(32, n_act - t0);
      int iw = 1;
      while (iw < nt) iw <<= 1;
      const int isub = lane & (iw - 1), jsub = lane / iw, jgroups = 32 / iw;
      const int li = t0 + (isub < nt ? isub : 0);
      const double4 p = S.a_pp[li], v = S.a_pv[li];
      Acc7 s;
      s.ax = s.ay = s.az = s.jx = s.jy = s.jz = s.pot = 0.0;
#pragma unroll 4
      for (int jj = warp * jgroups + jsub; jj < cnt; jj += CHIP_WARPS * jgroups)
        pair_interaction(ppos[jj], pvel[jj], eps2, p.x, p.y, p.z, v.x, v.y, v.z, s);
      for (int o = iw; o < 32; o <<= 1) {  // fixed butterfly over the lane bits above log2(iw)
        s.ax += __shfl_xor_sync(0xffffffffu, s.ax, o); s.ay += __shfl_xor_sync(0xffffffffu, s.ay, o);
        s.az += __shfl_xor_sync(0xffffffffu, s.az, o); s.jx += __shfl_xor_sync(0xffffffffu, s.jx, o);
        s.jy += __shfl_xor_sync(0xffffffffu, s.jy, o); s.jz += __shfl_xor_sync(0xffffffffu, s.jz, o);
        s.pot += __shfl_xor_sync(0xffffffffu, s.pot, o);
      }
      if (lane < iw) {
        S.red[warp][0][lane] = s.ax; S.red[warp][1][lane] = s.ay; S.red[warp][2][lane] = s.az;
        S.red[warp][3][lane] = s.jx; S.red[warp][4][lane] = s.jy; S.red[warp][5][lane] = s.jz;
        S.red[warp][6][lane] = s.pot;
      }
      __syncthreads();
      if (tid < 16 * nt) {  // 16 lanes per slot: fixed xor-butterfly over the warps' rows
        const int sl = tid >> 4, row = tid & 15;
        double r[7];
#pragma unroll
        for (int c = 0; c < 7; c++) r[c] = S.red[row][c][sl];
        const unsigned mask = __activemask();
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
          for (int c = 0; c < 7; c++) r[c] += __shfl_xor_sync(mask, r[c], o);
        }
        if (row == 0) {
          const long long o = (long long)(t0 + sl) * nc + me;
          g.part_a[o] = make_double4(r[0], r[1], r[2], r[6]);
          g.part_j[o] = make_double4(r[3], r[4], r[5], 0.0);
        }
      }
      __syncthreads();  // red[] is free for the next tile; all partial stores precede the flag
    }
    if (tid == 0) {
      fence_gpu();  // release: the partials before the flag
      stv_u32(&flag[me], seq);
    }
    PROF(1)
    // ---- 4: an owner sums its slots' partials, corrects them on the resident state and publishes its new header ----
    const int my_cnt = S.my_cnt, my_base = S.my_base;
    if (my_cnt > 0) {
      alive = chip_poll(g.hdr, [&]() { return tid >= nc || ldv_u32(&flag[tid]) == seq; });
      if (!alive) break;
      for (int q = warp; q < my_cnt; q += CHIP_WARPS) {
        const int slot = my_base + q;
        double r[7] = {0, 0, 0, 0, 0, 0, 0};
        const long long row = (long long)slot * nc;
        for (int c = lane; c < nc; c += 32) {
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
          const int i = S.a_idx[slot], k = i - j0;
          SlotIn in;
          in.i = i;
          in.a0 = acc[k]; in.j0 = jrk[k];
          in.xp = S.a_pp[slot]; in.vp = S.a_pv[slot];
          in.t = tt[k]; in.dt = dtt[k];
          unsigned long long c_bits = INF_BITS;
          NewState ns;
          correct_slot<MODE_STEP, false>(g, tn, in, r, c_bits, 0ull, Dmax, &ns);  // also writes the global records
          pos[k] = ns.pos; vel[k] = ns.vel; acc[k] = ns.acc; jrk[k] = ns.jrk;
          tt[k] = ns.t; dtt[k] = ns.dt;
        }
      }
      __syncthreads();
      publish(seq + 1u);
    }
    PROF(2)
    // ---- 5: everybody waits for the owners' headers (the other chunks' minima cannot have changed) ---------------
    alive = chip_poll(g.hdr, [&]() {
      if (tid >= n_own) return true;
      const int c = S.own_cta[tid];
      unsigned long long mb;
      int cc;
      if (!mail_get_header(mail + c, seq + 1u, mb, cc)) return false;
      S.cmin[c] = mb;
      S.ccount[c] = cc;
      return true;
    });
    PROF(3)
    seq += 1u;
    n_steps += 1;
    if (me == 0 && tid == 0) {
      int own = n_act;
      if (g.p2p) {  // replicated state: every rank steps all active particles, and accounts for the pairs of those it owns
        own = 0;
        for (int q = 0; q < n_act; q++) own += (S.a_idx[q] % g.world) == g.rank;
      }
      n_pairs += (long long)own * (long long)g.n_tot;
      int b = 0;
      while ((1 << (b + 1)) <= n_act && b < 8) b++;
      hist[b] += 1;
    }
  }
#undef PROF
  if (me == 0 && tid == 0) {
    if (alive) ctl->t_next_bits = tnb;  // the block time of the step that is still to be taken (or beyond the span)
    g.hdr->chip_seq = seq;
    g.hdr->bar_counter = 0u;  // chained: the loop kernel that follows counts its grid barriers from zero
    g.hdr->n_steps += n_steps;
    g.hdr->n_pairs += n_pairs;
    g.hdr->n_chip += n_steps;
    for (int k = 0; k < 6; k++) g.hdr->loop_cycles[k] += prof[k];
    for (int b = 0; b < 9; b++) g.hdr->nact_hist[b] += hist[b];
  }
}

// ---- host side ----
static bool g_chip_attr_set = false;

cudaError_t chip_kernel_setup(int max_smem_optin) {
  const cudaError_t e = cudaFuncSetAttribute(k_chip, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
  g_chip_attr_set = (e == cudaSuccess);
  return e;
}

// can n_ctas CTAs of the chip engine with p_cap particles each be co-resident (one per SM)?
bool chip_fits(int n_ctas, int p_cap, int sm_count, int max_smem_optin) {
  if (!g_chip_attr_set || chip_smem_bytes(p_cap) > max_smem_optin) return false;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chip, CHIP_T, (size_t)chip_smem_bytes(p_cap)) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return per_sm >= 1 && n_ctas <= per_sm * sm_count;
}

int launch_chip(const GravDev &g, int phase, cudaStream_t s, cudaError_t *err) {
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(g.chip_n);
  cfg.blockDim = dim3(CHIP_T);
  cfg.dynamicSmemBytes = (size_t)chip_smem_bytes(g.chip_p);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // co-residency of all CTAs: they wait for each other's flags
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, k_chip, g, phase, g.chip_p);
  if (err) *err = e;
  return 1;
}

}  // namespace al26
