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
constexpr int CHIP_FUSE_PREDICT = 4;  // blocks of at most this many particles predict j inside the pair loop

struct alignas(32) ChipShared {
  union {
    double red[CHIP_WARPS][7][32];    // force phase: the warps' partial sums
    double rowval[7][CHIP_MAX_CTAS];  // an owner, afterwards: the partial sums of one slot, by component and source CTA
  };
  double rsum[CHIP_WARPS][8];              // ... and their totals for a batch of slots
  double4 c_pp[CHIP_MAX_CTAS], c_pv[CHIP_MAX_CTAS];  // every chunk's first record, fetched together with its header
  int c_idx[CHIP_MAX_CTAS];
  unsigned c_seq[CHIP_MAX_CTAS];  // the step number of the cached first record (valid when equal to cseq)
  unsigned c_dte[CHIP_MAX_CTAS];  // ... and the exponent bits of that particle's timestep (a power of two)
  unsigned a_dte[CHIP_CAP];       // the same for the step's active set
  unsigned long long pred_bits;   // the block time ppos / pvel were predicted to ahead of time (0 = not valid)
  double4 a_pp[CHIP_CAP], a_pv[CHIP_CAP];  // the step's active set (predicted), gathered from the owners' mail
  int a_idx[CHIP_CAP];
  unsigned long long cmin[CHIP_MAX_CTAS];  // every chunk's min(t + dt) ...
  int ccount[CHIP_MAX_CTAS];               // ... how many of its particles attain it ...
  unsigned cseq[CHIP_MAX_CTAS];            // ... and the step number its header and records carry
  int own_cta[CHIP_MAX_CTAS], own_base[CHIP_MAX_CTAS];  // this step's owners, in CTA order, and their first slots
  unsigned long long wmin[CHIP_WARPS];
  unsigned long long tn_bits;
  int n_act, n_own, my_base, my_cnt, cand, acct_own, pad[2];
  int hist[12];  // CTA 0: block steps by floor(log2(n_act))
};
static_assert(sizeof(ChipShared) % 32 == 0, "ChipShared must keep the double4 arrays behind it aligned");

constexpr int CHIP_BYTES_PER_PARTICLE = 6 * 32 + 2 * 8;  // pos, vel, acc, jrk, ppos, pvel, t, dt

int chip_smem_bytes(int p_cap) { return (int)sizeof(ChipShared) + p_cap * CHIP_BYTES_PER_PARTICLE; }
size_t chip_mail_bytes(int n_ctas) { return (size_t)n_ctas * sizeof(ChipMail) + (size_t)CHIP_CAP * n_ctas * 16 * sizeof(unsigned long long); }

// n particles over n_ctas CTAs: per-CTA capacity, or false when the chunks do not fit `max_smem` bytes per block
bool chip_plan(int n, int n_ctas, int max_smem, int *p_cap_out) {
  if (n < 1 || n_ctas < 1 || n_ctas > CHIP_MAX_CTAS) return false;
  const int p_cap = (((n + n_ctas - 1) / n_ctas) + 7) & ~7;
  if (chip_smem_bytes(p_cap) > max_smem) return false;
  *p_cap_out = p_cap;
  return true;
}

namespace {

// Everything that crosses CTAs travels as 64-bit words {step number (high half) | 32 bits of payload}, written and read
// with volatile (L2) accesses: a reader polls the very words it needs until they carry the step number it expects, so a
// message needs neither a flag, nor a fence, nor a second round trip, and no word is ever paired with a stale
// neighbour.  Re-use of a buffer is ordered by data flow: a writer overwrites a word only after it has consumed values
// that its readers computed from the previous contents.
__device__ __forceinline__ void stv_2u64(unsigned long long *p, unsigned long long a, unsigned long long b) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ldv_2u64(const unsigned long long *p, unsigned long long &a, unsigned long long &b) {
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
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
constexpr unsigned long long TAG_MASK = 0xffffffff00000000ull;
__device__ __forceinline__ unsigned long long tag_of(unsigned seq) { return (unsigned long long)seq << 32; }
// one double as two tagged words {hi, lo} in one 16-byte store
__device__ __forceinline__ void put_double(unsigned long long *w, unsigned long long tag, double x) {
  stv_2u64(w, tag | (unsigned)__double2hiint(x), tag | (unsigned)__double2loint(x));
}
__device__ __forceinline__ bool get_double(const unsigned long long *w, unsigned long long tag, double &x) {
  unsigned long long a, b;
  ldv_2u64(w, a, b);
  x = __hiloint2double((int)(unsigned)a, (int)(unsigned)b);
  return (a & TAG_MASK) == tag && (b & TAG_MASK) == tag;
}

// Block-wide poll: every thread evaluates `ready()` (threads with nothing to wait for return true) until all do.
// Bounded: a protocol error raises hdr->loop_error (code 5) instead of hanging the GPU; returns false when the launch
// must be abandoned (uniform over the block).
template <class F>
__device__ __forceinline__ bool chip_poll(GravHeader *hdr, F ready) {
  unsigned spins = 0;
  while (true) {
    const bool ok = ready();
    if (__syncthreads_and(ok)) return true;
    if ((++spins & 1023u) == 0u) {
      int bad = 0;
      if (threadIdx.x == 0) {
        if (spins > CHIP_SPIN_LIMIT) atomicExch(&hdr->loop_error, 5);
        bad = ldv_u32((const unsigned *)&hdr->loop_error) != 0u;
      }
      if (__syncthreads_or(bad)) return false;
    }
  }
}

// header of one chunk: {step | hi(min t+dt)}, {step | lo}, {step | count}
__device__ __forceinline__ void mail_put_header(ChipMail *m, unsigned seq, unsigned long long min_bits, int count) {
  const unsigned long long tag = tag_of(seq);
  stv_2u64(&m->w[0], tag | (min_bits >> 32), tag | (min_bits & 0xffffffffull));
  stv_u64(&m->w[2], tag | (unsigned long long)(unsigned)count);
}
__device__ __forceinline__ bool mail_get_header(const ChipMail *m, unsigned seq, unsigned long long &min_bits, int &count) {
  const unsigned long long tag = tag_of(seq);
  unsigned long long w0, w1;
  ldv_2u64(&m->w[0], w0, w1);
  const unsigned long long w2 = ldv_u64(&m->w[2]);
  if ((w0 & TAG_MASK) != tag || (w1 & TAG_MASK) != tag || (w2 & TAG_MASK) != tag) return false;
  min_bits = (w0 << 32) | (w1 & 0xffffffffull);
  count = (int)(unsigned)(w2 & 0xffffffffull);
  return true;
}

// warp minimum of a 64-bit key with two 32-bit hardware reductions
__device__ __forceinline__ unsigned long long warp_min_u64_redux(unsigned long long v) {
  const unsigned hi = __reduce_min_sync(0xffffffffu, (unsigned)(v >> 32));
  const unsigned lo = __reduce_min_sync(0xffffffffu, (unsigned)(v >> 32) == hi ? (unsigned)v : 0xffffffffu);
  return ((unsigned long long)hi << 32) | lo;
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
  const long long t_entry = clock64();

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
  unsigned long long *rows = reinterpret_cast<unsigned long long *>(mail + nc);  // [slot][CTA][16]
  ChipMail *mine = mail + me;
  unsigned seq = g.hdr->chip_seq + 1u;  // step numbers go on from launch to launch (the last one is left in the header)

  const int per = (g.n_tot + nc - 1) / nc;
  const int j0 = min(me * per, g.n_tot), cnt = min(per, g.n_tot - j0);
  // peer-memory mode: the loop kernel's last block step was an exchanged one and nothing has passed over the state
  // since -- the records the peers staged for it are pulled into the local state here, as that kernel's next scheduler
  // pass would have done
  const bool pull = g.p2p && g.hdr->dist_prev_exch != 0;
  const unsigned pull_tag = (unsigned)g.hdr->dist_step;
  StagingView sv;
  if (pull) sv = staging_view(g.slab[g.rank], g.n_tot, (int)(g.hdr->dist_step & 1ull));
  for (int k = tid; k < cnt; k += CHIP_T) {
    const int i = j0 + k;
    if (pull && sv.tag[i] == pull_tag) {
      const double4 p = sv.pos[i], v = sv.vel[i], a = sv.acc[i], j = sv.jrk[i];
      const double ti = sv.t[i], dti = sv.dt[i];
      g.pos[i] = p; g.vel[i] = v; g.acc[i] = a; g.jrk[i] = j;
      g.t[i] = ti; g.dt[i] = dti;
      pos[k] = p; vel[k] = v; acc[k] = a; jrk[k] = j;
      tt[k] = ti; dtt[k] = dti;
    } else {
      pos[k] = g.pos[i]; vel[k] = g.vel[i]; acc[k] = g.acc[i]; jrk[k] = g.jrk[i];
      tt[k] = g.t[i]; dtt[k] = g.dt[i];
    }
  }
  __syncthreads();

  // this chunk's min(t + dt), the particles that attain it predicted to that time (tagged records), then the header
  auto publish = [&](const unsigned s) {
    unsigned long long v = INF_BITS;
    for (int k = tid; k < cnt; k += CHIP_T) {
      const unsigned long long cb = dbits(tt[k] + dtt[k]);
      v = cb < v ? cb : v;
    }
    v = warp_min_u64_redux(v);
    if (lane == 0) S.wmin[warp] = v;
    if (tid == 0) S.cand = 0;
    __syncthreads();
    const unsigned long long m = warp_min_u64_redux(S.wmin[lane & (CHIP_WARPS - 1)]);
    const double tm = bitsd(m);
    const unsigned long long tag = tag_of(s);
    for (int k = tid; k < cnt; k += CHIP_T) {
      if (dbits(tt[k] + dtt[k]) == m) {
        const int slot = atomicAdd(&S.cand, 1);  // the order of the slots is irrelevant to the results
        if (slot < CHIP_CAP) {
          double4 pp, pv;
          predict_to(tm, tt[k], pos[k], vel[k], acc[k], jrk[k], pp, pv);
          unsigned long long *r = mine->rec[slot];
          put_double(r + 0, tag, pp.x); put_double(r + 2, tag, pp.y); put_double(r + 4, tag, pp.z); put_double(r + 6, tag, pp.w);
          put_double(r + 8, tag, pv.x); put_double(r + 10, tag, pv.y); put_double(r + 12, tag, pv.z);
          stv_2u64(r + 14, tag | (unsigned)(j0 + k), tag | (unsigned)(dbits(dtt[k]) >> 52));
        }
      }
    }
    __syncthreads();
    if (tid == 0) mail_put_header(mine, s, m, S.cand);
  };

  publish(seq);
  // the first step needs everybody's header
  // one word of chunk c's first record into the cache; false while it does not carry step number s yet
  auto fetch_rec0_word = [&](const int c, const int w, const unsigned s) {
    const unsigned long long x = ldv_u64(&mail[c].rec[0][w]);
    if ((x & TAG_MASK) != tag_of(s)) return false;
    const unsigned half = (unsigned)x;
    if (w < 8) reinterpret_cast<unsigned *>(&S.c_pp[c])[(w & ~1) | ((w & 1) ^ 1)] = half;  // words are {hi, lo}
    else if (w < 14) reinterpret_cast<unsigned *>(&S.c_pv[c])[((w - 8) & ~1) | ((w & 1) ^ 1)] = half;
    else if (w == 14) S.c_idx[c] = (int)half;
    else {
      S.c_pv[c].w = 0.0;
      S.c_dte[c] = half;
      S.c_seq[c] = s;  // (the poll this runs in only ends when all 16 words were there in the same round)
    }
    return true;
  };
  bool alive = chip_poll(g.hdr, [&]() {
    if (tid >= nc) return true;
    unsigned long long mb;
    int c;
    if (!mail_get_header(mail + tid, seq, mb, c)) return false;
    S.cmin[tid] = mb;
    S.ccount[tid] = c;
    S.cseq[tid] = seq;
    S.c_seq[tid] = 0u;  // no record cached yet (step numbers start at 1)
    return true;
  });

  long long prof[6] = {0, 0, 0, 0, 0, 0}, tk = clock64();  // CTA 0 / thread 0: cycles per segment (al26_grav_loop_profile)
  prof[5] = tk - t_entry;                                  // [5]: the launch's prologue (state load, first publication, gather)
#define PROF(k)                      \
  if (me == 0 && tid == 0) {         \
    const long long now = clock64(); \
    prof[k] += now - tk;             \
    tk = now;                        \
  }
  // accounting (CTA 0 only) stays off the step's critical path: it lives in warp 2, which has nothing else to do
  // between the phases, not in the warp that schedules
  constexpr int ACCT_TID = 64;
  long long n_steps = 0, n_pairs = 0;
  if (tid < 12) S.hist[tid] = 0;
  if (tid == 0) {
    S.acct_own = 0;
    S.pred_bits = 0ull;
  }
  const double eps2 = g.eps2;
  const int ent = (nc + 31) / 32;
  unsigned long long tnb = INF_BITS;

  while (alive) {
    // ---- 1: block time, owners, slot numbering (warp 0, from the copies of the headers; the poll that filled them
    //         ended with a block barrier) ---------------------------------------------------------------------------
    if (warp == 0) {
      unsigned long long cm[CHIP_ENT];  // this lane's entries, all loads in flight at once
      int cc[CHIP_ENT];
      unsigned long long mymin = INF_BITS;
#pragma unroll
      for (int q = 0; q < CHIP_ENT; q++) {
        const int e = lane * ent + q;
        const bool in = q < ent && e < nc;
        cm[q] = in ? S.cmin[e] : INF_BITS;
        cc[q] = in ? S.ccount[e] : 0;
      }
#pragma unroll
      for (int q = 0; q < CHIP_ENT; q++) mymin = cm[q] < mymin ? cm[q] : mymin;
      const unsigned long long tn_b = warp_min_u64_redux(mymin);
      int c_l = 0, o_l = 0;
#pragma unroll
      for (int q = 0; q < CHIP_ENT; q++) {
        if (cm[q] == tn_b && tn_b != INF_BITS) {
          c_l += cc[q];
          o_l += 1;
        }
      }
      // inclusive scans over the lanes, both in one word (owners < 2^9, active particles < 2^22)
      int packed = c_l * 512 + o_l;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, packed, o);
        if (lane >= o) packed += a;
      }
      const int c_inc = packed >> 9, o_inc = packed & 511;
      int base = c_inc - c_l, opos = o_inc - o_l;
      if (lane == 0) {
        S.my_cnt = 0;
        S.my_base = 0;
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < CHIP_ENT; q++) {
        if (cm[q] == tn_b && tn_b != INF_BITS) {
          const int e = lane * ent + q;
          S.own_cta[opos] = e;
          S.own_base[opos] = base;
          if (e == me) {
            S.my_base = base;
            S.my_cnt = cc[q];
          }
          base += cc[q];
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
    PROF(0)
    // ---- 2: the predictor on the own chunk, then the owners' records: 16 threads per active particle poll its 16
    //         tagged words (published before the owner's header, so normally there at the first look) ---------------
    // tiny blocks (<= CHIP_FUSE_PREDICT active particles, more than half of all block steps): every j is visited by at most
    // that many lanes, so it is predicted on the fly in the pair loop -- the same arithmetic on the same operands, hence
    // the same bits (the self pair still cancels exactly) -- instead of in a pass of its own through shared memory
    const bool predicted = (S.pred_bits == tnb);  // (written two barriers ago)
    const bool fused_predict = !predicted && n_act <= CHIP_FUSE_PREDICT;
    if (!predicted && !fused_predict) {
      for (int k = tid; k < cnt; k += CHIP_T) {
        double4 pp, pv;
        predict_to(tn, tt[k], pos[k], vel[k], acc[k], jrk[k], pp, pv);
        ppos[k] = pp;
        pvel[k] = pv;
      }
    }
    for (int s0 = 0; s0 < n_act && alive; s0 += CHIP_T / 16) {
      const int slot = s0 + (tid >> 4), w = tid & 15;
      const unsigned long long *src = nullptr;
      unsigned long long want = 0;
      if (slot < n_act) {
        int k = 0;
        while (k + 1 < n_own && S.own_base[k + 1] <= slot) k++;
        const int c = S.own_cta[k];
        if (slot == S.own_base[k] && S.c_seq[c] == S.cseq[c]) {  // the chunk's first record came with its header
          if (w == 0) {
            S.a_pp[slot] = S.c_pp[c];
            S.a_pv[slot] = S.c_pv[c];
            S.a_idx[slot] = S.c_idx[c];
            S.a_dte[slot] = S.c_dte[c];
          }
        } else {
          src = &mail[c].rec[slot - S.own_base[k]][w];
          want = tag_of(S.cseq[c]);
        }
      }
      if (!__syncthreads_or(src != nullptr)) continue;  // nothing to fetch (the barrier also covers the predictor's stores)
      alive = chip_poll(g.hdr, [&]() {
        if (!src) return true;
        const unsigned long long x = ldv_u64(src);
        if ((x & TAG_MASK) != want) return false;
        const unsigned half = (unsigned)x;
        if (w < 8) reinterpret_cast<unsigned *>(&S.a_pp[slot])[(w & ~1) | ((w & 1) ^ 1)] = half;  // words are {hi, lo}
        else if (w < 14) reinterpret_cast<unsigned *>(&S.a_pv[slot])[((w - 8) & ~1) | ((w & 1) ^ 1)] = half;
        else if (w == 14) S.a_idx[slot] = (int)half;
        else {
          S.a_pv[slot].w = 0.0;
          S.a_dte[slot] = half;
        }
        return true;
      });
    }
    if (!alive) break;  // (the poll's barrier also covers the predictor's stores)
    PROF(1)
    // ---- 3: force of the active set against this chunk; one tagged row of partial sums per slot --------------------
    const unsigned long long tag = tag_of(seq);
    for (int t0 = 0; t0 < n_act; t0 += 32) {
      const int nt = min(32, n_act - t0);
      int iw = 1;
      while (iw < nt) iw <<= 1;
      const int isub = lane & (iw - 1), jsub = lane / iw, jgroups = 32 / iw;
      const int li = t0 + (isub < nt ? isub : 0);
      const double4 p = S.a_pp[li], v = S.a_pv[li];
      Acc7 s;
      s.ax = s.ay = s.az = s.jx = s.jy = s.jz = s.pot = 0.0;
      if (fused_predict) {
#pragma unroll 2
        for (int jj = warp * jgroups + jsub; jj < cnt; jj += CHIP_WARPS * jgroups) {
          double4 pp, pv;
          predict_to(tn, tt[jj], pos[jj], vel[jj], acc[jj], jrk[jj], pp, pv);
          pair_interaction(pp, pv, eps2, p.x, p.y, p.z, v.x, v.y, v.z, s);
        }
      } else {
#pragma unroll 4
        for (int jj = warp * jgroups + jsub; jj < cnt; jj += CHIP_WARPS * jgroups)
          pair_interaction(ppos[jj], pvel[jj], eps2, p.x, p.y, p.z, v.x, v.y, v.z, s);
      }
      for (int o = iw; o < 32; o <<= 1) {  // fixed butterfly over the lane bits above log2(iw)
        s.ax += __shfl_xor_sync(0xffffffffu, s.ax, o); s.ay += __shfl_xor_sync(0xffffffffu, s.ay, o);
        s.az += __shfl_xor_sync(0xffffffffu, s.az, o); s.jx += __shfl_xor_sync(0xffffffffu, s.jx, o);
        s.jy += __shfl_xor_sync(0xffffffffu, s.jy, o); s.jz += __shfl_xor_sync(0xffffffffu, s.jz, o);
        s.pot += __shfl_xor_sync(0xffffffffu, s.pot, o);
      }
      if (t0 > 0) __syncthreads();  // red[] of the previous tile has been read
      if (lane < iw) {
        S.red[warp][0][lane] = s.ax; S.red[warp][1][lane] = s.ay; S.red[warp][2][lane] = s.az;
        S.red[warp][3][lane] = s.jx; S.red[warp][4][lane] = s.jy; S.red[warp][5][lane] = s.jz;
        S.red[warp][6][lane] = s.pot;
      }
      __syncthreads();
      if (tid < 16 * nt) {  // 16 lanes per slot: fixed xor-butterfly over the warps' rows, then lane w stores word w
        const int sl = tid >> 4, w = tid & 15;
        double r[7];
#pragma unroll
        for (int c = 0; c < 7; c++) r[c] = S.red[w][c][sl];
        const unsigned mask = __activemask();
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
          for (int c = 0; c < 7; c++) r[c] += __shfl_xor_sync(mask, r[c], o);
        }
        double val = r[0];
#pragma unroll
        for (int c = 1; c < 7; c++) val = ((w >> 1) == c) ? r[c] : val;
        const unsigned half = (w & 1) ? (unsigned)__double2loint(val) : (unsigned)__double2hiint(val);
        if (w < 14) stv_u64(&rows[((size_t)(t0 + sl) * nc + me) * 16 + w], tag | half);
      }
    }
    if (me == 0 && g.p2p && tid < CHIP_CAP) {
      // replicated state: every rank steps all active particles, and accounts for the pairs of those it owns
      const unsigned b = __ballot_sync(0xffffffffu, tid < n_act && (S.a_idx[tid] % g.world) == g.rank);
      if (lane == 0 && b) atomicAdd(&S.acct_own, __popc(b));
    }
    PROF(2)
    // ---- 4: an owner collects the rows of its slots (polling the words themselves), sums them in CTA order, corrects
    //         the particles on the resident state and publishes its new header + records ----------------------------
    const int my_cnt = S.my_cnt, my_base = S.my_base;
    if (my_cnt > 0) {
      long long o_t0 = 0, o_t1 = 0, o_t2 = 0;  // the owner's own clock: rows complete, corrected, published
      if (tid == 0) o_t0 = clock64();
      for (int q0 = 0; q0 < my_cnt && alive; q0 += CHIP_WARPS) {
        const int nq = min(CHIP_WARPS, my_cnt - q0);
        for (int q = 0; q < nq && alive; q++) {
          if (q > 0) __syncthreads();  // the previous slot's sums have been taken out of rowval
          const unsigned long long *row = rows + (size_t)(my_base + q0 + q) * nc * 16;
          // item = (CTA c, component): thread t takes items t, t + 512, ...; every item is one 16-byte load
          alive = chip_poll(g.hdr, [&]() {
            bool ok = true;
            for (int it = tid; it < nc * 7; it += CHIP_T) {
              const int c = it / 7, comp = it - c * 7;
              double x;
              if (get_double(row + (size_t)c * 16 + 2 * comp, tag, x)) S.rowval[comp][c] = x;
              else ok = false;
            }
            return ok;
          });
          if (!alive) break;
          if (tid == 0 && q0 == 0 && q == 0) o_t1 = clock64();
          if (warp < 7) {  // component `warp`: lanes sum every 32nd CTA in order, then a fixed butterfly
            double a = 0.0;
            for (int c = lane; c < nc; c += 32) a += S.rowval[warp][c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) S.rsum[q][warp] = a;
          }
        }
        if (!alive) break;
        __syncthreads();
        if (lane == 0 && warp < nq) {  // one corrector per warp
          const int slot = my_base + q0 + warp;
          const int i = S.a_idx[slot], k = i - j0;
          double r[7];
#pragma unroll
          for (int c = 0; c < 7; c++) r[c] = S.rsum[warp][c];
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
        __syncthreads();
      }
      if (!alive) break;
      if (tid == 0) o_t2 = clock64();
      publish(seq + 1u);
      if (tid == 0) {  // diagnostic (al26_grav_fuse_profile): owner-side cycles per segment, over all owners
        const long long o_t3 = clock64();
        atomicAdd((unsigned long long *)&g.hdr->fuse_ns[0], (unsigned long long)(o_t1 - o_t0));
        atomicAdd((unsigned long long *)&g.hdr->fuse_ns[1], (unsigned long long)(o_t2 - o_t1));
        atomicAdd((unsigned long long *)&g.hdr->fuse_ns[2], (unsigned long long)(o_t3 - o_t2));
        atomicAdd((unsigned long long *)&g.hdr->fuse_ns[3], 1ull);
      }
    }
    PROF(3)
    // ---- 4b: while the owners correct and publish, predict the chunk to the LIKELY next block time: the smallest of the
    //          other chunks' minima and of t + dt of this step's active particles with their timesteps unchanged (they came
    //          with the records).  Right unless a timestep of the front runners changes; the next step checks ----------
    if (warp == 0) {
      unsigned long long g_min = INF_BITS;
      for (int e = lane; e < nc; e += 32) {
        const unsigned long long v = S.cmin[e];
        if (v != tnb) g_min = v < g_min ? v : g_min;
      }
      for (int q = lane; q < n_act; q += 32) {
        const unsigned long long v = dbits(tn + bitsd((unsigned long long)S.a_dte[q] << 52));
        g_min = v < g_min ? v : g_min;
      }
      g_min = warp_min_u64_redux(g_min);
      if (lane == 0) S.wmin[0] = g_min;
    }
    __syncthreads();
    {
      const unsigned long long gb = S.wmin[0];
      const double tg = bitsd(gb);
      for (int k = tid; k < cnt; k += CHIP_T) {
        double4 pp, pv;
        predict_to(tg, tt[k], pos[k], vel[k], acc[k], jrk[k], pp, pv);
        ppos[k] = pp;
        pvel[k] = pv;
      }
      if (tid == 0) S.pred_bits = gb;
    }
    // ---- 5: everybody waits for the owners' headers (the other chunks' minima cannot have changed) ---------------
    alive = chip_poll(g.hdr, [&]() {
      bool ok = true;
      for (int k = warp; k < n_own; k += CHIP_WARPS) {  // lane 0: the header; lanes 16-31: the first record
        const int c = S.own_cta[k];
        if (lane == 0) {
          unsigned long long mb;
          int cc;
          if (mail_get_header(mail + c, seq + 1u, mb, cc)) {
            S.cmin[c] = mb;
            S.ccount[c] = cc;
            S.cseq[c] = seq + 1u;
          } else {
            ok = false;
          }
        } else if (lane >= 16) {
          if (!fetch_rec0_word(c, lane - 16, seq + 1u)) ok = false;
        }
      }
      return ok;
    });
    PROF(4)
    seq += 1u;
    if (me == 0 && tid == ACCT_TID) {  // (the ballots above are separated from here by the barriers of the polls)
      n_steps += 1;
      n_pairs += (long long)(g.p2p ? S.acct_own : n_act) * (long long)g.n_tot;
      S.acct_own = 0;
      S.hist[31 - __clz(n_act)] += 1;  // n_act <= CHIP_CAP = 2^8
    }
  }
#undef PROF
  if (me == 0 && tid == ACCT_TID) {
    g.hdr->n_steps += n_steps;
    g.hdr->n_pairs += n_pairs;
    g.hdr->n_chip += n_steps;
    for (int b = 0; b < 9; b++) g.hdr->nact_hist[b] += S.hist[b];
  }
  if (me == 0 && tid == 0) {
    if (alive) ctl->t_next_bits = tnb;  // the block time of the step that is still to be taken (or beyond the span)
    g.hdr->chip_seq = seq;
    g.hdr->bar_counter = 0u;  // chained: the loop kernel that follows counts its grid barriers from zero
    if (pull) g.hdr->dist_prev_exch = 0;  // every CTA has read the flag: none leaves before all have published
    g.hdr->fuse_ns[4] += 1;  // diagnostic: launches that got past the span check
    for (int k = 0; k < 6; k++) g.hdr->loop_cycles[k] += prof[k];
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
  attr[0].id = cudaLaunchAttributeCooperative;  // co-residency of all CTAs: they wait for each other's words
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, k_chip, g, phase, g.chip_p);
  if (err) *err = e;
  return 1;
}

}  // namespace al26
