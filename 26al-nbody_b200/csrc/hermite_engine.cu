// hermite_engine.cu -- the cluster engine: runs of SMALL block steps for small N, entirely on chip.
//
// Why: a block step of a few particles is a chain of a dozen dependent hops (scheduler pass -> list -> force
// partials -> reduction -> corrector -> next block time).  Through global memory every hop costs ~0.4 us on B200
// (L2 round trip, atomic, fence; measured), so such a step costs 11-17 us however the kernels or
// grid barriers are arranged -- and for N <= 1e4 (the reference's own configurations) nearly every step is one.
// Inside ONE thread-block cluster the same hops are distributed-shared-memory accesses (~0.1 us) and hardware
// cluster barriers (~0.2 us).
//
// One cluster of 8 or 16 CTAs (16 needs the non-portable cluster size), one CTA per SM, 256 threads.  CTA r owns
// the contiguous chunk r of the particles and keeps their full state in its shared memory for the lifetime of the
// launch; corrected particles are also written through to the global records, so whatever runs next (the grid-wide
// kernels of a big block step, getters) sees a consistent state.  Per block step:
//   1  every CTA predicts its chunk (shared memory); an active particle takes a slot (atomicAdd on CTA 0's counter
//      through DSMEM) and its owner replicates its predicted state into every CTA's copy of the active set
//      -- cluster barrier --
//   2  every CTA: the <= 32 active particles against its own chunk (lanes split over i and j, fixed-order
//      reductions), one partial per slot, stored into the slot owner's shared memory
//      -- cluster barrier --
//   3  the owner sums the partials in CTA order, runs the corrector (the same correct_slot as every other path),
//      updates its resident state + the global records; every CTA's min(t + dt) goes to every CTA
//      -- cluster barrier --
// and the loop goes on until the next block step has more than 32 active particles (a job for the whole chip: the
// engine returns with the block time in the step's scheduler record, and the grid-wide predict / force / correct
// kernels that follow it in the CUDA graph take that step) or the call's span is exhausted.
// Same arithmetic as the other paths except the order of the j sum (CTA chunks), i.e. identical integer work,
// positions to rounding.  Stands in, like the rest, for ph4's evolve loop behind gravity.evolve_model
// (al26_nbody.py:833).
#include <cooperative_groups.h>

#include "hermite_force.cuh"
#include "hermite_step.cuh"

namespace cg = cooperative_groups;

namespace al26 {

constexpr int ENG_T = 512;  // measured: 256 and 1024 threads are both slower (fewer j in flight / costlier block barriers + spills)
constexpr int ENG_WARPS = ENG_T / 32;
static_assert(ENG_WARPS == 16, "the cross-warp sum below is a 16-lane butterfly");

struct alignas(32) EngShared {
  double part[ENG_MAX_ACT][ENG_CS_MAX][7];  // force partials of the slots this CTA owns, written by every CTA
  double red[ENG_WARPS][7][32];
  double4 a_pp[ENG_MAX_ACT], a_pv[ENG_MAX_ACT];  // the step's active set (predicted), replicated in every CTA
  int a_idx[ENG_MAX_ACT], a_owner[ENG_MAX_ACT];
  unsigned long long mins[ENG_CS_MAX];  // every CTA's min(t + dt)
  unsigned long long wmin[ENG_WARPS];
  int count[2];  // CTA 0's copy is the cluster's slot counter (two parities)
  int pad[2];
};
static_assert(sizeof(EngShared) % 32 == 0, "EngShared must keep the double4 arrays behind it aligned");

constexpr int ENG_BYTES_PER_PARTICLE = 6 * 32 + 2 * 8;  // pos, vel, acc, jrk, ppos, pvel, t, dt

int engine_smem_bytes(int p_cap) { return (int)sizeof(EngShared) + p_cap * ENG_BYTES_PER_PARTICLE; }

// The smallest cluster (8, then 16 CTAs) whose CTAs can hold n / cluster particles each within `max_smem` bytes of
// shared memory per block.  Pure arithmetic (host; also behind the host-only diagnostic al26_dbg_engine_plan).
bool engine_plan(int n, int max_smem, int *cs_out, int *p_cap_out) {
  if (n < 1) return false;
  for (int cs = 8; cs <= ENG_CS_MAX; cs *= 2) {
    const int p_cap = (((n + cs - 1) / cs) + 7) & ~7;
    if (engine_smem_bytes(p_cap) <= max_smem) {
      *cs_out = cs;
      *p_cap_out = p_cap;
      return true;
    }
  }
  return false;
}

__global__ void __launch_bounds__(ENG_T, 1) k_engine(const GravDev g, const int phase, const int p_cap) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), cs = (int)cluster.num_blocks();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  EngShared &E = *reinterpret_cast<EngShared *>(smem_raw);
  double4 *pos = reinterpret_cast<double4 *>(smem_raw + sizeof(EngShared));
  double4 *vel = pos + p_cap, *acc = vel + p_cap, *jrk = acc + p_cap, *ppos = jrk + p_cap, *pvel = ppos + p_cap;
  double *tt = reinterpret_cast<double *>(pvel + p_cap), *dtt = tt + p_cap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  StepCtrl *ctl = &g.ctrl[phase];
  unsigned long long tnb = ctl->t_next_bits;
  const double span = g.hdr->span;
  const double Dmax = g.hdr->D;  // not g.Dmax: the graph captured g by value at commit
  if (bitsd(tnb) > span) return;  // uniform over the cluster; the predict kernel that follows raises `done`

  const int per = (g.n_tot + cs - 1) / cs;
  const int j0 = min(rank * per, g.n_tot), cnt = min(per, g.n_tot - j0);
  for (int k = tid; k < cnt; k += ENG_T) {
    const int i = j0 + k;
    pos[k] = g.pos[i]; vel[k] = g.vel[i]; acc[k] = g.acc[i]; jrk[k] = g.jrk[i];
    tt[k] = g.t[i]; dtt[k] = g.dt[i];
  }
  if (tid < 2) E.count[tid] = 0;
  __syncthreads();
  cluster.sync();  // every CTA is up and its counters are zero before anybody touches remote shared memory

  int par = 0;
  long long prof[6] = {0, 0, 0, 0, 0, 0}, tk = clock64();  // CTA 0 / thread 0: cycles per phase (al26_grav_loop_profile)
#define PROF(k)                      \
  if (rank == 0 && tid == 0) {       \
    const long long now = clock64(); \
    prof[k] += now - tk;             \
    tk = now;                        \
  }
  long long n_steps = 0, n_pairs = 0;
  int hist[6] = {0, 0, 0, 0, 0, 0};
  const double eps2 = g.eps2;
  while (true) {
    const double tn = bitsd(tnb);
    // ---- 1: predictor + scheduler ----------------------------------------------------------------
    int *cnt0 = cluster.map_shared_rank(&E.count[par], 0);
    for (int k = tid; k < cnt; k += ENG_T) {
      const double ti = tt[k];
      const double s = tn - ti;
      const double s2 = s * s * 0.5, s3 = s * s * s * (1.0 / 6.0);
      const double4 p = pos[k], v = vel[k], a = acc[k], j = jrk[k];
      double4 pp, pv;
      pp.x = p.x + v.x * s + a.x * s2 + j.x * s3;
      pp.y = p.y + v.y * s + a.y * s2 + j.y * s3;
      pp.z = p.z + v.z * s + a.z * s2 + j.z * s3;
      pp.w = p.w;
      pv.x = v.x + a.x * s + j.x * s2;
      pv.y = v.y + a.y * s + j.y * s2;
      pv.z = v.z + a.z * s + j.z * s2;
      pv.w = 0.0;
      ppos[k] = pp;
      pvel[k] = pv;
      if (ti + dtt[k] == tn) {
        const int slot = atomicAdd(cnt0, 1);
        if (slot < ENG_MAX_ACT) {
          for (int r = 0; r < cs; r++) {
            EngShared *R = cluster.map_shared_rank(&E, r);
            R->a_pp[slot] = pp; R->a_pv[slot] = pv;
            R->a_idx[slot] = j0 + k; R->a_owner[slot] = rank;
          }
        }
      }
    }
    PROF(0)
    cluster.sync();
    const int n_act = *cnt0;  // the same for every thread of the cluster
    PROF(1)
    if (n_act > ENG_MAX_ACT) break;  // a block for the whole chip
    // ---- 2: force of the active set against this CTA's chunk -------------------------------------
    {
      int iw = 1;
      while (iw < n_act) iw <<= 1;
      const int isub = lane & (iw - 1), jsub = lane / iw, jgroups = 32 / iw;
      const int li = isub < n_act ? isub : 0;
      const double4 p = E.a_pp[li], v = E.a_pv[li];
      Acc7 s;
      s.ax = s.ay = s.az = s.jx = s.jy = s.jz = s.pot = 0.0;
#pragma unroll 4
      for (int jj = warp * jgroups + jsub; jj < cnt; jj += ENG_WARPS * jgroups)
        pair_interaction(ppos[jj], pvel[jj], eps2, p.x, p.y, p.z, v.x, v.y, v.z, s);
      for (int o = iw; o < 32; o <<= 1) {  // fixed butterfly over the lane bits above log2(iw)
        s.ax += __shfl_xor_sync(0xffffffffu, s.ax, o); s.ay += __shfl_xor_sync(0xffffffffu, s.ay, o);
        s.az += __shfl_xor_sync(0xffffffffu, s.az, o); s.jx += __shfl_xor_sync(0xffffffffu, s.jx, o);
        s.jy += __shfl_xor_sync(0xffffffffu, s.jy, o); s.jz += __shfl_xor_sync(0xffffffffu, s.jz, o);
        s.pot += __shfl_xor_sync(0xffffffffu, s.pot, o);
      }
      if (lane < iw) {
        E.red[warp][0][lane] = s.ax; E.red[warp][1][lane] = s.ay; E.red[warp][2][lane] = s.az;
        E.red[warp][3][lane] = s.jx; E.red[warp][4][lane] = s.jy; E.red[warp][5][lane] = s.jz;
        E.red[warp][6][lane] = s.pot;
      }
      __syncthreads();
      if (tid < 16 * n_act) {  // 16 lanes per slot: fixed xor-butterfly over the warps' rows, lane 0 -> the slot owner
        const int slot = tid >> 4, row = tid & 15;
        double r[7];
#pragma unroll
        for (int c = 0; c < 7; c++) r[c] = E.red[row][c][slot];
        const unsigned mask = __activemask();
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
          for (int c = 0; c < 7; c++) r[c] += __shfl_xor_sync(mask, r[c], o);
        }
        if (row == 0) {
          double *dst = cluster.map_shared_rank(&E.part[slot][rank][0], E.a_owner[slot]);
#pragma unroll
          for (int c = 0; c < 7; c++) dst[c] = r[c];
        }
      }
      if (rank == 0 && tid == 0) E.count[par ^ 1] = 0;
    }
    PROF(2)
    cluster.sync();
    PROF(3)
    // ---- 3: corrector on the slots this CTA owns; next block time ----------------------------------
    {
      // warp 0: the correctors of the slots this CTA owns; the other warps meanwhile scan the chunk's untouched
      // particles for min(t + dt)
      unsigned long long v = INF_BITS;
      if (warp == 0) {
        if (tid < n_act && E.a_owner[tid] == rank) {
          double r[7];
#pragma unroll
          for (int c = 0; c < 7; c++) {
            double a = E.part[tid][0][c];
            for (int q = 1; q < cs; q++) a += E.part[tid][q][c];
            r[c] = a;
          }
          const int i = E.a_idx[tid], k = i - j0;
          SlotIn in;
          in.i = i;
          in.a0 = acc[k]; in.j0 = jrk[k];
          in.xp = E.a_pp[tid]; in.vp = E.a_pv[tid];
          in.t = tt[k]; in.dt = dtt[k];
          unsigned long long c_bits = INF_BITS;
          NewState ns;
          correct_slot<MODE_STEP, false>(g, tn, in, r, c_bits, 0ull, Dmax, &ns);  // also writes the global records
          pos[k] = ns.pos; vel[k] = ns.vel; acc[k] = ns.acc; jrk[k] = ns.jrk;
          tt[k] = ns.t; dtt[k] = ns.dt;
          v = dbits(ns.t + ns.dt);
        }
      } else {
        for (int k = tid - 32; k < cnt; k += ENG_T - 32) {
          const double c = tt[k] + dtt[k];
          if (c != tn) {  // this step's active particles are warp 0's business
            const unsigned long long cb = dbits(c);
            v = cb < v ? cb : v;
          }
        }
      }
      v = warp_min_u64(v);
      if (lane == 0) E.wmin[warp] = v;
      __syncthreads();
      if (tid < cs) {  // thread q tells CTA q
        unsigned long long m = E.wmin[0];
#pragma unroll
        for (int w = 1; w < ENG_WARPS; w++) m = E.wmin[w] < m ? E.wmin[w] : m;
        *cluster.map_shared_rank(&E.mins[rank], tid) = m;
      }
    }
    PROF(4)
    cluster.sync();
    PROF(5)
    {
      unsigned long long m = E.mins[0];
      for (int q = 1; q < cs; q++) m = E.mins[q] < m ? E.mins[q] : m;
      tnb = m;
    }
    n_steps += 1;
    n_pairs += (long long)n_act * (long long)g.n_tot;
    {
      int b = 0;
      while ((1 << (b + 1)) <= n_act && b < 5) b++;
      hist[b] += 1;
    }
    par ^= 1;
    if (bitsd(tnb) > span) break;  // uniform
  }
#undef PROF
  cluster.sync();  // nobody leaves while a peer may still read its shared memory
  if (rank == 0 && tid == 0) {
    ctl->t_next_bits = tnb;  // the block time of the step that is still to be taken (or beyond the span)
    g.hdr->n_steps += n_steps;
    g.hdr->n_pairs += n_pairs;
    g.hdr->n_engine += n_steps;
    for (int k = 0; k < 6; k++) g.hdr->loop_cycles[k] += prof[k];
    for (int b = 0; b < 6; b++) g.hdr->nact_hist[b] += hist[b];
  }
}

// ---- host side ----
static bool g_engine_attr_set = false;

cudaError_t engine_kernel_setup(int max_smem_optin) {
  cudaError_t e = cudaFuncSetAttribute(k_engine, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_engine, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  g_engine_attr_set = (e == cudaSuccess);
  return e;
}

static void engine_config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr, int cs, int p_cap, cudaStream_t s) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(cs);
  cfg.blockDim = dim3(ENG_T);
  cfg.dynamicSmemBytes = (size_t)engine_smem_bytes(p_cap);
  cfg.stream = s;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

// Can one cluster of `cs` CTAs with p_cap particles each be resident?  (queried once per commit)
bool engine_fits(int cs, int p_cap, int max_smem_optin) {
  if (!g_engine_attr_set || engine_smem_bytes(p_cap) > max_smem_optin) return false;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  engine_config(cfg, attr, cs, p_cap, 0);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k_engine, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return n >= 1;
}

int launch_engine(const GravDev &g, int phase, int cs, int p_cap, cudaStream_t s, cudaError_t *err) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  engine_config(cfg, attr, cs, p_cap, s);
  const cudaError_t e = cudaLaunchKernelEx(&cfg, k_engine, g, phase, p_cap);
  if (err) *err = e;
  return 1;
}

}  // namespace al26
