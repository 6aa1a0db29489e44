// enrich.cu -- K5: the fused short-lived-radionuclide enrichment pass, one outer step of
// /root/reference/al26_nbody.py:878-1086 (interloper block excluded):
//   classify (:1194-1216) -> wind deposit, local + global bubble, 26Al + 60Fe (:642-702, :897-938)
//   -> supernova events + deposit (:945-967, :1326-1334) -> decay (:1048-1064) -> condense (:1071-1086).
// Compiled with --fmad=false and written in the reference's evaluation order: per disc the massive stars are
// visited in ascending index order (the sorted source table), each wind term is
// ((wind_ratio*mdot)*eta_bub)*dt with eta_bub = ((0.75*(r*r))*d_trav)/(R*(R*R)).  The reference evaluates the
// geometry four times per step (once per calc_wind_abs call); here once per (disc, source) pair.
//
// Two launches per outer step:
//   k_enrich_sources   all stars: m >= 13 Msun -> atomic append to the source list; the LAST CTA to finish sorts the
//                      list (ascending index = the reference's summation order), builds the source table
//                      {x,y,z,c26},{c60,sn26,sn60,event}, detects SN events (mdot == 0 and not kicked), sets kicked,
//                      emits the ordered event list, and (fast modes) the hoisted sums / the fast-test table / the cell
//                      grid.  More than ENR_MAX_SOURCES massive stars: flag raised, NOTHING is mutated (the disc kernel
//                      returns at once), the call fails with AL26_ECAP and the state is what it was.
//   k_enrich_discs<M>  one thread per star of this rank's slice, launched with programmatic stream serialization: all
//                      of the star's loads (inventories, kinematics, disc, flags -- the HBM traffic of the step) are
//                      issued before griddepcontrol.wait, i.e. while the source table is still being built; then the
//                      source loop, decay of all rows, condense flags.
// Modes (al26_enrich_set_mode):
//   0 exact   every (disc, source) pair in the reference's order: wind sums BIT-IDENTICAL to the reference's numba
//             kernel.  15 DP instructions per pair (6 of them the separable global model).
//   1 fast    tolerance mode (north_star: per-disc masses within 1e-10): the global model is separable, so its source
//             sum is hoisted into k_enrich_sources (relative difference ~1e-13: summation order); the local-bubble
//             test runs on |x|^2 - 2 s.x + (|s|^2 - q) < 0 with the source part precomputed: 3 DFMA + 1 DSETP per
//             pair.  Local deposits (same terms, same ascending order) and SN deposits (exact difference form over
//             the event list) stay bit-identical unless a pair sits within ~1e-12 (relative) of the bubble surface.
//   2 pruned  as 1, but from ENR_PRUNE_MIN_SOURCES massive stars up the local-bubble candidates come from per-cell
//             lists (uniform grid over the sources, cell >= bubble radius, every source entered into its 27
//             neighbouring cells: one lookup per disc) and are tested in the exact difference form: local rows
//             bit-identical to mode 0, the pair loop is gone, the kernel is HBM-bound at any source count.
#include "al26_internal.cuh"

namespace al26 {

constexpr int EN_T = 256;
constexpr int SRC_T = 1024;     // threads of k_enrich_sources
constexpr int SRC_TILE = 512;   // sources staged per shared-memory tile (2 x 16 KB)
constexpr int MATCH_CAP = 12;   // mode 2: local-bubble hits kept per disc before falling back to the full scan

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void star_pos(const EnrichDev &e, int i, double &x, double &y, double &z) {
  if (e.px) {
    const int k = i - e.pv_off;
    x = e.px[k]; y = e.py[k]; z = e.pz[k];
  } else {
    const double4 p = e.gpos[i];
    x = p.x * e.km_per_length; y = p.y * e.km_per_length; z = p.z * e.km_per_length;
  }
}
__device__ __forceinline__ void star_vel(const EnrichDev &e, int i, double &x, double &y, double &z) {
  if (e.px) {
    const int k = i - e.pv_off;
    x = e.pvx[k]; y = e.pvy[k]; z = e.pvz[k];
  } else {
    const double4 p = e.gvel[i];
    x = p.x * e.kms_per_speed; y = p.y * e.kms_per_speed; z = p.z * e.kms_per_speed;
  }
}

// fixed-order block reduction of NV values at once (warp butterflies, then every thread sums / mins the warp rows in
// order): two barriers for the lot; result valid in every thread.  OP 0: sum, 1: min.
template <int NV, int OP>
__device__ __forceinline__ void block_reduce_all(double (&v)[NV], double (*sh)[NV]) {
#pragma unroll
  for (int q = 0; q < NV; q++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double w = __shfl_xor_sync(0xffffffffu, v[q], o);
      v[q] = OP == 0 ? v[q] + w : fmin(v[q], w);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; q++) sh[warp][q] = v[q];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; q++) {
    double r = sh[0][q];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = OP == 0 ? r + sh[w][q] : fmin(r, sh[w][q]);
    v[q] = r;
  }
}

// phase 0: classify + tables; phase 1: classify + sort only (the host then gathers the massive stars' rows);
// phase 2: tables only, from the compact per-source rows e.hm_rows = [mdot, x, y, z][ENR_MAX_SOURCES]
__global__ void __launch_bounds__(SRC_T) k_enrich_sources(const EnrichDev e, const EnrichParams p, const int phase) {
  __shared__ int keys[ENR_MAX_SOURCES];
  __shared__ double shd[SRC_T / 32][6];
  __shared__ int shi[SRC_T / 32 + 1];
  __shared__ int is_last;
  pdl_launch_dependents();  // the disc kernel may start its (source-independent) loads
  const int tid = threadIdx.x;
  if (phase != 2) {
    // every mass once; four independent loads in flight per thread (the pass is one memory round trip, not four)
    const int span = gridDim.x * SRC_T;
    for (int i0 = blockIdx.x * SRC_T + tid; i0 < e.n_tot; i0 += 4 * span) {
      double mv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) mv[u] = (i0 + u * span < e.n_tot) ? e.mass_msun[i0 + u * span] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (mv[u] >= 13.0) {
          const int q = atomicAdd(&e.counters[0], 1);
          if (q < ENR_MAX_SOURCES) e.hm_list[q] = i0 + u * span;
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      is_last = (atomicAdd(&e.counters[4], 1) == (int)gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
  }
  long long tk = clock64();
#define SRC_PROF(k)                                   \
  if (tid == 0) {                                     \
    const long long now = clock64();                  \
    e.prof[k] = now - tk;                             \
    tk = now;                                         \
  }
  const int n_raw = (phase == 2) ? e.counters[3] : *(volatile int *)&e.counters[0];
  if (n_raw > ENR_MAX_SOURCES) {  // capacity exceeded: raise the flag, mutate nothing
    if (tid == 0) {
      e.counters[2] = 1;
      e.counters[1] = 0;
      e.counters[3] = 0;
    }
    return;
  }
  const int n_hm = n_raw;
  if (phase != 2) {
    if (n_hm <= 256) {
      // rank sort: the keys are distinct star indices; one barrier instead of the bitonic network's log^2 n
      int *raw = keys + ENR_MAX_SOURCES / 2;
      int key = 0;
      if (tid < n_hm) {
        key = __ldcg(&e.hm_list[tid]);
        raw[tid] = key;
      }
      __syncthreads();
      if (tid < n_hm) {
        int rank = 0;
        for (int j = 0; j < n_hm; j++) rank += (raw[j] < key) ? 1 : 0;
        keys[rank] = key;
      }
      __syncthreads();
    } else {
      int np2 = 1;
      while (np2 < n_hm) np2 <<= 1;
      for (int k = tid; k < np2; k += SRC_T) keys[k] = (k < n_hm) ? __ldcg(&e.hm_list[k]) : 0x7fffffff;
      __syncthreads();
      for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int k = tid; k < np2; k += SRC_T) {
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
    }
    for (int k = tid; k < n_hm; k += SRC_T) e.hm_list[k] = keys[k];
    if (tid == 0) e.counters[3] = n_hm;
    if (phase == 1) return;
  } else {
    for (int k = tid; k < n_hm; k += SRC_T) keys[k] = e.hm_list[k];
    __syncthreads();
  }
  SRC_PROF(0)

  // ---- source table, hoisted sums, bounding box, ordered SN event list -- one pass; origin of the fast test = the
  // first source (keeps |x'| at cluster scale) ----
  double ox = 0.0, oy = 0.0, oz = 0.0;
  if (n_hm > 0) {
    if (phase == 2) {
      ox = e.hm_rows[1 * ENR_MAX_SOURCES]; oy = e.hm_rows[2 * ENR_MAX_SOURCES]; oz = e.hm_rows[3 * ENR_MAX_SOURCES];
    } else {
      star_pos(e, keys[0], ox, oy, oz);
    }
  }
  double sums[2] = {0.0, 0.0};
  double box[6] = {1e300, 1e300, 1e300, 1e300, 1e300, 1e300};  // min x, y, z, min -x, -y, -z
  int ev_base = 0;
  for (int k0 = 0; k0 < n_hm; k0 += SRC_T) {
    const int k = k0 + tid;
    bool ev = false;
    double x = 0, y = 0, z = 0, y26 = 0, y60 = 0;
    int i = 0;
    if (k < n_hm) {
      i = keys[k];
      double mdot;
      if (phase == 2) {
        mdot = e.hm_rows[k];
        x = e.hm_rows[1 * ENR_MAX_SOURCES + k]; y = e.hm_rows[2 * ENR_MAX_SOURCES + k]; z = e.hm_rows[3 * ENR_MAX_SOURCES + k];
      } else {
        star_pos(e, i, x, y, z);
        mdot = e.mdot[i];
      }
      ev = (mdot == 0.0) && (e.kicked[i] == 0);
      const double c26 = e.wr26[i] * mdot, c60 = e.wr60[i] * mdot;
      const double q26 = e.sn26[i], q60 = e.sn60[i];  // unconditional: one round trip with the loads above
      y26 = ev ? q26 : 0.0;
      y60 = ev ? q60 : 0.0;
      e.src_a[k] = make_double4(x, y, z, c26);
      e.src_b[k] = make_double4(c60, y26, y60, ev ? 1.0 : 0.0);
      const double xs = x - ox, ys = y - oy, zs = z - oz;
      e.src_f[k] = make_double4(-2.0 * xs, -2.0 * ys, -2.0 * zs, (xs * xs + ys * ys + zs * zs) - p.q_local);
      sums[0] += c26;
      sums[1] += c60;
      box[0] = fmin(box[0], x); box[1] = fmin(box[1], y); box[2] = fmin(box[2], z);
      box[3] = fmin(box[3], -x); box[4] = fmin(box[4], -y); box[5] = fmin(box[5], -z);
    }
    // ordered compaction of this tile's events
    const unsigned m = __ballot_sync(0xffffffffu, ev);
    const int lane = tid & 31, warp = tid >> 5;
    if (lane == 0) shi[warp] = __popc(m);
    __syncthreads();
    int off = ev_base, tot = 0;
    for (int w = 0; w < SRC_T / 32; w++) {
      if (w < warp) off += shi[w];
      tot += shi[w];
    }
    if (ev) {
      const int slot = off + __popc(m & ((1u << lane) - 1u));
      e.sn_events[slot] = i;
      e.kicked[i] = 1;
      e.ev_a[slot] = make_double4(x, y, z, y26);
      e.ev_b[slot] = y60;
    }
    ev_base += tot;
    __syncthreads();
  }
  {
    double (*sh2)[2] = reinterpret_cast<double (*)[2]>(&shd[0][0]);
    block_reduce_all<2, 0>(sums, sh2);
  }
  if (tid == 0) {
    e.counters[1] = ev_base;
    e.fsum[0] = sums[0]; e.fsum[1] = sums[1];
    e.fsum[2] = ox; e.fsum[3] = oy; e.fsum[4] = oz;
  }
  SRC_PROF(1)
  if (p.mode != 2 || n_hm < ENR_PRUNE_MIN_SOURCES) return;  // few sources: the disc kernel scans them all (counters[8] stays 0)

  // ---- mode 2: per-cell candidate lists.  A uniform grid over the sources' bounding box plus a one-cell apron, cell
  // size h >= the local bubble radius; every source is entered into the lists of its 27 neighbouring cells, so a disc
  // needs ONE lookup: the list of its own cell holds every source that can be within R of it.  (Measured and rejected:
  // lists of {x, y, z, source} records instead of source numbers -- one dependent load fewer per candidate, but the
  // single table-building CTA then scatters 27 x 32 bytes per source, 39 us at 1000 sources.) ----
  __syncthreads();
  block_reduce_all<6, 1>(box, shd);
  double lo[3] = {box[0], box[1], box[2]};
  const double hi[3] = {-box[3], -box[4], -box[5]};
  int gd[3];
  double h = p.r_local * (1.0 + 1e-6);
  {
    const double ext = fmax(hi[0] - lo[0], fmax(hi[1] - lo[1], hi[2] - lo[2]));
    h = fmax(h, ext / (double)ENR_GRID_MAX * (1.0 + 1e-9));
    for (int c = 0; c < 3; c++) {
      int g = (int)floor((hi[c] - lo[c]) / h) + 1;
      g = g < 1 ? 1 : (g > ENR_GRID_MAX ? ENR_GRID_MAX : g);
      gd[c] = g + 2;         // apron
      lo[c] -= h;
    }
  }
  const double inv_h = 1.0 / h;
  const int ncell = gd[0] * gd[1] * gd[2];
  // counts -> starts -> scatter cursors all live in (dynamic) shared memory: one CTA, no global round trips
  extern __shared__ int cells[];  // [ncell + 1]
  for (int c = tid; c <= ncell; c += SRC_T) cells[c] = 0;
  __syncthreads();
  SRC_PROF(2)
  for (int pass = 0; pass < 2; pass++) {
    for (int k = tid; k < n_hm; k += SRC_T) {
      const double4 A = e.src_a[k];
      int cx = (int)floor((A.x - lo[0]) * inv_h), cy = (int)floor((A.y - lo[1]) * inv_h), cz = (int)floor((A.z - lo[2]) * inv_h);
      cx = min(max(cx, 1), gd[0] - 2); cy = min(max(cy, 1), gd[1] - 2); cz = min(max(cz, 1), gd[2] - 2);
      for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
          for (int dx = -1; dx <= 1; dx++) {
            const int cell = ((cz + dz) * gd[1] + (cy + dy)) * gd[0] + (cx + dx);
            if (pass == 0) atomicAdd(&cells[cell + 1], 1);  // counts shifted by one: the scan leaves the starts
            else e.cell_items[atomicAdd(&cells[cell], 1)] = k;
          }
    }
    __syncthreads();
    if (pass == 1) break;
    SRC_PROF(3)
    {  // inclusive scan of cells[1..ncell] in place (each thread a contiguous run, then the block's run totals)
      const int per = (ncell + SRC_T - 1) / SRC_T;
      const int b = 1 + tid * per, en = min(ncell + 1, b + per);
      int run = 0;
      for (int c = b; c < en; c++) run += cells[c];
      const int lane = tid & 31, warp = tid >> 5;
      int incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) shi[warp] = incl;
      __syncthreads();
      int off = incl - run;
      for (int w = 0; w < warp; w++) off += shi[w];
      for (int c = b; c < en; c++) {
        off += cells[c];
        cells[c] = off;
      }
    }
    __syncthreads();
    for (int c = tid; c <= ncell; c += SRC_T) e.cell_start[c] = cells[c];  // the scatter below turns cells[] into cursors
    __syncthreads();
    SRC_PROF(4)
  }
  SRC_PROF(5)
  if (tid == 0) {
    e.counters[5] = gd[0]; e.counters[6] = gd[1]; e.counters[7] = gd[2];
    e.counters[8] = 1;  // the lists are there
    e.fsum[5] = lo[0]; e.fsum[6] = lo[1]; e.fsum[7] = lo[2]; e.fsum[8] = inv_h;
  }
#undef SRC_PROF
}

// one thread per star.  MODE: see the file header.
template <int MODE>
__global__ void __launch_bounds__(EN_T, 4) k_enrich_discs(const EnrichDev e, const EnrichParams p) {
  __shared__ double4 sa[SRC_TILE];
  __shared__ double4 sb[SRC_TILE];
  const int li = blockIdx.x * EN_T + threadIdx.x;
  const bool valid = li < e.n_loc;
  const int gi = e.d0 + (valid ? li : 0);
  const int lj = valid ? li : 0;
  const size_t n = (size_t)e.n_loc;
  const bool agb = p.with_agb != 0;

  // ---- every load of this star up front: one memory round trip, none of it depends on the source kernel ----
  const double m = e.mass_msun[gi];
  double inv[ENR_NINV];
#pragma unroll
  for (int r = 0; r < ENR_NINV; r++) inv[r] = (agb || (r != 3 && r != 7)) ? e.inv[r * n + lj] : 0.0;
  double x, y, z, vx, vy, vz;
  star_pos(e, gi, x, y, z);
  star_vel(e, gi, vx, vy, vz);
  const double rd = e.r_disk[lj];
  const double tau = e.tau_disk[lj];
  const uint8_t was_alive = e.alive[lj];

  pdl_wait_primary();  // the source table, counters and event list are complete and visible
  if (e.counters[2] != 0) return;  // capacity exceeded: the call fails, nothing may change
  const int n_hm = e.counters[3];
  const int n_ev = e.counters[1];
  const bool is_lm = valid && (m >= 0.1) && (m <= 3.0);

  double eta_l = 0.0, eta_g = 0.0;
  if (is_lm && n_hm > 0) {
    const double spd = sqrt(vx * vx + vy * vy + vz * vz);   // (lm_vx**2 + lm_vy**2 + lm_vz**2)**0.5
    const double trav = spd * p.dt_s;                       // d_disk_trav = disk_spd * dt
    const double k0 = 0.75 * (rd * rd) * trav;              // 0.75 * (r_disk**2) * d_disk_trav
    eta_l = k0 / p.r_local3;                                // / (bubble_radius ** 3)
    eta_g = k0 / p.r_global3;
  }
  double l26 = 0.0, l60 = 0.0, g26 = 0.0, g60 = 0.0;
  double s26 = inv[2], s60 = inv[6];  // SN deposits add straight onto the inventories (:964-965)

  if (MODE == 0) {
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
  } else {
    // ---- fast modes: hoisted global sum; SN deposits over the (short, ordered) event list in the exact form ----
    if (is_lm && n_hm > 0) {
      g26 = (e.fsum[0] * eta_g) * p.dt_s;
      g60 = (e.fsum[1] * eta_g) * p.dt_s;
      for (int q = 0; q < n_ev; q++) {
        const double4 A = e.ev_a[q];
        const double dx = x - A.x, dy = y - A.y, dz = z - A.z;
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        const double eta = (0.5 * 0.7) * ((0.5 * (rd * rd)) / (4.0 * (d * d)));
        s26 += A.w * eta;
        s60 += e.ev_b[q] * eta;
      }
    }
    const bool have_lists = (MODE == 2) && __ldcg(&e.counters[8]) != 0;  // uniform; few sources -> the all-pairs loop below
    bool full_scan = false;
    if (have_lists && is_lm) {
      const int gx = __ldcg(&e.counters[5]), gy = __ldcg(&e.counters[6]), gz = __ldcg(&e.counters[7]);
      const double inv_h = e.fsum[8];
      const double ux = (x - e.fsum[5]) * inv_h, uy = (y - e.fsum[6]) * inv_h, uz = (z - e.fsum[7]) * inv_h;
      // outside the grid (apron included) no source can be within R; the comparisons also drop NaNs
      if (ux >= 0.0 && uy >= 0.0 && uz >= 0.0 && ux < (double)gx && uy < (double)gy && uz < (double)gz) {
        const int cell = ((int)uz * gy + (int)uy) * gx + (int)ux;
        const int tb = __ldg(&e.cell_start[cell]), te = __ldg(&e.cell_start[cell + 1]);
        int hits[MATCH_CAP];
        int nh = 0;
        for (int t = tb; t < te; t++) {
          const int k = __ldg(&e.cell_items[t]);
          const double4 A = e.src_a[k];
          const double dx = x - A.x, dy = y - A.y, dz = z - A.z;
          const double d2 = dx * dx + dy * dy + dz * dz;
          if (!(d2 >= p.q_local)) {
            if (nh < MATCH_CAP) hits[nh] = k;
            nh++;
          }
        }
        if (nh > MATCH_CAP) {
          full_scan = true;  // a disc inside more bubbles than the hit list holds: take the all-pairs scan for it
        } else {
          for (int a = 1; a < nh; a++) {  // ascending source order = the reference's summation order
            const int v = hits[a];
            int b = a - 1;
            while (b >= 0 && hits[b] > v) {
              hits[b + 1] = hits[b];
              b--;
            }
            hits[b + 1] = v;
          }
          for (int a = 0; a < nh; a++) {
            l26 += (e.src_a[hits[a]].w * eta_l) * p.dt_s;
            l60 += (e.src_b[hits[a]].x * eta_l) * p.dt_s;
          }
        }
      }
    }
    if (MODE == 1 || !have_lists) {
      // all pairs, 3 DFMA + 1 DSETP each: d^2 - q = |x'|^2 + (-2 s'.x' + |s'|^2 - q), positions relative to the origin
      const double xr = x - e.fsum[2], yr = y - e.fsum[3], zr = z - e.fsum[4];
      const double nx2 = -(xr * xr + yr * yr + zr * zr);
      for (int b = 0; b < n_hm; b += SRC_TILE) {
        const int cnt = min(SRC_TILE, n_hm - b);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += EN_T) {
          sa[k] = e.src_f[b + k];
          sb[k] = make_double4(e.src_a[b + k].w, e.src_b[b + k].x, 0.0, 0.0);
        }
        __syncthreads();
        if (is_lm) {
          for (int k0 = 0; k0 < cnt; k0 += 64) {
            const int gn = min(64, cnt - k0);
            unsigned long long marked = 0ull;
#pragma unroll 8
            for (int kk = 0; kk < gn; kk++) {
              const double4 F = sa[k0 + kk];
              const double t = fma(F.x, xr, fma(F.y, yr, fma(F.z, zr, F.w)));
              if (t < nx2) marked |= 1ull << kk;
            }
            while (marked) {
              const int kk = __ffsll((long long)marked) - 1;
              marked &= marked - 1ull;
              l26 += (sb[k0 + kk].x * eta_l) * p.dt_s;
              l60 += (sb[k0 + kk].y * eta_l) * p.dt_s;
            }
          }
        }
      }
    } else if (__syncthreads_or(full_scan ? 1 : 0)) {
      // mode 2 fallback (rare): exact all-pairs scan for the discs that overflowed their hit list
      for (int b = 0; b < n_hm; b += SRC_TILE) {
        const int cnt = min(SRC_TILE, n_hm - b);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += EN_T) {
          sa[k] = e.src_a[b + k];
          sb[k] = e.src_b[b + k];
        }
        __syncthreads();
        if (full_scan) {
          for (int kk = 0; kk < cnt; kk++) {
            const double4 A = sa[kk];
            const double dx = x - A.x, dy = y - A.y, dz = z - A.z;
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (!(d2 >= p.q_local)) {
              l26 += (A.w * eta_l) * p.dt_s;
              l60 += (sb[kk].x * eta_l) * p.dt_s;
            }
          }
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
  if (agb) {
    inv[3] *= p.decay26;
    inv[7] *= p.decay60;
  }
#pragma unroll
  for (int r = 0; r < ENR_NINV; r++)
    if (agb || (r != 3 && r != 7)) e.inv[r * n + li] = inv[r];  // the agb rows are untouched without an interloper
  // condense (:1071-1086)
  if (is_lm && was_alive) {
    if (tau >= p.t_new_myr) {
#pragma unroll
      for (int r = 0; r < ENR_NINV; r++)
        if (agb || (r != 3 && r != 7)) e.fin[r * n + li] = inv[r];
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

cudaError_t enrich_kernel_setup() {
  return cudaFuncSetAttribute(k_enrich_sources, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)((ENR_GRID_CELLS + 1) * sizeof(int)));
}

int enrich_sources_grid(int n_tot, int sm_count) {
  int g = (n_tot + SRC_T - 1) / SRC_T;
  const int cap = sm_count > 0 ? sm_count : 148;  // 1024 threads x 64 registers: one CTA per SM, one round
  return g < 1 ? 1 : (g > cap ? cap : g);
}

// classify + sort only (phase 1 of the sliced multi-GPU upload)
int launch_enrich_classify(const EnrichDev &e, const EnrichParams &p, int sm_count, cudaStream_t s) {
  k_enrich_sources<<<enrich_sources_grid(e.n_tot, sm_count), SRC_T, 0, s>>>(e, p, 1);
  return 1;
}

// tables_only: the source list is already on the device (sorted) and the compact rows e.hm_rows are filled
int launch_enrich(const EnrichDev &e, const EnrichParams &p, int sm_count, bool tables_only, cudaStream_t s, cudaError_t *err) {
  const size_t dsm = (p.mode == 2) ? (size_t)(ENR_GRID_CELLS + 1) * sizeof(int) : 0;  // the cell array of the list build
  if (tables_only) k_enrich_sources<<<1, SRC_T, dsm, s>>>(e, p, 2);
  else k_enrich_sources<<<enrich_sources_grid(e.n_tot, sm_count), SRC_T, dsm, s>>>(e, p, 0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((e.n_loc + EN_T - 1) / EN_T);
  cfg.blockDim = dim3(EN_T);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t rc;
  switch (p.mode) {
    case 1: rc = cudaLaunchKernelEx(&cfg, k_enrich_discs<1>, e, p); break;
    case 2: rc = cudaLaunchKernelEx(&cfg, k_enrich_discs<2>, e, p); break;
    default: rc = cudaLaunchKernelEx(&cfg, k_enrich_discs<0>, e, p); break;
  }
  if (err) *err = rc;
  return 2;
}

}  // namespace al26
