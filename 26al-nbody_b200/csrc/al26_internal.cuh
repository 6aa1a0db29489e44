// al26_internal.cuh -- device-side data layout and kernel launch prototypes shared by the
// translation units of libal26b200.so.  Not part of the public C-ABI (include/al26_b200.h).
//
// Layout in HBM (all fp64, "SoA of double4"): per local particle
//   pos = {x, y, z, m}   vel = {vx, vy, vz, -}   acc = {ax, ay, az, pot}   jrk = {jx, jy, jz, -}
//   t, dt (time relative to the start of the current evolve call; dt a power of two)
// and for all N (global) particles the predicted j-set the force kernel streams through
// shared memory by TMA bulk copies:
//   jpos = {xp, yp, zp, m}   jvel = {vxp, vyp, vzp, -}          (64 B per j)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace al26 {

constexpr unsigned long long INF_BITS = 0x7FF0000000000000ull;  // +inf as ordered bits

// one block step's scheduler record; three of them rotate (cur / next / being reset)
struct StepCtrl {
  unsigned long long t_next_bits;  // min over particles of (t + dt), as ordered bits of a double >= 0
  int n_act;                       // entries in the active list
  int work_counter;                // dynamic work-item counter of the force kernel
  int pad[4];
};

struct GravHeader {
  double span;  // length of the current evolve call
  double D;     // largest step of the call's dyadic ladder
  int done;     // next block time exceeds span
  int pad;
  long long n_steps;  // block steps taken (this rank)
  long long n_pairs;  // (i,j) pair evaluations (this rank)
  long long nact_hist[32];  // diagnostic: block steps by floor(log2(n_act)), since the last commit
};

enum StepMode { MODE_STEP = 0, MODE_INIT = 1, MODE_SYNC = 2, MODE_RAW = 3 };

struct GravDev {
  int n_loc;  // particles owned by this rank
  int n_tot;  // global particle count (the j-set)
  int i0;     // global index of local particle 0
  int grid_force;
  int variant;    // force-kernel configuration (hermite_force.cu)
  int force_ipt;  // its i-particles per lane for big blocks
  double eps2, eta, dt_max, dt_min;
  double4 *pos, *vel, *acc, *jrk;
  double *t, *dt;
  double4 *jpos, *jvel;
  int *list;
  double4 *part_a, *part_j;
  long long part_cap;
  StepCtrl *ctrl;
  GravHeader *hdr;
  // MODE_RAW outputs (parity hook al26_grav_force)
  double4 *raw_a, *raw_j;
};

// ---- work decomposition of one force evaluation, a pure function of (n_act, n_tot, grid) so
// the force kernel and the reduce/corrector kernel agree without communicating ----
constexpr int FORCE_TJ = 256;      // j per TMA tile (2 x 8 KB per stage)
constexpr int FORCE_STAGES = 3;
constexpr int FORCE_MIN_JCHUNK = 256;
constexpr int FORCE_MAX_ROUNDS = 16;             // work items per CTA at most (load balance for big blocks)
constexpr long long FORCE_ITEM_PAIRS = 250000;   // do not cut items finer than this many pairs (~150 us)
constexpr int FORCE_BIG_NACT_PER_IPT = 1024;      // n_act >= this x IPT: IPT i-particles per lane
constexpr int FORCE_SPLIT_MAX_NACT = 16;         // n_act <= this: lanes split over j as well (tiny blocks)
constexpr int FORCE_IPT_MAX = 4;

struct Decomp {
  int ipt;          // i-particles per lane (1, or the configuration's IPT for big blocks)
  int ti;           // i per work item = 32 * ipt
  int n_itiles;
  int n_jsplit;
  int jchunk;
  int slot_stride;  // n_itiles * ti
};

// Items = n_itiles x n_jsplit, handed out dynamically.  Small blocks: one round, at most `grid` items
// (every CTA gets one big item; per-item overhead -- TMA prologue, barrier, reduction -- is paid once).
// Big blocks: up to FORCE_MAX_ROUNDS rounds so the tail of the last round stays small.  The item
// count is kept just BELOW a multiple of the grid so the last round is full.
__host__ __device__ inline Decomp make_decomp(int n_act, int n_tot, int grid, int ipt_big) {
  Decomp d;
  d.ipt = (n_act >= FORCE_BIG_NACT_PER_IPT * ipt_big) ? ipt_big : 1;
  d.ti = 32 * d.ipt;
  d.n_itiles = (n_act + d.ti - 1) / d.ti;
  if (d.n_itiles < 1) d.n_itiles = 1;
  const long long total = (long long)d.n_itiles * d.ti * (long long)n_tot;
  long long rounds = total / ((long long)grid * FORCE_ITEM_PAIRS);
  if (rounds < 1) rounds = 1;
  if (rounds > FORCE_MAX_ROUNDS) rounds = FORCE_MAX_ROUNDS;
  int ns = (int)((rounds * grid) / d.n_itiles);
  int max_by_j = n_tot / FORCE_MIN_JCHUNK;
  if (max_by_j < 1) max_by_j = 1;
  if (ns > max_by_j) ns = max_by_j;
  if (ns < 1) ns = 1;
  int jc = (n_tot + ns - 1) / ns;
  jc = (jc + 7) & ~7;
  if (jc < 8) jc = 8;
  d.jchunk = jc;
  d.n_jsplit = (n_tot + jc - 1) / jc;
  if (d.n_jsplit < 1) d.n_jsplit = 1;
  d.slot_stride = d.n_itiles * d.ti;
  return d;
}

// partial-buffer entries that cover every n_act in [0, n_loc]
inline long long part_capacity(int n_loc, int grid) {
  return (long long)n_loc + 32 * FORCE_IPT_MAX + (long long)(FORCE_MAX_ROUNDS * grid + 1) * 32 * FORCE_IPT_MAX * 2;
}

// ---- launchers (each enqueues on `s`; returns the number of kernels launched) ----
int launch_begin(const GravDev &g, double span, double D, cudaStream_t s);
int launch_predict_list(const GravDev &g, int mode, int phase, cudaStream_t s);
int launch_force(const GravDev &g, int phase, cudaStream_t s);
int launch_correct(const GravDev &g, int mode, int phase, cudaStream_t s);
int launch_snapshot_j(const GravDev &g, cudaStream_t s);  // jpos/jvel := current state (s = 0)
int force_smem_bytes();
int force_variant_count();
int force_variant_info(int v, int *ctas_per_sm, int *ipt);
double launch_dfma_peak(int sm_count, int iters, double *scratch, cudaStream_t s);  // returns flops per launch
cudaError_t force_kernel_setup();

// energies (K4): per-rank partial sums over local i x all j
struct EnergyDev {
  int n_loc, n_tot, i0;
  double eps2;
  const double4 *pos, *vel;   // local
  const double4 *jpos;        // global snapshot
  double *block_part;         // [grid][3]
  double *out;                // [3] K, U, S
};
int launch_energies(const EnergyDev &e, cudaStream_t s);
int energy_grid(int n_loc);

// enrichment (K5)
constexpr int ENR_NINV = 8;
constexpr int ENR_MAX_SOURCES = 8192;
struct EnrichDev {
  int n_tot;      // global star count (classification, source table)
  int d0, n_loc;  // this rank's disc slice [d0, d0 + n_loc)
  // replicated per-star inputs (global length)
  const double *mass_msun, *mdot;
  const double *px, *py, *pz, *pvx, *pvy, *pvz;  // km, km/s  (explicit arrays) or null
  const double4 *gpos, *gvel;                    // gravity snapshot (global), used when px == null
  double km_per_length, kms_per_speed;
  const double *wr26, *wr60, *sn26, *sn60;       // global length
  uint8_t *kicked;                               // global length
  // sharded per-disc state (local length)
  const double *r_disk, *tau_disk;
  uint8_t *alive;
  double *inv, *fin;  // [8][n_loc]
  // scratch
  int *hm_list;       // [ENR_MAX_SOURCES]
  int *counters;      // [0] n_hm, [1] n_events, [2] overflow flag
  double4 *src_a;     // {x, y, z, c26}
  double4 *src_b;     // {c60, sn26, sn60, is_event}
  int *sn_events;     // [ENR_MAX_SOURCES]
};
struct EnrichParams {
  double dt_s, t_new_myr;
  double r_local, r_local3, q_local;  // bubble radius, its cube, and the d^2 threshold equivalent to R <= sqrt(d2)
  double r_global3;
  double decay26, decay60;
  int with_agb;
};
int launch_enrich(const EnrichDev &e, const EnrichParams &p, cudaStream_t s);

}  // namespace al26
