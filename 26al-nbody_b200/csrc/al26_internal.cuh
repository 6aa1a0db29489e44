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
#include <cstdlib>

namespace al26 {

constexpr unsigned long long INF_BITS = 0x7FF0000000000000ull;  // +inf as ordered bits

// one block step's scheduler record; three of them rotate (cur / next / being reset)
struct StepCtrl {
  unsigned long long t_next_bits;  // min over particles of (t + dt), as ordered bits of a double >= 0
  int n_act;                       // entries in the active list
  int work_counter;                // dynamic work-item counter of the force kernel; fused steps: CTAs whose partials are stored
  int pad[4];                      // [0] peer-memory mode: active particles this rank owns; [1] fused steps: slots corrected
};

constexpr int DIST_PROF_N = 12;
struct GravHeader {
  double span;  // length of the current evolve call
  double D;     // largest step of the call's dyadic ladder
  int done;     // next block time exceeds span
  int pad;
  long long n_steps;  // block steps taken (this rank)
  long long n_pairs;  // (i,j) pair evaluations (this rank)
  long long nact_hist[32];  // diagnostic: block steps by floor(log2(n_act)), since the last commit
  unsigned int bar_counter;  // grid barrier of the persistent loop kernel (zeroed before every launch)
  int loop_error;            // raised when a barrier spin limit is hit
  int phase;                 // StepCtrl phase after the last loop-kernel launch
  int pad2;
  unsigned long long dist_step;  // id of the last block step completed in peer-memory mode
  unsigned long long dist_tnext_bits;  // the global next block time after it
  int dist_prev_exch;                  // the last block step was an exchanged one (its staged records are still to be pulled)
  int pad3;
  long long loop_cycles[6];  // diagnostic: CTA 0's SM cycles in predict / barrier / force / barrier / correct / barrier
  long long n_fused;             // diagnostic: block steps taken through the fused path, since the last commit
  long long n_engine;            // diagnostic: block steps taken by the cluster engine, since the last commit
  long long fuse_ns[16];         // diagnostic (builds with -DAL26_FUSE_TIMING only): globaltimer ns per fused-step segment
  // peer-memory mode, CTA 0's SM cycles since the last commit: [0] fused redundant steps, [1] their number, [2] other
  // redundant steps, [3] their number, [4] their active particles; exchanged steps: [5] predictor + scheduler + barrier,
  // [6] force on the own share + barrier, [7] corrector + peer stores, [8] cross-GPU barrier, [9] their number,
  // [10] their active particles
  long long dist_prof[DIST_PROF_N];
  unsigned int chip_seq;  // chip engine: the last step number it used (goes on from launch to launch)
  int pad4;
  long long n_chip;       // diagnostic: block steps taken by the chip engine, since the last commit
};

enum StepMode { MODE_STEP = 0, MODE_INIT = 1, MODE_SYNC = 2, MODE_RAW = 3 };

// ---- fused small block steps of the persistent loop kernels (hermite_loop.cu): the scheduler pass leaves every
// active particle's predicted state and old force in this compact record, so the force phase and the correcting
// CTA fetch everything they need in one round trip.  One buffer is enough: a step's readers all finish before
// the step's release counter is complete.
constexpr int FUSE_CAP = 32;  // largest block handled by the fused path
constexpr int FUSE_AUTO_MAX_N = 32768;  // automatic setting: fused path on for N <= this many particles
struct ActBuf {
  double4 pos[FUSE_CAP], vel[FUSE_CAP];  // predicted {x,y,z,m}, {vx,vy,vz,-}
  double4 acc[FUSE_CAP], jrk[FUSE_CAP];  // force at the start of the step
  double2 tdt[FUSE_CAP];                 // {t, dt}
  int idx[FUSE_CAP];
};

// ---- peer-memory multi-GPU mode: every rank holds the full state; a rank corrects the
// active particles it owns (i % world == rank) and stores their new state straight into EVERY rank's staging
// slab over NVLink; staged records are pulled into the local state by the next predictor pass.  One slab per
// rank (one cudaMalloc, exported with CUDA IPC): two parities of {pos, vel, acc, jrk, t, dt, tag} for all N
// particles, then the mailbox used by the cross-GPU barrier.
constexpr int MAX_PEERS = 8;
struct StagingView {
  double4 *pos, *vel, *acc, *jrk;
  double *t, *dt;
  unsigned int *tag;  // id of the block step that wrote the record
};
struct MboxEntry {  // two self-validating words: (step id low 32 bits) << 32 | one half of the sender's candidate time
  unsigned long long tmin_bits;  // ... | high half of min(t+dt)
  unsigned long long tag;        // ... | low half
};
__host__ __device__ inline size_t staging_parity_bytes(int n) {
  return ((size_t)n * (4 * 32 + 2 * 8 + 4) + 255) & ~(size_t)255;
}
__host__ __device__ inline StagingView staging_view(void *slab, int n, int parity) {
  char *b = (char *)slab + (size_t)parity * staging_parity_bytes(n);
  StagingView v;
  v.pos = (double4 *)b;
  v.vel = v.pos + n;
  v.acc = v.vel + n;
  v.jrk = v.acc + n;
  v.t = (double *)(v.jrk + n);
  v.dt = v.t + n;
  v.tag = (unsigned int *)(v.dt + n);
  return v;
}
__host__ __device__ inline MboxEntry *mbox_of(void *slab, int n) {
  return (MboxEntry *)((char *)slab + 2 * staging_parity_bytes(n));
}
inline size_t slab_bytes(int n) { return 2 * staging_parity_bytes(n) + 2 * MAX_PEERS * sizeof(MboxEntry) + 256; }

struct GravDev {
  int n_loc;  // particles owned by this rank
  int n_tot;  // global particle count (the j-set)
  int i0;     // global index of local particle 0
  int grid_force;
  int variant;    // force-kernel configuration (hermite_force.cu)
  int force_ipt;  // its i-particles per lane for big blocks
  const int *decomp_tab;  // n_jsplit by number of i-tiles (fill_decomp_table)
  int big_nact;           // blocks of at least this many particles use force_ipt i-particles per lane
  double eps2, eta, dt_max, dt_min;
  double Dmax;  // largest step of the current call's dyadic ladder (= hdr->D; set by the host at begin_evolve)
  int keep_dt;  // MODE_INIT only: recompute acc / jerk / pot, leave every particle's timestep alone (mass-only update)
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
  // peer-memory mode
  int rank, world, p2p;
  // fused small block steps (loop kernels; needs n_loc == n_tot and a j-chunk per CTA that fits the stage buffers)
  ActBuf *act;
  int fuse_max;   // block steps of at most this many active particles take the fused path (0 = off)
  int fuse_jc;    // particles per CTA: CTA b owns [b * fuse_jc, (b + 1) * fuse_jc)
  int *list_own;   // active particles this rank owns (i % world == rank); `list` holds ALL active ones
  int split_min;   // block steps with fewer active particles are computed redundantly by every rank (no exchange)
  void *slab[MAX_PEERS];  // slab[q]: rank q's staging slab as mapped in this process (slab[rank] = own)
  // chip engine (hermite_chip.cu): per-CTA mail + flags, CTAs, particles per CTA, largest block it steps (0 = off)
  void *chip_mail;
  int chip_n, chip_p, chip_max;
};

// ---- the cluster engine (hermite_engine.cu): runs of small block steps for small N inside one thread-block cluster
constexpr int ENG_MAX_ACT = 32;  // largest block the engine steps itself
constexpr int ENG_CS_MAX = 16;   // cluster size: 8 (portable) or 16
int engine_smem_bytes(int p_cap);
bool engine_plan(int n, int max_smem, int *cs_out, int *p_cap_out);
cudaError_t engine_kernel_setup(int max_smem_optin);
bool engine_fits(int cs, int p_cap, int max_smem_optin);

// ---- the chip engine (hermite_chip.cu): runs of small block steps with the particle set resident in the shared
// memory of the whole chip, one CTA per SM; no grid barrier, no atomics: single-writer mail, flags and partial rows
constexpr int CHIP_CAP = 256;       // largest block the engine can be asked to step (records per CTA)
constexpr int CHIP_MAX_CTAS = 256;  // CTAs at most
constexpr int CHIP_MAX_ACT_DEFAULT = 32;
struct alignas(128) ChipMail {
  unsigned long long w[16];            // header: {step | hi(min t+dt)}, {step | lo}, {step | count}; the rest pads to a line
  unsigned long long rec[CHIP_CAP][16];  // the particles of this chunk that attain the chunk's minimum, predicted to that time:
                                         // {step | half} words: x, y, z, m, vx, vy, vz as {hi, lo} pairs, then the index
};
int chip_smem_bytes(int p_cap);
size_t chip_mail_bytes(int n_ctas);
bool chip_plan(int n, int n_ctas, int max_smem, int *p_cap_out);
cudaError_t chip_kernel_setup(int max_smem_optin);
bool chip_fits(int n_ctas, int p_cap, int sm_count, int max_smem_optin);

// ---- work decomposition of one force evaluation, a pure function of (n_act, n_tot, grid) so
// the force kernel and the reduce/corrector kernel agree without communicating ----
constexpr int FORCE_TJ = 256;      // j per TMA tile (2 x 8 KB per stage)
constexpr int FORCE_STAGES = 3;
constexpr int FORCE_MIN_JCHUNK = 64;
constexpr int FORCE_MAX_ROUNDS = 32;             // default: work items per CTA at most (load balance for big blocks)
constexpr int FORCE_MAX_ROUNDS_CAP = 64;         // upper limit of the tunable (sizes the partial buffers)
constexpr double FORCE_ITEM_OVERHEAD_PAIRS = 3200.0;  // fixed cost of one item (TMA prologue, barriers, reduction) in pair units
constexpr int FORCE_BIG_NACT_DEFAULT = 2048;      // n_act >= GravDev::big_nact: the configuration's IPT i-particles per lane
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

// Items = n_itiles x n_jsplit, handed out dynamically.  n_jsplit is chosen among the candidates
// floor(k * grid / n_itiles), k = 1..FORCE_MAX_ROUNDS (item count just below a multiple of the grid, so the
// last round is full) to maximise  (fill of the last round) x (item work / (item work + fixed item cost)):
// small blocks get one round of big items, big blocks get many rounds so the tail stays short.
// The search runs on the HOST once per commit (decomp_table); kernels only look the answer up.
inline int choose_jsplit(int n_itiles, int ti, int n_tot, int grid, int max_rounds = FORCE_MAX_ROUNDS,
                         double overhead_pairs = FORCE_ITEM_OVERHEAD_PAIRS) {
  int max_by_j = n_tot / FORCE_MIN_JCHUNK;
  if (max_by_j < 1) max_by_j = 1;
  int best_ns = 1, prev = 0;
  double best_eff = -1.0;
  for (int k = 1; k <= max_rounds; k++) {
    int ns = (int)(((long long)k * grid) / n_itiles);
    if (ns > max_by_j) ns = max_by_j;
    if (ns < 1) ns = 1;
    if (ns == prev) continue;
    prev = ns;
    const long long items = (long long)n_itiles * ns;
    const long long rounds = (items + grid - 1) / grid;
    const double w = (double)ti * (double)((n_tot + ns - 1) / ns);
    const double eff = ((double)items / (double)(rounds * grid)) * (w / (w + overhead_pairs));
    if (eff > best_eff) {
      best_eff = eff;
      best_ns = ns;
    }
  }
  return best_ns;
}

// table layout: entries [0, n_small) for ipt = 1 (index n_itiles - 1), then entries for ipt = ipt_big
inline int decomp_small_entries(int big_nact) { return (big_nact + 31) / 32 + 1; }
inline int decomp_table_entries(int n_loc, int ipt_big, int big_nact) {
  return decomp_small_entries(big_nact) + (n_loc + 32 * ipt_big - 1) / (32 * ipt_big) + 2;
}
inline void fill_decomp_table(int *tab, int n_loc, int n_tot, int grid, int ipt_big, int big_nact,
                              int max_rounds = FORCE_MAX_ROUNDS, double overhead_pairs = FORCE_ITEM_OVERHEAD_PAIRS) {
  const int ns_small = decomp_small_entries(big_nact);
  for (int t = 1; t <= ns_small; t++) tab[t - 1] = choose_jsplit(t, 32, n_tot, grid, max_rounds, overhead_pairs);
  const int nb = decomp_table_entries(n_loc, ipt_big, big_nact) - ns_small;
  for (int t = 1; t <= nb; t++) tab[ns_small + t - 1] = choose_jsplit(t, 32 * ipt_big, n_tot, grid, max_rounds, overhead_pairs);
}

__host__ __device__ inline Decomp make_decomp(int n_act, int n_tot, const int *__restrict__ tab, int ipt_big, int big_nact) {
  Decomp d;
  d.ipt = (n_act >= big_nact) ? ipt_big : 1;
  d.ti = 32 * d.ipt;
  d.n_itiles = (n_act + d.ti - 1) / d.ti;
  if (d.n_itiles < 1) d.n_itiles = 1;
  const int small = (big_nact + 31) / 32 + 1;
  const int ns = tab[(d.ipt == 1 ? 0 : small) + d.n_itiles - 1];
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
  return (long long)n_loc + 32 * FORCE_IPT_MAX + (long long)(FORCE_MAX_ROUNDS_CAP * grid + 1) * 32 * FORCE_IPT_MAX * 2;
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
double launch_dfma_mufu_mix(int sm_count, int iters, double *scratch, cudaStream_t s);  // + 1 MUFU.RSQ64H per 32 DFMA
double launch_fp64_rate(int variant, int sm_count, int iters, double *scratch, cudaStream_t s);  // returns DP lane-instructions per launch
cudaError_t force_kernel_setup();
int launch_loop(const GravDev &g, int phase, int max_steps, cudaStream_t s, cudaError_t *err);
int launch_loop_dist(const GravDev &g, int mode, int phase, int max_steps, unsigned long long step_id0,
                     cudaStream_t s, cudaError_t *err);
int launch_pull(const GravDev &g, unsigned long long step_id, cudaStream_t s);
int launch_engine(const GravDev &g, int phase, int cs, int p_cap, cudaStream_t s, cudaError_t *err);
int launch_chip(const GravDev &g, int phase, cudaStream_t s, cudaError_t *err);
cudaError_t loop_kernel_setup();
int loop_max_ctas_per_sm(int variant);

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
constexpr int ENR_GRID_MAX = 32;                                           // mode 2: cells per dimension at most (plus a one-cell apron)
constexpr int ENR_GRID_CELLS = (ENR_GRID_MAX + 2) * (ENR_GRID_MAX + 2) * (ENR_GRID_MAX + 2);
constexpr int ENR_PRUNE_MIN_SOURCES = 48;                                  // mode 2: below this many massive stars every pair is tested
constexpr int ENR_NCOUNTERS = 16;
struct EnrichDev {
  int n_tot;      // global star count (classification, source table)
  int d0, n_loc;  // this rank's disc slice [d0, d0 + n_loc)
  // per-star inputs: mass always global length; mdot global length or (sliced upload) unused
  const double *mass_msun, *mdot;
  const double *px, *py, *pz, *pvx, *pvy, *pvz;  // km, km/s  (explicit arrays) or null
  int pv_off;                                    // the explicit arrays start at this global index (0, or d0 when only the slice was uploaded)
  const double4 *gpos, *gvel;                    // gravity snapshot (global), used when px == null
  double km_per_length, kms_per_speed;
  const double *wr26, *wr60, *sn26, *sn60;       // global length
  uint8_t *kicked;                               // global length
  // sharded per-disc state (local length)
  const double *r_disk, *tau_disk;
  uint8_t *alive;
  double *inv, *fin;  // [8][n_loc]
  // scratch
  int *hm_list;       // [ENR_MAX_SOURCES] massive stars, ascending index after the sort
  int *counters;      // [0] massive stars found, [1] n_events, [2] capacity flag, [3] n_hm (sorted list), [4] CTAs done,
                      // [5..7] grid cells per dimension (mode 2), [8] candidate lists built
  double4 *src_a;     // {x, y, z, c26}
  double4 *src_b;     // {c60, sn26, sn60, is_event}
  double4 *src_f;     // fast test: {-2 x', -2 y', -2 z', |s'|^2 - q}, positions relative to fsum[2..4]
  double4 *ev_a;      // this step's supernovae, ascending: {x, y, z, sn26}
  double *ev_b;       //                                     sn60
  double *fsum;       // [0] sum c26, [1] sum c60, [2..4] origin of the fast test, [5..7] grid corner, [8] 1 / cell size
  double *hm_rows;    // sliced upload: [mdot, x, y, z][ENR_MAX_SOURCES] of the listed massive stars, gathered by the host
  int *cell_start;    // mode 2: [ENR_GRID_CELLS + 1] list starts
  int *cell_items;    // [27 * ENR_MAX_SOURCES] per-cell candidate lists (every source sits in 27 of them)
  long long *prof;    // [8] diagnostic: SM cycles of the last CTA of k_enrich_sources per phase (al26_enrich_profile)
  int *sn_events;     // [ENR_MAX_SOURCES]
};
struct EnrichParams {
  double dt_s, t_new_myr;
  double r_local, r_local3, q_local;  // bubble radius, its cube, and the d^2 threshold equivalent to R <= sqrt(d2)
  double r_global3;
  double decay26, decay60;
  int with_agb;
  int mode;  // 0 exact (reference order, every pair), 1 fast (hoisted global sum, 4-DP pair test), 2 fast + cell-grid pruning
};
int launch_enrich(const EnrichDev &e, const EnrichParams &p, int sm_count, bool tables_only, cudaStream_t s, cudaError_t *err);
int launch_enrich_classify(const EnrichDev &e, const EnrichParams &p, int sm_count, cudaStream_t s);
cudaError_t enrich_kernel_setup();

// AGB interloper deposit (al26_nbody.py:985-1028, calc_intersection :1156-1190)
struct InterloperParams {
  int k_int;                 // global index of the interloper
  const double *old_pc, *new_pc;  // [3][n_tot] positions before / after the gravity step, in pc
  double q_test;             // largest d^2 with sqrt(d^2) <= r_test (pc^2)
  double r_bub3;             // interloper_bubble_radius cubed (km^3)
  double km_per_pc, rate26, rate60, dt_s;
  double *raw;               // [2][n_loc] mass_{26al,60fe}_agb_raw
};
int launch_interloper(const EnrichDev &e, const InterloperParams &p, cudaStream_t s);

// post-processing (analysis.cu)
int launch_local_density(int n, const double *x, const double *y, const double *z, const double *m, double *rho,
                         cudaStream_t s);

}  // namespace al26
