/* include/al26_b200.h -- C-ABI of libal26b200.so, the B200 (sm_100a) drop-in for the
 * hot path of jweatson/26al-nbody.
 *
 * The reference has no C-ABI for this path: the boundary is a Python object protocol
 * (AMUSE GravitationalDynamics, driven over RPC to MPI workers) plus in-process numpy /
 * numba code.  Each entry point below names the reference interface it stands in for
 * (file:line in /root/reference).  The Python host shim (26al-nbody_b200/gravity.py,
 * enrichment.py) binds these with ctypes and re-creates the reference-side surface;
 * INTEGRATION.md shows the stub a maintainer adds to al26_nbody.py.
 *
 * Conventions
 *   - every pointer is a HOST pointer to a caller-owned, contiguous array; the library
 *     copies in / out; device memory is library-owned;
 *   - every function returns 0 on success, a negative AL26_E* code on failure, and
 *     never aborts; al26_last_error() gives the message of the last failure;
 *   - one context = one GPU = one host thread at a time; multi-GPU = one process and one
 *     context per GPU, joined by al26_dist_init() (NCCL over NVLink);
 *   - gravity works in N-body units, G = 1; enrichment works in the units of the
 *     reference's numba kernel call (km, km/s, kg/s, s, kg; al26_nbody.py:886-895,904-905);
 *   - there is NO CPU fallback: with no usable CUDA device al26_create() fails.
 */
#ifndef AL26_B200_H
#define AL26_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct al26_ctx al26_ctx;

enum {
  AL26_OK = 0,
  AL26_EINVAL = -1,   /* bad argument (size mismatch, negative parameter, null pointer) */
  AL26_ESTATE = -2,   /* call not legal in the current state (e.g. evolve before commit) */
  AL26_ETIME = -3,    /* t_end earlier than model time */
  AL26_ECAP = -4,     /* caller buffer too small / internal capacity exceeded */
  AL26_ECUDA = -5,    /* CUDA runtime error (message holds the CUDA error string) */
  AL26_ENCCL = -6,    /* NCCL error or NCCL not loadable */
  AL26_ENODEV = -7    /* no sm_100 device */
};

#define AL26_NINV 8 /* inventory rows: local26, global26, sne26, agb26, local60, global60, sne60, agb60
                       (cluster.mass_{26al,60fe}_{local,global,sne,agb}, al26_nbody.py:1556-1577) */

/* ---- context -------------------------------------------------------------------------
 * replaces: the worker constructors `ph4(converter, number_of_workers=workers)` /
 * `Hermite(...)` / `BHTree(...)` (al26_nbody.py:1709-1722) and `gravity.stop()` (:1762). */
al26_ctx *al26_create(int device_id);
void al26_destroy(al26_ctx *ctx);
const char *al26_last_error(al26_ctx *ctx); /* ctx may be NULL: message of a failed al26_create */
int al26_version(void);
int al26_device_count(void); /* usable CUDA devices (0 without a driver / GPU) */
/* device info for the host shim / bench: SM count, clock (kHz), free and total bytes */
int al26_device_info(al26_ctx *ctx, int *sm_count, int *clock_khz, int64_t *free_bytes, int64_t *total_bytes);

/* ---- multi-GPU: i-particles and discs partitioned by contiguous index range;
 * every block step each rank predicts its own j-slice and the slices are all-gathered.
 * `nccl_unique_id` is the 128-byte ncclUniqueId made by rank 0 (al26_dist_unique_id) and
 * broadcast by the host (torch.distributed).  replaces: `number_of_workers=8` MPI ranks of
 * the reference's worker (al26_nbody.py:57,1711-1720). */
int al26_dist_unique_id(void *out128);
int al26_dist_init(al26_ctx *ctx, int rank, int world, const void *nccl_unique_id);
/* how ranks exchange data inside evolve (before commit).  1 (default): peer memory -- every rank holds the
 * full state, corrects the active particles it owns (i % world == rank) and stores them straight into every
 * rank's staging slab over NVLink from inside the persistent loop kernel; one cross-GPU flag barrier per
 * block step, no NCCL in the loop.  0: contiguous i-slices, NCCL all-gather of the predicted j-set + an
 * 8-byte min-reduce per block step. */
int al26_dist_set_mode(al26_ctx *ctx, int mode);
/* peer-memory mode (before commit): block steps with fewer than n_act_min active particles are not exchanged --
 * the state is replicated, so every rank computes them itself, bit-identically, with no NVLink traffic and no
 * cross-GPU barrier; 0 = automatic (the block size at which the force work saved outweighs an exchange, ~2e7/N) */
int al26_dist_set_split_min(al26_ctx *ctx, int n_act_min);
/* peer-memory mode, after every al26_grav_commit: export this rank's slab as a 64-byte CUDA IPC handle, let
 * the host all-gather the handles (torch.distributed), import the world x 64 bytes */
int al26_dist_p2p_export(al26_ctx *ctx, void *out64);
int al26_dist_p2p_import(al26_ctx *ctx, const void *handles, int world);

/* the same peer-memory protocol between contexts of ONE process (what al26_group below does): join without NCCL,
 * then hand every rank the others' slab pointers; peer access is enabled here (cudaDeviceEnablePeerAccess).
 * Energies then return this rank's partial sums (the caller adds them up). */
int al26_dist_init_local(al26_ctx *ctx, int rank, int world);
int al26_dist_p2p_local_slab(al26_ctx *ctx, void **slab);
int al26_dist_p2p_attach(al26_ctx *ctx, void *const *slabs, const int *devices, int world);

/* ---- several GPUs from one process ------------------------------------------------------
 * replaces: `ph4(converter, number_of_workers=workers)` (al26_nbody.py:57,1711-1720) -- ONE script process, `workers`
 * worker ranks behind one object.  A group owns n_gpus contexts, one per GPU, each on its own host thread; every
 * call below fans out to all of them and joins (the ranks' kernels must run at the same time: they exchange
 * corrected particles and barrier flags over NVLink).  Arguments as for the single-context calls; host arrays are
 * global length.  device_ids may be NULL (GPUs 0..n_gpus-1).  Any n (a rank owns the particles i % n_gpus == rank). */
typedef struct al26_group al26_group;
al26_group *al26_group_create(int n_gpus, const int *device_ids);
void al26_group_destroy(al26_group *grp);
const char *al26_group_last_error(al26_group *grp); /* grp may be NULL: message of a failed al26_group_create */
int al26_group_size(al26_group *grp);
al26_ctx *al26_group_ctx(al26_group *grp, int rank); /* tuning hooks / diagnostics of one rank (before commit) */
int al26_group_grav_set_params(al26_group *grp, double eps2, double eta, double dt_max, double dt_min);
int al26_group_grav_set_reinit_policy(al26_group *grp, int policy);
int al26_group_grav_commit(al26_group *grp, int64_t n, const double *m, const double *x, const double *y, const double *z,
                           const double *vx, const double *vy, const double *vz);
int al26_group_grav_set_mass(al26_group *grp, int64_t n, const double *m);
int al26_group_grav_set_time(al26_group *grp, double t);
int al26_group_grav_get_time(al26_group *grp, double *t);
/* n_block_steps: block steps of the call; n_pairs: pair evaluations summed over the GPUs */
int al26_group_grav_evolve(al26_group *grp, double t_end, int64_t *n_block_steps, int64_t *n_pairs);
int al26_group_grav_get_state(al26_group *grp, int64_t n, double *m, double *x, double *y, double *z, double *vx,
                              double *vy, double *vz);
int al26_group_grav_energies(al26_group *grp, double *kinetic, double *potential, double *sum_mm_over_r);
/* device time of the last evolve / enrich_step: the slowest GPU's; kernel launches: all GPUs' */
int al26_group_last_device_ms(al26_group *grp, double *ms, int64_t *kernel_launches);
int al26_group_enrich_commit(al26_group *grp, int64_t n, const double *r_disk_km, const double *tau_disk_myr,
                             const uint8_t *disk_alive, const uint8_t *kicked, const double *wr26, const double *wr60,
                             const double *sn26_kg, const double *sn60_kg);
int al26_group_enrich_set_units(al26_group *grp, double km_per_length, double kms_per_speed);
int al26_group_enrich_set_mode(al26_group *grp, int mode);
int al26_group_enrich_set_inventories(al26_group *grp, int64_t n, const double *inv, const double *fin);
int al26_group_enrich_step(al26_group *grp, int64_t n, const double *mass_msun, const double *mdot_kg_s,
                           const double *pos_vel, double dt_s, double t_new_myr, double r_bub_local_km,
                           double r_bub_global_km, double decay26, double decay60, int with_agb, int32_t *sn_events,
                           int64_t sn_cap, int64_t *n_sn_events);
int al26_group_enrich_get(al26_group *grp, int64_t n, double *inv, double *fin, uint8_t *disk_alive, uint8_t *kicked);

/* diagnostic, peer-memory mode: where a rank's time goes, in SM cycles of CTA 0 since the last commit (12 values):
 * [0] fused redundant steps, [1] their number, [2] other redundant steps, [3] their number, [4] their active
 * particles; exchanged steps: [5] predictor + scheduler + barrier, [6] force on the own share + barrier,
 * [7] corrector + peer stores, [8] cross-GPU barrier, [9] their number, [10] their active particles */
int al26_dist_profile(al26_ctx *ctx, int64_t *cycles12);

/* ---- gravity -------------------------------------------------------------------------*/
/* replaces: `gravity.parameters.epsilon_squared / timestep_parameter` (never set by the
 * script; defaults eps2 = 0, eta = 0.14).  dt_max / dt_min are rounded down to powers of 2. */
int al26_grav_set_params(al26_ctx *ctx, double eps2, double eta, double dt_max, double dt_min);
/* replaces: `gravity.particles.add_particles(cluster)` (al26_nbody.py:1728).  n is the GLOBAL
 * particle count; index order is preserved (the script checks keys by index, :781-783). */
int al26_grav_commit(al26_ctx *ctx, int64_t n, const double *m, const double *x, const double *y,
                     const double *z, const double *vx, const double *vy, const double *vz);
/* replaces: `stel_to_grav.copy_attributes(["mass"])` (al26_nbody.py:871,874): legal between
 * evolve calls; forces and timesteps are re-initialised at the next evolve. */
int al26_grav_set_mass(al26_ctx *ctx, int64_t n, const double *m);
/* what such a mass-only update costs at the next evolve (every particle is synchronised at model time then):
 * 0 (default): acc / jerk / pot are recomputed with the new masses, every particle keeps the timestep the
 *    synchronisation step gave it -- ph4's recommit_particles ("recompute forces ... we don't recompute the time
 *    steps" [upstream amuse_ph4 interface.cc, from memory]) behind the AMUSE state machine that the script's
 *    per-step mass channel triggers (al26_nbody.py:871,874);
 * 1: forces AND initial timesteps eta/16 |a|/|jerk| again, as after a commit (the round-1 behaviour; costs the
 *    block steps every particle then needs to climb back up the ladder). */
int al26_grav_set_reinit_policy(al26_ctx *ctx, int policy);
/* replaces: `gravity.model_time` setter / getter (al26_nbody.py:763,1101,1736). */
int al26_grav_set_time(al26_ctx *ctx, double t);
int al26_grav_get_time(al26_ctx *ctx, double *t);
/* replaces: `gravity.evolve_model(t_new)` (al26_nbody.py:833).  On return every particle is
 * synchronised at t_end.  n_block_steps / n_pairs (may be NULL) count this call's block steps
 * and (i,j) pair evaluations on this rank. */
int al26_grav_evolve(al26_ctx *ctx, double t_end, int64_t *n_block_steps, int64_t *n_pairs);
/* replaces: the bulk getters `gravity.particles.{mass,x,y,z,vx,vy,vz}`, `.copy()` and the
 * channel `grav_to_clus.copy()` (al26_nbody.py:831,872,876,886-891). n = global count. */
int al26_grav_get_state(al26_ctx *ctx, int64_t n, double *m, double *x, double *y, double *z,
                        double *vx, double *vy, double *vz);
/* replaces: `gravity.kinetic_energy`, `gravity.potential_energy` (plotting/al26_plot.py:288-289)
 * and `cluster.virial_radius()` (al26_nbody.py:770): R_vir = M^2 / (2 * sum_mm_over_r), G = 1.
 * sum_mm_over_r is the UNSOFTENED sum over pairs i<j of m_i m_j / r_ij. */
int al26_grav_energies(al26_ctx *ctx, double *kinetic, double *potential, double *sum_mm_over_r);

/* parity hooks (no reference counterpart; used by tests/ against oracle/) */
int al26_grav_initialize(al26_ctx *ctx); /* forces + initial timesteps now, without advancing */
int al26_grav_get_acc_jerk(al26_ctx *ctx, int64_t n, double *ax, double *ay, double *az, double *jx,
                           double *jy, double *jz, double *pot);
int al26_grav_get_timesteps(al26_ctx *ctx, int64_t n, double *t, double *dt);
int al26_grav_set_timesteps(al26_ctx *ctx, int64_t n, const double *t, const double *dt);
/* active set of the next block step, ascending index order (bit-exact), and its time */
int al26_grav_get_active(al26_ctx *ctx, int64_t cap, int32_t *idx, int64_t *n_active, double *tau_next);
/* the active list the DEVICE scheduler built for the block step executed last inside
 * dbg_begin/dbg_finish (ballot compaction in the predict kernel), sorted ascending */
int al26_grav_get_last_active(al26_ctx *ctx, int64_t cap, int32_t *idx, int64_t *n_active);
/* stepwise evolve: begin(t_end); advance(max_steps) any number of times; finish() */
int al26_grav_dbg_begin(al26_ctx *ctx, double t_end);
int al26_grav_dbg_advance(al26_ctx *ctx, int64_t max_steps, int64_t *n_done, int *finished);
int al26_grav_dbg_finish(al26_ctx *ctx);
/* diagnostic (between dbg_begin and dbg_finish, one GPU): run up to `reps` block steps as individual launches
 * with CUDA events around each kernel; us3 = mean microseconds of predict+schedule, force, correct */
int al26_grav_dbg_profile_steps(al26_ctx *ctx, int reps, double *us3);
/* one force evaluation (K1) on caller arrays: acc, jerk, pot on the n_act listed particles */
int al26_grav_force(al26_ctx *ctx, int64_t n, double eps2, const double *m, const double *x, const double *y,
                    const double *z, const double *vx, const double *vy, const double *vz, int64_t n_act,
                    const int32_t *idx, double *ax, double *ay, double *az, double *jx, double *jy,
                    double *jz, double *pot);
/* timing hook for bench.py: device time (ms) of the last al26_grav_evolve / al26_enrich_step,
 * measured with CUDA events on the library's own stream, and kernels launched in it */
int al26_last_device_ms(al26_ctx *ctx, double *ms, int64_t *kernel_launches);
/* bench hook: device time (ms) of the three enrichment kernels alone in the last al26_enrich_step
 * (CUDA events around the kernels; host<->device copies excluded) */
int al26_enrich_last_kernel_ms(al26_ctx *ctx, double *ms);
/* bench hook: time `reps` full force evaluations (all i x all j, K1 only) on the committed
 * state with CUDA events on the library stream; returns average ms per evaluation */
int al26_grav_bench_force(al26_ctx *ctx, int reps, double *avg_ms, int64_t *pairs_per_eval);
/* same with only the first n_act (<= 0: all) local particles active: the small-block regime */
int al26_grav_bench_force_n(al26_ctx *ctx, int64_t n_act, int reps, double *avg_ms, int64_t *pairs_per_eval);

/* host-only diagnostic (needs no GPU): the force kernel's work decomposition for a block of n_act active
 * particles among n_tot: out8 = {i per lane, i per item, i-tiles, j-chunks, j per chunk, slot stride,
 * partial-buffer capacity, CTAs} */
int al26_dbg_decomposition(int n_act, int n_tot, int sm_count, int variant, int big_nact, int64_t *out8);
/* tuning hook: pick one of the compiled force-kernel configurations (0 = default); applies at the
 * next al26_grav_commit */
int al26_set_force_variant(al26_ctx *ctx, int variant);
/* tuning hook: active blocks of at least n_act_min particles use the configuration's several i-particles
 * per lane (default 2048); applies at the next al26_grav_commit */
int al26_set_big_block(al26_ctx *ctx, int n_act_min);
/* tuning hook: parameters of the force kernel's work decomposition (al26_internal.cuh: choose_jsplit): at most
 * max_rounds work items per CTA (default 32) and the fixed cost of one item in pair units (default 3200);
 * applies at the next al26_grav_commit */
int al26_set_decomposition(al26_ctx *ctx, int max_rounds, double item_overhead_pairs);
/* tuning hook: how block steps are driven on one GPU.  -1 (default): automatic -- 2 when the particles fit one
 * cluster (N <= ~13000), else 0.  0: a CUDA graph of three kernels per block
 * step, relaunched until the device reports the call done; 1: one persistent cooperative kernel runs the
 * whole predict -> force -> correct loop with grid barriers (the form the multi-GPU peer-memory mode uses).
 * Bit-identical results (except the loop kernel's fused small steps, al26_set_fuse_max: identical integer work,
 * positions to rounding); measured on B200 the graph is 3 % (N=1e5) faster, the loop with fused steps 3-7 % faster at
 * N = 3e3..1e4, per block step. */
int al26_set_step_mode(al26_ctx *ctx, int mode);
/* step mode 2 = the graph with the CLUSTER ENGINE in front of every block step: one thread-block cluster (8 or 16
 * CTAs, one per SM) keeps the whole particle set in distributed shared memory and takes every run of small block
 * steps (<= 32 active particles) on chip -- DSMEM hops and hardware cluster barriers (~0.1-0.2 us) instead of L2
 * round trips, atomics and fences (~0.4 us each) -- writing corrected particles through to the global records; a
 * bigger block is left to the grid-wide kernels that follow in the graph.  Applies when the particles fit one
 * cluster (N <= ~13000 on B200; otherwise the mode behaves like 0).  Identical integer work, positions to rounding;
 * measured 1.15-1.4x faster per block step than mode 0 at N = 1e3..1e4.
 * diagnostic: block steps the engine took since the last commit, and its cluster size (0 = engine not in use;
 * cluster_size may be NULL) */
int al26_grav_engine_steps(al26_ctx *ctx, int64_t *n_engine, int *cluster_size);
/* host-only diagnostic (needs no GPU): the cluster the engine would use for n particles when a block may opt in to
 * max_smem_per_block bytes of shared memory (232448 on B200) -- cluster size (8 or 16; 0 = the particles do not fit),
 * particles per CTA, shared memory per CTA */
int al26_dbg_engine_plan(int n, int max_smem_per_block, int *cluster_size, int *particles_per_cta, int *smem_bytes);
/* step mode 3 = the graph with the CHIP ENGINE in front of every block step, for the particle sets that do not fit one
 * cluster (up to N ~ 1.1e5 on B200: 144 + 64 B per particle in the shared memory of 148 SMs; on request only on one GPU,
 * automatic in the peer-memory multi-GPU mode):
 * one CTA per SM keeps a contiguous chunk of the particles resident for a whole run of small block steps; the CTAs
 * talk through single-writer records in L2 that carry their step number (no grid barrier, no atomics): the chunk's
 * min(t + dt) with the particles that attain it already predicted, one force partial per (active particle, chunk),
 * one "partials stored" flag per CTA.  A step is 5-6 dependent L2 hops instead of 10-18 and never passes over the
 * state in L2.  In the peer-memory multi-GPU mode (al26_dist_set_mode 1) the same kernel takes every run of small block
 * steps -- which every rank computes redundantly -- between the launches of the loop kernel.  Identical integer work,
 * positions to rounding.  al26_set_chip_max: largest block the engine steps (1..256; -1 = default 32; 0 = engine
 * off), before or after commit; al26_grav_chip_steps: block steps it took since the last commit, its CTAs (0 = not in
 * use) and its block limit (the last two may be NULL).  Stands in, like every step mode, for ph4's evolve loop behind
 * gravity.evolve_model (al26_nbody.py:833). */
int al26_set_chip_max(al26_ctx *ctx, int n_act_max);
int al26_grav_chip_steps(al26_ctx *ctx, int64_t *n_chip, int *n_ctas, int *n_act_max);
/* host-only diagnostic (needs no GPU): the chip engine's layout for n particles on n_ctas CTAs (one per SM) when a block
 * may opt in to max_smem_per_block bytes -- particles per CTA (0 = the chunks do not fit), shared memory per CTA, and the
 * bytes of the mail + partial-row buffers in HBM */
int al26_dbg_chip_plan(int n, int n_ctas, int max_smem_per_block, int *particles_per_cta, int *smem_bytes, int64_t *mail_bytes);
/* tuning hook of the persistent loop kernels (step mode 1 and the peer-memory multi-GPU mode), before commit:
 * block steps of at most n_act_max active particles (0..32; 0 = off; -1 = default: 32 when N <= 32768,
 * else off) take the fused small-step path
 * -- one grid barrier instead of three, force from the CTA's own shared-memory chunk, one corrector CTA per slot.
 * Available when every CTA's share of the particles fits its stage buffers (N <= ~2.2e5 on B200). */
int al26_set_fuse_max(al26_ctx *ctx, int n_act_max);
/* diagnostic: block steps taken through the fused path since the last commit */
int al26_grav_fused_steps(al26_ctx *ctx, int64_t *n_fused);
/* diagnostic, non-zero only in a library built with -DAL26_FUSE_TIMING: nanoseconds (globaltimer) CTA 0 spent, summed
 * over the fused steps since the last commit, in [0] force + partial store + arrival, [1] waiting for all partials,
 * [2] reduce + corrector, [4] waiting for the release counter, [5] the whole fused part; over all loop steps:
 * [6] scan, [7] barrier (16 values) */
int al26_grav_fuse_profile(al26_ctx *ctx, int64_t *ns16);
/* diagnostic: number of block steps by floor(log2(n_active)) since the last commit (32 bins) */
int al26_grav_block_histogram(al26_ctx *ctx, int64_t *hist32);
/* diagnostic: SM cycles CTA 0 of the persistent loop kernel spent in predict / barrier / force / barrier /
 * correct / barrier since the last commit (6 values) */
int al26_grav_loop_profile(al26_ctx *ctx, int64_t *cycles6);
/* bench hook: measured FP64 FMA throughput (TFLOP/s) of a DFMA-only microkernel on this GPU:
 * the roofline denominator of the force kernel */
int al26_bench_fp64_peak(al26_ctx *ctx, double *tflops);
/* the FP64 issue rate (lane-instructions per second; nominal 148 SMs x 64 lanes x clock) by operand pattern:
 * 0 x=fma(x,a,b) with a, b shared by the chains (the roofline denominator above); 1 x=fma(x,imm,b); 2 x=x+b (DADD);
 * 3 x=x*imm (DMUL); 4 x=fma(x,x,b); 5 x=fma(x,a_k,b_k), three distinct registers per instruction */
int al26_bench_fp64_rate(al26_ctx *ctx, int variant, double *lane_inst_per_s);
/* the same DFMA chains with one independent MUFU.RSQ64H per 32 DFMAs (the force kernel's ratio): the DFMA TFLOP/s
 * that remain -- i.e. whether the 64-bit reciprocal square root takes FP64 issue slots */
int al26_bench_fp64_with_rsqrt(al26_ctx *ctx, double *tflops);

/* ---- post-processing ------------------------------------------------------------------
 * replaces: `local_densities_numba(x, y, z, masses)` (plotting/al26_plot.py:324-359, called by
 * calc_local_densities :361-371): 10-nearest-neighbour local density of every star, rho = sum of the ten nearest
 * masses / (4.18879020479 d10^3), positions in pc, masses in Msun -> Msun/pc^3.  O(N) memory instead of the
 * reference's N x N matrix; bit-identical results.  n >= 11. */
int al26_local_densities(al26_ctx *ctx, int64_t n, const double *x_pc, const double *y_pc, const double *z_pc,
                         const double *mass_msun, double *rho);

/* ---- enrichment (state lives on the device between calls) ---------------------------*/
/* replaces: the per-star cluster attributes set in init_cluster (al26_nbody.py:1543-1603):
 * r_disk [km], tau_disk [Myr], disk_alive, kicked, wind_ratio_26al/60fe, sn_yield_26al/60fe [kg].
 * n = global star count.  Inventories and finals start at zero (:1556-1577). */
int al26_enrich_commit(al26_ctx *ctx, int64_t n, const double *r_disk_km, const double *tau_disk_myr,
                       const uint8_t *disk_alive, const uint8_t *kicked, const double *wr26,
                       const double *wr60, const double *sn26_kg, const double *sn60_kg);
/* overwrite inventories / finals (checkpoint resume, al26_nbody.py:1641-1656); either may be NULL */
int al26_enrich_set_inventories(al26_ctx *ctx, int64_t n, const double *inv /*[8][n]*/, const double *fin /*[8][n]*/);
/* conversion applied when al26_enrich_step is given pos_vel == NULL and reads the gravity state
 * in place: x_km = x_nbody * km_per_length, v_kms = v_nbody * kms_per_speed
 * (replaces `.value_in(units.km)` / `.value_in(units.km/units.s)`, al26_nbody.py:886-891) */
int al26_enrich_set_units(al26_ctx *ctx, double km_per_length, double kms_per_speed);
/* how the disc kernel treats the (disc, massive star) pairs -- no reference counterpart, a knob of this library:
 *   0 (default) exact: every pair in the reference's order; wind sums bit-identical to the reference's numba kernel;
 *   1 fast: the north_star's tolerance mode (per-disc masses within 1e-10): the separable global-bubble source sum is
 *     hoisted (relative difference ~1e-13), the local-bubble test runs in expanded form, 3 DFMA + 1 compare per pair;
 *     local and SN deposits stay bit-identical unless a pair sits within ~1e-12 (relative) of the bubble surface;
 *   2 fast + pruned: as 1, with the local-bubble candidates taken from a cell grid over the massive stars and tested
 *     in the exact form (local rows bit-identical to mode 0); no pair loop, HBM-bound at any source count. */
int al26_enrich_set_mode(al26_ctx *ctx, int mode);
/* diagnostic: SM cycles the table-building CTA of the last al26_enrich_step spent in [0] the sort, [1] the source table +
 * sums + event list, and (mode 2) [2] clearing the cell array, [3] counting, [4] the scan + start table, [5] the scatter */
int al26_enrich_profile(al26_ctx *ctx, int64_t *cycles8);
/* replaces one pass of al26_nbody.py:878-1086 (interloper block excluded):
 *   classify (:1194-1216) on mass_msun; 4x calc_wind_abs (:642-702, :897-938) with local bubble
 *   r_bub_local_km (distance-tested) and global bubble r_bub_global_km (= virial radius, no
 *   test); SN events (mdot == 0 and not kicked, :946-967) with calc_eta_disk_sne (:1326-1334);
 *   decay by decay26 / decay60 (host computes exp() as :1050-1051); condense (:1071-1086).
 * mass_msun: the masses the reference classifies on (cluster.mass before this step's stellar
 * evolve, :767).  mdot: -wind_mass_loss_rate in kg/s (:892).  pos_vel: [6][n] x,y,z (km),
 * vx,vy,vz (km/s), or NULL to use the gravity state.  with_agb != 0 also decays / condenses the
 * agb rows (:1062-1064,:1080-1082).  sn_events receives the ascending indices of this step's
 * supernovae (capacity sn_cap). */
int al26_enrich_step(al26_ctx *ctx, int64_t n, const double *mass_msun, const double *mdot_kg_s,
                     const double *pos_vel, double dt_s, double t_new_myr, double r_bub_local_km,
                     double r_bub_global_km, double decay26, double decay60, int with_agb,
                     int32_t *sn_events, int64_t sn_cap, int64_t *n_sn_events);
/* replaces the AGB interloper deposit of al26_nbody.py:985-1028 (optional `-i` flag): per disc
 * (0.1 <= mass_msun <= 3, not the interloper itself) the fraction of the step spent within r_test_pc of the
 * interloper by calc_intersection's recipe (:1156-1190: 1024 np.linspace samples of both straight-line paths,
 * positions in pc before / after the gravity step, [3][n] each), then
 * rate * 0.75 r_disk^2 (|dx_disc| frac) / r_bub^3 * dt onto the agb rows and the never-decayed agb_raw
 * accumulators.  The caller has already checked interloper_time > 0 and rate > 0 (:985,:989).  Call it BEFORE
 * al26_enrich_step(..., with_agb = 1) of the same outer step (the reference deposits, then decays). */
int al26_enrich_interloper(al26_ctx *ctx, int64_t n, const double *mass_msun, const double *pos_old_pc,
                           const double *pos_new_pc, int64_t interloper_index, double r_test_pc, double r_bub_km,
                           double km_per_pc, double rate26_kg_s, double rate60_kg_s, double dt_s);
/* cluster.mass_{26al,60fe}_agb_raw (al26_nbody.py:1560,1572): raw[2][n] */
int al26_enrich_get_agb_raw(al26_ctx *ctx, int64_t n, double *raw);
/* replaces: reading cluster.mass_*_{local,global,sne,agb}[_final], disk_alive, kicked when
 * yields / checkpoints are written (al26_nbody.py:1097-1105).  Any pointer may be NULL. */
int al26_enrich_get(al26_ctx *ctx, int64_t n, double *inv /*[8][n]*/, double *fin /*[8][n]*/,
                    uint8_t *disk_alive, uint8_t *kicked);

#ifdef __cplusplus
}
#endif
#endif /* AL26_B200_H */
