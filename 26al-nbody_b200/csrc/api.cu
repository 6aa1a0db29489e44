// api.cu -- the C-ABI of libal26b200.so (include/al26_b200.h): context, memory, CUDA-graph block
// stepping, NCCL plumbing (dlopen'ed, so the library loads on a box without NCCL or a GPU),
// host <-> device marshalling.  No CPU fallback anywhere: without a CUDA device al26_create fails.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>  // types only; every NCCL symbol is resolved with dlsym

#include <algorithm>
#include <condition_variable>
#include <cstdarg>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/al26_b200.h"
#include "al26_internal.cuh"

using namespace al26;

// ------------------------------------------------------------------------------------------
// NCCL through dlopen
// ------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
std::string g_create_error;

bool load_nccl(std::string &why) {
  if (g_nccl.ok) return true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) {
    why = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
    return false;
  }
#define L(sym)                                                              \
  *(void **)(&g_nccl.sym) = dlsym(g_nccl.h, "nccl" #sym);                   \
  if (!g_nccl.sym) {                                                        \
    why = "libnccl lacks nccl" #sym;                                        \
    return false;                                                           \
  }
  L(GetUniqueId) L(CommInitRank) L(CommDestroy) L(AllGather) L(AllReduce) L(GroupStart) L(GroupEnd) L(GetErrorString)
#undef L
  g_nccl.ok = true;
  return true;
}
}  // namespace

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct al26_ctx {
  int device = 0;
  int sm_count = 0;
  int clock_khz = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  double last_kernel_ms = 0.0;  // kernels only (no host<->device copies) of the last al26_enrich_step
  std::string err;
  double last_ms = 0.0;
  int64_t last_launches = 0;
  int64_t launches = 0;  // running counter of kernel launches (graph nodes included)

  // dist
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  int dist_mode = 1;             // world > 1: 1 = peer-memory (NVLink) exchange inside the loop kernel, 0 = NCCL graph path
  void *slab = nullptr;          // own staging slab (peer-memory mode)
  bool p2p_ready = false;        // peers' slabs imported
  bool p2p_ipc = false;          // ... as CUDA IPC mappings (one process per GPU); false: plain peer pointers (al26_group, one process)
  bool local_group = false;      // joined by al26_dist_init_local: no NCCL communicator, energies return this rank's partial sums
  unsigned long long dist_step = 0;
  int split_min = 0;             // peer-memory mode: exchange only block steps with at least this many active particles (0 = auto)

  // gravity
  GravDev g{};
  bool committed = false, in_evolve = false;
  // 0 = forces and timesteps valid; 1 = masses changed (forces stale, timesteps still those of the last
  // synchronisation); 2 = nothing valid yet (after commit / set_params)
  int dirty = 2;
  int reinit_policy = 0;  // what a mass-only update costs: 0 = forces only, timesteps kept; 1 = forces + initial timesteps
  double t_model = 0.0, t_end_pending = 0.0;
  double eps2 = 0.0, eta = 0.14, dt_max = 0.125, dt_min = 9.094947017729282e-13 /* 2^-40 */;
  int64_t n_tot = 0;
  int dbg_phase = 0;
  int force_variant = 0;
  int big_nact = FORCE_BIG_NACT_DEFAULT;  // tuning: block size from which the force kernel holds several i per lane
  int max_rounds = FORCE_MAX_ROUNDS;      // tuning: work items per CTA at most
  int fuse_max = -1;                      // loop kernels: block steps of at most this many particles take the fused path (0 = off, -1 = by N)
  double item_overhead = FORCE_ITEM_OVERHEAD_PAIRS;  // tuning: fixed cost of a work item, in pair units
  int step_mode = -1;     // 1 GPU: -1 = automatic (2 when the particles fit one cluster, else 0), 0 = CUDA graph of 3 kernels per block step, 1 = persistent cooperative loop kernel, 2 = graph + cluster engine, 3 = graph + chip engine
  bool coop_ok = false;   // device supports cooperative launch
  int max_smem_optin = 0;  // largest dynamic shared memory a block may opt in to
  bool engine_ok = false;  // the cluster-engine kernel's attributes could be set
  bool engine_on = false;  // this commit's graph carries the engine (step mode 2 and the particles fit one cluster)
  int engine_cs = 0, engine_p = 0;  // its cluster size and per-CTA particle capacity
  bool chip_ok = false;    // the chip-engine kernel's attributes could be set
  bool chip_on = false;    // this commit runs its small block steps in the chip engine (hermite_chip.cu)
  int chip_max = -1;       // largest block the chip engine steps (-1 = default, 0 = engine off)
  cudaGraphExec_t graph = nullptr;
  int graph_steps = 0;
  bool graph_stale = false;  // parameters changed since the graph captured them by value
  int expected_graph_launches = 1;
  GravHeader *h_hdr = nullptr;  // pinned
  double *scratch = nullptr;    // device staging, >= 16 * n_tot doubles
  size_t scratch_doubles = 0;
  double *en_scratch = nullptr;
  size_t en_scratch_doubles = 0;
  double *en_out = nullptr;  // device [3]
  double *h_small = nullptr; // pinned [16]

  // enrichment
  EnrichDev e{};
  bool e_committed = false;
  double *e_glob = nullptr;     // device: mass, mdot, px..pvz (8 n), wr26, wr60, sn26, sn60 (4 n)
  double *e_loc = nullptr;      // device: r_disk, tau (2 nloc), inv (8 nloc), fin (8 nloc)
  uint8_t *e_flags = nullptr;   // device: kicked (n), alive (nloc)
  int *e_ints = nullptr;        // device: counters, hm_list, sn_events, cell_items, cell_start + cursors
  double4 *e_src = nullptr;     // device: src_a, src_b, src_f, ev_a
  double *e_dbl = nullptr;      // device: ev_b, fsum, hm_rows
  int *h_events = nullptr;      // pinned [ENR_NCOUNTERS + ENR_MAX_SOURCES]
  double *h_rows = nullptr;     // pinned [4][ENR_MAX_SOURCES]: mdot, x, y, z of the massive stars (sliced upload)
  int enrich_mode = 0;          // 0 exact, 1 fast, 2 fast + pruned (al26_enrich_set_mode)
  double km_per_length = 1.0, kms_per_speed = 1.0;
};

namespace {

inline bool is_p2p(const al26_ctx *c) { return c->world > 1 && c->dist_mode == 1; }

int fail(al26_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  else g_create_error = buf;
  return code;
}

#define CU(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t _e = (call);                                                                           \
    if (_e != cudaSuccess) return fail(c, AL26_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define NC(call)                                                                                       \
  do {                                                                                                 \
    ncclResult_t _r = (call);                                                                          \
    if (_r != ncclSuccess) return fail(c, AL26_ENCCL, "%s failed: %s", #call, g_nccl.GetErrorString(_r)); \
  } while (0)

double pow2floor_h(double x) {
  int e;
  frexp(x, &e);
  return ldexp(1.0, e - 1);
}

// ---- small marshalling kernels ----
__global__ void k_pack(int n, const double *m, const double *x, const double *y, const double *z, const double *vx,
                       const double *vy, const double *vz, double4 *pos, double4 *vel) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    pos[i] = make_double4(x[i], y[i], z[i], m[i]);
    vel[i] = make_double4(vx[i], vy[i], vz[i], 0.0);
  }
}
__global__ void k_unpack(int n, const double4 *pos, const double4 *vel, double *m, double *x, double *y, double *z,
                         double *vx, double *vy, double *vz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double4 p = pos[i], v = vel[i];
    x[i] = p.x; y[i] = p.y; z[i] = p.z; m[i] = p.w;
    vx[i] = v.x; vy[i] = v.y; vz[i] = v.z;
  }
}
__global__ void k_unpack_aj(int n, const double4 *acc, const double4 *jrk, double *out /*[7][n]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double4 a = acc[i], j = jrk[i];
    out[0 * (size_t)n + i] = a.x; out[1 * (size_t)n + i] = a.y; out[2 * (size_t)n + i] = a.z;
    out[3 * (size_t)n + i] = j.x; out[4 * (size_t)n + i] = j.y; out[5 * (size_t)n + i] = j.z;
    out[6 * (size_t)n + i] = a.w;
  }
}
__global__ void k_set_mass(int n, const double *m, double4 *pos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pos[i].w = m[i];
}
__global__ void k_reset_ctrl(StepCtrl *ctrl, GravHeader *hdr, int zero_counters) {
  const int i = threadIdx.x;
  if (i < 3) {
    ctrl[i].t_next_bits = INF_BITS;
    ctrl[i].n_act = 0;
    ctrl[i].work_counter = 0;
    ctrl[i].pad[0] = 0;
    ctrl[i].pad[1] = 0;
  }
  if (i == 0 && zero_counters) {
    hdr->n_steps = 0;
    hdr->n_pairs = 0;
    hdr->done = 0;
  }
}
__global__ void k_reset_work(StepCtrl *ctrl) { ctrl->work_counter = 0; }
__global__ void k_loop_prepare(GravHeader *hdr) {
  if (threadIdx.x == 0) {
    hdr->bar_counter = 0u;
    hdr->loop_error = 0;
  }
}
__global__ void k_set_nact(StepCtrl *ctrl, int n_act) {
  ctrl->n_act = n_act;
  ctrl->work_counter = 0;
}
__global__ void k_set_list(int n_act, const int *idx, int *list, StepCtrl *ctrl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_act) list[i] = idx[i];
  if (i == 0) {
    ctrl->n_act = n_act;
    ctrl->work_counter = 0;
  }
}
// scheduler parity hook: min(t+dt), then the list of particles that attain it
__global__ void k_min_tnext(int n, const double *t, const double *dt, unsigned long long *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v = INF_BITS;
  if (i < n) v = (unsigned long long)__double_as_longlong(t[i] + dt[i]);
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  if ((threadIdx.x & 31) == 0 && v != INF_BITS) atomicMin(out, v);
}
__global__ void k_list_at(int n, const double *t, const double *dt, const unsigned long long *tn, int *list, int *cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (unsigned long long)__double_as_longlong(t[i] + dt[i]) == *tn) list[atomicAdd(cnt, 1)] = i;
}

inline int nblk(int64_t n, int t = 256) { return (int)((n + t - 1) / t); }

int ensure_scratch(al26_ctx *c, size_t doubles) {
  if (c->scratch_doubles >= doubles) return 0;
  if (c->scratch) cudaFree(c->scratch);
  c->scratch = nullptr;
  c->scratch_doubles = 0;
  CU(cudaMalloc(&c->scratch, doubles * sizeof(double)));
  c->scratch_doubles = doubles;
  return 0;
}

void free_gravity(al26_ctx *c) {
  if (c->graph) cudaGraphExecDestroy(c->graph);
  c->graph = nullptr;
  GravDev &g = c->g;
  for (int q = 0; q < MAX_PEERS; q++)
    if (g.slab[q] && q != c->rank && c->p2p_ready && c->p2p_ipc) cudaIpcCloseMemHandle(g.slab[q]);
  if (c->slab) cudaFree(c->slab);
  c->slab = nullptr;
  c->p2p_ready = false;
  void *ptrs[] = {g.pos, g.vel, g.acc, g.jrk, g.t, g.dt, g.jpos, g.jvel, g.list, g.part_a, g.part_j, g.ctrl, g.hdr,
                  (void *)g.decomp_tab, g.list_own, g.act, g.chip_mail};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  g = GravDev{};
  c->committed = false;
}

void free_enrich(al26_ctx *c) {
  void *ptrs[] = {c->e_glob, c->e_loc, c->e_flags, c->e_ints, c->e_src, c->e_dbl};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  c->e_glob = c->e_loc = c->e_dbl = nullptr;
  c->e_flags = nullptr;
  c->e_ints = nullptr;
  c->e_src = nullptr;
  c->e = EnrichDev{};
  c->e_committed = false;
}

void slice_of(int64_t n, int rank, int world, int64_t &i0, int64_t &nloc) {
  i0 = n * rank / world;
  const int64_t i1 = n * (rank + 1) / world;
  nloc = i1 - i0;
}

// NCCL mode: all-gather the predicted / snapshot j-set, one in-place ncclAllGather per array;
// slices are equal (n % world == 0 is enforced at commit).  No-op on one GPU and in peer-memory mode.
int gather_j(al26_ctx *c) {
  if (c->world == 1 || is_p2p(c)) return 0;  // peer-memory mode: the state is replicated, nothing to gather
  const GravDev &g = c->g;
  const size_t cnt = (size_t)g.n_loc * 4;  // doubles per rank
  NC(g_nccl.GroupStart());
  NC(g_nccl.AllGather(g.jpos + g.i0, g.jpos, cnt, ncclDouble, c->comm, c->stream));
  NC(g_nccl.AllGather(g.jvel + g.i0, g.jvel, cnt, ncclDouble, c->comm, c->stream));
  NC(g_nccl.GroupEnd());
  return 0;
}
// global minimum of the next block time (one 8-byte all-reduce); no-op on one GPU
int reduce_tnext(al26_ctx *c, int phase) {
  if (c->world == 1 || is_p2p(c)) return 0;
  unsigned long long *p = &c->g.ctrl[phase].t_next_bits;
  NC(g_nccl.AllReduce(p, p, 1, ncclUint64, ncclMin, c->comm, c->stream));
  return 0;
}

// step mode 2: does the particle set fit one cluster?  8 CTAs when that is enough, else 16.
void decide_engine(al26_ctx *c) {
  c->engine_on = false;
  if ((c->step_mode != 2 && c->step_mode != -1) || c->world != 1 || !c->engine_ok || !c->committed) return;
  int cs = 0, p_cap = 0;
  if (!engine_plan(c->g.n_tot, c->max_smem_optin, &cs, &p_cap)) return;
  if (!engine_fits(cs, p_cap, c->max_smem_optin)) return;  // the driver's say on co-residency of the cluster
  c->engine_on = true;
  c->engine_cs = cs;
  c->engine_p = p_cap;
}

// The chip engine (hermite_chip.cu) for the particle sets that do not fit one cluster: one GPU in step modes -1 / 3
// (in front of every block step of the graph, like the cluster engine), and the peer-memory multi-GPU mode (between the
// launches of the loop kernel, which hands every run of small steps over to it).  Decided once per commit.
void decide_chip(al26_ctx *c) {
  c->chip_on = false;
  GravDev &g = c->g;
  g.chip_max = 0;
  if (!c->chip_ok || !c->coop_ok || !c->committed || c->chip_max == 0 || g.n_loc != g.n_tot) return;
  // one GPU: on request only (step mode 3) -- measured at N = 1e5 the ~12 us it needs per small block step (graph: 14 us)
  // do not pay for a cooperative launch + state load in front of every bigger step, the runs being 3-8 steps long;
  // peer-memory mode: automatic -- there a small step costs 21 us inside the loop kernel, on every rank
  if (c->world == 1 ? c->step_mode != 3 : !is_p2p(c)) return;
  int p_cap = 0;
  const int nc = c->sm_count < CHIP_MAX_CTAS ? c->sm_count : CHIP_MAX_CTAS;
  if (!chip_plan(g.n_tot, nc, c->max_smem_optin, &p_cap) || !chip_fits(nc, p_cap, c->sm_count, c->max_smem_optin)) return;
  if (!g.chip_mail) {
    if (cudaMalloc(&g.chip_mail, chip_mail_bytes(nc)) != cudaSuccess) {
      cudaGetLastError();
      g.chip_mail = nullptr;
      return;
    }
    cudaMemsetAsync(g.chip_mail, 0, chip_mail_bytes(nc), c->stream);
  }
  g.chip_n = nc;
  g.chip_p = p_cap;
  // default: one GPU 32; peer-memory mode every block step that is not exchanged (below the split threshold)
  int dflt = CHIP_MAX_ACT_DEFAULT;
  if (c->world > 1) dflt = g.split_min - 1 < CHIP_CAP ? g.split_min - 1 : CHIP_CAP;
  g.chip_max = c->chip_max > 0 ? (c->chip_max < CHIP_CAP ? c->chip_max : CHIP_CAP) : dflt;
  if (g.chip_max < 1) {
    g.chip_max = 0;
    return;
  }
  c->chip_on = true;
}

// one block step (or the init / sync variant) enqueued on the stream
int enqueue_step(al26_ctx *c, int mode, int phase, bool with_engine = false) {
  if (with_engine && mode == MODE_STEP && c->engine_on) {  // first every small step up to the next big one, on chip
    cudaError_t e = cudaSuccess;
    c->launches += launch_engine(c->g, phase, c->engine_cs, c->engine_p, c->stream, &e);
    if (e != cudaSuccess) return fail(c, AL26_ECUDA, "cluster-engine launch failed: %s", cudaGetErrorString(e));
  } else if (with_engine && mode == MODE_STEP && c->chip_on) {
    cudaError_t e = cudaSuccess;
    c->launches += launch_chip(c->g, phase, c->stream, &e);
    if (e != cudaSuccess) return fail(c, AL26_ECUDA, "chip-engine launch failed: %s", cudaGetErrorString(e));
  }
  c->launches += launch_predict_list(c->g, mode, phase, c->stream);
  int rc = gather_j(c);
  if (rc) return rc;
  c->launches += launch_force(c->g, phase, c->stream);
  c->launches += launch_correct(c->g, mode, phase, c->stream);
  if (mode == MODE_STEP) {
    rc = reduce_tnext(c, (phase + 1) % 3);
    if (rc) return rc;
  }
  return 0;
}

constexpr int GRAPH_ROUNDS = 8;  // 3 phases x 8 = 24 block steps per graph launch
// with the cluster engine a graph slot covers a whole run of small steps plus one big step, and a call often ends
// inside the first slots: a shorter graph leaves fewer no-op launches behind (AL26_GRAPH_ROUNDS_ENGINE overrides)
constexpr int GRAPH_ROUNDS_ENGINE = 2;

int build_graph_once(al26_ctx *c);
int build_graph(al26_ctx *c) {
  if (c->graph) cudaGraphExecDestroy(c->graph);
  c->graph = nullptr;
  decide_engine(c);
  decide_chip(c);
  if (is_p2p(c)) return 0;  // the peer-memory path runs inside the loop kernel (and the chip engine)
  int rc = build_graph_once(c);
  if (rc && c->chip_on) {  // a driver that cannot capture the cooperative launch: the plain graph
    cudaGetLastError();
    c->chip_on = false;
    c->g.chip_max = 0;
    rc = build_graph_once(c);
  }
  return rc;
}
int build_graph_once(al26_ctx *c) {
  cudaGraph_t graph = nullptr;
  CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  const int64_t l0 = c->launches;
  int rc = 0;
  int rounds = GRAPH_ROUNDS;
  if (c->engine_on) {
    rounds = GRAPH_ROUNDS_ENGINE;
    if (const char *e = getenv("AL26_GRAPH_ROUNDS_ENGINE")) rounds = atoi(e) > 0 ? atoi(e) : rounds;
  }
  for (int r = 0; r < rounds && !rc; r++)
    for (int ph = 0; ph < 3 && !rc; ph++) rc = enqueue_step(c, MODE_STEP, ph, true);
  c->launches = l0;
  cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail(c, AL26_ECUDA, "graph capture failed: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(&c->graph, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(c, AL26_ECUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
  c->graph_steps = 3 * rounds;
  return 0;
}

int reset_ctrl(al26_ctx *c, int zero_counters) {
  k_reset_ctrl<<<1, 32, 0, c->stream>>>(c->g.ctrl, c->g.hdr, zero_counters);
  c->launches++;
  return 0;
}

// forces + initial timesteps on every local particle (the "dirty" path)
int dist_single(al26_ctx *c, int mode);
int initialise_forces(al26_ctx *c) {
  int rc;
  // mass-only update (al26_grav_set_mass between two evolve calls): every particle sits at the synchronisation time
  // with the timestep the synchronisation step gave it; under policy 0 only acc / jerk / pot are recomputed
  c->g.keep_dt = (c->dirty == 1 && c->reinit_policy == 0) ? 1 : 0;
  if (is_p2p(c)) {
    rc = dist_single(c, MODE_INIT);
  } else {
    reset_ctrl(c, 0);
    rc = enqueue_step(c, MODE_INIT, 0);
  }
  if (rc) return rc;
  c->dirty = 0;
  return 0;
}

int read_header(al26_ctx *c) {
  CU(cudaMemcpyAsync(c->h_hdr, c->g.hdr, sizeof(GravHeader), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

bool use_loop(al26_ctx *c) {
  if (c->step_mode != 1 || c->world != 1 || !c->coop_ok) return false;
  int minb = 2, ipt = 2;
  force_variant_info(c->force_variant, &minb, &ipt);
  return loop_max_ctas_per_sm(c->force_variant) >= minb;
}

// up to max_steps block steps inside one cooperative launch; refreshes the host copy of the header
int run_loop(al26_ctx *c, int max_steps) {
  k_loop_prepare<<<1, 32, 0, c->stream>>>(c->g.hdr);
  cudaError_t e = cudaSuccess;
  c->launches += 1 + launch_loop(c->g, c->dbg_phase, max_steps, c->stream, &e);
  if (e != cudaSuccess) return fail(c, AL26_ECUDA, "cooperative launch of the loop kernel failed: %s", cudaGetErrorString(e));
  int rc = read_header(c);
  if (rc) return rc;
  if (c->h_hdr->loop_error) return fail(c, AL26_ECUDA, "loop kernel: grid barrier spin limit hit");
  c->dbg_phase = c->h_hdr->phase;
  return 0;
}

// peer-memory mode: up to max_steps block steps (or the single init / sync step) in one cooperative launch
int run_dist(al26_ctx *c, int mode, int max_steps) {
  if (!c->p2p_ready) return fail(c, AL26_ESTATE, "peer-memory mode: peers' slabs not imported (al26_dist_p2p_import)");
  k_loop_prepare<<<1, 32, 0, c->stream>>>(c->g.hdr);
  cudaError_t e = cudaSuccess;
  c->launches += 1 + launch_loop_dist(c->g, mode, c->dbg_phase, max_steps, c->dist_step, c->stream, &e);
  if (e != cudaSuccess) return fail(c, AL26_ECUDA, "cooperative launch of the peer-memory loop kernel failed: %s", cudaGetErrorString(e));
  int rc = read_header(c);
  if (rc) return rc;
  if (c->h_hdr->loop_error) return fail(c, AL26_ECUDA, "peer-memory loop kernel: barrier spin limit hit (code %d)", c->h_hdr->loop_error);
  c->dbg_phase = c->h_hdr->phase;
  c->dist_step = c->h_hdr->dist_step;
  return 0;
}
// peer-memory mode with the chip engine: [k_chip, k_loop_dist] pairs queued back to back -- the chip engine takes every run
// of small block steps (redundantly on every rank, no traffic), the loop kernel the bigger ones and hands the next run
// back; phase and exchange id travel through the header, so the host only looks at it once per batch
int run_dist_chained(al26_ctx *c) {
  if (!c->p2p_ready) return fail(c, AL26_ESTATE, "peer-memory mode: peers' slabs not imported (al26_dist_p2p_import)");
  k_loop_prepare<<<1, 32, 0, c->stream>>>(c->g.hdr);
  c->launches++;
  int pairs_needed = 0;
  int batch = c->expected_graph_launches > 8 ? c->expected_graph_launches - 4 : 8;
  while (true) {
    for (int b = 0; b < batch; b++) {
      cudaError_t e = cudaSuccess;
      c->launches += launch_chip(c->g, -1, c->stream, &e);
      if (e != cudaSuccess) return fail(c, AL26_ECUDA, "chip-engine launch failed: %s", cudaGetErrorString(e));
      // ONE step per launch: a big block is almost always followed by a run of small ones (it sits on a coarse level of
      // the block-time hierarchy), and finding that out inside the loop kernel costs a whole scheduler pass
      c->launches += launch_loop_dist(c->g, MODE_STEP, -1, 1, 0ull, c->stream, &e);
      if (e != cudaSuccess) return fail(c, AL26_ECUDA, "cooperative launch of the peer-memory loop kernel failed: %s", cudaGetErrorString(e));
      pairs_needed++;
    }
    int rc = read_header(c);
    if (rc) return rc;
    if (c->h_hdr->loop_error) return fail(c, AL26_ECUDA, "peer-memory loop / chip engine: spin limit hit (code %d)", c->h_hdr->loop_error);
    if (c->h_hdr->done) break;
    batch = 8;
  }
  c->expected_graph_launches = pairs_needed;
  c->dbg_phase = c->h_hdr->phase;
  c->dist_step = c->h_hdr->dist_step;
  return 0;
}

// one init or sync step in peer-memory mode, then pull what was staged into the local state
int dist_single(al26_ctx *c, int mode) {
  reset_ctrl(c, 0);
  c->dbg_phase = 0;
  int rc = run_dist(c, mode, 1);
  if (rc) return rc;
  c->launches += launch_pull(c->g, c->dist_step, c->stream);
  return 0;
}

int begin_evolve(al26_ctx *c, double t_end) {
  if (!c->committed) return fail(c, AL26_ESTATE, "evolve before commit");
  if (c->in_evolve) return fail(c, AL26_ESTATE, "evolve already in progress");
  const double span = t_end - c->t_model;
  if (!(span > 0.0)) return fail(c, AL26_ETIME, "t_end %.17g is not after model time %.17g", t_end, c->t_model);
  c->g.eps2 = c->eps2; c->g.eta = c->eta; c->g.dt_max = c->dt_max; c->g.dt_min = c->dt_min;
  int rc;
  if (c->graph_stale) {
    if ((rc = build_graph(c))) return rc;
    c->graph_stale = false;
  }
  reset_ctrl(c, 1);  // zero this call's step / pair counters (the init force counts its pairs)
  if (c->dirty && (rc = initialise_forces(c))) return rc;
  double D = pow2floor_h(span);
  if (D > c->dt_max) D = c->dt_max;
  c->g.Dmax = D;
  reset_ctrl(c, 0);
  c->launches += launch_begin(c->g, span, D, c->stream);
  if ((rc = reduce_tnext(c, 0))) return rc;
  c->in_evolve = true;
  c->t_end_pending = t_end;
  c->dbg_phase = 0;
  return 0;
}

int finish_evolve(al26_ctx *c) {
  int rc;
  if (is_p2p(c)) {
    rc = dist_single(c, MODE_SYNC);
  } else {
    reset_ctrl(c, 0);
    rc = enqueue_step(c, MODE_SYNC, 0);
  }
  if (rc) return rc;
  // times are relative to the start of an evolve call: everybody is synchronised, so tau = 0
  CU(cudaMemsetAsync(c->g.t, 0, (size_t)c->g.n_loc * sizeof(double), c->stream));
  if ((rc = read_header(c))) return rc;
  c->t_model = c->t_end_pending;
  c->in_evolve = false;
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int al26_version(void) { return 200; }

int al26_device_count(void) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return ndev;
}

const char *al26_last_error(al26_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

al26_ctx *al26_create(int device_id) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    fail(nullptr, AL26_ENODEV, "no CUDA device: %s", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return nullptr;
  }
  if (device_id < 0 || device_id >= ndev) {
    fail(nullptr, AL26_EINVAL, "device %d out of range (%d devices)", device_id, ndev);
    return nullptr;
  }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) {
    fail(nullptr, AL26_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return nullptr;
  }
  if (prop.major != 10) {
    fail(nullptr, AL26_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", device_id, prop.major,
         prop.minor);
    return nullptr;
  }
  al26_ctx *c = new al26_ctx();
  c->device = device_id;
  c->sm_count = prop.multiProcessorCount;
  c->clock_khz = prop.clockRate;
  bool ok = cudaSetDevice(device_id) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess &&
            cudaEventCreate(&c->ev2) == cudaSuccess && cudaEventCreate(&c->ev3) == cudaSuccess &&
            cudaMallocHost(&c->h_hdr, sizeof(GravHeader)) == cudaSuccess &&
            cudaMallocHost(&c->h_small, 16 * sizeof(double)) == cudaSuccess &&
            cudaMallocHost(&c->h_events, (ENR_NCOUNTERS + ENR_MAX_SOURCES) * sizeof(int)) == cudaSuccess &&
            cudaMallocHost(&c->h_rows, 4 * ENR_MAX_SOURCES * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&c->en_out, 4 * sizeof(double)) == cudaSuccess && force_kernel_setup() == cudaSuccess &&
            loop_kernel_setup() == cudaSuccess && enrich_kernel_setup() == cudaSuccess;
  if (!ok) {
    fail(nullptr, AL26_ECUDA, "context setup failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return nullptr;
  }
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device_id);
  c->coop_ok = coop != 0;
  cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device_id);
  c->engine_ok = engine_kernel_setup(c->max_smem_optin) == cudaSuccess;
  if (!c->engine_ok) cudaGetLastError();
  c->chip_ok = chip_kernel_setup(c->max_smem_optin) == cudaSuccess;
  if (!c->chip_ok) cudaGetLastError();
  return c;
}

void al26_destroy(al26_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  free_gravity(c);
  free_enrich(c);
  if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
  if (c->scratch) cudaFree(c->scratch);
  if (c->en_scratch) cudaFree(c->en_scratch);
  if (c->en_out) cudaFree(c->en_out);
  if (c->h_hdr) cudaFreeHost(c->h_hdr);
  if (c->h_small) cudaFreeHost(c->h_small);
  if (c->h_events) cudaFreeHost(c->h_events);
  if (c->h_rows) cudaFreeHost(c->h_rows);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev2) cudaEventDestroy(c->ev2);
  if (c->ev3) cudaEventDestroy(c->ev3);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int al26_device_info(al26_ctx *c, int *sm_count, int *clock_khz, int64_t *free_bytes, int64_t *total_bytes) {
  if (!c) return AL26_EINVAL;
  CU(cudaSetDevice(c->device));
  size_t f = 0, t = 0;
  CU(cudaMemGetInfo(&f, &t));
  if (sm_count) *sm_count = c->sm_count;
  if (clock_khz) *clock_khz = c->clock_khz;
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
  return 0;
}

int al26_dist_unique_id(void *out128) {
  std::string why;
  if (!out128) return AL26_EINVAL;
  if (!load_nccl(why)) {
    g_create_error = why;
    return AL26_ENCCL;
  }
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) {
    g_create_error = "ncclGetUniqueId failed";
    return AL26_ENCCL;
  }
  memcpy(out128, &id, sizeof(id));
  return 0;
}

int al26_dist_init(al26_ctx *c, int rank, int world, const void *uid) {
  if (!c || world < 1 || rank < 0 || rank >= world) return fail(c, AL26_EINVAL, "bad rank/world %d/%d", rank, world);
  if (c->committed || c->e_committed) return fail(c, AL26_ESTATE, "al26_dist_init must precede commit");
  if (world == 1) {
    c->rank = 0;
    c->world = 1;
    return 0;
  }
  if (world > MAX_PEERS && c->dist_mode == 1) return fail(c, AL26_EINVAL, "peer-memory mode supports at most %d ranks", MAX_PEERS);
  if (!uid) return fail(c, AL26_EINVAL, "null nccl unique id");
  std::string why;
  if (!load_nccl(why)) return fail(c, AL26_ENCCL, "%s", why.c_str());
  CU(cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, uid, sizeof(id));
  NC(g_nccl.CommInitRank(&c->comm, world, id, rank));
  c->rank = rank;
  c->world = world;
  return 0;
}

int al26_dist_set_mode(al26_ctx *c, int mode) {
  if (!c) return AL26_EINVAL;
  if (mode != 0 && mode != 1) return fail(c, AL26_EINVAL, "dist mode must be 0 (NCCL) or 1 (peer memory)");
  if (c->committed || c->e_committed) return fail(c, AL26_ESTATE, "al26_dist_set_mode must precede commit");
  c->dist_mode = mode;
  return 0;
}

int al26_dist_set_split_min(al26_ctx *c, int n_act_min) {
  if (!c) return AL26_EINVAL;
  if (n_act_min < 0) return fail(c, AL26_EINVAL, "split threshold must be >= 0 (0 = automatic)");
  if (c->committed) return fail(c, AL26_ESTATE, "al26_dist_set_split_min must precede commit");
  c->split_min = n_act_min;
  return 0;
}

int al26_dist_p2p_export(al26_ctx *c, void *out64) {
  if (!c || !out64) return AL26_EINVAL;
  if (!is_p2p(c) || !c->committed || !c->slab) return fail(c, AL26_ESTATE, "p2p_export needs a committed peer-memory context");
  CU(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, c->slab));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out64, &h, 64);
  return 0;
}

int al26_dist_p2p_import(al26_ctx *c, const void *handles, int world) {
  if (!c || !handles) return AL26_EINVAL;
  if (!is_p2p(c) || !c->committed) return fail(c, AL26_ESTATE, "p2p_import needs a committed peer-memory context");
  if (world != c->world || world > MAX_PEERS) return fail(c, AL26_EINVAL, "p2p_import: world %d (context %d, max %d)", world, c->world, MAX_PEERS);
  CU(cudaSetDevice(c->device));
  for (int q = 0; q < world; q++) {
    if (q == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + 64 * q, 64);
    void *p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->g.slab[q] = p;
  }
  c->p2p_ready = true;
  c->p2p_ipc = true;
  return 0;
}

int al26_dist_init_local(al26_ctx *c, int rank, int world) {
  if (!c || world < 1 || rank < 0 || rank >= world) return fail(c, AL26_EINVAL, "bad rank/world %d/%d", rank, world);
  if (c->committed || c->e_committed) return fail(c, AL26_ESTATE, "al26_dist_init_local must precede commit");
  if (world > MAX_PEERS) return fail(c, AL26_EINVAL, "peer-memory mode supports at most %d ranks", MAX_PEERS);
  c->rank = rank;
  c->world = world;
  c->dist_mode = 1;
  c->local_group = world > 1;
  return 0;
}

int al26_dist_p2p_local_slab(al26_ctx *c, void **slab) {
  if (!c || !slab) return AL26_EINVAL;
  if (!is_p2p(c) || !c->committed || !c->slab) return fail(c, AL26_ESTATE, "p2p_local_slab needs a committed peer-memory context");
  *slab = c->slab;
  return 0;
}

int al26_dist_p2p_attach(al26_ctx *c, void *const *slabs, const int *devices, int world) {
  if (!c || !slabs || !devices) return AL26_EINVAL;
  if (!is_p2p(c) || !c->committed) return fail(c, AL26_ESTATE, "p2p_attach needs a committed peer-memory context");
  if (world != c->world || world > MAX_PEERS) return fail(c, AL26_EINVAL, "p2p_attach: world %d (context %d, max %d)", world, c->world, MAX_PEERS);
  CU(cudaSetDevice(c->device));
  for (int q = 0; q < world; q++) {
    if (q == c->rank) continue;
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, c->device, devices[q]));
    if (!can) return fail(c, AL26_ECUDA, "device %d cannot access device %d's memory (no NVLink / P2P path)", c->device, devices[q]);
    const cudaError_t e = cudaDeviceEnablePeerAccess(devices[q], 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(c, AL26_ECUDA, "cudaDeviceEnablePeerAccess(%d): %s", devices[q], cudaGetErrorString(e));
    cudaGetLastError();
    c->g.slab[q] = slabs[q];
  }
  c->p2p_ready = true;
  c->p2p_ipc = false;
  return 0;
}

int al26_grav_set_params(al26_ctx *c, double eps2, double eta, double dt_max, double dt_min) {
  if (!c) return AL26_EINVAL;
  if (!(eps2 >= 0.0) || !(eta > 0.0) || !(dt_max > 0.0) || !(dt_min > 0.0) || dt_min > dt_max)
    return fail(c, AL26_EINVAL, "bad gravity parameters eps2=%g eta=%g dt_max=%g dt_min=%g", eps2, eta, dt_max, dt_min);
  if (c->in_evolve) return fail(c, AL26_ESTATE, "set_params during evolve");
  if (dt_min < 4.909093465297727e-91 /* 2^-300 */)
    return fail(c, AL26_EINVAL, "dt_min %g below 2^-300 (the corrector forms dt^-3)", dt_min);
  c->eps2 = eps2;
  c->eta = eta;
  c->dt_max = pow2floor_h(dt_max);
  c->dt_min = pow2floor_h(dt_min);
  c->dirty = 2;
  c->graph_stale = c->committed;
  return 0;
}

int al26_grav_set_reinit_policy(al26_ctx *c, int policy) {
  if (!c) return AL26_EINVAL;
  if (policy != 0 && policy != 1) return fail(c, AL26_EINVAL, "re-initialisation policy must be 0 (keep timesteps) or 1 (initial timesteps)");
  if (c->in_evolve) return fail(c, AL26_ESTATE, "set_reinit_policy during evolve");
  c->reinit_policy = policy;
  return 0;
}

int al26_grav_commit(al26_ctx *c, int64_t n, const double *m, const double *x, const double *y, const double *z,
                     const double *vx, const double *vy, const double *vz) {
  if (!c) return AL26_EINVAL;
  if (n <= 0 || n > 0x7fffffff / 2) return fail(c, AL26_EINVAL, "particle count %lld out of range", (long long)n);
  if (!m || !x || !y || !z || !vx || !vy || !vz) return fail(c, AL26_EINVAL, "null array");
  if (c->world > 1 && !is_p2p(c) && n % c->world)  // the all-gather wants equal slices; the peer-memory mode owns i % world and takes any n
    return fail(c, AL26_EINVAL, "NCCL mode: particle count %lld not divisible by world size %d", (long long)n, c->world);
  if (c->in_evolve) return fail(c, AL26_ESTATE, "commit during evolve");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  free_gravity(c);
  int64_t i0, nloc;
  slice_of(n, c->rank, c->world, i0, nloc);
  if (is_p2p(c)) {  // replicated state: every rank holds all n particles, owns i % world == rank
    i0 = 0;
    nloc = n;
  }
  GravDev &g = c->g;
  g.n_tot = (int)n; g.n_loc = (int)nloc; g.i0 = (int)i0;
  g.rank = c->rank; g.world = c->world; g.p2p = is_p2p(c) ? 1 : 0;
  if (is_p2p(c)) {
    CU(cudaMalloc(&c->slab, slab_bytes((int)n)));
    CU(cudaMemset(c->slab, 0, slab_bytes((int)n)));
    g.slab[c->rank] = c->slab;
    c->dist_step = 0;
    CU(cudaMalloc(&g.list_own, (size_t)nloc * sizeof(int)));
    // splitting a block step over the ranks pays once the force work saved exceeds the ~20 us an exchange costs:
    // (1 - 1/P) n_act N / (4.7e11 pairs/s) > 20 us
    long long auto_min = (long long)(20e-6 * 4.7e11 / (double)n * (double)c->world / (double)(c->world - 1)) + 1;
    if (auto_min < 8) auto_min = 8;
    if (auto_min > 8192) auto_min = 8192;
    g.split_min = c->split_min > 0 ? c->split_min : (int)auto_min;
  }
  {
    int minb = 2, ipt = 2;
    force_variant_info(c->force_variant, &minb, &ipt);
    g.variant = c->force_variant;
    g.force_ipt = ipt;
    g.grid_force = minb * c->sm_count;
  }
  g.eps2 = c->eps2; g.eta = c->eta; g.dt_max = c->dt_max; g.dt_min = c->dt_min;
  g.part_cap = part_capacity(g.n_loc, g.grid_force);
  const size_t nl = (size_t)nloc, nt = (size_t)n;
  CU(cudaMalloc(&g.pos, nl * sizeof(double4)));
  CU(cudaMalloc(&g.vel, nl * sizeof(double4)));
  CU(cudaMalloc(&g.acc, nl * sizeof(double4)));
  CU(cudaMalloc(&g.jrk, nl * sizeof(double4)));
  CU(cudaMalloc(&g.t, nl * sizeof(double)));
  CU(cudaMalloc(&g.dt, nl * sizeof(double)));
  CU(cudaMalloc(&g.jpos, nt * sizeof(double4)));
  CU(cudaMalloc(&g.jvel, nt * sizeof(double4)));
  CU(cudaMalloc(&g.list, nl * sizeof(int)));
  CU(cudaMalloc(&g.part_a, (size_t)g.part_cap * sizeof(double4)));
  CU(cudaMalloc(&g.part_j, (size_t)g.part_cap * sizeof(double4)));
  CU(cudaMalloc(&g.ctrl, 3 * sizeof(StepCtrl)));
  CU(cudaMalloc(&g.hdr, sizeof(GravHeader)));
  CU(cudaMalloc(&g.act, sizeof(ActBuf)));
  CU(cudaMemsetAsync(g.act, 0, sizeof(ActBuf), c->stream));
  {  // fused small steps: one contiguous chunk of particles per CTA, held in the force kernel's stage buffers
    int jc = (int)((nloc + g.grid_force - 1) / g.grid_force);
    jc = (jc + 7) & ~7;
    g.fuse_jc = jc;
    const bool ok = (g.n_loc == g.n_tot) && jc <= FORCE_STAGES * FORCE_TJ;
    // automatic: on where block steps are latency-bound (measured on B200, one and two GPUs: from N ~ 5e4 up the ~4 us
    // saved per small step are cancelled by the fused kernel's ~1 % slower big-block force loop)
    const int want = c->fuse_max >= 0 ? c->fuse_max : (n <= FUSE_AUTO_MAX_N ? FUSE_CAP : 0);
    g.fuse_max = ok ? want : 0;
  }
  {
    g.big_nact = c->big_nact;
    std::vector<int> tab(decomp_table_entries(g.n_loc, g.force_ipt, g.big_nact));
    fill_decomp_table(tab.data(), g.n_loc, g.n_tot, g.grid_force, g.force_ipt, g.big_nact, c->max_rounds, c->item_overhead);
    int *d_tab = nullptr;
    CU(cudaMalloc(&d_tab, tab.size() * sizeof(int)));
    CU(cudaMemcpy(d_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
    g.decomp_tab = d_tab;
  }
  CU(cudaMemsetAsync(g.acc, 0, nl * sizeof(double4), c->stream));
  CU(cudaMemsetAsync(g.jrk, 0, nl * sizeof(double4), c->stream));
  CU(cudaMemsetAsync(g.t, 0, nl * sizeof(double), c->stream));
  CU(cudaMemsetAsync(g.dt, 0, nl * sizeof(double), c->stream));
  CU(cudaMemsetAsync(g.hdr, 0, sizeof(GravHeader), c->stream));
  int rc = ensure_scratch(c, 16 * nt);
  if (rc) return rc;
  const double *src[7] = {m, x, y, z, vx, vy, vz};
  for (int k = 0; k < 7; k++)
    CU(cudaMemcpyAsync(c->scratch + k * nl, src[k] + i0, nl * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  double *s = c->scratch;
  k_pack<<<nblk(nloc), 256, 0, c->stream>>>((int)nloc, s, s + nl, s + 2 * nl, s + 3 * nl, s + 4 * nl, s + 5 * nl,
                                            s + 6 * nl, g.pos, g.vel);
  c->launches++;
  reset_ctrl(c, 1);
  CU(cudaStreamSynchronize(c->stream));
  c->n_tot = n;
  c->committed = true;
  c->dirty = 2;
  rc = build_graph(c);
  if (rc) {
    free_gravity(c);
    return rc;
  }
  return 0;
}

int al26_grav_set_mass(al26_ctx *c, int64_t n, const double *m) {
  if (!c) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "set_mass before commit");
  if (c->in_evolve) return fail(c, AL26_ESTATE, "set_mass during evolve");
  if (n != c->n_tot || !m) return fail(c, AL26_EINVAL, "set_mass: n=%lld, committed %lld", (long long)n, (long long)c->n_tot);
  CU(cudaSetDevice(c->device));
  const size_t nl = (size_t)c->g.n_loc;
  CU(cudaMemcpyAsync(c->scratch, m + c->g.i0, nl * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  k_set_mass<<<nblk(nl), 256, 0, c->stream>>>((int)nl, c->scratch, c->g.pos);
  c->launches++;
  CU(cudaStreamSynchronize(c->stream));  // the host buffer may be reused by the caller
  if (c->dirty < 1) c->dirty = 1;
  return 0;
}

int al26_grav_set_time(al26_ctx *c, double t) {
  if (!c) return AL26_EINVAL;
  if (c->in_evolve) return fail(c, AL26_ESTATE, "set_time during evolve");
  c->t_model = t;
  return 0;
}
int al26_grav_get_time(al26_ctx *c, double *t) {
  if (!c || !t) return AL26_EINVAL;
  *t = c->t_model;
  return 0;
}

int al26_grav_initialize(al26_ctx *c) {
  if (!c) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "initialize before commit");
  if (c->in_evolve) return fail(c, AL26_ESTATE, "initialize during evolve");
  CU(cudaSetDevice(c->device));
  c->g.eps2 = c->eps2; c->g.eta = c->eta; c->g.dt_max = c->dt_max; c->g.dt_min = c->dt_min;
  if (c->dirty) {
    int rc = initialise_forces(c);
    if (rc) return rc;
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// the part of al26_grav_evolve between begin_evolve and the end of the call: any failure in here leaves through the
// caller, which clears in_evolve (so a CUDA error does not wedge the context in "evolve already in progress")
static int evolve_run(al26_ctx *c, const int64_t l0, int64_t *n_block_steps, int64_t *n_pairs) {
  int rc;
  if (is_p2p(c) && c->chip_on) {
    if ((rc = run_dist_chained(c))) return rc;
  } else if (is_p2p(c)) {
    while (true) {
      if ((rc = run_dist(c, MODE_STEP, 1 << 30))) return rc;
      if (c->h_hdr->done) break;
    }
  } else if (use_loop(c)) {
    while (true) {
      if ((rc = run_loop(c, 1 << 30))) return rc;
      if (c->h_hdr->done) break;
    }
  } else {
    if (c->chip_on) {  // the engine's error flag lives in the header: clear what an earlier call may have left
      k_loop_prepare<<<1, 32, 0, c->stream>>>(c->g.hdr);
      c->launches++;
    }
    int launches_needed = 0;
    int batch = c->expected_graph_launches > 1 ? c->expected_graph_launches - 1 : 1;
    while (true) {
      for (int b = 0; b < batch; b++) {
        CU(cudaGraphLaunch(c->graph, c->stream));
        c->launches += (c->engine_on || c->chip_on ? 4 : 3) * c->graph_steps;
        launches_needed++;
      }
      batch = 1;
      if ((rc = read_header(c))) return rc;
      if (c->chip_on && c->h_hdr->loop_error)
        return fail(c, AL26_ECUDA, "chip engine: spin limit hit (code %d)", c->h_hdr->loop_error);
      if (c->h_hdr->done) break;
    }
    c->expected_graph_launches = launches_needed;
  }
  if ((rc = finish_evolve(c))) return rc;
  CU(cudaEventRecord(c->ev1, c->stream));
  CU(cudaEventSynchronize(c->ev1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->last_ms = ms;
  c->last_launches = c->launches - l0;
  if (n_block_steps) *n_block_steps = c->h_hdr->n_steps;
  if (n_pairs) *n_pairs = c->h_hdr->n_pairs;
  return 0;
}

int al26_grav_evolve(al26_ctx *c, double t_end, int64_t *n_block_steps, int64_t *n_pairs) {
  if (!c) return AL26_EINVAL;
  if (n_block_steps) *n_block_steps = 0;
  if (n_pairs) *n_pairs = 0;
  if (c->committed && !c->in_evolve && t_end == c->t_model) return 0;
  CU(cudaSetDevice(c->device));
  const int64_t l0 = c->launches;
  CU(cudaEventRecord(c->ev0, c->stream));
  int rc = begin_evolve(c, t_end);
  if (rc) return rc;
  rc = evolve_run(c, l0, n_block_steps, n_pairs);
  if (rc) c->in_evolve = false;  // every error exit: the context stays usable (the particle state is undefined after a CUDA error)
  return rc;
}

int al26_grav_dbg_begin(al26_ctx *c, double t_end) {
  if (!c) return AL26_EINVAL;
  CU(cudaSetDevice(c->device));
  int rc = begin_evolve(c, t_end);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int al26_grav_dbg_advance(al26_ctx *c, int64_t max_steps, int64_t *n_done, int *finished) {
  if (!c) return AL26_EINVAL;
  if (!c->in_evolve) return fail(c, AL26_ESTATE, "dbg_advance outside begin/finish");
  CU(cudaSetDevice(c->device));
  int64_t done = 0;
  int fin = 0;
  while (max_steps < 0 || done < max_steps) {
    int rc = read_header(c);
    if (rc) return rc;
    const long long before = c->h_hdr->n_steps;
    if (is_p2p(c) || use_loop(c)) {
      if ((rc = (is_p2p(c) ? run_dist(c, MODE_STEP, 1) : run_loop(c, 1)))) return rc;
      if (c->h_hdr->done) {
        fin = 1;
        break;
      }
      done += c->h_hdr->n_steps - before;
      continue;
    }
    if ((rc = enqueue_step(c, MODE_STEP, c->dbg_phase))) return rc;
    if ((rc = read_header(c))) return rc;
    if (c->h_hdr->done) {
      // the step was a no-op; do not advance the phase: the no-op carried t_next forward into
      // the next record, so keep rotating to stay consistent
      c->dbg_phase = (c->dbg_phase + 1) % 3;
      fin = 1;
      break;
    }
    c->dbg_phase = (c->dbg_phase + 1) % 3;
    done += c->h_hdr->n_steps - before;
  }
  if (n_done) *n_done = done;
  if (finished) *finished = fin;
  return 0;
}

int al26_grav_dbg_profile_steps(al26_ctx *c, int reps, double *us3) {
  if (!c || !us3) return AL26_EINVAL;
  if (!c->in_evolve) return fail(c, AL26_ESTATE, "dbg_profile_steps outside begin/finish");
  if (c->world != 1) return fail(c, AL26_ESTATE, "dbg_profile_steps is a single-GPU diagnostic");
  CU(cudaSetDevice(c->device));
  cudaEvent_t ev[4];
  for (auto &e : ev) CU(cudaEventCreate(&e));
  double acc[3] = {0, 0, 0};
  int done_steps = 0;
  for (int r = 0; r < reps; r++) {
    CU(cudaEventRecord(ev[0], c->stream));
    c->launches += launch_predict_list(c->g, MODE_STEP, c->dbg_phase, c->stream);
    CU(cudaEventRecord(ev[1], c->stream));
    c->launches += launch_force(c->g, c->dbg_phase, c->stream);
    CU(cudaEventRecord(ev[2], c->stream));
    c->launches += launch_correct(c->g, MODE_STEP, c->dbg_phase, c->stream);
    CU(cudaEventRecord(ev[3], c->stream));
    CU(cudaEventSynchronize(ev[3]));
    c->dbg_phase = (c->dbg_phase + 1) % 3;
    int rc = read_header(c);
    if (rc) return rc;
    if (c->h_hdr->done) break;
    for (int k = 0; k < 3; k++) {
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
      acc[k] += ms * 1e3;
    }
    done_steps++;
  }
  for (auto &e : ev) cudaEventDestroy(e);
  for (int k = 0; k < 3; k++) us3[k] = done_steps ? acc[k] / done_steps : 0.0;
  return 0;
}

int al26_grav_dbg_finish(al26_ctx *c) {
  if (!c) return AL26_EINVAL;
  if (!c->in_evolve) return fail(c, AL26_ESTATE, "dbg_finish outside begin");
  CU(cudaSetDevice(c->device));
  const int rc = finish_evolve(c);
  if (rc) c->in_evolve = false;
  return rc;
}

int al26_grav_get_state(al26_ctx *c, int64_t n, double *m, double *x, double *y, double *z, double *vx, double *vy,
                        double *vz) {
  if (!c) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "get_state before commit");
  if (n != c->n_tot) return fail(c, AL26_EINVAL, "get_state: n=%lld, committed %lld", (long long)n, (long long)c->n_tot);
  CU(cudaSetDevice(c->device));
  const GravDev &g = c->g;
  c->launches += launch_snapshot_j(g, c->stream);
  int rc = gather_j(c);
  if (rc) return rc;
  const size_t nt = (size_t)n;
  double *s = c->scratch;
  k_unpack<<<nblk(n), 256, 0, c->stream>>>((int)n, g.jpos, g.jvel, s, s + nt, s + 2 * nt, s + 3 * nt, s + 4 * nt,
                                           s + 5 * nt, s + 6 * nt);
  c->launches++;
  double *dst[7] = {m, x, y, z, vx, vy, vz};
  for (int k = 0; k < 7; k++)
    if (dst[k]) CU(cudaMemcpyAsync(dst[k], s + k * nt, nt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int al26_grav_get_acc_jerk(al26_ctx *c, int64_t n, double *ax, double *ay, double *az, double *jx, double *jy,
                           double *jz, double *pot) {
  if (!c) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "get_acc_jerk before commit");
  if (n != c->n_tot) return fail(c, AL26_EINVAL, "get_acc_jerk: size mismatch");
  CU(cudaSetDevice(c->device));
  const GravDev &g = c->g;
  const size_t nl = (size_t)g.n_loc;
  k_unpack_aj<<<nblk(nl), 256, 0, c->stream>>>((int)nl, g.acc, g.jrk, c->scratch);
  c->launches++;
  double *dst[7] = {ax, ay, az, jx, jy, jz, pot};
  for (int k = 0; k < 7; k++)
    if (dst[k])
      CU(cudaMemcpyAsync(dst[k] + g.i0, c->scratch + k * nl, nl * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int al26_grav_get_timesteps(al26_ctx *c, int64_t n, double *t, double *dt) {
  if (!c) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "get_timesteps before commit");
  if (n != c->n_tot) return fail(c, AL26_EINVAL, "get_timesteps: size mismatch");
  CU(cudaSetDevice(c->device));
  const GravDev &g = c->g;
  if (t) CU(cudaMemcpyAsync(t + g.i0, g.t, (size_t)g.n_loc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (dt) CU(cudaMemcpyAsync(dt + g.i0, g.dt, (size_t)g.n_loc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int al26_grav_set_timesteps(al26_ctx *c, int64_t n, const double *t, const double *dt) {
  if (!c) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "set_timesteps before commit");
  if (n != c->n_tot || !t || !dt) return fail(c, AL26_EINVAL, "set_timesteps: size mismatch");
  for (int64_t i = 0; i < n; i++)  // the block-step machinery relies on it (exact dyadic times, exponent-flip reciprocals)
    if (!(dt[i] >= 4.909093465297727e-91) || !(dt[i] <= 1.0e300) || dt[i] != pow2floor_h(dt[i]))
      return fail(c, AL26_EINVAL, "set_timesteps: dt[%lld] = %.17g is not a power of two", (long long)i, dt[i]);
  CU(cudaSetDevice(c->device));
  const GravDev &g = c->g;
  CU(cudaMemcpyAsync(g.t, t + g.i0, (size_t)g.n_loc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(g.dt, dt + g.i0, (size_t)g.n_loc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int al26_grav_get_active(al26_ctx *c, int64_t cap, int32_t *idx, int64_t *n_active, double *tau_next) {
  if (!c) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "get_active before commit");
  if (!idx || !n_active) return fail(c, AL26_EINVAL, "null output");
  if (c->world != 1) return fail(c, AL26_ESTATE, "get_active is a single-GPU parity hook");
  CU(cudaSetDevice(c->device));
  const GravDev &g = c->g;
  // scratch: [0] t_next bits, [1] count (as int), list after
  unsigned long long *d_tn = reinterpret_cast<unsigned long long *>(c->scratch);
  int *d_cnt = reinterpret_cast<int *>(c->scratch + 1);
  int *d_list = reinterpret_cast<int *>(c->scratch + 2);
  const unsigned long long inf = INF_BITS;
  CU(cudaMemcpyAsync(d_tn, &inf, sizeof(inf), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(d_cnt, 0, sizeof(double), c->stream));
  k_min_tnext<<<nblk(g.n_loc), 256, 0, c->stream>>>(g.n_loc, g.t, g.dt, d_tn);
  k_list_at<<<nblk(g.n_loc), 256, 0, c->stream>>>(g.n_loc, g.t, g.dt, d_tn, d_list, d_cnt);
  c->launches += 2;
  unsigned long long tn_bits = 0;
  int cnt = 0;
  CU(cudaMemcpyAsync(&tn_bits, d_tn, sizeof(tn_bits), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (cnt > cap) return fail(c, AL26_ECAP, "active list of %d exceeds capacity %lld", cnt, (long long)cap);
  CU(cudaMemcpy(idx, d_list, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost));
  std::sort(idx, idx + cnt);
  *n_active = cnt;
  if (tau_next) memcpy(tau_next, &tn_bits, sizeof(double));
  return 0;
}

int al26_grav_get_last_active(al26_ctx *c, int64_t cap, int32_t *idx, int64_t *n_active) {
  if (!c) return AL26_EINVAL;
  if (!c->in_evolve) return fail(c, AL26_ESTATE, "get_last_active outside dbg_begin/dbg_finish");
  if (!idx || !n_active) return fail(c, AL26_EINVAL, "null output");
  CU(cudaSetDevice(c->device));
  const int ph = (c->dbg_phase + 2) % 3;  // the step executed last
  StepCtrl rec;
  CU(cudaMemcpyAsync(&rec, &c->g.ctrl[ph], sizeof(rec), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (rec.n_act > cap) return fail(c, AL26_ECAP, "active list of %d exceeds capacity %lld", rec.n_act, (long long)cap);
  CU(cudaMemcpy(idx, c->g.list, (size_t)rec.n_act * sizeof(int), cudaMemcpyDeviceToHost));
  std::sort(idx, idx + rec.n_act);
  for (int k = 0; k < rec.n_act; k++) idx[k] += c->g.i0;
  *n_active = rec.n_act;
  return 0;
}

int al26_grav_energies(al26_ctx *c, double *kinetic, double *potential, double *sum_mm_over_r) {
  if (!c) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "energies before commit");
  if (c->in_evolve) return fail(c, AL26_ESTATE, "energies during evolve");
  CU(cudaSetDevice(c->device));
  const GravDev &g = c->g;
  c->launches += launch_snapshot_j(g, c->stream);
  int rc = gather_j(c);
  if (rc) return rc;
  const size_t need = (size_t)energy_grid(g.n_loc) * 3 + 2 * ((size_t)1184 * 256 + (size_t)g.n_loc + 512) + 64;  // sized for n_loc >= the slice
  if (c->en_scratch_doubles < need) {
    if (c->en_scratch) cudaFree(c->en_scratch);
    c->en_scratch = nullptr;
    c->en_scratch_doubles = 0;
    CU(cudaMalloc(&c->en_scratch, need * sizeof(double)));
    c->en_scratch_doubles = need;
  }
  EnergyDev e;
  e.n_loc = g.n_loc; e.n_tot = g.n_tot; e.i0 = g.i0; e.eps2 = c->eps2;
  e.pos = g.pos; e.vel = g.vel; e.jpos = g.jpos;
  if (is_p2p(c)) {  // replicated state: each rank sums its contiguous share of the i-particles
    int64_t s0, sl;
    slice_of(g.n_tot, c->rank, c->world, s0, sl);
    e.n_loc = (int)sl; e.i0 = (int)s0; e.pos = g.pos + s0; e.vel = g.vel + s0;
  }
  e.block_part = c->en_scratch; e.out = c->en_out;
  c->launches += launch_energies(e, c->stream);
  if (c->world > 1 && !c->local_group) NC(g_nccl.AllReduce(c->en_out, c->en_out, 3, ncclDouble, ncclSum, c->comm, c->stream));
  CU(cudaMemcpyAsync(c->h_small, c->en_out, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (kinetic) *kinetic = c->h_small[0];
  if (potential) *potential = c->h_small[1];
  if (sum_mm_over_r) *sum_mm_over_r = c->h_small[2];
  return 0;
}

int al26_grav_force(al26_ctx *c, int64_t n, double eps2, const double *m, const double *x, const double *y,
                    const double *z, const double *vx, const double *vy, const double *vz, int64_t n_act,
                    const int32_t *idx, double *ax, double *ay, double *az, double *jx, double *jy, double *jz,
                    double *pot) {
  if (!c) return AL26_EINVAL;
  if (n <= 0 || n_act < 0 || n_act > n || !m || !x || !y || !z || !vx || !vy || !vz || (n_act && !idx))
    return fail(c, AL26_EINVAL, "al26_grav_force: bad arguments");
  if (n_act == 0) return 0;
  for (int64_t k = 0; k < n_act; k++)
    if (idx[k] < 0 || idx[k] >= n) return fail(c, AL26_EINVAL, "al26_grav_force: index %d out of range", idx[k]);
  CU(cudaSetDevice(c->device));
  GravDev g{};
  g.n_loc = g.n_tot = (int)n; g.i0 = 0; g.eps2 = eps2;
  {
    int minb = 2, ipt = 2;
    force_variant_info(c->force_variant, &minb, &ipt);
    g.variant = c->force_variant;
    g.force_ipt = ipt;
    g.grid_force = minb * c->sm_count;
  }
  g.part_cap = part_capacity((int)n, g.grid_force);
  const size_t nt = (size_t)n, na = (size_t)n_act;
  double *stage = nullptr;
  int *d_idx = nullptr;
  int rc = 0;
  cudaError_t e = cudaSuccess;
#define TRY(call) if (e == cudaSuccess) e = (call)
  TRY(cudaMalloc(&stage, 7 * nt * sizeof(double)));
  TRY(cudaMalloc(&g.jpos, nt * sizeof(double4)));
  TRY(cudaMalloc(&g.jvel, nt * sizeof(double4)));
  TRY(cudaMalloc(&g.list, nt * sizeof(int)));
  TRY(cudaMalloc(&d_idx, na * sizeof(int)));
  TRY(cudaMalloc(&g.part_a, (size_t)g.part_cap * sizeof(double4)));
  TRY(cudaMalloc(&g.part_j, (size_t)g.part_cap * sizeof(double4)));
  TRY(cudaMalloc(&g.raw_a, na * sizeof(double4)));
  TRY(cudaMalloc(&g.raw_j, na * sizeof(double4)));
  TRY(cudaMalloc(&g.ctrl, 3 * sizeof(StepCtrl)));
  TRY(cudaMalloc(&g.hdr, sizeof(GravHeader)));
  int *d_tab = nullptr;
  {
    g.big_nact = c->big_nact;
    std::vector<int> tab(decomp_table_entries(g.n_loc, g.force_ipt, g.big_nact));
    fill_decomp_table(tab.data(), g.n_loc, g.n_tot, g.grid_force, g.force_ipt, g.big_nact, c->max_rounds, c->item_overhead);
    TRY(cudaMalloc(&d_tab, tab.size() * sizeof(int)));
    TRY(cudaMemcpy(d_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
    g.decomp_tab = d_tab;
  }
  const double *src[7] = {m, x, y, z, vx, vy, vz};
  for (int k = 0; k < 7; k++) TRY(cudaMemcpyAsync(stage + k * nt, src[k], nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TRY(cudaMemcpyAsync(d_idx, idx, na * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if (e == cudaSuccess) {
    k_pack<<<nblk(n), 256, 0, c->stream>>>((int)n, stage, stage + nt, stage + 2 * nt, stage + 3 * nt, stage + 4 * nt,
                                           stage + 5 * nt, stage + 6 * nt, g.jpos, g.jvel);
    k_reset_ctrl<<<1, 32, 0, c->stream>>>(g.ctrl, g.hdr, 1);
    k_set_list<<<nblk(n_act), 256, 0, c->stream>>>((int)n_act, d_idx, g.list, g.ctrl);
    launch_force(g, 0, c->stream);
    launch_correct(g, MODE_RAW, 0, c->stream);
    k_unpack_aj<<<nblk(n_act), 256, 0, c->stream>>>((int)n_act, g.raw_a, g.raw_j, stage);
    c->launches += 6;
    double *dst[7] = {ax, ay, az, jx, jy, jz, pot};
    for (int k = 0; k < 7; k++)
      if (dst[k]) TRY(cudaMemcpyAsync(dst[k], stage + k * na, na * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    TRY(cudaStreamSynchronize(c->stream));
    TRY(cudaGetLastError());
  }
#undef TRY
  if (e != cudaSuccess) rc = fail(c, AL26_ECUDA, "al26_grav_force: %s", cudaGetErrorString(e));
  void *ptrs[] = {stage, g.jpos, g.jvel, g.list, d_idx, g.part_a, g.part_j, g.raw_a, g.raw_j, g.ctrl, g.hdr, d_tab};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  return rc;
}

int al26_last_device_ms(al26_ctx *c, double *ms, int64_t *kernel_launches) {
  if (!c) return AL26_EINVAL;
  if (ms) *ms = c->last_ms;
  if (kernel_launches) *kernel_launches = c->last_launches;
  return 0;
}

int al26_enrich_last_kernel_ms(al26_ctx *c, double *ms) {
  if (!c || !ms) return AL26_EINVAL;
  *ms = c->last_kernel_ms;
  return 0;
}

int al26_grav_bench_force_n(al26_ctx *c, int64_t n_act, int reps, double *avg_ms, int64_t *pairs_per_eval) {
  if (!c) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "bench_force before commit");
  if (c->in_evolve) return fail(c, AL26_ESTATE, "bench_force during evolve");
  if (reps < 1) return fail(c, AL26_EINVAL, "reps must be >= 1");
  if (n_act <= 0 || n_act > c->g.n_loc) n_act = c->g.n_loc;
  CU(cudaSetDevice(c->device));
  c->g.eps2 = c->eps2;
  const int64_t l0 = c->launches;
  reset_ctrl(c, 0);
  c->launches += launch_predict_list(c->g, MODE_INIT, 0, c->stream);  // list = all, jpos = current
  int rc = gather_j(c);
  if (rc) return rc;
  k_set_nact<<<1, 1, 0, c->stream>>>(&c->g.ctrl[0], (int)n_act);
  c->launches += 1 + launch_force(c->g, 0, c->stream);  // warm-up
  CU(cudaEventRecord(c->ev0, c->stream));
  for (int r = 0; r < reps; r++) {
    k_reset_work<<<1, 1, 0, c->stream>>>(&c->g.ctrl[0]);
    c->launches += 1 + launch_force(c->g, 0, c->stream);
  }
  CU(cudaEventRecord(c->ev1, c->stream));
  CU(cudaEventSynchronize(c->ev1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  CU(cudaGetLastError());
  if (avg_ms) *avg_ms = (double)ms / reps;
  if (pairs_per_eval) *pairs_per_eval = n_act * (int64_t)c->g.n_tot;
  c->last_ms = ms;
  c->last_launches = c->launches - l0;
  return 0;
}

int al26_grav_bench_force(al26_ctx *c, int reps, double *avg_ms, int64_t *pairs_per_eval) {
  return al26_grav_bench_force_n(c, 0, reps, avg_ms, pairs_per_eval);
}

int al26_dbg_decomposition(int n_act, int n_tot, int sm_count, int variant, int big_nact, int64_t *out8) {
  // host-only (no device, no context): the work decomposition the kernels would use, for CPU-side invariant tests
  int minb = 2, ipt = 2;
  if (!out8 || n_act < 1 || n_tot < 1 || n_act > n_tot || sm_count < 1 || big_nact < 33) return AL26_EINVAL;
  if (force_variant_info(variant, &minb, &ipt)) return AL26_EINVAL;
  const int grid = minb * sm_count;
  std::vector<int> tab(decomp_table_entries(n_tot, ipt, big_nact));
  fill_decomp_table(tab.data(), n_tot, n_tot, grid, ipt, big_nact);
  const Decomp d = make_decomp(n_act, n_tot, tab.data(), ipt, big_nact);
  out8[0] = d.ipt; out8[1] = d.ti; out8[2] = d.n_itiles; out8[3] = d.n_jsplit; out8[4] = d.jchunk;
  out8[5] = d.slot_stride; out8[6] = part_capacity(n_tot, grid); out8[7] = grid;
  return 0;
}

int al26_dbg_engine_plan(int n, int max_smem_per_block, int *cluster_size, int *particles_per_cta, int *smem_bytes) {
  if (!cluster_size || !particles_per_cta || !smem_bytes || n < 1 || max_smem_per_block < 0) return AL26_EINVAL;
  int cs = 0, p_cap = 0;
  if (!engine_plan(n, max_smem_per_block, &cs, &p_cap)) {
    *cluster_size = *particles_per_cta = *smem_bytes = 0;
    return 0;
  }
  *cluster_size = cs;
  *particles_per_cta = p_cap;
  *smem_bytes = engine_smem_bytes(p_cap);
  return 0;
}

int al26_dbg_chip_plan(int n, int n_ctas, int max_smem_per_block, int *particles_per_cta, int *smem_bytes, int64_t *mail_bytes) {
  if (!particles_per_cta || !smem_bytes || !mail_bytes || n < 1 || n_ctas < 1 || max_smem_per_block < 0) return AL26_EINVAL;
  int p_cap = 0;
  if (!chip_plan(n, n_ctas, max_smem_per_block, &p_cap)) {
    *particles_per_cta = *smem_bytes = 0;
    *mail_bytes = 0;
    return 0;
  }
  *particles_per_cta = p_cap;
  *smem_bytes = chip_smem_bytes(p_cap);
  *mail_bytes = (int64_t)chip_mail_bytes(n_ctas);
  return 0;
}

int al26_set_force_variant(al26_ctx *c, int variant) {
  if (!c) return AL26_EINVAL;
  if (variant < 0 || variant >= force_variant_count()) return fail(c, AL26_EINVAL, "force variant %d out of range", variant);
  if (c->in_evolve) return fail(c, AL26_ESTATE, "set_force_variant during evolve");
  c->force_variant = variant;  // takes effect at the next commit (buffers and graph depend on it)
  return 0;
}

int al26_set_big_block(al26_ctx *c, int n_act_min) {
  if (!c) return AL26_EINVAL;
  if (n_act_min < 33 || n_act_min > (1 << 20)) return fail(c, AL26_EINVAL, "big-block threshold %d out of range", n_act_min);
  if (c->in_evolve) return fail(c, AL26_ESTATE, "set_big_block during evolve");
  c->big_nact = n_act_min;  // takes effect at the next commit
  return 0;
}

int al26_set_decomposition(al26_ctx *c, int max_rounds, double item_overhead_pairs) {
  if (!c) return AL26_EINVAL;
  if (max_rounds < 1 || max_rounds > FORCE_MAX_ROUNDS_CAP || !(item_overhead_pairs >= 0.0))
    return fail(c, AL26_EINVAL, "decomposition: max_rounds in [1, %d], overhead >= 0", FORCE_MAX_ROUNDS_CAP);
  if (c->in_evolve) return fail(c, AL26_ESTATE, "set_decomposition during evolve");
  c->max_rounds = max_rounds;  // takes effect at the next commit
  c->item_overhead = item_overhead_pairs;
  return 0;
}

int al26_set_fuse_max(al26_ctx *c, int n_act_max) {
  if (!c) return AL26_EINVAL;
  if (n_act_max < -1 || n_act_max > FUSE_CAP) return fail(c, AL26_EINVAL, "fuse_max in [-1, %d]", FUSE_CAP);
  if (c->in_evolve) return fail(c, AL26_ESTATE, "set_fuse_max during evolve");
  c->fuse_max = n_act_max;  // takes effect at the next commit
  return 0;
}

int al26_grav_fused_steps(al26_ctx *c, int64_t *n_fused) {
  if (!c || !n_fused) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "fused_steps before commit");
  CU(cudaSetDevice(c->device));
  int rc = read_header(c);
  if (rc) return rc;
  *n_fused = c->h_hdr->n_fused;
  return 0;
}

int al26_grav_fuse_profile(al26_ctx *c, int64_t *ns8) {  // 16 values
  if (!c || !ns8) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "fuse_profile before commit");
  CU(cudaSetDevice(c->device));
  int rc = read_header(c);
  if (rc) return rc;
  for (int k = 0; k < 16; k++) ns8[k] = c->h_hdr->fuse_ns[k];
  return 0;
}

int al26_set_step_mode(al26_ctx *c, int mode) {
  if (!c) return AL26_EINVAL;
  if (mode < -1 || mode > 3)
    return fail(c, AL26_EINVAL, "step mode must be -1 (automatic), 0 (graph), 1 (persistent loop), 2 (graph + cluster engine) or 3 (graph + chip engine)");
  if (c->in_evolve) return fail(c, AL26_ESTATE, "set_step_mode during evolve");
  if (mode != c->step_mode) c->graph_stale = c->committed;  // the graph may gain / lose the engine nodes
  c->step_mode = mode;
  return 0;
}

int al26_set_chip_max(al26_ctx *c, int n_act_max) {
  if (!c) return AL26_EINVAL;
  if (n_act_max < -1 || n_act_max > CHIP_CAP) return fail(c, AL26_EINVAL, "chip_max must be -1 (default), 0 (off) or 1..%d", CHIP_CAP);
  if (c->in_evolve) return fail(c, AL26_ESTATE, "set_chip_max during evolve");
  if (n_act_max != c->chip_max) c->graph_stale = c->committed;
  c->chip_max = n_act_max;
  return 0;
}

int al26_grav_chip_steps(al26_ctx *c, int64_t *n_chip, int *n_ctas, int *n_act_max) {
  if (!c || !n_chip) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "chip_steps before commit");
  CU(cudaSetDevice(c->device));
  int rc = read_header(c);
  if (rc) return rc;
  *n_chip = c->h_hdr->n_chip;
  if (n_ctas) *n_ctas = c->chip_on ? c->g.chip_n : 0;
  if (n_act_max) *n_act_max = c->chip_on ? c->g.chip_max : 0;
  return 0;
}

int al26_grav_engine_steps(al26_ctx *c, int64_t *n_engine, int *cluster_size) {
  if (!c || !n_engine) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "engine_steps before commit");
  CU(cudaSetDevice(c->device));
  int rc = read_header(c);
  if (rc) return rc;
  *n_engine = c->h_hdr->n_engine;
  if (cluster_size) *cluster_size = c->engine_on ? c->engine_cs : 0;
  return 0;
}

int al26_grav_block_histogram(al26_ctx *c, int64_t *hist32) {
  if (!c || !hist32) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "block_histogram before commit");
  CU(cudaSetDevice(c->device));
  int rc = read_header(c);
  if (rc) return rc;
  for (int b = 0; b < 32; b++) hist32[b] = c->h_hdr->nact_hist[b];
  return 0;
}

int al26_dist_profile(al26_ctx *c, int64_t *out12) {
  if (!c || !out12) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "dist_profile before commit");
  CU(cudaSetDevice(c->device));
  int rc = read_header(c);
  if (rc) return rc;
  for (int k = 0; k < DIST_PROF_N; k++) out12[k] = c->h_hdr->dist_prof[k];
  return 0;
}

int al26_grav_loop_profile(al26_ctx *c, int64_t *cycles6) {
  if (!c || !cycles6) return AL26_EINVAL;
  if (!c->committed) return fail(c, AL26_ESTATE, "loop_profile before commit");
  CU(cudaSetDevice(c->device));
  int rc = read_header(c);
  if (rc) return rc;
  for (int k = 0; k < 6; k++) cycles6[k] = c->h_hdr->loop_cycles[k];
  return 0;
}

static int bench_fp64(al26_ctx *c, double *tflops, bool with_mufu) {
  if (!c || !tflops) return AL26_EINVAL;
  CU(cudaSetDevice(c->device));
  int rc = ensure_scratch(c, 1024);
  if (rc) return rc;
  (with_mufu ? launch_dfma_mufu_mix : launch_dfma_peak)(c->sm_count, 2000, c->scratch, c->stream);  // warm-up
  double best = 0.0;
  for (int r = 0; r < 5; r++) {
    CU(cudaEventRecord(c->ev0, c->stream));
    const double flops = (with_mufu ? launch_dfma_mufu_mix : launch_dfma_peak)(c->sm_count, 20000, c->scratch, c->stream);
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
    c->launches++;
  }
  CU(cudaGetLastError());
  *tflops = best;
  return 0;
}

int al26_bench_fp64_rate(al26_ctx *c, int variant, double *lane_inst_per_s) {
  if (!c || !lane_inst_per_s || variant < 0 || variant > 5) return AL26_EINVAL;
  CU(cudaSetDevice(c->device));
  int rc = ensure_scratch(c, 1024);
  if (rc) return rc;
  launch_fp64_rate(variant, c->sm_count, 2000, c->scratch, c->stream);  // warm-up
  double best = 0.0;
  for (int r = 0; r < 5; r++) {
    CU(cudaEventRecord(c->ev0, c->stream));
    const double inst = launch_fp64_rate(variant, c->sm_count, 20000, c->scratch, c->stream);
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    const double rate = inst / (ms * 1e-3);
    if (rate > best) best = rate;
    c->launches++;
  }
  CU(cudaGetLastError());
  *lane_inst_per_s = best;
  return 0;
}

int al26_bench_fp64_peak(al26_ctx *c, double *tflops) { return bench_fp64(c, tflops, false); }
int al26_bench_fp64_with_rsqrt(al26_ctx *c, double *tflops) { return bench_fp64(c, tflops, true); }

int al26_local_densities(al26_ctx *c, int64_t n, const double *x_pc, const double *y_pc, const double *z_pc,
                         const double *mass_msun, double *rho) {
  if (!c) return AL26_EINVAL;
  if (n < 11 || n > 0x7fffffff / 2) return fail(c, AL26_EINVAL, "local_densities needs at least 11 stars (10 neighbours), got %lld", (long long)n);
  if (!x_pc || !y_pc || !z_pc || !mass_msun || !rho) return fail(c, AL26_EINVAL, "null array");
  CU(cudaSetDevice(c->device));
  const size_t nt = (size_t)n;
  double *d = nullptr;
  CU(cudaMalloc(&d, 5 * nt * sizeof(double)));
  cudaError_t e = cudaSuccess;
  const double *src[4] = {x_pc, y_pc, z_pc, mass_msun};
  for (int k = 0; k < 4 && e == cudaSuccess; k++) e = cudaMemcpyAsync(d + k * nt, src[k], nt * sizeof(double), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    c->launches += launch_local_density((int)n, d, d + nt, d + 2 * nt, d + 3 * nt, d + 4 * nt, c->stream);
    e = cudaMemcpyAsync(rho, d + 4 * nt, nt * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(d);
  if (e != cudaSuccess) return fail(c, AL26_ECUDA, "al26_local_densities: %s", cudaGetErrorString(e));
  return 0;
}

// ---- enrichment ---------------------------------------------------------------------------

int al26_enrich_commit(al26_ctx *c, int64_t n, const double *r_disk_km, const double *tau_disk_myr,
                       const uint8_t *disk_alive, const uint8_t *kicked, const double *wr26, const double *wr60,
                       const double *sn26, const double *sn60) {
  if (!c) return AL26_EINVAL;
  if (n <= 0 || n > 0x7fffffff / 2) return fail(c, AL26_EINVAL, "star count %lld out of range", (long long)n);
  if (!r_disk_km || !tau_disk_myr || !disk_alive || !kicked || !wr26 || !wr60 || !sn26 || !sn60)
    return fail(c, AL26_EINVAL, "null array");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  free_enrich(c);
  int64_t d0, nloc;
  slice_of(n, c->rank, c->world, d0, nloc);
  const size_t nt = (size_t)n, nl = (size_t)nloc;
  CU(cudaMalloc(&c->e_glob, 12 * nt * sizeof(double)));
  CU(cudaMalloc(&c->e_loc, 20 * nl * sizeof(double)));  // r_disk, tau, inv[8], fin[8], agb_raw[2]
  CU(cudaMalloc(&c->e_flags, nt + nl));
  const size_t n_ints = ENR_NCOUNTERS + 29 * (size_t)ENR_MAX_SOURCES + ((size_t)ENR_GRID_CELLS + 1);
  CU(cudaMalloc(&c->e_ints, n_ints * sizeof(int)));
  CU(cudaMemsetAsync(c->e_ints, 0, n_ints * sizeof(int), c->stream));
  CU(cudaMalloc(&c->e_src, 4 * (size_t)ENR_MAX_SOURCES * sizeof(double4)));
  CU(cudaMalloc(&c->e_dbl, (5 * (size_t)ENR_MAX_SOURCES + 16 + 8) * sizeof(double)));
  CU(cudaMemsetAsync(c->e_dbl, 0, (5 * (size_t)ENR_MAX_SOURCES + 16 + 8) * sizeof(double), c->stream));
  EnrichDev &e = c->e;
  e.n_tot = (int)n; e.d0 = (int)d0; e.n_loc = (int)nloc;
  double *G = c->e_glob;
  e.mass_msun = G; e.mdot = G + nt;
  e.px = nullptr; e.py = e.pz = e.pvx = e.pvy = e.pvz = nullptr;
  e.pv_off = 0;
  e.wr26 = G + 8 * nt; e.wr60 = G + 9 * nt; e.sn26 = G + 10 * nt; e.sn60 = G + 11 * nt;
  double *Lc = c->e_loc;
  e.r_disk = Lc; e.tau_disk = Lc + nl; e.inv = Lc + 2 * nl; e.fin = Lc + 10 * nl;
  e.kicked = c->e_flags; e.alive = c->e_flags + nt;
  e.counters = c->e_ints;
  e.hm_list = c->e_ints + ENR_NCOUNTERS;
  e.sn_events = e.hm_list + ENR_MAX_SOURCES;
  e.cell_items = e.sn_events + ENR_MAX_SOURCES;
  e.cell_start = e.cell_items + 27 * ENR_MAX_SOURCES;
  e.src_a = c->e_src; e.src_b = c->e_src + ENR_MAX_SOURCES; e.src_f = c->e_src + 2 * ENR_MAX_SOURCES;
  e.ev_a = c->e_src + 3 * ENR_MAX_SOURCES;
  e.ev_b = c->e_dbl; e.fsum = c->e_dbl + ENR_MAX_SOURCES; e.hm_rows = c->e_dbl + ENR_MAX_SOURCES + 16;
  e.prof = reinterpret_cast<long long *>(c->e_dbl + 5 * (size_t)ENR_MAX_SOURCES + 16);
  CU(cudaMemcpyAsync(G + 8 * nt, wr26, nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(G + 9 * nt, wr60, nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(G + 10 * nt, sn26, nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(G + 11 * nt, sn60, nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(Lc, r_disk_km + d0, nl * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(Lc + nl, tau_disk_myr + d0, nl * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(Lc + 2 * nl, 0, 18 * nl * sizeof(double), c->stream));
  CU(cudaMemcpyAsync(c->e_flags, kicked, nt, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->e_flags + nt, disk_alive + d0, nl, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->e_committed = true;
  return 0;
}

int al26_enrich_set_inventories(al26_ctx *c, int64_t n, const double *inv, const double *fin) {
  if (!c) return AL26_EINVAL;
  if (!c->e_committed) return fail(c, AL26_ESTATE, "enrich_set_inventories before enrich_commit");
  if (n != c->e.n_tot) return fail(c, AL26_EINVAL, "size mismatch");
  CU(cudaSetDevice(c->device));
  const size_t nl = (size_t)c->e.n_loc;
  for (int r = 0; r < ENR_NINV; r++) {
    if (inv) CU(cudaMemcpyAsync(c->e.inv + r * nl, inv + (size_t)r * n + c->e.d0, nl * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (fin) CU(cudaMemcpyAsync(c->e.fin + r * nl, fin + (size_t)r * n + c->e.d0, nl * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int al26_enrich_set_units(al26_ctx *c, double km_per_length, double kms_per_speed) {
  if (!c) return AL26_EINVAL;
  if (!(km_per_length > 0.0) || !(kms_per_speed > 0.0)) return fail(c, AL26_EINVAL, "unit factors must be positive");
  c->km_per_length = km_per_length;
  c->kms_per_speed = kms_per_speed;
  return 0;
}

int al26_enrich_profile(al26_ctx *c, int64_t *cycles8) {
  if (!c || !cycles8) return AL26_EINVAL;
  if (!c->e_committed) return fail(c, AL26_ESTATE, "enrich_profile before enrich_commit");
  CU(cudaSetDevice(c->device));
  long long h[8];
  CU(cudaMemcpy(h, c->e.prof, sizeof(h), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 8; k++) cycles8[k] = h[k];
  return 0;
}

int al26_enrich_set_mode(al26_ctx *c, int mode) {
  if (!c) return AL26_EINVAL;
  if (mode < 0 || mode > 2) return fail(c, AL26_EINVAL, "enrichment mode must be 0 (exact), 1 (fast) or 2 (fast, pruned)");
  c->enrich_mode = mode;
  return 0;
}

int al26_enrich_step(al26_ctx *c, int64_t n, const double *mass_msun, const double *mdot, const double *pos_vel,
                     double dt_s, double t_new_myr, double r_bub_local_km, double r_bub_global_km, double decay26,
                     double decay60, int with_agb, int32_t *sn_events, int64_t sn_cap, int64_t *n_sn_events) {
  if (!c) return AL26_EINVAL;
  if (!c->e_committed) return fail(c, AL26_ESTATE, "enrich_step before enrich_commit");
  if (n != c->e.n_tot || !mass_msun || !mdot) return fail(c, AL26_EINVAL, "enrich_step: size mismatch or null array");
  if (!(r_bub_local_km > 0.0) || !(r_bub_global_km > 0.0)) return fail(c, AL26_EINVAL, "bubble radii must be positive");
  if (!pos_vel && (!c->committed || c->n_tot != n))
    return fail(c, AL26_ESTATE, "enrich_step with pos_vel == NULL needs a committed gravity state of the same size");
  if (c->in_evolve) return fail(c, AL26_ESTATE, "enrich_step during evolve");
  CU(cudaSetDevice(c->device));
  const int64_t l0 = c->launches;
  CU(cudaEventRecord(c->ev0, c->stream));
  const size_t nt = (size_t)n;
  EnrichDev &e = c->e;
  double *G = c->e_glob;
  EnrichParams p;
  p.dt_s = dt_s; p.t_new_myr = t_new_myr;
  p.r_local = r_bub_local_km;
  p.r_local3 = r_bub_local_km * (r_bub_local_km * r_bub_local_km);
  p.r_global3 = r_bub_global_km * (r_bub_global_km * r_bub_global_km);
  // smallest double q with sqrt(q) >= R: `R <= sqrt(d2)` <=> `d2 >= q` (sqrt is monotone and correctly rounded)
  double q = r_bub_local_km * r_bub_local_km;
  while (sqrt(q) >= r_bub_local_km) q = nextafter(q, -INFINITY);
  while (sqrt(q) < r_bub_local_km) q = nextafter(q, INFINITY);
  p.q_local = q;
  p.decay26 = decay26; p.decay60 = decay60; p.with_agb = with_agb; p.mode = c->enrich_mode;
  CU(cudaMemsetAsync(e.counters, 0, ENR_NCOUNTERS * sizeof(int), c->stream));
  CU(cudaMemcpyAsync(G, mass_msun, nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  // Several ranks, explicit positions: every rank needs all N masses (classification) but only ITS discs' kinematics
  // and the massive stars' rows -- classify on the device, read the (short, sorted) list back, gather those stars'
  // mdot / x / y / z on the host (an index gather, no arithmetic) and upload them with the slice:
  // 8 N + 48 N / P bytes per rank instead of 64 N.
  const bool sliced = pos_vel && c->world > 1;
  bool tables_only = false;
  float ms_cls = 0.f;
  if (sliced) {
    CU(cudaEventRecord(c->ev2, c->stream));
    c->launches += launch_enrich_classify(e, p, c->sm_count, c->stream);
    CU(cudaEventRecord(c->ev3, c->stream));
    CU(cudaMemcpyAsync(c->h_events, e.counters, ENR_NCOUNTERS * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_events + ENR_NCOUNTERS, e.hm_list, ENR_MAX_SOURCES * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventElapsedTime(&ms_cls, c->ev2, c->ev3));
    if (c->h_events[0] > ENR_MAX_SOURCES) return fail(c, AL26_ECAP, "more than %d massive stars", ENR_MAX_SOURCES);
    const int n_hm = c->h_events[3];
    const int *list = c->h_events + ENR_NCOUNTERS;
    for (int k = 0; k < n_hm; k++) {
      const size_t i = (size_t)list[k];
      c->h_rows[k] = mdot[i];
      for (int d = 0; d < 3; d++) c->h_rows[(size_t)(d + 1) * ENR_MAX_SOURCES + k] = pos_vel[(size_t)d * nt + i];
    }
    if (n_hm > 0)
      for (int r = 0; r < 4; r++)
        CU(cudaMemcpyAsync(e.hm_rows + (size_t)r * ENR_MAX_SOURCES, c->h_rows + (size_t)r * ENR_MAX_SOURCES,
                           (size_t)n_hm * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const size_t nl = (size_t)e.n_loc;
    for (int r = 0; r < 6; r++)
      CU(cudaMemcpyAsync(G + (2 + r) * nt, pos_vel + (size_t)r * nt + e.d0, nl * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    e.px = G + 2 * nt; e.py = G + 3 * nt; e.pz = G + 4 * nt;
    e.pvx = G + 5 * nt; e.pvy = G + 6 * nt; e.pvz = G + 7 * nt;
    e.pv_off = e.d0;
    e.gpos = e.gvel = nullptr;
    tables_only = true;
  } else {
    CU(cudaMemcpyAsync(G + nt, mdot, nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    e.pv_off = 0;
    if (pos_vel) {
      CU(cudaMemcpyAsync(G + 2 * nt, pos_vel, 6 * nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      e.px = G + 2 * nt; e.py = G + 3 * nt; e.pz = G + 4 * nt;
      e.pvx = G + 5 * nt; e.pvy = G + 6 * nt; e.pvz = G + 7 * nt;
      e.gpos = e.gvel = nullptr;
    } else {
      c->launches += launch_snapshot_j(c->g, c->stream);
      int rc = gather_j(c);
      if (rc) return rc;
      e.px = e.py = e.pz = e.pvx = e.pvy = e.pvz = nullptr;
      e.gpos = c->g.jpos;
      e.gvel = c->g.jvel;
      e.km_per_length = c->km_per_length;
      e.kms_per_speed = c->kms_per_speed;
    }
  }
  CU(cudaEventRecord(c->ev2, c->stream));
  cudaError_t le = cudaSuccess;
  c->launches += launch_enrich(e, p, c->sm_count, tables_only, c->stream, &le);
  if (le != cudaSuccess) return fail(c, AL26_ECUDA, "enrichment launch failed: %s", cudaGetErrorString(le));
  CU(cudaEventRecord(c->ev3, c->stream));
  CU(cudaMemcpyAsync(c->h_events, e.counters, ENR_NCOUNTERS * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(c->h_events + ENR_NCOUNTERS, e.sn_events, ENR_MAX_SOURCES * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaEventRecord(c->ev1, c->stream));
  CU(cudaEventSynchronize(c->ev1));
  CU(cudaGetLastError());
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->last_ms = ms;
  CU(cudaEventElapsedTime(&ms, c->ev2, c->ev3));
  c->last_kernel_ms = ms + ms_cls;
  c->last_launches = c->launches - l0;
  // capacity: detected by the source kernel BEFORE anything was mutated (the disc kernel returned at once)
  if (c->h_events[2]) return fail(c, AL26_ECAP, "more than %d massive stars", ENR_MAX_SOURCES);
  const int ne = c->h_events[1];
  if (n_sn_events) *n_sn_events = ne;
  if (ne > 0) {
    if (!sn_events || sn_cap < ne) return fail(c, AL26_ECAP, "%d supernova events exceed sn_cap %lld", ne, (long long)sn_cap);
    memcpy(sn_events, c->h_events + ENR_NCOUNTERS, (size_t)ne * sizeof(int));
  }
  return 0;
}

int al26_enrich_interloper(al26_ctx *c, int64_t n, const double *mass_msun, const double *pos_old_pc,
                           const double *pos_new_pc, int64_t interloper_index, double r_test_pc, double r_bub_km,
                           double km_per_pc, double rate26_kg_s, double rate60_kg_s, double dt_s) {
  if (!c) return AL26_EINVAL;
  if (!c->e_committed) return fail(c, AL26_ESTATE, "enrich_interloper before enrich_commit");
  if (n != c->e.n_tot || !mass_msun || !pos_old_pc || !pos_new_pc) return fail(c, AL26_EINVAL, "enrich_interloper: size mismatch or null array");
  if (interloper_index < 0 || interloper_index >= n) return fail(c, AL26_EINVAL, "interloper index out of range");
  if (!(r_test_pc > 0.0) || !(r_bub_km > 0.0) || !(km_per_pc > 0.0)) return fail(c, AL26_EINVAL, "radii must be positive");
  CU(cudaSetDevice(c->device));
  const size_t nt = (size_t)n, nl = (size_t)c->e.n_loc;
  double *G = c->e_glob;
  CU(cudaMemcpyAsync(G, mass_msun, nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(G + 2 * nt, pos_old_pc, 3 * nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(G + 5 * nt, pos_new_pc, 3 * nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  InterloperParams p;
  p.k_int = (int)interloper_index;
  p.old_pc = G + 2 * nt;
  p.new_pc = G + 5 * nt;
  // largest q with sqrt(q) <= r: `sqrt(d2) <= r` <=> `d2 <= q`
  double q = r_test_pc * r_test_pc;
  while (sqrt(q) > r_test_pc) q = nextafter(q, -INFINITY);
  while (sqrt(nextafter(q, INFINITY)) <= r_test_pc) q = nextafter(q, INFINITY);
  p.q_test = q;
  p.r_bub3 = r_bub_km * (r_bub_km * r_bub_km);
  p.km_per_pc = km_per_pc; p.rate26 = rate26_kg_s; p.rate60 = rate60_kg_s; p.dt_s = dt_s;
  p.raw = c->e_loc + 18 * nl;
  c->launches += launch_interloper(c->e, p, c->stream);
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return 0;
}

int al26_enrich_get_agb_raw(al26_ctx *c, int64_t n, double *raw) {
  if (!c || !raw) return AL26_EINVAL;
  if (!c->e_committed) return fail(c, AL26_ESTATE, "enrich_get_agb_raw before enrich_commit");
  if (n != c->e.n_tot) return fail(c, AL26_EINVAL, "size mismatch");
  CU(cudaSetDevice(c->device));
  const size_t nl = (size_t)c->e.n_loc;
  for (int r = 0; r < 2; r++)
    CU(cudaMemcpyAsync(raw + (size_t)r * n + c->e.d0, c->e_loc + (18 + r) * nl, nl * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int al26_enrich_get(al26_ctx *c, int64_t n, double *inv, double *fin, uint8_t *disk_alive, uint8_t *kicked) {
  if (!c) return AL26_EINVAL;
  if (!c->e_committed) return fail(c, AL26_ESTATE, "enrich_get before enrich_commit");
  if (n != c->e.n_tot) return fail(c, AL26_EINVAL, "size mismatch");
  CU(cudaSetDevice(c->device));
  const EnrichDev &e = c->e;
  const size_t nl = (size_t)e.n_loc;
  for (int r = 0; r < ENR_NINV; r++) {
    if (inv) CU(cudaMemcpyAsync(inv + (size_t)r * n + e.d0, e.inv + r * nl, nl * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (fin) CU(cudaMemcpyAsync(fin + (size_t)r * n + e.d0, e.fin + r * nl, nl * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  if (disk_alive) CU(cudaMemcpyAsync(disk_alive + e.d0, e.alive, nl, cudaMemcpyDeviceToHost, c->stream));
  if (kicked) CU(cudaMemcpyAsync(kicked, e.kicked, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// al26_group: several GPUs driven from ONE host process -- what `ph4(converter, number_of_workers=k)` is to the
// reference script (al26_nbody.py:57,1711-1720: one Python process, k MPI worker ranks behind one object).  k contexts,
// one per GPU, each on its own persistent host thread; every call fans out to the threads and joins.  The ranks run the
// peer-memory protocol (replicated state, owner-computes, corrected particles stored into the peers' slabs over NVLink,
// DESIGN.md section 5) on plain peer pointers (cudaDeviceEnablePeerAccess; no IPC, no NCCL).
// ------------------------------------------------------------------------------------------
struct al26_group {
  int n = 0;
  std::vector<int> dev;
  std::vector<al26_ctx *> ctx;
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::function<int(int)> job;
  unsigned long long job_id = 0;
  int pending = 0;
  std::vector<int> rc;
  bool quit = false;
  std::string err;
  int64_t n_tot = 0;
};

namespace {
std::string g_group_error;

void group_worker(al26_group *g, int r) {
  cudaSetDevice(g->dev[r]);
  unsigned long long seen = 0;
  while (true) {
    std::function<int(int)> fn;
    {
      std::unique_lock<std::mutex> lk(g->mu);
      g->cv_job.wait(lk, [&] { return g->quit || g->job_id != seen; });
      if (g->quit) return;
      seen = g->job_id;
      fn = g->job;
    }
    const int rc = fn(r);
    {
      std::lock_guard<std::mutex> lk(g->mu);
      g->rc[r] = rc;
      if (--g->pending == 0) g->cv_done.notify_all();
    }
  }
}

// run fn(rank) on every GPU's thread at the same time; first failing rank's code and message
int group_run(al26_group *g, std::function<int(int)> fn) {
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->job = std::move(fn);
    g->pending = g->n;
    g->job_id++;
  }
  g->cv_job.notify_all();
  {
    std::unique_lock<std::mutex> lk(g->mu);
    g->cv_done.wait(lk, [&] { return g->pending == 0; });
  }
  for (int r = 0; r < g->n; r++)
    if (g->rc[r]) {
      g->err = "GPU " + std::to_string(g->dev[r]) + ": " + (g->ctx[r] ? g->ctx[r]->err : std::string("no context"));
      return g->rc[r];
    }
  return 0;
}
int gfail(al26_group *g, int code, const char *msg) {
  if (g) g->err = msg;
  else g_group_error = msg;
  return code;
}
}  // namespace

extern "C" {

al26_group *al26_group_create(int n_gpus, const int *device_ids) {
  if (n_gpus < 1 || n_gpus > MAX_PEERS) {
    gfail(nullptr, AL26_EINVAL, "al26_group_create: between 1 and 8 GPUs");
    return nullptr;
  }
  al26_group *g = new al26_group();
  g->n = n_gpus;
  g->dev.resize(n_gpus);
  g->ctx.assign(n_gpus, nullptr);
  g->rc.assign(n_gpus, 0);
  for (int r = 0; r < n_gpus; r++) g->dev[r] = device_ids ? device_ids[r] : r;
  for (int r = 0; r < n_gpus; r++) {
    g->ctx[r] = al26_create(g->dev[r]);
    if (!g->ctx[r] || al26_dist_init_local(g->ctx[r], r, n_gpus) != 0) {
      g_group_error = std::string("al26_group_create: GPU ") + std::to_string(g->dev[r]) + ": " +
                      (g->ctx[r] ? g->ctx[r]->err : g_create_error);
      for (int q = 0; q <= r; q++)
        if (g->ctx[q]) al26_destroy(g->ctx[q]);
      delete g;
      return nullptr;
    }
  }
  for (int r = 0; r < n_gpus; r++) g->th.emplace_back(group_worker, g, r);
  return g;
}

void al26_group_destroy(al26_group *g) {
  if (!g) return;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->quit = true;
  }
  g->cv_job.notify_all();
  for (auto &t : g->th) t.join();
  for (al26_ctx *c : g->ctx)
    if (c) al26_destroy(c);
  delete g;
}

const char *al26_group_last_error(al26_group *g) { return g ? g->err.c_str() : g_group_error.c_str(); }
int al26_group_size(al26_group *g) { return g ? g->n : 0; }
al26_ctx *al26_group_ctx(al26_group *g, int rank) { return (g && rank >= 0 && rank < g->n) ? g->ctx[rank] : nullptr; }

int al26_group_grav_set_params(al26_group *g, double eps2, double eta, double dt_max, double dt_min) {
  if (!g) return AL26_EINVAL;
  return group_run(g, [=](int r) { return al26_grav_set_params(g->ctx[r], eps2, eta, dt_max, dt_min); });
}

int al26_group_grav_set_reinit_policy(al26_group *g, int policy) {
  if (!g) return AL26_EINVAL;
  return group_run(g, [=](int r) { return al26_grav_set_reinit_policy(g->ctx[r], policy); });
}

int al26_group_grav_commit(al26_group *g, int64_t n, const double *m, const double *x, const double *y, const double *z,
                           const double *vx, const double *vy, const double *vz) {
  if (!g) return AL26_EINVAL;
  int rc = group_run(g, [=](int r) { return al26_grav_commit(g->ctx[r], n, m, x, y, z, vx, vy, vz); });
  if (rc || g->n == 1) {
    if (!rc) g->n_tot = n;
    return rc;
  }
  std::vector<void *> slabs(g->n, nullptr);
  for (int r = 0; r < g->n; r++)
    if ((rc = al26_dist_p2p_local_slab(g->ctx[r], &slabs[r]))) {
      g->err = g->ctx[r]->err;
      return rc;
    }
  rc = group_run(g, [&](int r) { return al26_dist_p2p_attach(g->ctx[r], slabs.data(), g->dev.data(), g->n); });
  if (!rc) g->n_tot = n;
  return rc;
}

int al26_group_grav_set_mass(al26_group *g, int64_t n, const double *m) {
  if (!g) return AL26_EINVAL;
  return group_run(g, [=](int r) { return al26_grav_set_mass(g->ctx[r], n, m); });
}

int al26_group_grav_set_time(al26_group *g, double t) {
  if (!g) return AL26_EINVAL;
  return group_run(g, [=](int r) { return al26_grav_set_time(g->ctx[r], t); });
}

int al26_group_grav_get_time(al26_group *g, double *t) { return g ? al26_grav_get_time(g->ctx[0], t) : AL26_EINVAL; }

int al26_group_grav_evolve(al26_group *g, double t_end, int64_t *n_block_steps, int64_t *n_pairs) {
  if (!g) return AL26_EINVAL;
  std::vector<int64_t> steps(g->n, 0), pairs(g->n, 0);
  const int rc = group_run(g, [&](int r) { return al26_grav_evolve(g->ctx[r], t_end, &steps[r], &pairs[r]); });
  if (rc) return rc;
  int64_t tot = 0;
  for (int r = 0; r < g->n; r++) tot += pairs[r];
  if (n_block_steps) *n_block_steps = steps[0];  // every rank takes every block step
  if (n_pairs) *n_pairs = tot;                   // each rank counts the pairs of the particles it owns
  return 0;
}

int al26_group_grav_get_state(al26_group *g, int64_t n, double *m, double *x, double *y, double *z, double *vx, double *vy,
                              double *vz) {
  if (!g) return AL26_EINVAL;
  const int rc = al26_grav_get_state(g->ctx[0], n, m, x, y, z, vx, vy, vz);  // the state is replicated: rank 0 has it all
  if (rc) g->err = g->ctx[0]->err;
  return rc;
}

int al26_group_grav_energies(al26_group *g, double *kinetic, double *potential, double *sum_mm_over_r) {
  if (!g) return AL26_EINVAL;
  std::vector<double> k(g->n, 0.0), u(g->n, 0.0), s(g->n, 0.0);
  const int rc = group_run(g, [&](int r) { return al26_grav_energies(g->ctx[r], &k[r], &u[r], &s[r]); });
  if (rc) return rc;
  double K = 0.0, U = 0.0, S = 0.0;
  for (int r = 0; r < g->n; r++) {  // fixed rank order: deterministic
    K += k[r]; U += u[r]; S += s[r];
  }
  if (kinetic) *kinetic = K;
  if (potential) *potential = U;
  if (sum_mm_over_r) *sum_mm_over_r = S;
  return 0;
}

int al26_group_last_device_ms(al26_group *g, double *ms, int64_t *kernel_launches) {
  if (!g) return AL26_EINVAL;
  double mx = 0.0;
  int64_t nl = 0;
  for (int r = 0; r < g->n; r++) {
    mx = std::max(mx, g->ctx[r]->last_ms);
    nl += g->ctx[r]->last_launches;
  }
  if (ms) *ms = mx;
  if (kernel_launches) *kernel_launches = nl;
  return 0;
}

int al26_group_enrich_commit(al26_group *g, int64_t n, const double *r_disk_km, const double *tau_disk_myr,
                             const uint8_t *disk_alive, const uint8_t *kicked, const double *wr26, const double *wr60,
                             const double *sn26_kg, const double *sn60_kg) {
  if (!g) return AL26_EINVAL;
  return group_run(g, [=](int r) {
    return al26_enrich_commit(g->ctx[r], n, r_disk_km, tau_disk_myr, disk_alive, kicked, wr26, wr60, sn26_kg, sn60_kg);
  });
}

int al26_group_enrich_set_units(al26_group *g, double km_per_length, double kms_per_speed) {
  if (!g) return AL26_EINVAL;
  return group_run(g, [=](int r) { return al26_enrich_set_units(g->ctx[r], km_per_length, kms_per_speed); });
}

int al26_group_enrich_set_mode(al26_group *g, int mode) {
  if (!g) return AL26_EINVAL;
  return group_run(g, [=](int r) { return al26_enrich_set_mode(g->ctx[r], mode); });
}

int al26_group_enrich_set_inventories(al26_group *g, int64_t n, const double *inv, const double *fin) {
  if (!g) return AL26_EINVAL;
  return group_run(g, [=](int r) { return al26_enrich_set_inventories(g->ctx[r], n, inv, fin); });
}

int al26_group_enrich_step(al26_group *g, int64_t n, const double *mass_msun, const double *mdot_kg_s, const double *pos_vel,
                           double dt_s, double t_new_myr, double r_bub_local_km, double r_bub_global_km, double decay26,
                           double decay60, int with_agb, int32_t *sn_events, int64_t sn_cap, int64_t *n_sn_events) {
  if (!g) return AL26_EINVAL;
  // every rank detects the same events; rank 0 reports them, the others use scratch of their own
  std::vector<std::vector<int32_t>> scratch(g->n);
  std::vector<int64_t> ne(g->n, 0);
  const int rc = group_run(g, [&](int r) {
    int32_t *ev = sn_events;
    int64_t cap = sn_cap;
    if (r != 0) {
      scratch[r].resize(ENR_MAX_SOURCES);
      ev = scratch[r].data();
      cap = ENR_MAX_SOURCES;
    }
    return al26_enrich_step(g->ctx[r], n, mass_msun, mdot_kg_s, pos_vel, dt_s, t_new_myr, r_bub_local_km, r_bub_global_km,
                            decay26, decay60, with_agb, ev, cap, &ne[r]);
  });
  if (rc) return rc;
  if (n_sn_events) *n_sn_events = ne[0];
  return 0;
}

int al26_group_enrich_get(al26_group *g, int64_t n, double *inv, double *fin, uint8_t *disk_alive, uint8_t *kicked) {
  if (!g) return AL26_EINVAL;
  // every rank writes its own disc slice of the caller's arrays; `kicked` is replicated: rank 0 writes it
  return group_run(g, [=](int r) { return al26_enrich_get(g->ctx[r], n, inv, fin, disk_alive, r == 0 ? kicked : nullptr); });
}

}  // extern "C"
