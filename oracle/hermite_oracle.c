/* oracle/hermite_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (fp64, OpenMP) of the gravity hot path behind
 * `gravity.evolve_model(t_new)` (reference call site al26_nbody.py:833; ctor
 * :1709-1722; add_particles :1728; mass channel :871-874; getters :886-891;
 * model_time :763,:1736; energies plotting/al26_plot.py:288-289; virial radius
 * al26_nbody.py:770).
 *
 * PARITY UNPINNED: the arithmetic the reference runs at that call site lives in
 * the third-party AMUSE community code `amuse_ph4==2023.5.0` (requirements.txt:23),
 * which is neither vendored under /root/reference nor installable here (no MPI,
 * no network) and the reference holds no tests / golden vectors for it.  This file
 * therefore restates the *published* algorithm ph4 implements -- Makino & Aarseth
 * (1992) 4th-order Hermite predictor-corrector with individual power-of-two block
 * timesteps and the Aarseth step criterion with ph4's eta convention -- and the
 * explicit choices in SURVEY.md section 8(c).  It is anchored by analytic
 * known-answer tests (Kepler, Plummer virial identities, momentum conservation,
 * 4th-order convergence) and by a long-double direct sum for acc/jerk/pot.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (26al-nbody_b200/) never does.
 *
 * THE SCHEME (shared spec with the CUDA path; all in N-body units, G = 1):
 *   state per particle: m, x, v, a, jerk, pot, t (relative to the start of the
 *   current evolve call, "tau"), dt (power of two).
 *   init (first evolve after commit / set_params):
 *       a, jerk, pot on all particles from current x, v;
 *       dt0 = eta * 0.0625 * |a|/|jerk|  (dt_max if |a| or |jerk| == 0),
 *       dt  = max(dt_min, pow2floor(min(dt0, 2^-5, dt_max)))
 *   mass-only update (set_mass between two evolve calls; the script does this every outer step,
 *       al26_nbody.py:874): a, jerk, pot recomputed from the synchronised x, v with the new masses;
 *       every dt_i keeps the value the synchronisation step gave it (reinit_policy 0, the default --
 *       ph4's recommit_particles "recompute forces ... we don't recompute the time steps"
 *       [upstream, from memory]); reinit_policy 1 re-derives the initial timesteps as above.
 *   begin(t_end): span = t_end - t_model; D = min(dt_max, pow2floor(span));
 *       dt_i = min(dt_i, D); tau_i = 0
 *   block step: tau_next = min_i(tau_i + dt_i); stop if tau_next > span;
 *       active = { i : tau_i + dt_i == tau_next }   (exact fp compare; all dyadic)
 *       predict every j to tau_next:
 *           xp = x + v s + a s^2/2 + jerk s^3/6,  vp = v + a s + jerk s^2/2
 *       a1, j1, pot on active i from predicted j (pairs with r^2 + eps2 == 0 skipped)
 *       corrector (s = dt_i): alpha = -3(a0-a1) - s(2 j0 + j1),
 *           beta = 2(a0-a1) + s(j0 + j1);
 *           x1 = xp + s^2 (alpha/12 + beta/20);  v1 = vp + s (alpha/3 + beta/4)
 *           a2 = (2 alpha + 6 beta)/s^2 (at end of step), a3 = 6 beta / s^3
 *           dt_A = eta sqrt((|a1||a2| + |j1|^2)/(|j1||a3| + |a2|^2))
 *       ladder: dt_A < dt -> dt/2 (not below dt_min);
 *               dt_A >= 2dt and tau_next mod 2dt == 0 and 2dt <= D -> 2dt; else dt.
 *   finish: sync step to tau = span for every i with tau_i < span (s = span - tau_i,
 *       not dyadic); afterwards dt_i = max(dt_min, pow2floor(min(dt_A, dt_max)));
 *       t_model = t_end.
 *
 * Summation order of the oracle force: j ascending in double (default), or long
 * double (check path).  The CUDA kernel sums in tiles, so acc/jerk agree to
 * rounding (<= 1e-12 relative), not bitwise.  The corrector/ladder/scheduler code
 * below is compiled without FMA contraction so that, given identical inputs, the
 * integer decisions are bit-exact against the CUDA path.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NOFMA __attribute__((optimize("fp-contract=off")))

typedef struct {
  int64_t n;
  double eps2, eta, dt_max, dt_min;
  double t_model;
  int dirty;      /* 0 = forces and timesteps valid; 1 = masses changed since the last synchronisation; 2 = nothing valid */
  int reinit_policy; /* mass-only update: 0 = recompute forces, keep timesteps; 1 = forces + initial timesteps */
  int in_evolve;  /* between begin and finish */
  double span, D;
  int use_long_double;
  double *m, *x, *y, *z, *vx, *vy, *vz;
  double *ax, *ay, *az, *jx, *jy, *jz, *pot;
  double *t, *dt;
  double *px, *py, *pz, *pvx, *pvy, *pvz;
  double *nax, *nay, *naz, *njx, *njy, *njz, *npot;
  int32_t *active;
  int64_t n_active;
  int64_t n_block_steps, n_pairs;
} orc_t;

static double *dalloc(int64_t n) { return (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }

orc_t *orc_create(int64_t n) {
  orc_t *o = (orc_t *)calloc(1, sizeof(orc_t));
  o->n = n;
  o->eps2 = 0.0;
  o->eta = 0.14;
  o->dt_max = 0.125;
  o->dt_min = ldexp(1.0, -40);
  o->dirty = 2;
  double **arrs[] = {&o->m, &o->x, &o->y, &o->z, &o->vx, &o->vy, &o->vz, &o->ax, &o->ay, &o->az,
                     &o->jx, &o->jy, &o->jz, &o->pot, &o->t, &o->dt, &o->px, &o->py, &o->pz,
                     &o->pvx, &o->pvy, &o->pvz, &o->nax, &o->nay, &o->naz, &o->njx, &o->njy,
                     &o->njz, &o->npot};
  for (size_t k = 0; k < sizeof(arrs) / sizeof(arrs[0]); k++) *arrs[k] = dalloc(n);
  o->active = (int32_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
  return o;
}

void orc_destroy(orc_t *o) {
  if (!o) return;
  double *arrs[] = {o->m, o->x, o->y, o->z, o->vx, o->vy, o->vz, o->ax, o->ay, o->az,
                    o->jx, o->jy, o->jz, o->pot, o->t, o->dt, o->px, o->py, o->pz,
                    o->pvx, o->pvy, o->pvz, o->nax, o->nay, o->naz, o->njx, o->njy,
                    o->njz, o->npot};
  for (size_t k = 0; k < sizeof(arrs) / sizeof(arrs[0]); k++) free(arrs[k]);
  free(o->active);
  free(o);
}

int orc_set_params(orc_t *o, double eps2, double eta, double dt_max, double dt_min) {
  if (eps2 < 0 || eta <= 0 || dt_max <= 0 || dt_min <= 0 || dt_min > dt_max) return -1;
  o->eps2 = eps2;
  o->eta = eta;
  /* the ladder is dyadic: round the limits down / up to powers of two */
  int e;
  frexp(dt_max, &e);
  o->dt_max = ldexp(1.0, e - 1);
  frexp(dt_min, &e);
  o->dt_min = ldexp(1.0, e - 1);
  o->dirty = 2;
  return 0;
}
int orc_set_reinit_policy(orc_t *o, int policy) {
  if (policy != 0 && policy != 1) return -1;
  o->reinit_policy = policy;
  return 0;
}
void orc_set_long_double(orc_t *o, int on) { o->use_long_double = on; }

int orc_commit(orc_t *o, int64_t n, const double *m, const double *x, const double *y, const double *z,
               const double *vx, const double *vy, const double *vz) {
  if (n != o->n) return -1;
  size_t b = (size_t)n * sizeof(double);
  memcpy(o->m, m, b); memcpy(o->x, x, b); memcpy(o->y, y, b); memcpy(o->z, z, b);
  memcpy(o->vx, vx, b); memcpy(o->vy, vy, b); memcpy(o->vz, vz, b);
  o->dirty = 2;
  return 0;
}
int orc_set_mass(orc_t *o, int64_t n, const double *m) {
  if (n != o->n) return -1;
  memcpy(o->m, m, (size_t)n * sizeof(double));
  if (o->dirty < 1) o->dirty = 1;
  return 0;
}
int orc_set_time(orc_t *o, double t) { o->t_model = t; return 0; }
double orc_get_time(orc_t *o) { return o->t_model; }

/* ---- force: acc, jerk, pot on the listed i from predicted j (this is G3 of SURVEY section 8a) ---- */
static void force_double(const orc_t *o, int64_t n_act, const int32_t *idx) {
  const int64_t n = o->n;
  const double eps2 = o->eps2;
  const double *restrict px = o->px, *restrict py = o->py, *restrict pz = o->pz;
  const double *restrict pvx = o->pvx, *restrict pvy = o->pvy, *restrict pvz = o->pvz;
  const double *restrict m = o->m;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t k = 0; k < n_act; k++) {
    const int64_t i = idx[k];
    const double xi = px[i], yi = py[i], zi = pz[i];
    const double vxi = pvx[i], vyi = pvy[i], vzi = pvz[i];
    double ax = 0, ay = 0, az = 0, jx = 0, jy = 0, jz = 0, pot = 0;
#pragma omp simd reduction(+ : ax, ay, az, jx, jy, jz, pot)
    for (int64_t j = 0; j < n; j++) {
      const double dx = px[j] - xi, dy = py[j] - yi, dz = pz[j] - zi;
      const double dvx = pvx[j] - vxi, dvy = pvy[j] - vyi, dvz = pvz[j] - vzi;
      const double r2 = dx * dx + dy * dy + dz * dz + eps2;
      const double rv = dx * dvx + dy * dvy + dz * dvz;
      const double rinv = (r2 > 0.0 && j != i) ? 1.0 / sqrt(r2) : 0.0; /* self pair: out, also when softened */
      const double rinv2 = rinv * rinv;
      const double mrinv = m[j] * rinv;
      const double mrinv3 = mrinv * rinv2;
      const double al = -3.0 * rv * rinv2;
      pot -= mrinv;
      ax += mrinv3 * dx; ay += mrinv3 * dy; az += mrinv3 * dz;
      jx += mrinv3 * (dvx + al * dx);
      jy += mrinv3 * (dvy + al * dy);
      jz += mrinv3 * (dvz + al * dz);
    }
    o->nax[k] = ax; o->nay[k] = ay; o->naz[k] = az;
    o->njx[k] = jx; o->njy[k] = jy; o->njz[k] = jz;
    o->npot[k] = pot;
  }
}

static void force_long_double(const orc_t *o, int64_t n_act, const int32_t *idx) {
  const int64_t n = o->n;
  const long double eps2 = o->eps2;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t k = 0; k < n_act; k++) {
    const int64_t i = idx[k];
    const long double xi = o->px[i], yi = o->py[i], zi = o->pz[i];
    const long double vxi = o->pvx[i], vyi = o->pvy[i], vzi = o->pvz[i];
    long double ax = 0, ay = 0, az = 0, jx = 0, jy = 0, jz = 0, pot = 0;
    for (int64_t j = 0; j < n; j++) {
      const long double dx = o->px[j] - xi, dy = o->py[j] - yi, dz = o->pz[j] - zi;
      const long double dvx = o->pvx[j] - vxi, dvy = o->pvy[j] - vyi, dvz = o->pvz[j] - vzi;
      const long double r2 = dx * dx + dy * dy + dz * dz + eps2;
      if (!(r2 > 0.0L) || j == i) continue;
      const long double rv = dx * dvx + dy * dvy + dz * dvz;
      const long double rinv = 1.0L / sqrtl(r2);
      const long double rinv2 = rinv * rinv;
      const long double mrinv = (long double)o->m[j] * rinv;
      const long double mrinv3 = mrinv * rinv2;
      const long double al = -3.0L * rv * rinv2;
      pot -= mrinv;
      ax += mrinv3 * dx; ay += mrinv3 * dy; az += mrinv3 * dz;
      jx += mrinv3 * (dvx + al * dx);
      jy += mrinv3 * (dvy + al * dy);
      jz += mrinv3 * (dvz + al * dz);
    }
    o->nax[k] = (double)ax; o->nay[k] = (double)ay; o->naz[k] = (double)az;
    o->njx[k] = (double)jx; o->njy[k] = (double)jy; o->njz[k] = (double)jz;
    o->npot[k] = (double)pot;
  }
}

static void force(orc_t *o, int64_t n_act, const int32_t *idx) {
  if (o->use_long_double) force_long_double(o, n_act, idx);
  else force_double(o, n_act, idx);
  o->n_pairs += n_act * o->n;
}

/* stand-alone force evaluation on caller-supplied arrays: the parity hook for K1 */
int orc_force(int64_t n, double eps2, const double *m, const double *x, const double *y, const double *z,
              const double *vx, const double *vy, const double *vz, int64_t n_act, const int32_t *idx,
              int use_long_double, double *ax, double *ay, double *az, double *jx, double *jy, double *jz,
              double *pot) {
  orc_t o;
  memset(&o, 0, sizeof(o));
  o.n = n; o.eps2 = eps2; o.use_long_double = use_long_double;
  o.m = (double *)m; o.px = (double *)x; o.py = (double *)y; o.pz = (double *)z;
  o.pvx = (double *)vx; o.pvy = (double *)vy; o.pvz = (double *)vz;
  o.nax = ax; o.nay = ay; o.naz = az; o.njx = jx; o.njy = jy; o.njz = jz; o.npot = pot;
  force(&o, n_act, idx);
  return 0;
}

/* ---- everything below: no FMA contraction (bit-exact decisions vs the CUDA path) ---- */

NOFMA static double pow2floor(double x) {
  int e;
  frexp(x, &e);
  return ldexp(1.0, e - 1);
}

NOFMA static void predict_all(orc_t *o, double tau_next) {
  const int64_t n = o->n;
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; j++) {
    const double s = tau_next - o->t[j];
    const double s2 = s * s * 0.5, s3 = s * s * s * (1.0 / 6.0);
    o->px[j] = o->x[j] + o->vx[j] * s + o->ax[j] * s2 + o->jx[j] * s3;
    o->py[j] = o->y[j] + o->vy[j] * s + o->ay[j] * s2 + o->jy[j] * s3;
    o->pz[j] = o->z[j] + o->vz[j] * s + o->az[j] * s2 + o->jz[j] * s3;
    o->pvx[j] = o->vx[j] + o->ax[j] * s + o->jx[j] * s2;
    o->pvy[j] = o->vy[j] + o->ay[j] * s + o->jy[j] * s2;
    o->pvz[j] = o->vz[j] + o->az[j] * s + o->jz[j] * s2;
  }
}

/* Aarseth estimate from (a1, j1, a2, a3); returns dt_A (huge if the denominator vanishes) */
NOFMA static double aarseth(double eta, const double a1[3], const double j1[3], const double a2[3],
                            const double a3[3]) {
  const double sa = a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2];
  const double sj = j1[0] * j1[0] + j1[1] * j1[1] + j1[2] * j1[2];
  const double s2 = a2[0] * a2[0] + a2[1] * a2[1] + a2[2] * a2[2];
  const double s3 = a3[0] * a3[0] + a3[1] * a3[1] + a3[2] * a3[2];
  const double num = sqrt(sa * s2) + sj;
  const double den = sqrt(sj * s3) + s2;
  if (!(den > 0.0) || !(num > 0.0)) return 1.0e300;
  return eta * sqrt(num / den);
}

/* corrector for slot k (particle i) with step s; returns dt_A */
NOFMA static double correct_one(orc_t *o, int64_t k, int64_t i, double s) {
  const double a0[3] = {o->ax[i], o->ay[i], o->az[i]};
  const double j0[3] = {o->jx[i], o->jy[i], o->jz[i]};
  const double a1[3] = {o->nax[k], o->nay[k], o->naz[k]};
  const double j1[3] = {o->njx[k], o->njy[k], o->njz[k]};
  const double xp[3] = {o->px[i], o->py[i], o->pz[i]};
  const double vp[3] = {o->pvx[i], o->pvy[i], o->pvz[i]};
  double x1[3], v1[3], a2[3], a3[3];
  const double s2 = s * s;
  const double is2 = 1.0 / s2, is3 = 1.0 / (s2 * s);
  for (int c = 0; c < 3; c++) {
    const double da = a0[c] - a1[c];
    const double alpha = -3.0 * da - s * (2.0 * j0[c] + j1[c]);
    const double beta = 2.0 * da + s * (j0[c] + j1[c]);
    x1[c] = xp[c] + s2 * (alpha * (1.0 / 12.0) + beta * (1.0 / 20.0));
    v1[c] = vp[c] + s * (alpha * (1.0 / 3.0) + beta * 0.25);
    a2[c] = (2.0 * alpha + 6.0 * beta) * is2;
    a3[c] = (6.0 * beta) * is3;
  }
  o->x[i] = x1[0]; o->y[i] = x1[1]; o->z[i] = x1[2];
  o->vx[i] = v1[0]; o->vy[i] = v1[1]; o->vz[i] = v1[2];
  o->ax[i] = a1[0]; o->ay[i] = a1[1]; o->az[i] = a1[2];
  o->jx[i] = j1[0]; o->jy[i] = j1[1]; o->jz[i] = j1[2];
  o->pot[i] = o->npot[k];
  return aarseth(o->eta, a1, j1, a2, a3);
}

NOFMA static void initialise(orc_t *o) {
  const int64_t n = o->n;
  for (int64_t i = 0; i < n; i++) {
    o->t[i] = 0.0;
    o->active[i] = (int32_t)i;
  }
  predict_all(o, 0.0); /* s = 0: predicted == current */
  force(o, n, o->active);
  const double lim = ldexp(1.0, -5);
  const int keep_dt = (o->dirty == 1 && o->reinit_policy == 0);
  for (int64_t i = 0; i < n; i++) {
    o->ax[i] = o->nax[i]; o->ay[i] = o->nay[i]; o->az[i] = o->naz[i];
    o->jx[i] = o->njx[i]; o->jy[i] = o->njy[i]; o->jz[i] = o->njz[i];
    o->pot[i] = o->npot[i];
    if (keep_dt) continue; /* mass-only update: the timestep of the last synchronisation step stands */
    const double sa = o->ax[i] * o->ax[i] + o->ay[i] * o->ay[i] + o->az[i] * o->az[i];
    const double sj = o->jx[i] * o->jx[i] + o->jy[i] * o->jy[i] + o->jz[i] * o->jz[i];
    double dt0 = o->dt_max;
    if (sa > 0.0 && sj > 0.0) dt0 = o->eta * 0.0625 * sqrt(sa / sj);
    if (dt0 > lim) dt0 = lim;
    if (dt0 > o->dt_max) dt0 = o->dt_max;
    double d = pow2floor(dt0);
    if (d < o->dt_min) d = o->dt_min;
    o->dt[i] = d;
  }
  o->dirty = 0;
}

/* make forces and timesteps valid without advancing (parity hook) */
int orc_initialize(orc_t *o) {
  if (o->in_evolve) return -2;
  if (o->dirty) initialise(o);
  return 0;
}

NOFMA int orc_begin(orc_t *o, double t_end) {
  if (o->in_evolve) return -2;
  const double span = t_end - o->t_model;
  if (!(span > 0.0)) return -3;
  if (o->dirty) initialise(o);
  double D = pow2floor(span);
  if (D > o->dt_max) D = o->dt_max;
  o->span = span;
  o->D = D;
  for (int64_t i = 0; i < o->n; i++) {
    o->t[i] = 0.0;
    if (o->dt[i] > D) o->dt[i] = D;
  }
  o->in_evolve = 1;
  return 0;
}

/* scheduler (G5): tau_next and the active list, ascending index order */
NOFMA static double schedule(orc_t *o) {
  double tn = 1.0e300;
  for (int64_t i = 0; i < o->n; i++) {
    const double c = o->t[i] + o->dt[i];
    if (c < tn) tn = c;
  }
  int64_t k = 0;
  for (int64_t i = 0; i < o->n; i++)
    if (o->t[i] + o->dt[i] == tn) o->active[k++] = (int32_t)i;
  o->n_active = k;
  return tn;
}

/* at most max_steps block steps; *finished = 1 when the next block time exceeds span */
NOFMA int orc_advance(orc_t *o, int64_t max_steps, int64_t *n_done, int *finished) {
  if (!o->in_evolve) return -2;
  int64_t done = 0;
  *finished = 0;
  while (max_steps < 0 || done < max_steps) {
    const double tn = schedule(o);
    if (tn > o->span) { *finished = 1; break; }
    predict_all(o, tn);
    force(o, o->n_active, o->active);
    for (int64_t k = 0; k < o->n_active; k++) {
      const int64_t i = o->active[k];
      const double dt = o->dt[i];
      const double dtA = correct_one(o, k, i, dt);
      double nd = dt;
      if (dtA < dt) {
        if (0.5 * dt >= o->dt_min) nd = 0.5 * dt;
      } else if (dtA >= 2.0 * dt && 2.0 * dt <= o->D) {
        const double q = tn / (2.0 * dt);
        if (q == floor(q)) nd = 2.0 * dt;
      }
      o->t[i] = tn;
      o->dt[i] = nd;
    }
    o->n_block_steps++;
    done++;
  }
  if (n_done) *n_done = done;
  return 0;
}

NOFMA int orc_finish(orc_t *o) {
  if (!o->in_evolve) return -2;
  const double span = o->span;
  int64_t k = 0;
  for (int64_t i = 0; i < o->n; i++)
    if (o->t[i] < span) o->active[k++] = (int32_t)i;
  o->n_active = k;
  if (k > 0) {
    predict_all(o, span);
    force(o, k, o->active);
    for (int64_t q = 0; q < k; q++) {
      const int64_t i = o->active[q];
      const double s = span - o->t[i];
      double dtA = correct_one(o, q, i, s);
      if (dtA > o->dt_max) dtA = o->dt_max;
      double d = pow2floor(dtA);
      if (d < o->dt_min) d = o->dt_min;
      o->dt[i] = d;
    }
    o->n_block_steps++;
  }
  for (int64_t i = 0; i < o->n; i++) o->t[i] = 0.0;
  o->t_model += span;
  o->in_evolve = 0;
  return 0;
}

int orc_evolve(orc_t *o, double t_end, int64_t *n_block_steps, int64_t *n_pairs) {
  const int64_t s0 = o->n_block_steps, p0 = o->n_pairs;
  if (t_end == o->t_model) {
    if (n_block_steps) *n_block_steps = 0;
    if (n_pairs) *n_pairs = 0;
    return 0;
  }
  int rc = orc_begin(o, t_end);
  if (rc) return rc;
  int fin = 0;
  rc = orc_advance(o, -1, NULL, &fin);
  if (rc) return rc;
  rc = orc_finish(o);
  o->t_model = t_end; /* exact, not t_model + span */
  if (n_block_steps) *n_block_steps = o->n_block_steps - s0;
  if (n_pairs) *n_pairs = o->n_pairs - p0;
  return rc;
}

int orc_get_state(orc_t *o, int64_t n, double *m, double *x, double *y, double *z, double *vx, double *vy,
                  double *vz) {
  if (n != o->n) return -1;
  size_t b = (size_t)n * sizeof(double);
  memcpy(m, o->m, b); memcpy(x, o->x, b); memcpy(y, o->y, b); memcpy(z, o->z, b);
  memcpy(vx, o->vx, b); memcpy(vy, o->vy, b); memcpy(vz, o->vz, b);
  return 0;
}
int orc_get_acc_jerk(orc_t *o, int64_t n, double *ax, double *ay, double *az, double *jx, double *jy,
                     double *jz, double *pot) {
  if (n != o->n) return -1;
  size_t b = (size_t)n * sizeof(double);
  memcpy(ax, o->ax, b); memcpy(ay, o->ay, b); memcpy(az, o->az, b);
  memcpy(jx, o->jx, b); memcpy(jy, o->jy, b); memcpy(jz, o->jz, b);
  memcpy(pot, o->pot, b);
  return 0;
}
int orc_get_timesteps(orc_t *o, int64_t n, double *t, double *dt) {
  if (n != o->n) return -1;
  memcpy(t, o->t, (size_t)n * sizeof(double));
  memcpy(dt, o->dt, (size_t)n * sizeof(double));
  return 0;
}
int orc_set_timesteps(orc_t *o, int64_t n, const double *t, const double *dt) {
  if (n != o->n) return -1;
  memcpy(o->t, t, (size_t)n * sizeof(double));
  memcpy(o->dt, dt, (size_t)n * sizeof(double));
  return 0;
}
/* active set of the next block step from the current (t, dt): ascending indices */
int orc_get_active(orc_t *o, int64_t cap, int32_t *idx, int64_t *n_active, double *tau_next) {
  const double tn = schedule(o);
  if (o->n_active > cap) return -4;
  memcpy(idx, o->active, (size_t)o->n_active * sizeof(int32_t));
  *n_active = o->n_active;
  if (tau_next) *tau_next = tn;
  return 0;
}

/* G9: K = 1/2 sum m v^2; U = -sum_{i<j} m_i m_j / sqrt(r^2 + eps2); S = sum_{i<j} m_i m_j / r
 * (R_vir = M^2 / (2 S), al26_nbody.py:770).  long double accumulation. */
int orc_energies(orc_t *o, double *kinetic, double *potential, double *sum_mm_over_r) {
  const int64_t n = o->n;
  long double K = 0, U = 0, S = 0;
  for (int64_t i = 0; i < n; i++)
    K += 0.5L * o->m[i] * ((long double)o->vx[i] * o->vx[i] + (long double)o->vy[i] * o->vy[i] +
                           (long double)o->vz[i] * o->vz[i]);
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : U, S)
  for (int64_t i = 0; i < n; i++) {
    long double u = 0, s = 0;
    for (int64_t j = i + 1; j < n; j++) {
      const long double dx = o->x[j] - o->x[i], dy = o->y[j] - o->y[i], dz = o->z[j] - o->z[i];
      const long double r2 = dx * dx + dy * dy + dz * dz;
      if (r2 + o->eps2 > 0.0L) u += (long double)o->m[j] / sqrtl(r2 + (long double)o->eps2);
      if (r2 > 0.0L) s += (long double)o->m[j] / sqrtl(r2);
    }
    U -= o->m[i] * u;
    S += o->m[i] * s;
  }
  *kinetic = (double)K;
  *potential = (double)U;
  *sum_mm_over_r = (double)S;
  return 0;
}

/* torchrun exports OMP_NUM_THREADS=1 to its children; the CPU arm of bench.py asks for all host cores */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_counters(orc_t *o, int64_t *steps, int64_t *pairs) { *steps = o->n_block_steps; *pairs = o->n_pairs; }
