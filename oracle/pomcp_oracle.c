/*
 * pomcp_oracle.c — TEST INFRASTRUCTURE ONLY (see pomcp_oracle.h).
 *
 * Sequential CPU restatement of the POMCP.jl hot path.  Every function cites
 * the reference lines it follows (paths relative to the POMCP.jl tree).
 * PARITY: pinned against the golden vectors recoverable from the reference's
 * notebooks (tests/golden/); UNPINNED for the parts that live in un-vendored
 * third-party packages (SIR filter, RolloutSimulator, Tiger/Baby/RockSample).
 *
 * Conventions that are OURS (the reference leaves them to Julia's RNG / Dict):
 *  - randomness is Philox4x32-10, counter = (index, tag|rank<<8, query, epoch),
 *    key = 64-bit seed; uniforms are word * 2^-32;
 *  - actions are enumerated in ascending index order, so the reference's
 *    "last maximal child wins" (>= at src/solver.jl:64,119) means highest index;
 *  - log/exp/cos are the fdlibm algorithms written in plain IEEE ops (compile
 *    with -ffp-contract=off) so that the CUDA kernels can reproduce every bit.
 */
#include "pomcp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011)                                  */
/* ------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { TAG_ROOT = 0, TAG_TREE = 1, TAG_ROLLOUT = 2, TAG_PF = 3, TAG_RESAMPLE = 4, TAG_INIT = 5 };

static inline double u01(uint32_t w) { return (double)w * 2.3283064365386963e-10; /* 2^-32 */ }

/* ------------------------------------------------------------------ */
/* deterministic math: fdlibm algorithms, plain IEEE mul/add/div only  */
/* ------------------------------------------------------------------ */
static inline uint64_t d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

double orc_det_log(double x) {
  static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                      Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                      Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                      Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                      Lg7 = 1.479819860511658591e-01;
  if (x == 0.0) return -INFINITY;
  if (x < 0.0 || x != x) return NAN;
  if (x == INFINITY) return x;
  uint64_t ux = d2u(x);
  int k = 0;
  if ((ux >> 52) == 0) { /* subnormal */
    x *= 18014398509481984.0; /* 2^54 */
    ux = d2u(x);
    k -= 54;
  }
  uint32_t hx = (uint32_t)(ux >> 32);
  k += (int)(hx >> 20) - 1023;
  hx &= 0x000fffffu;
  uint32_t i = (hx + 0x95f64u) & 0x100000u;
  ux = ((uint64_t)(hx | (i ^ 0x3ff00000u)) << 32) | (ux & 0xffffffffu);
  k += (int)(i >> 20);
  double m = u2d(ux);
  double f = m - 1.0;
  double s = f / (2.0 + f);
  double dk = (double)k;
  double z = s * s;
  double w = z * z;
  double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  double R = t2 + t1;
  double hfsq = 0.5 * f * f;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

double orc_det_exp(double x) {
  static const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
                      invln2 = 1.44269504088896338700e+00, P1 = 1.66666666666666019037e-01,
                      P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                      P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  if (x != x) return x;
  if (x > 709.0) return INFINITY;
  if (x < -745.0) return 0.0;
  int k = (int)(invln2 * x + (x < 0.0 ? -0.5 : 0.5));
  double t = (double)k;
  double hi = x - t * ln2HI;
  double lo = t * ln2LO;
  double r = hi - lo;
  double tt = r * r;
  double c = r - tt * (P1 + tt * (P2 + tt * (P3 + tt * (P4 + tt * P5))));
  double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
  /* scale by 2^k in two exact steps so that subnormal results round once */
  if (k >= -1000) return y * u2d((uint64_t)(k + 1023) << 52);
  return (y * u2d((uint64_t)(k + 1000 + 1023) << 52)) * u2d((uint64_t)(-1000 + 1023) << 52);
}

/* sin / cos on [0, pi/4] (fdlibm k_sin.c / k_cos.c polynomials) */
static double ksin(double y) {
  static const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                      S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                      S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = y * y;
  double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return y + y * (z * (S1 + z * r));
}
static double kcos(double y) {
  static const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                      C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                      C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = y * y;
  double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  return (1.0 - 0.5 * z) + z * r;
}
/* cos(2*pi*u), u in [0,1): quadrant split, fold to [0, pi/4] */
double orc_det_cos2pi(double u) {
  static const double HALF_PI = 1.57079632679489661923;
  double t = u * 4.0;
  int q = (int)t;
  double f = t - (double)q; /* [0,1): angle in quadrant = f*pi/2 */
  double sn, cs;
  if (f <= 0.5) {
    double y = f * HALF_PI;
    sn = ksin(y); cs = kcos(y);
  } else {
    double y = (1.0 - f) * HALF_PI;
    sn = kcos(y); cs = ksin(y);
  }
  switch (q & 3) {
    case 0: return cs;
    case 1: return -sn;
    case 2: return -cs;
    default: return sn;
  }
}
/* Box-Muller: randn(rng) stand-in (notebooks/Minimal_Example.ipynb:91,142) */
double orc_normal_from_uniforms(double u0, double u1) {
  return sqrt(-2.0 * orc_det_log(1.0 - u0)) * orc_det_cos2pi(u1);
}

/* ------------------------------------------------------------------ */
/* models                                                              */
/* ------------------------------------------------------------------ */
typedef struct ostate { int64_t i; double y; } ostate; /* discrete: i = packed u32; LightDark: (status, y) */

typedef struct model {
  pomcp_problem p;
  int nA, nO;      /* nO == 0: continuous observation */
  int state_bytes; /* 4 or 16 */
  int draws_per_rollout_step; /* 1 or 2 Philox words */
  int8_t rock_at[256];                 /* cell (y*16+x) -> rock index or -1 */
  double eta[256 * POMCP_MAX_ROCKS];   /* P(check_i correct) at cell */
} model;

static int model_init(model* m, const pomcp_problem* p) {
  memset(m, 0, sizeof(*m));
  m->p = *p;
  switch (p->problem_id) {
    case POMCP_TIGER: m->nA = 3; m->nO = 2; m->state_bytes = 4; m->draws_per_rollout_step = 2; break;
    case POMCP_BABY: m->nA = 2; m->nO = 2; m->state_bytes = 4; m->draws_per_rollout_step = 2; break;
    case POMCP_ROCKSAMPLE: {
      if (p->rs_n < 1 || p->rs_n > 15 || p->rs_k < 0 || p->rs_k > POMCP_MAX_ROCKS) return POMCP_ERR_INVALID_ARG;
      m->nA = 5 + p->rs_k; m->nO = 3; m->state_bytes = 4; m->draws_per_rollout_step = 1;
      memset(m->rock_at, -1, sizeof(m->rock_at));
      for (int i = 0; i < p->rs_k; ++i) {
        if (p->rs_rock_x[i] < 0 || p->rs_rock_x[i] >= p->rs_n || p->rs_rock_y[i] < 0 || p->rs_rock_y[i] >= p->rs_n)
          return POMCP_ERR_INVALID_ARG;
        m->rock_at[p->rs_rock_y[i] * 16 + p->rs_rock_x[i]] = (int8_t)i;
      }
      for (int y = 0; y < p->rs_n; ++y)
        for (int x = 0; x < p->rs_n; ++x)
          for (int i = 0; i < p->rs_k; ++i) {
            double dx = (double)(x - p->rs_rock_x[i]), dy = (double)(y - p->rs_rock_y[i]);
            double d = sqrt(dx * dx + dy * dy);
            /* Smith & Simmons sensor: eta = (1 + 2^(-d/d0)) / 2 */
            m->eta[(y * 16 + x) * POMCP_MAX_ROCKS + i] = 0.5 * (1.0 + exp2(-d / p->rs_half_eff_dist));
          }
      break;
    }
    case POMCP_LIGHTDARK1D: m->nA = 3; m->nO = 0; m->state_bytes = 16; m->draws_per_rollout_step = 1; break;
    default: return POMCP_ERR_INVALID_ARG;
  }
  if (!(p->discount > 0.0) || !(p->discount < 1.0)) return POMCP_ERR_INVALID_ARG; /* gamma==1: src/solver.jl:97-102 bug */
  return POMCP_OK;
}

static inline int is_terminal(const model* m, ostate s) {
  switch (m->p.problem_id) {
    case POMCP_ROCKSAMPLE: return ((uint32_t)s.i & POMCP_RS_TERMINAL) != 0;
    case POMCP_LIGHTDARK1D: return s.i < 0; /* notebooks/Minimal_Example.ipynb:77 */
    default: return 0;
  }
}

static inline double ld_sigma(double x) { /* notebooks/Minimal_Example.ipynb:89 */
  return fabs(x - 5.0) / 1.4142135623730951 + 1e-2;
}

/* generate_sor(problem, s, a, rng) (src/solver.jl:126): transition draws first,
 * observation draws second (POMDPs.jl default). */
static inline void model_step(const model* m, ostate s, int a, double ut0, double uo0, double uo1, int need_obs,
                              ostate* sp_out, uint64_t* obs_out, double* r_out) {
  const pomcp_problem* p = &m->p;
  ostate sp = s;
  uint64_t o = 0;
  double r = 0.0;
  switch (p->problem_id) {
    case POMCP_TIGER: {
      int left = (int)s.i; /* 1: tiger behind left door */
      if (a == 0) {
        r = p->tiger_r_listen;
        if (need_obs) o = (uo0 < p->tiger_p_listen_correctly) ? (uint64_t)left : (uint64_t)(1 - left);
      } else {
        int opened_tiger = (a == 1) ? left : (1 - left);
        r = opened_tiger ? p->tiger_r_findtiger : p->tiger_r_escapetiger;
        sp.i = (ut0 < 0.5) ? 1 : 0;
        if (need_obs) o = (uo0 < 0.5) ? 1u : 0u;
      }
      break;
    }
    case POMCP_BABY: {
      int hungry = (int)s.i;
      r = (hungry ? p->baby_r_hungry : 0.0) + (a ? p->baby_r_feed : 0.0);
      if (a) sp.i = 0;
      else if (hungry) sp.i = 1;
      else sp.i = (ut0 < p->baby_p_become_hungry) ? 1 : 0;
      if (need_obs) {
        double pc = sp.i ? p->baby_p_cry_when_hungry : p->baby_p_cry_when_not_hungry;
        o = (uo0 < pc) ? 1u : 0u;
      }
      break;
    }
    case POMCP_ROCKSAMPLE: {
      uint32_t st = (uint32_t)s.i;
      int x = (int)(st & 15u), y = (int)((st >> 4) & 15u);
      o = 2; /* none */
      if (a == 0) { if (y < p->rs_n - 1) ++y; }
      else if (a == 2) { if (y > 0) --y; }
      else if (a == 3) { if (x > 0) --x; }
      else if (a == 1) {
        if (x == p->rs_n - 1) { st |= POMCP_RS_TERMINAL; r = p->rs_r_exit; }
        else ++x;
      } else if (a == 4) {
        int ri = m->rock_at[y * 16 + x];
        if (ri >= 0) {
          if ((st >> (8 + ri)) & 1u) { r = p->rs_r_good; st &= ~(1u << (8 + ri)); }
          else r = p->rs_r_bad;
        }
      } else {
        int ri = a - 5;
        if (need_obs) {
          int good = (int)((st >> (8 + ri)) & 1u);
          double e = m->eta[(y * 16 + x) * POMCP_MAX_ROCKS + ri];
          int correct = uo0 < e;
          o = (uint64_t)((good == correct) ? 0 : 1); /* 0 good, 1 bad */
        } else o = 0;
      }
      st = (st & ~0xffu) | (uint32_t)x | ((uint32_t)y << 4);
      sp.i = (int64_t)st;
      break;
    }
    case POMCP_LIGHTDARK1D: { /* notebooks/Minimal_Example.ipynb:94-120 */
      if (s.i < 0) { r = 0.0; }
      else if (a == 1) {
        r = (fabs(s.y) < 1.0) ? p->ld_correct_r : p->ld_incorrect_r;
        sp.i = -1;
      } else {
        sp.y = s.y + (double)(a - 1);
      }
      if (need_obs) {
        double z = orc_normal_from_uniforms(uo0, uo1);
        double ob = sp.y + z * ld_sigma(sp.y);
        o = d2u(ob);
      }
      break;
    }
  }
  (void)ut0;
  *sp_out = sp; *obs_out = o; *r_out = r;
}

/* obs_weight = pdf(observation(model, a, sp), o) (ParticleFilters.jl; LightDark pdf:
 * notebooks/Minimal_Example.ipynb:231-240) */
static double obs_weight(const model* m, int a, ostate sp, uint64_t obs) {
  const pomcp_problem* p = &m->p;
  switch (p->problem_id) {
    case POMCP_TIGER:
      if (a == 0) return ((uint64_t)sp.i == obs) ? p->tiger_p_listen_correctly : 1.0 - p->tiger_p_listen_correctly;
      return 0.5;
    case POMCP_BABY: {
      double pc = sp.i ? p->baby_p_cry_when_hungry : p->baby_p_cry_when_not_hungry;
      return obs ? pc : 1.0 - pc;
    }
    case POMCP_ROCKSAMPLE: {
      if (a < 5) return obs == 2 ? 1.0 : 0.0;
      if (obs > 1) return 0.0;
      uint32_t st = (uint32_t)sp.i;
      int x = (int)(st & 15u), y = (int)((st >> 4) & 15u), ri = a - 5;
      int good = (int)((st >> (8 + ri)) & 1u);
      double e = m->eta[(y * 16 + x) * POMCP_MAX_ROCKS + ri];
      int says_good = obs == 0;
      return (good == says_good) ? e : 1.0 - e;
    }
    case POMCP_LIGHTDARK1D: {
      double x = u2d(obs), mu = sp.y, sd = ld_sigma(sp.y);
      double d = x - mu;
      return orc_det_exp(-(d * d) / (2.0 * (sd * sd))) / (sd * 2.5066282746310002);
    }
  }
  return 0.0;
}

/* Exact Bayes filter of a two-state problem (POMDPModels' BabyBeliefUpdater / Tiger's discrete updater), the
 * node_belief_updater of notebooks/Sanity_Checks.ipynb:47,70 (`updater(problem)`; src/solver.jl:138,
 * src/updater.jl:33): b = P(s = 1), b' ~ O(o | a, s') * sum_s T(s' | s, a) b(s).  Reproduces the notebooks'
 * belief chain (tests/golden D1, D2). */
static double trans_p1(const model* m, int s, int a) {
  if (m->p.problem_id == POMCP_TIGER) return a == 0 ? (s ? 1.0 : 0.0) : 0.5;
  return a ? 0.0 : (s ? 1.0 : m->p.baby_p_become_hungry); /* Baby */
}
static double bayes2(const model* m, double b, int a, uint64_t obs) {
  ostate s1; s1.i = 1; s1.y = 0.0;
  ostate s0; s0.i = 0; s0.y = 0.0;
  const double p1 = b * trans_p1(m, 1, a) + (1.0 - b) * trans_p1(m, 0, a);
  const double n1 = obs_weight(m, a, s1, obs) * p1;
  const double n0 = obs_weight(m, a, s0, obs) * (1.0 - p1);
  const double den = n0 + n1;
  return den > 0.0 ? n1 / den : p1;
}

static ostate load_state(const model* m, const void* arr, int64_t i) {
  ostate s;
  if (m->state_bytes == 4) { s.i = (int64_t)((const uint32_t*)arr)[i]; s.y = 0.0; }
  else { const pomcp_lightdark_state* a = (const pomcp_lightdark_state*)arr; s.i = a[i].status; s.y = a[i].y; }
  return s;
}
static void store_state(const model* m, void* arr, int64_t i, ostate s) {
  if (m->state_bytes == 4) ((uint32_t*)arr)[i] = (uint32_t)s.i;
  else { pomcp_lightdark_state* a = (pomcp_lightdark_state*)arr; a[i].status = s.i; a[i].y = s.y; }
}

int32_t orc_model_step(const pomcp_problem* p, const void* state, int32_t a, const double ut[2], const double uo[2],
                       void* next_state, uint64_t* obs_bits, double* reward) {
  model* m = (model*)malloc(sizeof(model));
  int rc = model_init(m, p);
  if (rc) { free(m); return rc; }
  if (a < 0 || a >= m->nA) { free(m); return POMCP_ERR_INVALID_ARG; }
  ostate s = load_state(m, state, 0), sp;
  model_step(m, s, a, ut[0], uo[0], uo[1], 1, &sp, obs_bits, reward);
  store_state(m, next_state, 0, sp);
  free(m);
  return POMCP_OK;
}
double orc_obs_weight(const pomcp_problem* p, int32_t a, const void* next_state, uint64_t obs_bits) {
  model* m = (model*)malloc(sizeof(model));
  if (model_init(m, p)) { free(m); return NAN; }
  double w = obs_weight(m, a, load_state(m, next_state, 0), obs_bits);
  free(m);
  return w;
}
double orc_bayes2(const pomcp_problem* p, double b, int32_t a, uint64_t obs_bits) {
  model* m = (model*)malloc(sizeof(model));
  double r = model_init(m, p) == POMCP_OK ? bayes2(m, b, a, obs_bits) : NAN;
  free(m);
  return r;
}
int32_t orc_is_terminal(const pomcp_problem* p, const void* state) {
  model* m = (model*)malloc(sizeof(model));
  if (model_init(m, p)) { free(m); return -1; }
  int t = is_terminal(m, load_state(m, state, 0));
  free(m);
  return t;
}

/* ------------------------------------------------------------------ */
/* rollout: POMDPToolbox.RolloutSimulator loop under RandomPolicy       */
/* (src/rollout.jl:77-84; loop restated in SURVEY.md App. B.1)          */
/* ------------------------------------------------------------------ */
static double rollout_random(const model* m, ostate s, int64_t steps, const uint32_t key[2], uint32_t tagword,
                             uint32_t q, uint32_t epoch, int64_t* nsteps) {
  const double gamma = m->p.discount;
  double disc = 1.0, total = 0.0;
  int64_t step = 0;
  uint32_t w[4];
  int64_t blk = -1;
  const int dps = m->draws_per_rollout_step;
  while (disc > 0.0 && !is_terminal(m, s) && step < steps) {
    /* RockSample draws its action from 16 bits (8 steps per Philox block, low half of a word first):
     * the rollout kernel is bound by instruction issue and the Philox rounds are a third of a step */
    const int halves = m->p.problem_id == POMCP_ROCKSAMPLE;
    int64_t word = halves ? step >> 1 : step * dps;
    if ((word >> 2) != blk) {
      blk = word >> 2;
      uint32_t ctr[4] = {(uint32_t)blk, tagword, q, epoch};
      orc_philox4x32_10(ctr, key, w);
    }
    uint32_t wa = w[word & 3];
    if (halves) wa = (step & 1) ? (wa & 0xffff0000u) : (wa << 16);
    int a = (int)(((uint64_t)wa * (uint64_t)m->nA) >> 32); /* rand(rng, actions): floor(u*nA) */
    double ut0 = dps == 2 ? u01(w[(word & 3) + 1]) : 0.0;
    ostate sp; uint64_t o; double r;
    model_step(m, s, a, ut0, 0.0, 0.0, 0, &sp, &o, &r);
    total += disc * r;
    s = sp;
    disc *= gamma;
    ++step;
  }
  if (nsteps) *nsteps += step;
  return total;
}

/* Rollout under POMDPModels.FeedWhenCrying: feed iff the previous observation was crying; the
 * "belief" starts as the label of the leaf node (extract_belief for PreviousObservationUpdater,
 * src/rollout.jl:105-106).  Two Philox words per step: transition, observation. */
static double rollout_fwc(const model* m, ostate s, int64_t steps, uint64_t last_obs, const uint32_t key[2],
                          uint32_t tagword, uint32_t q, uint32_t epoch, int64_t* nsteps) {
  const double gamma = m->p.discount;
  double disc = 1.0, total = 0.0;
  int64_t step = 0;
  uint32_t w[4];
  int64_t blk = -1;
  while (disc > 0.0 && !is_terminal(m, s) && step < steps) {
    int64_t word = step * 2;
    if ((word >> 2) != blk) {
      blk = word >> 2;
      uint32_t ctr[4] = {(uint32_t)blk, tagword, q, epoch};
      orc_philox4x32_10(ctr, key, w);
    }
    int a = last_obs ? 1 : 0;
    ostate sp; uint64_t o; double r;
    model_step(m, s, a, u01(w[word & 3]), u01(w[(word & 3) + 1]), 0.0, 1, &sp, &o, &r);
    total += disc * r;
    s = sp;
    last_obs = o;
    disc *= gamma;
    ++step;
  }
  if (nsteps) *nsteps += step;
  return total;
}

/* FORollout(FeedWhenCrying()) on Baby (src/rollout.jl:86-92 -> POMDPToolbox MDP simulate): the policy is
 * handed the state where it expects the previous observation, so it feeds iff the baby is hungry.
 * One Philox word per step (transition); generate_sr draws no observation. */
static double rollout_fo_fwc(const model* m, ostate s, int64_t steps, const uint32_t key[2], uint32_t tagword,
                             uint32_t q, uint32_t epoch, int64_t* nsteps) {
  const double gamma = m->p.discount;
  double disc = 1.0, total = 0.0;
  int64_t step = 0;
  uint32_t w[4];
  int64_t blk = -1;
  while (disc > 0.0 && !is_terminal(m, s) && step < steps) {
    if ((step >> 2) != blk) {
      blk = step >> 2;
      uint32_t ctr[4] = {(uint32_t)blk, tagword, q, epoch};
      orc_philox4x32_10(ctr, key, w);
    }
    int a = s.i ? 1 : 0;
    ostate sp; uint64_t o; double r;
    model_step(m, s, a, u01(w[step & 3]), 0.0, 0.0, 0, &sp, &o, &r);
    total += disc * r;
    s = sp;
    disc *= gamma;
    ++step;
  }
  if (nsteps) *nsteps += step;
  return total;
}

static int64_t state_index(const model* m, ostate s) {
  if (m->p.problem_id == POMCP_ROCKSAMPLE) return (int64_t)((uint64_t)s.i & 0xffffffull);
  return s.i;
}

/* ------------------------------------------------------------------ */
/* planner / tree (src/tree.jl:7-26 as flat arrays)                     */
/* ------------------------------------------------------------------ */
struct orc_planner {
  model m;
  pomcp_solver_params sp;
  int nA, nO;
  /* belief nodes */
  int64_t n_nodes, cap_nodes;
  int64_t* node_N;
  uint8_t* node_expanded;
  uint32_t* node_amask; /* action widening: bit a = ActNode a exists */
  double* node_bel;     /* POMCP_BELIEF_EXACT2: the node's exact two-state belief P(s = 1) */
  int64_t* node_phead; int64_t* node_ptail; int64_t* node_pcount; /* ParticleCollection per node */
  int64_t* node_parent_slot; uint64_t* node_obs;                 /* label */
  /* action nodes: slot = node*nA + a */
  int64_t* act_N; double* act_V;
  int32_t* child; /* discrete: [slot*nO + o] */
  /* continuous observations: open-addressed (slot, obs) -> node */
  int64_t ht_cap, ht_used; int32_t* ht;
  /* DPW observation widening: per action node the generated (child, state) pairs in generation order */
  int64_t* ev_n; int64_t* ev_cap; int64_t** ev_node; ostate** ev_state; int64_t* act_nc;
  /* particle log: linked per node, in push order */
  int64_t n_log, cap_log; ostate* log_state; int64_t* log_next;
  /* root */
  int64_t root;
  int belief_kind; /* 0 none, 1 initial distribution, 2 explicit particles, 3 node particles, 4 exact two-state belief */
  int64_t n_bel; ostate* bel;
  /* horizon tables */
  int64_t horizon;    /* first depth with gamma^depth < eps */
  double* gpow;       /* gamma^d, d = 0..horizon */
  double log_ratio;   /* log(eps)/log(gamma) */
  uint32_t epoch;
  double* state_values; int64_t n_state_values; /* FOValue table */
  pomcp_counters ctr;
  /* scratch path */
  int64_t cap_path; int64_t* path_slot; int64_t* path_node; double* path_r; ostate* path_sp;
};

#define GROW(ptr, type, n) ptr = (type*)realloc(ptr, sizeof(type) * (size_t)(n))

static void grow_nodes(orc_planner* pl, int64_t need) {
  if (need <= pl->cap_nodes) return;
  int64_t nc = pl->cap_nodes ? pl->cap_nodes : 1024;
  while (nc < need) nc *= 2;
  GROW(pl->node_N, int64_t, nc); GROW(pl->node_expanded, uint8_t, nc); GROW(pl->node_amask, uint32_t, nc);
  GROW(pl->node_bel, double, nc);
  GROW(pl->node_phead, int64_t, nc); GROW(pl->node_ptail, int64_t, nc); GROW(pl->node_pcount, int64_t, nc);
  GROW(pl->node_parent_slot, int64_t, nc); GROW(pl->node_obs, uint64_t, nc);
  GROW(pl->act_N, int64_t, nc * pl->nA); GROW(pl->act_V, double, nc * pl->nA);
  if (pl->nO > 0) GROW(pl->child, int32_t, nc * pl->nA * pl->nO);
  if (pl->sp.dpw_observation) {
    GROW(pl->ev_n, int64_t, nc * pl->nA); GROW(pl->ev_cap, int64_t, nc * pl->nA); GROW(pl->act_nc, int64_t, nc * pl->nA);
    GROW(pl->ev_node, int64_t*, nc * pl->nA); GROW(pl->ev_state, ostate*, nc * pl->nA);
    for (int64_t k = pl->cap_nodes * pl->nA; k < nc * pl->nA; ++k) { pl->ev_n[k] = pl->ev_cap[k] = pl->act_nc[k] = 0; pl->ev_node[k] = NULL; pl->ev_state[k] = NULL; }
  }
  pl->cap_nodes = nc;
}

static int64_t new_node(orc_planner* pl, int64_t parent_slot, uint64_t obs) {
  grow_nodes(pl, pl->n_nodes + 1);
  int64_t n = pl->n_nodes++;
  pl->node_N[n] = 0; pl->node_expanded[n] = 0; pl->node_amask[n] = 0; pl->node_bel[n] = 0.0;
  pl->node_phead[n] = -1; pl->node_ptail[n] = -1; pl->node_pcount[n] = 0;
  pl->node_parent_slot[n] = parent_slot; pl->node_obs[n] = obs;
  for (int a = 0; a < pl->nA; ++a) { pl->act_N[n * pl->nA + a] = 0; pl->act_V[n * pl->nA + a] = 0.0; }
  if (pl->nO > 0)
    for (int j = 0; j < pl->nA * pl->nO; ++j) pl->child[n * pl->nA * pl->nO + j] = -1;
  pl->ctr.nodes = pl->n_nodes;
  return n;
}

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
  return x;
}
static void ht_rehash(orc_planner* pl, int64_t ncap) {
  int32_t* nt = (int32_t*)malloc(sizeof(int32_t) * (size_t)ncap);
  for (int64_t i = 0; i < ncap; ++i) nt[i] = -1;
  for (int64_t i = 0; i < pl->ht_cap; ++i) {
    int32_t id = pl->ht[i];
    if (id < 0) continue;
    uint64_t h = mix64((uint64_t)pl->node_parent_slot[id] * 0x9E3779B97F4A7C15ull ^ pl->node_obs[id]) & (uint64_t)(ncap - 1);
    while (nt[h] >= 0) h = (h + 1) & (uint64_t)(ncap - 1);
    nt[h] = id;
  }
  free(pl->ht); pl->ht = nt; pl->ht_cap = ncap;
}
/* haskey(best_node.children, o) / best_node.children[o] (src/solver.jl:132-142) */
static int64_t find_child(orc_planner* pl, int64_t slot, uint64_t obs) {
  if (pl->nO > 0) {
    if (obs >= (uint64_t)pl->nO) return -1;
    return pl->child[slot * pl->nO + (int64_t)obs];
  }
  if (pl->ht_cap == 0) return -1;
  uint64_t h = mix64((uint64_t)slot * 0x9E3779B97F4A7C15ull ^ obs) & (uint64_t)(pl->ht_cap - 1);
  for (;;) {
    int32_t id = pl->ht[h];
    if (id < 0) return -1;
    if (pl->node_parent_slot[id] == slot && pl->node_obs[id] == obs) return id;
    h = (h + 1) & (uint64_t)(pl->ht_cap - 1);
  }
}
static int64_t insert_child(orc_planner* pl, int64_t slot, uint64_t obs) {
  int64_t n = new_node(pl, slot, obs);
  if (pl->nO > 0) { pl->child[slot * pl->nO + (int64_t)obs] = (int32_t)n; return n; }
  if ((pl->ht_used + 1) * 2 > pl->ht_cap) ht_rehash(pl, pl->ht_cap ? pl->ht_cap * 2 : 4096);
  uint64_t h = mix64((uint64_t)slot * 0x9E3779B97F4A7C15ull ^ obs) & (uint64_t)(pl->ht_cap - 1);
  while (pl->ht[h] >= 0) h = (h + 1) & (uint64_t)(pl->ht_cap - 1);
  pl->ht[h] = (int32_t)n; pl->ht_used++;
  return n;
}

/* push!(hao.B, sp) (src/particle_filter.jl:30, called at src/solver.jl:146-148) */
static void push_particle(orc_planner* pl, int64_t node, ostate s) {
  if (pl->n_log + 1 > pl->cap_log) {
    int64_t nc = pl->cap_log ? pl->cap_log * 2 : 4096;
    GROW(pl->log_state, ostate, nc); GROW(pl->log_next, int64_t, nc);
    pl->cap_log = nc;
  }
  int64_t e = pl->n_log++;
  pl->log_state[e] = s; pl->log_next[e] = -1;
  if (pl->node_ptail[node] < 0) pl->node_phead[node] = e; else pl->log_next[pl->node_ptail[node]] = e;
  pl->node_ptail[node] = e;
  pl->node_pcount[node]++;
  pl->ctr.log_entries = pl->n_log;
}

static void free_tree(orc_planner* pl) {
  if (pl->ev_node) {
    for (int64_t k = 0; k < pl->cap_nodes * pl->nA; ++k) { free(pl->ev_node[k]); free(pl->ev_state[k]); }
    free(pl->ev_n); free(pl->ev_cap); free(pl->ev_node); free(pl->ev_state); free(pl->act_nc);
    pl->ev_n = pl->ev_cap = pl->act_nc = NULL; pl->ev_node = NULL; pl->ev_state = NULL;
  }
  free(pl->node_N); free(pl->node_expanded); free(pl->node_amask); pl->node_amask = NULL; free(pl->node_bel); pl->node_bel = NULL; free(pl->node_phead); free(pl->node_ptail); free(pl->node_pcount);
  free(pl->node_parent_slot); free(pl->node_obs); free(pl->act_N); free(pl->act_V); free(pl->child); free(pl->ht);
  free(pl->log_state); free(pl->log_next);
  pl->node_N = NULL; pl->node_expanded = NULL; pl->node_phead = pl->node_ptail = pl->node_pcount = NULL;
  pl->node_parent_slot = NULL; pl->node_obs = NULL; pl->act_N = NULL; pl->act_V = NULL; pl->child = NULL; pl->ht = NULL;
  pl->log_state = NULL; pl->log_next = NULL;
  pl->n_nodes = pl->cap_nodes = 0; pl->ht_cap = pl->ht_used = 0; pl->n_log = pl->cap_log = 0;
}

/* RootNode(0, b, {}) (src/solver.jl:278-281, src/updater.jl:45-47) */
static void fresh_root(orc_planner* pl) {
  free_tree(pl);
  pl->root = new_node(pl, -1, 0);
}

orc_planner* orc_create(const pomcp_problem* p, const pomcp_solver_params* sp) {
  orc_planner* pl = (orc_planner*)calloc(1, sizeof(orc_planner));
  if (model_init(&pl->m, p) != POMCP_OK) { free(pl); return NULL; }
  pl->sp = *sp;
  if (sp->num_sparse_actions > 0 || !(sp->eps > 0.0) || !(sp->eps < 1.0)) { free(pl); return NULL; }
  if (sp->rollouts_per_leaf < 0 || sp->rollouts_per_leaf > 64) { free(pl); return NULL; }
  if (sp->dpw_observation && (!(sp->k_observation >= 0.0) || !(sp->alpha_observation >= 0.0))) { free(pl); return NULL; }
  if (sp->dpw_action && (!(sp->k_action >= 0.0) || !(sp->alpha_action >= 0.0) || pl->m.nA > 32)) { free(pl); return NULL; }
  if ((sp->estimator == POMCP_EST_FWC_ROLLOUT || sp->estimator == POMCP_EST_FO_FWC_ROLLOUT) &&
      p->problem_id != POMCP_BABY) { free(pl); return NULL; }
  if (sp->estimator == POMCP_EST_STATE_VALUE && p->problem_id == POMCP_LIGHTDARK1D) { free(pl); return NULL; }
  if (sp->belief_mode == POMCP_BELIEF_EXACT2 && ((p->problem_id != POMCP_TIGER && p->problem_id != POMCP_BABY) || sp->dpw_action)) { free(pl); return NULL; }
  pl->nA = pl->m.nA; pl->nO = pl->m.nO;
  /* cut-off gamma^depth < eps (src/solver.jl:82) and steps_to_eps (src/solver.jl:100) */
  pl->log_ratio = log(sp->eps) / log(p->discount);
  int64_t d = 0;
  while (pow(p->discount, (double)d) >= sp->eps) ++d;
  pl->horizon = d;
  pl->gpow = (double*)malloc(sizeof(double) * (size_t)(d + 1));
  for (int64_t i = 0; i <= d; ++i) pl->gpow[i] = pow(p->discount, (double)i);
  pl->cap_path = d + 2;
  pl->path_slot = (int64_t*)malloc(sizeof(int64_t) * (size_t)pl->cap_path);
  pl->path_node = (int64_t*)malloc(sizeof(int64_t) * (size_t)pl->cap_path);
  pl->path_r = (double*)malloc(sizeof(double) * (size_t)pl->cap_path);
  pl->path_sp = (ostate*)malloc(sizeof(ostate) * (size_t)pl->cap_path);
  fresh_root(pl);
  return pl;
}

void orc_destroy(orc_planner* pl) {
  if (!pl) return;
  free_tree(pl);
  free(pl->bel); free(pl->gpow); free(pl->path_slot); free(pl->path_node); free(pl->path_r); free(pl->path_sp);
  free(pl->state_values);
  free(pl);
}

int32_t orc_set_belief_particles(orc_planner* pl, const void* states, int64_t n) {
  if (n <= 0 || n >= ((int64_t)1 << 32)) return POMCP_ERR_INVALID_ARG;
  fresh_root(pl);
  free(pl->bel);
  pl->bel = (ostate*)malloc(sizeof(ostate) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) pl->bel[i] = load_state(&pl->m, states, i);
  pl->n_bel = n; pl->belief_kind = 2;
  return POMCP_OK;
}
int32_t orc_set_state_values(orc_planner* pl, const double* values, int64_t n) {
  if (!values || n <= 0) return POMCP_ERR_INVALID_ARG;
  free(pl->state_values);
  pl->state_values = (double*)malloc(sizeof(double) * (size_t)n);
  memcpy(pl->state_values, values, sizeof(double) * (size_t)n);
  pl->n_state_values = n;
  return POMCP_OK;
}
int32_t orc_set_belief_initial(orc_planner* pl) {
  fresh_root(pl);
  free(pl->bel); pl->bel = NULL; pl->n_bel = 0; pl->belief_kind = 1;
  return POMCP_OK;
}
/* initialize_belief(up, BoolDistribution(p)) = RootNode(0, b, {}) (src/updater.jl:45-47) */
int32_t orc_set_belief_exact2(orc_planner* pl, double p1) {
  if (pl->sp.belief_mode != POMCP_BELIEF_EXACT2) return POMCP_ERR_UNSUPPORTED;
  if (!(p1 >= 0.0) || !(p1 <= 1.0)) return POMCP_ERR_INVALID_ARG;
  fresh_root(pl);
  free(pl->bel); pl->bel = NULL; pl->n_bel = 0; pl->belief_kind = 4;
  pl->node_bel[pl->root] = p1;
  return POMCP_OK;
}
int32_t orc_get_belief_exact2(orc_planner* pl, double* p1) {
  if (pl->belief_kind != 4) return POMCP_ERR_INVALID_ARG;
  *p1 = pl->node_bel[pl->root];
  return POMCP_OK;
}
static int64_t find_child(orc_planner* pl, int64_t slot, uint64_t obs);
/* update(::RootUpdater, b_old, a, o) with a user-supplied belief updater, src/updater.jl:30-39 */
int32_t orc_update_exact2(orc_planner* pl, int32_t a, uint64_t obs) {
  if (pl->belief_kind != 4) return POMCP_ERR_UNSUPPORTED;
  if (a < 0 || a >= pl->nA || obs >= (uint64_t)pl->nO) return POMCP_ERR_INVALID_ARG;
  int64_t hao = pl->node_expanded[pl->root] ? find_child(pl, pl->root * pl->nA + a, obs) : -1;
  if (hao >= 0) { pl->root = hao; return POMCP_OK; } /* the child with its subtree */
  const double nb = bayes2(&pl->m, pl->node_bel[pl->root], a, obs);
  fresh_root(pl); /* ObsNode(o, 0, new_belief, {}) */
  pl->node_bel[pl->root] = nb;
  return POMCP_OK;
}

int32_t orc_get_belief_particles(orc_planner* pl, void* out, int64_t cap, int64_t* n) {
  if (pl->belief_kind == 2) {
    *n = pl->n_bel;
    for (int64_t i = 0; i < pl->n_bel && i < cap; ++i) store_state(&pl->m, out, i, pl->bel[i]);
    return POMCP_OK;
  }
  if (pl->belief_kind == 3) {
    *n = pl->node_pcount[pl->root];
    int64_t i = 0;
    for (int64_t e = pl->node_phead[pl->root]; e >= 0 && i < cap; e = pl->log_next[e], ++i)
      store_state(&pl->m, out, i, pl->log_state[e]);
    return POMCP_OK;
  }
  *n = 0;
  return POMCP_OK;
}

/* ---- randomness plumbing ---- */
typedef struct orng {
  int replay; const double* u; int64_t nu, cur; int overflow;
  const double* ret; int64_t nret, rcur;
  uint32_t key[2]; uint32_t rank; uint32_t epoch;
} orng;

static void draws(orng* g, uint32_t tag, uint32_t idx, uint32_t q, int n, double* out) {
  if (g->replay) {
    for (int i = 0; i < n; ++i) {
      if (g->cur < g->nu) out[i] = g->u[g->cur++]; else { out[i] = 0.0; g->overflow = 1; }
    }
    return;
  }
  uint32_t ctr[4] = {idx, tag | (g->rank << 8), q, g->epoch}, w[4];
  orc_philox4x32_10(ctr, g->key, w);
  for (int i = 0; i < n; ++i) out[i] = u01(w[i]);
}

/* rand(rng, initial_state_distribution(pomdp)) from two uniforms */
static ostate initial_state(const pomcp_problem* p, double u0, double u1) {
  ostate s; s.i = 0; s.y = 0.0;
  switch (p->problem_id) {
    case POMCP_TIGER: s.i = u0 < 0.5 ? 1 : 0; break;
    case POMCP_BABY: s.i = 0; break; /* BoolDistribution(0.0) */
    case POMCP_ROCKSAMPLE: {
      uint32_t rocks = (uint32_t)(u0 * (double)(1u << p->rs_k));
      s.i = (int64_t)((uint32_t)p->rs_start_x | ((uint32_t)p->rs_start_y << 4) | (rocks << 8));
      break;
    }
    case POMCP_LIGHTDARK1D: s.i = 0; s.y = p->ld_init_mean + orc_normal_from_uniforms(u0, u1) * p->ld_init_std; break;
  }
  return s;
}

/* rand(rng, b) at src/solver.jl:48 -> src/updater.jl:53-55 */
static ostate sample_root(orc_planner* pl, orng* g, uint32_t q) {
  double u[2];
  draws(g, TAG_ROOT, 0, q, 2, u);
  const pomcp_problem* p = &pl->m.p;
  if (pl->belief_kind == 2) {
    return pl->bel[(int64_t)(u[0] * (double)pl->n_bel)];
  } else if (pl->belief_kind == 4) { /* rand(rng, BoolDistribution(p)) */
    ostate s; s.i = u[0] < pl->node_bel[pl->root] ? 1 : 0; s.y = 0.0;
    return s;
  } else if (pl->belief_kind == 3) {
    int64_t k = (int64_t)(u[0] * (double)pl->node_pcount[pl->root]);
    int64_t e = pl->node_phead[pl->root];
    while (k-- > 0) e = pl->log_next[e];
    return pl->log_state[e];
  }
  return initial_state(p, u[0], u[1]);
}

/* simulate() for POMCPSolver, src/solver.jl:78-157, recursion unrolled into a
 * descent loop plus a bottom-up backup loop (same order of side effects). */
static void simulate(orc_planner* pl, orng* g, ostate s, uint32_t q) {
  const model* m = &pl->m;
  const pomcp_solver_params* sol = &pl->sp;
  const double gamma = m->p.discount;
  int64_t h = pl->root;
  int64_t depth = 0;
  double leaf = 0.0;
  for (;;) {
    /* src/solver.jl:82-86 */
    if (depth >= pl->horizon || is_terminal(m, s) || depth >= sol->max_depth) { leaf = 0.0; break; }
    int leaf_now = 0;
    int64_t totalN = pl->node_N[h];
    if (sol->dpw_action) {
      /* src/solver.jl:172-186: action progressive widening */
      uint32_t mask = pl->node_amask[h];
      totalN = 0;
      for (int a = 0; a < pl->nA; ++a) if ((mask >> a) & 1u) totalN += pl->act_N[h * pl->nA + a];
      int nch = __builtin_popcount(mask);
      double lim = totalN <= 0 ? 0.0
                 : sol->alpha_action == 0.5 ? sol->k_action * sqrt((double)totalN)
                 : sol->k_action * orc_det_exp(sol->alpha_action * orc_det_log((double)totalN));
      if ((double)nch <= lim) {
        double ua[1];
        draws(g, TAG_TREE, (uint32_t)depth | 0x80000000u, q, 1, ua); /* next_action: rand(rng, actions(pomdp)) */
        int an = (int)(ua[0] * (double)pl->nA);
        if (!((mask >> an) & 1u)) {
          pl->act_N[h * pl->nA + an] = sol->init_N; pl->act_V[h * pl->nA + an] = sol->init_V;
          mask |= 1u << an;
          pl->node_amask[h] = mask;
        }
        if (__builtin_popcount(mask) <= 1) leaf_now = 1; /* src/solver.jl:180-186 */
      }
    } else if (!pl->node_expanded[h]) {
      /* src/solver.jl:87-94: one ActNode per action, N=init_N, V=init_V */
      for (int a = 0; a < pl->nA; ++a) { pl->act_N[h * pl->nA + a] = sol->init_N; pl->act_V[h * pl->nA + a] = sol->init_V; }
      pl->node_expanded[h] = 1;
      leaf_now = 1;
    }
    if (leaf_now) {
      if (depth > 0) {
        /* src/solver.jl:97-103 */
        int64_t steps_to_eps = (int64_t)ceil(pl->log_ratio - (double)depth);
        int64_t steps = sol->max_depth - depth < steps_to_eps ? sol->max_depth - depth : steps_to_eps;
        /* POMCPDPWSolver hands `depth` to estimate_value as the step budget (src/solver.jl:182,199) */
        if (sol->dpw_solver || sol->dpw_observation || sol->dpw_action) steps = depth;
        double est;
        if (g->replay) {
          if (g->rcur < g->nret) est = g->ret[g->rcur++]; else { est = 0.0; g->overflow = 1; }
        } else if (sol->estimator == POMCP_EST_CONSTANT) {
          est = sol->estimator_value; /* src/rollout.jl:10 */
        } else if (sol->estimator == POMCP_EST_STATE_VALUE) { /* value(policy, s), src/rollout.jl:67-69 */
          int64_t si = state_index(m, s);
          est = (is_terminal(m, s) || si < 0 || si >= pl->n_state_values) ? 0.0 : pl->state_values[si];
        } else {
          /* R rollouts per leaf (device extension; R = 1 is the reference).  Rollout r draws from the
           * stream tagged r<<16; the R returns are added in the butterfly order of a 32-lane warp. */
          int R = sol->rollouts_per_leaf > 0 ? sol->rollouts_per_leaf : 1;
          double x[64], y[32];
          for (int r = 0; r < 64; ++r)
            x[r] = r >= R ? 0.0
                   : sol->estimator == POMCP_EST_FWC_ROLLOUT
                       ? rollout_fwc(m, s, steps, pl->node_obs[h], g->key,
                                     TAG_ROLLOUT | (g->rank << 8) | ((uint32_t)r << 16), q, g->epoch,
                                     &pl->ctr.rollout_steps)
                   : sol->estimator == POMCP_EST_FO_FWC_ROLLOUT
                       ? rollout_fo_fwc(m, s, steps, g->key, TAG_ROLLOUT | (g->rank << 8) | ((uint32_t)r << 16), q,
                                        g->epoch, &pl->ctr.rollout_steps)
                       : rollout_random(m, s, steps, g->key, TAG_ROLLOUT | (g->rank << 8) | ((uint32_t)r << 16), q,
                                        g->epoch, &pl->ctr.rollout_steps);
          if (R > 32) /* lane l of the warp holds rollouts l and l + 32 */
            for (int i = 0; i < 32; ++i) x[i] = x[i] + x[i + 32];
          for (int off = 16; off > 0; off >>= 1) {
            for (int i = 0; i < 32; ++i) y[i] = x[i] + x[i ^ off];
            memcpy(x, y, sizeof(y));
          }
          est = x[0] / (double)R;
          pl->ctr.rollouts += R - 1;
        }
        pl->ctr.rollouts++;
        leaf = pl->gpow[depth] * est;
      } else {
        leaf = 0.0; /* src/solver.jl:104-106 */
      }
      break;
    }
    /* UCB1, src/solver.jl:109-124 */
    double best = -INFINITY;
    int best_a = -1;
    int64_t hN = totalN; /* h.N, or the sum of the children's N under action widening (src/solver.jl:171) */
    for (int a = 0; a < pl->nA; ++a) {
      if (sol->dpw_action && !((pl->node_amask[h] >> a) & 1u)) continue; /* only existing children */
      int64_t nN = pl->act_N[h * pl->nA + a];
      double nV = pl->act_V[h * pl->nA + a];
      double crit;
      if (nN == 0 && ((sol->dpw_solver || sol->dpw_observation || sol->dpw_action) ? hN <= 1 : hN == 1)) crit = nV; /* src/solver.jl:112 / :210 */
      else if (nN == 0 && nV == -INFINITY) crit = INFINITY;
      else crit = nV + sol->c * sqrt(orc_det_log((double)hN) / (double)nN);
      if (crit >= best) { best = crit; best_a = a; }
    }
    if (best_a < 0) best_a = pl->nA - 1; /* all-NaN criteria: Julia would hit an undefined local; never happens with c>0 */
    /* src/solver.jl:126 */
    double u[4];
    draws(g, TAG_TREE, (uint32_t)depth, q, 4, u);
    ostate sp; uint64_t o; double r;
    model_step(m, s, best_a, u[0], u[2], u[3], 1, &sp, &o, &r);
    pl->ctr.tree_steps++;
    int64_t slot = h * pl->nA + best_a;
    int64_t hao;
    int generated = 1;
    if (sol->dpw_observation) {
      /* src/solver.jl:223: length(best_node.children) <= k_observation * best_node.N^alpha_observation */
      int64_t bn = pl->act_N[slot];
      double lim = bn <= 0 ? 0.0
                 : sol->alpha_observation == 0.5 ? sol->k_observation * sqrt((double)bn)
                 : sol->k_observation * orc_det_exp(sol->alpha_observation * orc_det_log((double)bn));
      generated = (double)pl->act_nc[slot] <= lim;
    }
    if (generated) {
      hao = find_child(pl, slot, o);
      if (hao < 0) { /* src/solver.jl:132-142 */
        /* new_belief = update(pomcp.node_belief_updater, h.B, a, o) (src/solver.jl:138) */
        const double nb = sol->belief_mode == POMCP_BELIEF_EXACT2 ? bayes2(m, pl->node_bel[h], best_a, o) : 0.0;
        hao = insert_child(pl, slot, o);
        pl->node_bel[hao] = nb;
        if (sol->dpw_observation) pl->act_nc[slot]++;
      }
      if (sol->dpw_observation) {
        /* state_was_generated: push!(hao.B, sp); hao.N += 1 BEFORE the recursive call (src/solver.jl:257-262) */
        if (pl->ev_n[slot] == pl->ev_cap[slot]) {
          int64_t ncap = pl->ev_cap[slot] ? pl->ev_cap[slot] * 2 : 4;
          pl->ev_node[slot] = (int64_t*)realloc(pl->ev_node[slot], sizeof(int64_t) * (size_t)ncap);
          pl->ev_state[slot] = (ostate*)realloc(pl->ev_state[slot], sizeof(ostate) * (size_t)ncap);
          pl->ev_cap[slot] = ncap;
        }
        pl->ev_node[slot][pl->ev_n[slot]] = hao; pl->ev_state[slot][pl->ev_n[slot]] = sp; pl->ev_n[slot]++;
        if (sol->belief_mode == POMCP_BELIEF_UNWEIGHTED) push_particle(pl, hao, sp);
        pl->node_N[hao] += 1;
      }
    } else {
      /* src/solver.jl:249-255: hao ~ children weighted by N, sp ~ hao.B, r = reward(s, a, sp).  Every generation
       * event added one to its child's N and one state to its B, so a uniform draw over the events of this
       * action node is that joint distribution. */
      int64_t e = (int64_t)(u[1] * (double)pl->ev_n[slot]); /* word 1 of the level's block: unused by every model */
      hao = pl->ev_node[slot][e];
      sp = pl->ev_state[slot][e]; /* r of model_step above is reward(s, a, .): it does not depend on sp in these models */
    }
    pl->path_slot[depth] = slot; pl->path_node[depth] = generated ? hao : -hao - 1; pl->path_r[depth] = r; pl->path_sp[depth] = sp;
    h = hao; s = sp; ++depth;
  }
  if (depth > pl->ctr.max_depth_seen) pl->ctr.max_depth_seen = depth;
  /* src/solver.jl:144-156, deepest level first */
  double R = leaf;
  for (int64_t d = depth - 1; d >= 0; --d) {
    R = pl->path_r[d] + gamma * R;
    int64_t hao = pl->path_node[d], slot = pl->path_slot[d];
    if (!sol->dpw_observation) { /* DPW did both before descending, and only for generated states */
      if (sol->belief_mode == POMCP_BELIEF_UNWEIGHTED) push_particle(pl, hao, pl->path_sp[d]);
      pl->node_N[hao] += 1;
    }
    pl->act_N[slot] += 1;
    if (pl->act_V[slot] != -INFINITY) pl->act_V[slot] += (R - pl->act_V[slot]) / (double)pl->act_N[slot];
  }
}

static int32_t root_argmax(orc_planner* pl, int32_t* action_out, int64_t* root_N_out, double* root_V_out, int32_t nA) {
  /* src/solver.jl:60-70: `for node in values(b.children)` - only ActNodes that exist.  Under action
   * widening a slot that was never added is not a child (its N / V are reported as 0 / init_V). */
  int64_t h = pl->root;
  const uint32_t mask = pl->sp.dpw_action ? pl->node_amask[h] : 0xffffffffu;
  double best_V = 0.0;
  int best = -1;
  for (int a = 0; a < pl->nA; ++a) {
    if (!((mask >> a) & 1u)) continue;
    double v = pl->act_V[h * pl->nA + a];
    if (best < 0 && v != v) return POMCP_ERR_NAN_VALUE; /* @assert !isnan(best_V) on the first child, src/solver.jl:62 */
    if (best < 0 || v >= best_V) { best_V = v; best = a; }
  }
  if (action_out) *action_out = best < 0 ? 0 : best;
  for (int a = 0; a < pl->nA && a < nA; ++a) {
    const int exists = (mask >> a) & 1u;
    if (root_N_out) root_N_out[a] = exists ? pl->act_N[h * pl->nA + a] : 0;
    if (root_V_out) root_V_out[a] = exists ? pl->act_V[h * pl->nA + a] : pl->sp.init_V;
  }
  return POMCP_OK;
}

static int32_t search_impl(orc_planner* pl, orng* g, int64_t Q, int32_t* action_out, int64_t* root_N_out,
                           double* root_V_out, int32_t nA) {
  if (pl->belief_kind == 0) return POMCP_ERR_INVALID_ARG;
  if (Q < 0) Q = pl->sp.tree_queries;
  int all_terminal = 1;
  for (int64_t q = 0; q < Q; ++q) { /* src/solver.jl:47-54 */
    ostate s = sample_root(pl, g, (uint32_t)q);
    if (is_terminal(&pl->m, s)) { pl->ctr.terminal_skips++; continue; }
    simulate(pl, g, s, (uint32_t)q);
    pl->node_N[pl->root] += 1;
    pl->ctr.simulations++;
    all_terminal = 0;
  }
  pl->epoch++;
  if (all_terminal) return POMCP_ERR_ALL_SAMPLES_TERMINAL; /* src/solver.jl:56-58 */
  return root_argmax(pl, action_out, root_N_out, root_V_out, nA);
}

int32_t orc_search(orc_planner* pl, int64_t Q, int32_t* action_out, int64_t* root_N_out, double* root_V_out, int32_t nA) {
  orng g; memset(&g, 0, sizeof(g));
  g.key[0] = (uint32_t)pl->sp.seed; g.key[1] = (uint32_t)(pl->sp.seed >> 32);
  g.rank = (uint32_t)pl->sp.rank; g.epoch = pl->epoch;
  return search_impl(pl, &g, Q, action_out, root_N_out, root_V_out, nA);
}

int32_t orc_search_replay(orc_planner* pl, int64_t Q, const double* uniforms, int64_t nu, const double* rets, int64_t nr,
                          int32_t* action_out, int64_t* root_N_out, double* root_V_out, int32_t nA, int64_t* u_used,
                          int64_t* r_used) {
  orng g; memset(&g, 0, sizeof(g));
  g.replay = 1; g.u = uniforms; g.nu = nu; g.ret = rets; g.nret = nr;
  int32_t rc = search_impl(pl, &g, Q, action_out, root_N_out, root_V_out, nA);
  if (u_used) *u_used = g.cur;
  if (r_used) *r_used = g.rcur;
  if (g.overflow) return POMCP_ERR_INVALID_ARG;
  return rc;
}

/* update(::RootUpdater{<:ParticleReinvigorator}, b_old, a, o) (src/updater.jl:13-27) with the
 * DeadReinvigorator (src/particle_filter.jl:85-101) */
int32_t orc_update_unweighted(orc_planner* pl, int32_t a, uint64_t obs_bits) {
  if (a < 0 || a >= pl->nA) return POMCP_ERR_INVALID_ARG;
  if (pl->sp.belief_mode != POMCP_BELIEF_UNWEIGHTED) return POMCP_ERR_UNSUPPORTED;
  if (!pl->node_expanded[pl->root] && !pl->node_amask[pl->root]) return POMCP_ERR_INVALID_ARG; /* KeyError on b_old.children[a] */
  int64_t hao = find_child(pl, pl->root * pl->nA + a, obs_bits);
  if (hao < 0) return POMCP_ERR_PARTICLE_DEPLETION;          /* handle_unseen_observation */
  if (pl->node_pcount[hao] == 0) return POMCP_ERR_PARTICLE_DEPLETION; /* reinvigorate! on empty collection */
  pl->root = hao; pl->belief_kind = 3;
  return POMCP_OK;
}

/* ------------------------------------------------------------------ */
/* SIR particle filter (ParticleFilters.jl SimpleParticleFilter +        */
/* LowVarianceResampler; call site test/runtests.jl:57) */
/* ------------------------------------------------------------------ */
int32_t orc_resample_systematic_f64(const double* w, int64_t n, double r01, int64_t n_out, int64_t* idx_out) {
  /* upstream running sums: r = rand*W/n; c = w[1]; U = r; while U > c: i++, c += w[i]; U += W/n */
  if (n <= 0 || n_out <= 0) return POMCP_ERR_INVALID_ARG;
  double W = 0.0;
  for (int64_t i = 0; i < n; ++i) W += w[i];
  if (!(W > 0.0)) return POMCP_ERR_PARTICLE_DEPLETION;
  double step = W / (double)n_out;
  double U = r01 * W / (double)n_out;
  double c = w[0];
  int64_t i = 0;
  for (int64_t mth = 0; mth < n_out; ++mth) {
    while (U > c) {
      if (i + 1 >= n) break; /* rounding at the very end: upstream would throw BoundsError */
      ++i; c += w[i];
    }
    U += step;
    idx_out[mth] = i;
  }
  return POMCP_OK;
}

static int ceil_log2_i64(int64_t n) { int L = 0; while (((int64_t)1 << L) < n) ++L; return L; }
/* analytic ceiling of obs_weight(): weights < 2^(E + 1) */
static int weight_ceil_exp(const model* m) { return m->p.problem_id == POMCP_LIGHTDARK1D ? 5 : 0; }

/* Fixed-point restatement: q_i = trunc(w_i * 2^s); output m takes the first i with
 * C_i * n_out * 2^32 >= (ru + m*2^32) * T  (ru = floor(r*2^32)).  Integer sums are associative, so a parallel
 * scan gives the same indices in any order.
 * Scale rule: s = 61 - ceil(log2 n) - E keeps sum q < 2^62 when every weight is below 2^(E+1).
 *   - weights of a model have an analytic ceiling (probabilities: E_ceil = 0; LightDark's Normal pdf with
 *     sigma >= 0.01: at most 39.9 < 2^6, E_ceil = 5).  That a-priori scale is used when no weight exceeds the ceiling
 *     and the total keeps at least 40 bits (sum q >= 2^40): the device then needs no pass over the data to find
 *     the largest weight;
 *   - otherwise (and for replayed weights, which have no ceiling) E = floor(log2 max w). */
#define ORC_NO_CEIL (-100000)
static int32_t resample_fixed(const double* w, const uint8_t* alive, int64_t n, uint32_t ru, int64_t n_out,
                              int64_t* idx_out, int e_ceil) {
  if (n <= 0 || n_out <= 0 || n_out >= ((int64_t)1 << 32) || n >= ((int64_t)1 << 32)) return POMCP_ERR_INVALID_ARG;
  double maxw = 0.0;
  for (int64_t i = 0; i < n; ++i) if ((!alive || alive[i]) && w[i] > maxw) maxw = w[i];
  if (!(maxw > 0.0) || maxw == INFINITY) return POMCP_ERR_PARTICLE_DEPLETION;
  int e = (int)((d2u(maxw) >> 52) & 0x7ff) - 1023; /* floor(log2 maxw) for normal doubles */
  int s = 61 - ceil_log2_i64(n) - e;
  uint64_t T = 0;
  if (e_ceil != ORC_NO_CEIL && e <= e_ceil) {
    int sc = 61 - ceil_log2_i64(n) - e_ceil;
    uint64_t Tc = 0;
    for (int64_t i = 0; i < n; ++i) Tc += ((!alive || alive[i]) && w[i] > 0.0) ? (uint64_t)ldexp(w[i], sc) : 0;
    if (Tc >= ((uint64_t)1 << 40)) { s = sc; T = Tc; }
  }
  if (T == 0)
    for (int64_t i = 0; i < n; ++i) T += ((!alive || alive[i]) && w[i] > 0.0) ? (uint64_t)ldexp(w[i], s) : 0;
  if (T == 0) return POMCP_ERR_PARTICLE_DEPLETION;
  int64_t i = -1;
  uint64_t C = 0;
  /* first alive particle */
  for (int64_t mth = 0; mth < n_out; ++mth) {
    unsigned __int128 rhs = ((unsigned __int128)ru + ((unsigned __int128)mth << 32)) * T;
    for (;;) {
      if (i >= 0) {
        unsigned __int128 lhs = ((unsigned __int128)C * (uint64_t)n_out) << 32;
        if (lhs >= rhs) break;
      }
      /* advance to next alive particle */
      int64_t j = i + 1;
      while (j < n && alive && !alive[j]) ++j;
      if (j >= n) break;
      i = j;
      C += (w[i] > 0.0) ? (uint64_t)ldexp(w[i], s) : 0;
    }
    idx_out[mth] = i;
  }
  return POMCP_OK;
}

int32_t orc_resample_systematic_fixed(const double* w, int64_t n, double r01, int64_t n_out, int64_t* idx_out) {
  if (!(r01 >= 0.0) || !(r01 < 1.0)) return POMCP_ERR_INVALID_ARG;
  return resample_fixed(w, NULL, n, (uint32_t)(r01 * 4294967296.0), n_out, idx_out, ORC_NO_CEIL);
}

static void pf_propagate(const model* m, const void* states, int64_t n, int a, uint64_t obs, uint64_t seed,
                         ostate* sp_out, double* w_out, uint8_t* alive_out) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int64_t i = 0; i < n; ++i) {
    ostate s = load_state(m, states, i);
    if (is_terminal(m, s)) { alive_out[i] = 0; w_out[i] = 0.0; sp_out[i] = s; continue; }
    uint32_t ctr[4] = {0, TAG_PF, (uint32_t)i, (uint32_t)((uint64_t)i >> 32)}, wd[4];
    orc_philox4x32_10(ctr, key, wd);
    ostate sp; uint64_t o; double r;
    model_step(m, s, a, u01(wd[0]), 0.0, 0.0, 0, &sp, &o, &r); /* generate_s */
    sp_out[i] = sp; alive_out[i] = 1;
    w_out[i] = obs_weight(m, a, sp, obs);
  }
}

int32_t orc_pf_weights(const pomcp_problem* p, const void* states, int64_t n, int32_t a, uint64_t obs, uint64_t seed,
                       double* weights_out, void* next_states_out) {
  model* m = (model*)malloc(sizeof(model));
  int rc = model_init(m, p);
  if (rc) { free(m); return rc; }
  ostate* sp = (ostate*)malloc(sizeof(ostate) * (size_t)n);
  uint8_t* alive = (uint8_t*)malloc((size_t)n);
  pf_propagate(m, states, n, a, obs, seed, sp, weights_out, alive);
  if (next_states_out) for (int64_t i = 0; i < n; ++i) store_state(m, next_states_out, i, sp[i]);
  free(sp); free(alive); free(m);
  return POMCP_OK;
}

int32_t orc_pf_update_sir(orc_planner* pl, int32_t a, uint64_t obs, uint64_t seed, int64_t n_out, int32_t mode) {
  if (pl->belief_kind != 2) return POMCP_ERR_INVALID_ARG;
  if (a < 0 || a >= pl->nA || n_out <= 0) return POMCP_ERR_INVALID_ARG;
  const model* m = &pl->m;
  int64_t n = pl->n_bel;
  void* tmp = malloc((size_t)m->state_bytes * (size_t)n);
  for (int64_t i = 0; i < n; ++i) store_state(m, tmp, i, pl->bel[i]);
  ostate* sp = (ostate*)malloc(sizeof(ostate) * (size_t)n);
  double* w = (double*)malloc(sizeof(double) * (size_t)n);
  uint8_t* alive = (uint8_t*)malloc((size_t)n);
  pf_propagate(m, tmp, n, a, obs, seed, sp, w, alive);
  free(tmp);
  int any_alive = 0;
  for (int64_t i = 0; i < n; ++i) any_alive |= alive[i];
  int32_t rc = POMCP_OK;
  int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_out);
  if (!any_alive) rc = POMCP_ERR_ALL_SAMPLES_TERMINAL; /* "all states in the particle collection were terminal" */
  else {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {0, TAG_RESAMPLE, 0, 0}, wd[4];
    orc_philox4x32_10(ctr, key, wd);
    if (mode == 0) rc = resample_fixed(w, alive, n, wd[0], n_out, idx, weight_ceil_exp(m));
    else {
      /* terminal particles are dropped before weighting: compact, resample, map back */
      int64_t na = 0;
      int64_t* map = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
      double* wc = (double*)malloc(sizeof(double) * (size_t)n);
      for (int64_t i = 0; i < n; ++i) if (alive[i]) { map[na] = i; wc[na] = w[i]; ++na; }
      rc = orc_resample_systematic_f64(wc, na, u01(wd[0]), n_out, idx);
      if (rc == POMCP_OK) for (int64_t k = 0; k < n_out; ++k) idx[k] = map[idx[k]];
      free(map); free(wc);
    }
  }
  if (rc == POMCP_OK) {
    ostate* nb = (ostate*)malloc(sizeof(ostate) * (size_t)n_out);
    for (int64_t k = 0; k < n_out; ++k) nb[k] = sp[idx[k]];
    fresh_root(pl); /* a ParticleCollection is not a BeliefNode: new RootNode, src/solver.jl:278-281 */
    free(pl->bel); pl->bel = nb; pl->n_bel = n_out; pl->belief_kind = 2;
  }
  free(idx); free(sp); free(w); free(alive);
  return rc;
}

/* update(::RootUpdater{R}, b_old, a, o) with a device-side reinvigorator R = SIRReinvigorator(n_min)
 * (the hooks of src/particle_filter.jl:37-73, src/updater.jl:13-27):
 *   handle_unseen_observation: child (a,o) missing or empty -> new belief = n_min particles from the SIR
 *                              posterior of the old root belief (propagate, weight, systematic resample);
 *   reinvigorate!:             child has fewer than n_min particles -> top it up from the same posterior.
 * outcome: 0 plain re-root, 1 topped up, 2 rebuilt (tree discarded). */
int32_t orc_update_reinvigorated(orc_planner* pl, int32_t a, uint64_t obs, uint64_t seed, int64_t n_min,
                                 int32_t* outcome) {
  if (a < 0 || a >= pl->nA || n_min <= 0) return POMCP_ERR_INVALID_ARG;
  if (pl->sp.belief_mode != POMCP_BELIEF_UNWEIGHTED) return POMCP_ERR_UNSUPPORTED;
  if (pl->belief_kind == 0) return POMCP_ERR_INVALID_ARG;
  const model* m = &pl->m;
  int64_t hao = (pl->node_expanded[pl->root] || pl->node_amask[pl->root]) ? find_child(pl, pl->root * pl->nA + a, obs) : -1;
  int64_t cnt = hao >= 0 ? pl->node_pcount[hao] : 0;
  int have = hao >= 0 && cnt > 0;
  if (outcome) *outcome = 0;
  if (have && cnt >= n_min) { pl->root = hao; pl->belief_kind = 3; return POMCP_OK; }
  int64_t extra = n_min - cnt;
  /* the old root belief as an array */
  int64_t n_old;
  ostate* old;
  if (pl->belief_kind == 2) {
    n_old = pl->n_bel; old = (ostate*)malloc(sizeof(ostate) * (size_t)n_old);
    memcpy(old, pl->bel, sizeof(ostate) * (size_t)n_old);
  } else if (pl->belief_kind == 3) {
    n_old = pl->node_pcount[pl->root]; old = (ostate*)malloc(sizeof(ostate) * (size_t)n_old);
    int64_t i = 0;
    for (int64_t e = pl->node_phead[pl->root]; e >= 0; e = pl->log_next[e]) old[i++] = pl->log_state[e];
  } else { /* initial distribution: n_min draws */
    n_old = n_min; old = (ostate*)malloc(sizeof(ostate) * (size_t)n_old);
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int64_t i = 0; i < n_old; ++i) {
      uint32_t ctr[4] = {0, TAG_INIT, (uint32_t)i, (uint32_t)((uint64_t)i >> 32)}, wd[4];
      orc_philox4x32_10(ctr, key, wd);
      old[i] = initial_state(&m->p, u01(wd[0]), u01(wd[1]));
    }
  }
  void* packed = malloc((size_t)m->state_bytes * (size_t)n_old);
  for (int64_t i = 0; i < n_old; ++i) store_state(m, packed, i, old[i]);
  free(old);
  ostate* sp = (ostate*)malloc(sizeof(ostate) * (size_t)n_old);
  double* w = (double*)malloc(sizeof(double) * (size_t)n_old);
  uint8_t* alive = (uint8_t*)malloc((size_t)n_old);
  pf_propagate(m, packed, n_old, a, obs, seed, sp, w, alive);
  free(packed);
  int any_alive = 0;
  for (int64_t i = 0; i < n_old; ++i) any_alive |= alive[i];
  int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)extra);
  int32_t rc = POMCP_ERR_PARTICLE_DEPLETION;
  if (any_alive) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {0, TAG_RESAMPLE, 0, 0}, wd[4];
    orc_philox4x32_10(ctr, key, wd);
    rc = resample_fixed(w, alive, n_old, wd[0], extra, idx, weight_ceil_exp(m));
  }
  if (rc == POMCP_OK) {
    if (have) {
      pl->root = hao; pl->belief_kind = 3;
      for (int64_t k = 0; k < extra; ++k) push_particle(pl, hao, sp[idx[k]]);
      if (outcome) *outcome = 1;
    } else {
      ostate* nb = (ostate*)malloc(sizeof(ostate) * (size_t)extra);
      for (int64_t k = 0; k < extra; ++k) nb[k] = sp[idx[k]];
      fresh_root(pl);
      free(pl->bel); pl->bel = nb; pl->n_bel = extra; pl->belief_kind = 2;
      if (outcome) *outcome = 2;
    }
  } else if (have) { /* no posterior mass to draw from: keep what the planner pushed */
    pl->root = hao; pl->belief_kind = 3;
    rc = POMCP_OK;
  } else {
    rc = POMCP_ERR_PARTICLE_DEPLETION;
  }
  free(idx); free(sp); free(w); free(alive);
  return rc;
}

int32_t orc_rollout_batch(const pomcp_problem* p, const pomcp_solver_params* sp, const void* states, int64_t n,
                          int32_t steps, uint64_t seed, double* returns_out, int32_t* steps_out) {
  (void)sp;
  model* m = (model*)malloc(sizeof(model));
  int rc = model_init(m, p);
  if (rc) { free(m); return rc; }
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int64_t i = 0; i < n; ++i) {
    int64_t ns = 0;
    returns_out[i] = rollout_random(m, load_state(m, states, i), steps, key, TAG_ROLLOUT, (uint32_t)i, 0, &ns);
    if (steps_out) steps_out[i] = (int32_t)ns;
  }
  free(m);
  return POMCP_OK;
}

/* ------------------------------------------------------------------ */
/* inspection                                                          */
/* ------------------------------------------------------------------ */
int32_t orc_get_root_stats(orc_planner* pl, int64_t* root_N, int64_t* act_N, double* act_V, int32_t nA) {
  if (root_N) *root_N = pl->node_N[pl->root];
  for (int a = 0; a < pl->nA && a < nA; ++a) {
    if (act_N) act_N[a] = pl->act_N[pl->root * pl->nA + a];
    if (act_V) act_V[a] = pl->act_V[pl->root * pl->nA + a];
  }
  return POMCP_OK;
}
int32_t orc_get_node_stats(orc_planner* pl, const int32_t* actions, const uint64_t* obs_bits, int32_t depth,
                           int64_t* node_N, int64_t* n_particles, int64_t* act_N, double* act_V, int32_t nA) {
  int64_t h = pl->root;
  for (int d = 0; d < depth; ++d) {
    if (actions[d] < 0 || actions[d] >= pl->nA) return POMCP_ERR_INVALID_ARG;
    h = find_child(pl, h * pl->nA + actions[d], obs_bits[d]);
    if (h < 0) { if (node_N) *node_N = -1; return POMCP_OK; }
  }
  if (node_N) *node_N = pl->node_N[h];
  if (n_particles) *n_particles = pl->node_pcount[h];
  for (int a = 0; a < pl->nA && a < nA; ++a) {
    if (act_N) act_N[a] = pl->act_N[h * pl->nA + a];
    if (act_V) act_V[a] = pl->act_V[h * pl->nA + a];
  }
  return POMCP_OK;
}
int32_t orc_dump_tree(orc_planner* pl, int64_t cap, int64_t* n_nodes, int64_t* node_N, int64_t* act_N, double* act_V,
                      int32_t* child) {
  *n_nodes = pl->n_nodes;
  int64_t n = pl->n_nodes < cap ? pl->n_nodes : cap;
  for (int64_t i = 0; i < n; ++i) {
    if (node_N) node_N[i] = pl->node_N[i];
    for (int a = 0; a < pl->nA; ++a) {
      if (act_N) act_N[i * pl->nA + a] = pl->act_N[i * pl->nA + a];
      if (act_V) act_V[i * pl->nA + a] = pl->act_V[i * pl->nA + a];
    }
    if (child && pl->nO > 0)
      for (int j = 0; j < pl->nA * pl->nO; ++j) child[i * pl->nA * pl->nO + j] = pl->child[i * pl->nA * pl->nO + j];
  }
  return POMCP_OK;
}
int32_t orc_get_counters(orc_planner* pl, pomcp_counters* out) { *out = pl->ctr; return POMCP_OK; }

/* ------------------------------------------------------------------ */
/* root-parallel CPU search (the multi-core baseline)                   */
/* ------------------------------------------------------------------ */
int32_t orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int32_t orc_search_rootparallel(const pomcp_problem* p, const pomcp_solver_params* sp, const void* belief_states,
                                int64_t n_belief, int32_t n_trees, int32_t n_threads, int64_t Q, int32_t* action_out,
                                int64_t* root_N_out, double* root_V_out, int32_t nA, pomcp_counters* csum) {
  if (n_trees <= 0 || n_trees > 4096) return POMCP_ERR_INVALID_ARG;
  model* m0 = (model*)malloc(sizeof(model));
  int rc0 = model_init(m0, p);
  int mA = m0->nA;
  free(m0);
  if (rc0) return rc0;
  double* sumN = (double*)calloc((size_t)mA, sizeof(double));
  double* sumNV = (double*)calloc((size_t)mA, sizeof(double));
  int64_t* cntN = (int64_t*)calloc((size_t)mA, sizeof(int64_t));
  int32_t* rcs = (int32_t*)calloc((size_t)n_trees, sizeof(int32_t));
  pomcp_counters* cs = (pomcp_counters*)calloc((size_t)n_trees, sizeof(pomcp_counters));
  int64_t* tN = (int64_t*)calloc((size_t)n_trees * mA, sizeof(int64_t));
  double* tV = (double*)calloc((size_t)n_trees * mA, sizeof(double));
  uint32_t* tE = (uint32_t*)calloc((size_t)n_trees, sizeof(uint32_t)); /* existing root ActNodes of tree t */
  uint32_t anyE = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int t = 0; t < n_trees; ++t) {
    pomcp_solver_params spt = *sp;
    spt.rank = t;
    orc_planner* pl = orc_create(p, &spt);
    if (!pl) { rcs[t] = POMCP_ERR_INVALID_ARG; continue; }
    if (belief_states) orc_set_belief_particles(pl, belief_states, n_belief); else orc_set_belief_initial(pl);
    int32_t act;
    rcs[t] = orc_search(pl, Q, &act, tN + (size_t)t * mA, tV + (size_t)t * mA, mA);
    cs[t] = pl->ctr;
    tE[t] = sp->dpw_action ? pl->node_amask[pl->root] : 0xffffffffu;
    orc_destroy(pl);
  }
  (void)n_threads;
  int32_t rc = POMCP_OK;
  if (csum) memset(csum, 0, sizeof(*csum));
  int n_ok = 0;
  for (int t = 0; t < n_trees; ++t) {
    if (rcs[t] == POMCP_ERR_ALL_SAMPLES_TERMINAL) continue;
    if (rcs[t] != POMCP_OK) { rc = rcs[t]; continue; }
    ++n_ok;
    anyE |= tE[t];
    for (int a = 0; a < mA; ++a) {
      if (!((tE[t] >> a) & 1u)) continue;
      cntN[a] += tN[(size_t)t * mA + a];
      sumN[a] += (double)tN[(size_t)t * mA + a];
      sumNV[a] += (double)tN[(size_t)t * mA + a] * tV[(size_t)t * mA + a];
    }
    if (csum) {
      csum->simulations += cs[t].simulations; csum->rollouts += cs[t].rollouts;
      csum->rollout_steps += cs[t].rollout_steps; csum->tree_steps += cs[t].tree_steps;
      csum->nodes += cs[t].nodes; csum->terminal_skips += cs[t].terminal_skips;
    }
  }
  if (rc == POMCP_OK && n_ok == 0) rc = POMCP_ERR_ALL_SAMPLES_TERMINAL;
  if (rc == POMCP_OK) {
    int best = -1; double bestV = 0.0;
    for (int a = 0; a < mA; ++a) {
      double V = sumN[a] > 0.0 ? sumNV[a] / sumN[a] : sp->init_V;
      if (a < nA) { if (root_N_out) root_N_out[a] = cntN[a]; if (root_V_out) root_V_out[a] = V; }
      if (!((anyE >> a) & 1u)) continue; /* no tree holds this ActNode (action widening) */
      if (best < 0 || V >= bestV) { bestV = V; best = a; }
    }
    if (action_out) *action_out = best < 0 ? 0 : best;
  }
  free(sumN); free(sumNV); free(cntN); free(rcs); free(cs); free(tN); free(tV); free(tE);
  return rc;
}
