/*
 * ORACLE (test infrastructure, NOT product code) -- plain-C twin of oracle/mbpo_oracle.py.
 *
 * CPU restatement of the reference's iCEM planning hot path, used (a) to cross-check the NumPy
 * oracle with an independently written implementation and (b) as the timed CPU baseline
 * (bench.py cpu_baseline / --impl reference: kind "port", all host cores via OpenMP).
 * The reference itself is pure Python/JAX and cannot be compiled or imported here (no jax in
 * the image), so there is no oracle/_ref; float parity against a real JAX run is UNPINNED.
 *
 * Reference lines restated (paths relative to /root/reference):
 *   mbpo/optimizers/trajectory_optimizers/icem_optimizer.py:134-252   iCemTO.optimize
 *   mbpo/utils/general_utils.py:81-208                                powerlaw_psd_gaussian
 *   mbpo/utils/optimizer_utils.py:11-59                               rollout_actions
 *   mbpo/systems/pendulum_system.py:18-39, dynamics/pendulum_dynamics.py:29-63,
 *   rewards/pendulum_reward.py:27-42                                  PendulumSystem.step
 *   mbpo/optimizers/policy_optimizers/brax_utils/training.py:91-137,
 *   sac/acting.py:35-55                                               wrapped env step
 * JAX PRNG (third party, absent): threefry2x32 / split / random_bits / uniform / normal /
 * XLA ErfInv32 as published (see oracle/jax_prng.py header).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may load
 * this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_H 128
#define ORC_MAX_F (ORC_MAX_H / 2 + 1)

/* ---------------------------------------------------------------- threefry2x32 ---------- */
static inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static void threefry(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* y0, uint32_t* y1) {
  static const int rot[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
  for (int g = 0; g < 5; ++g) {
    for (int i = 0; i < 4; ++i) {
      x0 += x1;
      x1 = rotl(x1, rot[g & 1][i]);
      x1 ^= x0;
    }
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
  *y0 = x0;
  *y1 = x1;
}

void orc_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* out2) {
  threefry(k0, k1, x0, x1, &out2[0], &out2[1]);
}

/* threefry_2x32(key, iota(n)) flattened: zero pad to even, halves, concat (legacy layout) */
static void legacy_stream(const uint32_t key[2], int n, uint32_t* out) {
  const int npad = n + (n & 1), h = npad / 2;
  for (int j = 0; j < h; ++j) {
    uint32_t c1 = (uint32_t)(h + j), y0, y1;
    if ((int)c1 >= n) c1 = 0;
    threefry(key[0], key[1], (uint32_t)j, c1, &y0, &y1);
    out[j] = y0;
    if (h + j < n) out[h + j] = y1;
  }
}

void orc_split(const uint32_t key[2], int num, int partitionable, uint32_t* out /*[num,2]*/) {
  if (partitionable) {
    for (int i = 0; i < num; ++i) threefry(key[0], key[1], 0u, (uint32_t)i, &out[2 * i], &out[2 * i + 1]);
  } else {
    legacy_stream(key, 2 * num, out);
  }
}

void orc_random_bits(const uint32_t key[2], int n, int partitionable, uint32_t* out) {
  if (partitionable) {
    for (int i = 0; i < n; ++i) {
      uint32_t y0, y1;
      threefry(key[0], key[1], 0u, (uint32_t)i, &y0, &y1);
      out[i] = y0 ^ y1;
    }
  } else {
    legacy_stream(key, n, out);
  }
}

/* ---------------------------------------------------------------- uniform / normal -------- */
static inline float u32_as_f32(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static float erf_inv32(float x) {
  static const float a[9] = {2.81022636e-08f, 3.43273939e-07f, -3.5233877e-06f, -4.39150654e-06f, 0.00021858087f,
                             -0.00125372503f, -0.00417768164f, 0.246640727f, 1.50140941f};
  static const float b[9] = {-0.000200214257f, 0.000100950558f, 0.00134934322f, -0.00367342844f, 0.00573950773f,
                             -0.0076224613f, 0.00943887047f, 1.00167406f, 2.83297682f};
  if (fabsf(x) == 1.0f) return x * INFINITY;
  float w = -log1pf(-(x * x));
  const float* c;
  if (w < 5.0f) {
    w = w - 2.5f;
    c = a;
  } else {
    w = sqrtf(w) - 3.0f;
    c = b;
  }
  float p = c[0];
  for (int i = 1; i < 9; ++i) {
    const float t = p * w; /* -ffp-contract=off: unfused, as the NumPy oracle */
    p = c[i] + t;
  }
  return p * x;
}

static inline float bits_to_normal(uint32_t bits) {
  const float lo = -0.99999994f; /* nextafter(-1, 0) */
  const float f = u32_as_f32((bits >> 9) | 0x3F800000u) - 1.0f;
  const float t = f * 2.0f;   /* hi - lo rounds to 2.0f */
  float u = t + lo;
  if (u < lo) u = lo;
  return 1.41421356f * erf_inv32(u);
}

void orc_normal(const uint32_t key[2], int n, int partitionable, float* out) {
  uint32_t* bits = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n + 1));
  orc_random_bits(key, n, partitionable, bits);
  for (int i = 0; i < n; ++i) out[i] = bits_to_normal(bits[i]);
  free(bits);
}

/* ---------------------------------------------------------------- colored noise ----------- */
typedef struct {
  int H, F;
  float s_scale[ORC_MAX_F];
  float sigma;
  float cr[ORC_MAX_F][ORC_MAX_H]; /* weight * cos(2 pi k t / H) / H */
  float ci[ORC_MAX_F][ORC_MAX_H]; /* -weight * sin(2 pi k t / H) / H */
} NoiseTables;

static void noise_tables(NoiseTables* nt, int H, float exponent) {
  const int F = H / 2 + 1;
  nt->H = H;
  nt->F = F;
  float f[ORC_MAX_F];
  const float fmin = (float)(1.0 / H);
  for (int i = 0; i < F; ++i) f[i] = (float)i / (float)H;
  int ix = 0;
  for (int i = 0; i < F; ++i) ix += f[i] < fmin;
  if (ix && ix < F)
    for (int i = 0; i < ix; ++i) f[i] = f[ix];
  float sumsq = 0.0f;
  for (int i = 0; i < F; ++i) {
    nt->s_scale[i] = powf(f[i], (float)(-(double)exponent / 2.0));
    if (i >= 1) {
      float w = nt->s_scale[i];
      if (i == F - 1) w = w * (float)((1 + (H % 2)) / 2.0);
      sumsq += w * w;
    }
  }
  nt->sigma = 2.0f * sqrtf(sumsq) / (float)H;
  for (int k = 0; k < F; ++k) {
    double wgt = (k == 0 || (H % 2 == 0 && k == F - 1)) ? 1.0 : 2.0;
    for (int t = 0; t < H; ++t) {
      const double ang = 2.0 * M_PI * (double)((k * t) % H) / (double)H;
      nt->cr[k][t] = (float)(wgt * cos(ang) / H);
      nt->ci[k][t] = (float)(-wgt * sin(ang) / H);
    }
    if (k == 0 || (H % 2 == 0 && k == F - 1))
      for (int t = 0; t < H; ++t) nt->ci[k][t] = 0.0f;
  }
}

/* powerlaw_psd_gaussian(exponent, H, rng) -> y[H]  (general_utils.py:189-207) */
static void powerlaw_row(const NoiseTables* nt, const uint32_t rng[2], int partitionable, float* y) {
  const int H = nt->H, F = nt->F;
  uint32_t sub[6], bits[ORC_MAX_F + 1];
  float sr[ORC_MAX_F], si[ORC_MAX_F];
  orc_split(rng, 3, partitionable, sub); /* key_sr, key_si, _ */
  orc_random_bits(&sub[0], F, partitionable, bits);
  for (int k = 0; k < F; ++k) sr[k] = bits_to_normal(bits[k]) * nt->s_scale[k];
  orc_random_bits(&sub[2], F, partitionable, bits);
  for (int k = 0; k < F; ++k) si[k] = bits_to_normal(bits[k]) * nt->s_scale[k];
  if (H % 2 == 0) {
    si[F - 1] = 0.0f;
    sr[F - 1] *= 1.41421356f;
  }
  si[0] = 0.0f;
  sr[0] *= 1.41421356f;
  for (int t = 0; t < H; ++t) y[t] = 0.0f;
  for (int k = 0; k < F; ++k) {
    const float a = sr[k], b = si[k];
    const float* cr = nt->cr[k];
    const float* ci = nt->ci[k];
    for (int t = 0; t < H; ++t) y[t] += a * cr[t] + b * ci[t];
  }
  for (int t = 0; t < H; ++t) y[t] /= nt->sigma;
}

void orc_powerlaw_noise(const uint32_t* keys, int M, int H, float exponent, int partitionable, float* out) {
  NoiseTables* nt = (NoiseTables*)malloc(sizeof(NoiseTables));
  noise_tables(nt, H, exponent);
  for (int i = 0; i < M; ++i) powerlaw_row(nt, &keys[2 * i], partitionable, out + (size_t)i * H);
  free(nt);
}

/* ---------------------------------------------------------------- pendulum ---------------- */
typedef struct {
  float max_speed, max_torque, dt, g, m, l, control_cost, angle_cost, target_angle;
} Pend;

static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* one PendulumSystem.step: x[3], u -> x_next[3], reward */
static inline float pend_step(const Pend* p, const float* x, float u, float* xn) {
  const float PI = 3.14159274f, TWO_PI = 6.28318548f;
  const float th = atan2f(x[1], x[0]);
  const float thdot = x[2];
  const float uu = clampf(u, -1.0f, 1.0f) * p->max_torque;
  const float c_g = (3.0f * p->g) / (2.0f * p->l);
  const float c_u = 3.0f / (p->m * (p->l * p->l));
  const float t1 = c_g * sinf(th);
  const float t2 = c_u * uu;
  const float thdd = t1 + t2;
  const float t3 = thdd * p->dt;
  const float nthd = clampf(thdot + t3, -p->max_speed, p->max_speed);
  const float t4 = nthd * p->dt;
  const float newth = th + t4;
  xn[0] = cosf(newth);
  xn[1] = sinf(newth);
  xn[2] = nthd;
  float d = th - p->target_angle;
  float r = fmodf(d + PI, TWO_PI);
  if (r < 0.0f) r += TWO_PI; /* floored modulo */
  d = r - PI;
  const float a1 = p->angle_cost * (d * d);
  const float a2 = 0.1f * (thdot * thdot);
  const float a3 = p->control_cost * (u * u);
  return -(a1 + a2) - a3;
}

void orc_pendulum_step(const float* params9, const float* x, const float* u, int R, float* xn, float* rew) {
  Pend p;
  memcpy(&p, params9, sizeof(p));
  for (int i = 0; i < R; ++i) rew[i] = pend_step(&p, x + 3 * i, u[i], xn + 3 * i);
}

static float rollout_return(const Pend* p, const float* x0, const float* acts, int H) {
  float x[3] = {x0[0], x0[1], x0[2]}, xn[3], acc = 0.0f;
  for (int t = 0; t < H; ++t) {
    acc += pend_step(p, x, acts[t], xn);
    x[0] = xn[0]; x[1] = xn[1]; x[2] = xn[2];
  }
  return acc / (float)H;
}

void orc_rollout_returns(const float* params9, const float* x0, const float* actions, int B, int M, int H,
                         float* ret) {
  Pend p;
  memcpy(&p, params9, sizeof(p));
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)B * M; ++i) ret[i] = rollout_return(&p, x0 + 3 * (i / M), actions + (size_t)i * H, H);
}

/* ---------------------------------------------------------------- iCEM -------------------- */
typedef struct {
  int horizon, num_samples, num_elites, num_prev_elites, num_particles, num_steps, warm_start, partitionable,
      summarize_max;
  float init_std, alpha, exponent, u_min, u_max;
} OrcIcemCfg;

static inline uint32_t total_order_key(float v) {
  uint32_t b;
  memcpy(&b, &v, 4);
  if (v == 0.0f) b = 0u;
  if (v != v) b = 0x7FC00000u;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

typedef struct {
  uint32_t key;
  int idx;
} SortItem;

static int cmp_item(const void* a, const void* b) {
  const SortItem* x = (const SortItem*)a;
  const SortItem* y = (const SortItem*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

/* iCemTO.optimize for one problem.  Work buffers are caller provided. */
static void icem_optimize_one(const OrcIcemCfg* c, const Pend* p, const NoiseTables* nt, const float* x0,
                              const uint32_t key_in[2], const float* best_seq_in, float* best_seq_out,
                              float* best_val_out, uint32_t key_out[2], float* acts /*[M,H]*/, float* vals /*[M]*/,
                              SortItem* items /*[M]*/, uint32_t* skeys /*[2(N+1)]*/) {
  const int H = c->horizon, N = c->num_samples, Np = c->num_prev_elites, M = N + Np, K = c->num_elites;
  float mean[ORC_MAX_H], std[ORC_MAX_H], best_seq[ORC_MAX_H], noise[ORC_MAX_H];
  for (int t = 0; t < H; ++t) {
    float m = 0.0f;
    if (c->warm_start) m = best_seq_in[t + 1 < H ? t + 1 : H - 1];
    mean[t] = m;
    std[t] = c->init_std;
    best_seq[t] = m;
  }
  float best_val = -INFINITY;
  uint32_t ks[4], carry[2];
  orc_split(key_in, 2, c->partitionable, ks); /* optimizer_key, key */
  carry[0] = ks[0]; carry[1] = ks[1];
  key_out[0] = ks[2]; key_out[1] = ks[3];
  const float one_m = (float)(1.0 - (double)c->alpha);
  float zero_ret = 0.0f;
  for (int it = 0; it < c->num_steps; ++it) {
    orc_split(carry, 2, c->partitionable, ks);           /* sampling_rng, particles_rng (dead) */
    orc_split(&ks[0], N + 1, c->partitionable, skeys);   /* :175 */
    carry[0] = skeys[0]; carry[1] = skeys[1];            /* key = sampling_rng[0] */
    for (int n = 0; n < N; ++n) {
      uint32_t dk[2];
      orc_split(&skeys[2 * (n + 1)], 1, c->partitionable, dk); /* vmap(split(x, action_dim)), A = 1 */
      powerlaw_row(nt, dk, c->partitionable, noise);
      float* row = acts + (size_t)n * H;
      for (int t = 0; t < H; ++t) {
        const float q = noise[t] * std[t];
        row[t] = clampf(mean[t] + q, c->u_min, c->u_max);
      }
      vals[n] = rollout_return(p, x0, row, H);
    }
    if (it == 0) {
      memset(acts + (size_t)N * H, 0, sizeof(float) * (size_t)Np * H);
      zero_ret = rollout_return(p, x0, acts + (size_t)N * H, H);
    }
    for (int n = N; n < M; ++n) vals[n] = zero_ret;
    if (c->num_particles > 1 && !c->summarize_max) { /* mean over P identical particles */
      for (int n = 0; n < M; ++n) {
        float acc = 0.0f;
        for (int q = 0; q < c->num_particles; ++q) acc += vals[n];
        vals[n] = acc / (float)c->num_particles;
      }
    }
    for (int n = 0; n < M; ++n) {
      items[n].key = total_order_key(vals[n]);
      items[n].idx = n;
    }
    qsort(items, (size_t)M, sizeof(SortItem), cmp_item); /* (key, idx) pairs are unique: order = stable argsort */
    const SortItem* el = items + (M - K);
    for (int t = 0; t < H; ++t) {
      float acc = 0.0f;
      for (int e = 0; e < K; ++e) acc += acts[(size_t)el[e].idx * H + t];
      const float em = acc / (float)K;
      acc = 0.0f;
      for (int e = 0; e < K; ++e) {
        const float d = acts[(size_t)el[e].idx * H + t] - em;
        const float dd = d * d;
        acc += dd;
      }
      const float ev = acc / (float)K;
      const float m1 = mean[t] * c->alpha, m2 = one_m * em;
      const float s2 = std[t] * std[t];
      const float v1 = s2 * c->alpha, v2 = one_m * ev;
      mean[t] = m1 + m2;
      std[t] = sqrtf(v1 + v2);
    }
    const float best_elite = vals[el[K - 1].idx];
    if (best_val <= best_elite) {
      best_val = best_elite;
      memcpy(best_seq, acts + (size_t)el[K - 1].idx * H, sizeof(float) * (size_t)H);
    }
  }
  memcpy(best_seq_out, best_seq, sizeof(float) * (size_t)H);
  *best_val_out = best_val;
}

typedef struct {
  float* acts;
  float* vals;
  SortItem* items;
  uint32_t* skeys;
} Work;

static void work_alloc(Work* w, const OrcIcemCfg* c) {
  const size_t M = (size_t)c->num_samples + c->num_prev_elites;
  w->acts = (float*)malloc(sizeof(float) * M * c->horizon);
  w->vals = (float*)malloc(sizeof(float) * M);
  w->items = (SortItem*)malloc(sizeof(SortItem) * M);
  w->skeys = (uint32_t*)malloc(sizeof(uint32_t) * 2 * ((size_t)c->num_samples + 2));
}

static void work_free(Work* w) {
  free(w->acts); free(w->vals); free(w->items); free(w->skeys);
}

/* vmap(iCemTO.optimize) over B problems, OpenMP over problems.  Returns the thread count used. */
int orc_icem_optimize_batch(const OrcIcemCfg* c, const float* params9, const float* x0, const uint32_t* key_in,
                            const float* best_seq_in, int B, float* best_seq_out, float* best_val_out,
                            uint32_t* key_out, int num_threads) {
  Pend p;
  memcpy(&p, params9, sizeof(p));
  NoiseTables* nt = (NoiseTables*)malloc(sizeof(NoiseTables));
  noise_tables(nt, c->horizon, c->exponent);
  int used = 1;
#ifdef _OPENMP
  if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
#pragma omp parallel
  {
#ifdef _OPENMP
#pragma omp single
    used = omp_get_num_threads();
#endif
    Work w;
    work_alloc(&w, c);
#pragma omp for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b)
      icem_optimize_one(c, &p, nt, x0 + 3 * b, key_in + 2 * b, best_seq_in + (size_t)b * c->horizon,
                        best_seq_out + (size_t)b * c->horizon, best_val_out + b, key_out + 2 * b, w.acts, w.vals,
                        w.items, w.skeys);
    work_free(&w);
  }
  free(nt);
  return used;
}

/* tests/test_icemopt.py:19-32 for one problem: T x { plan; true system.step; warm start } */
void orc_icem_closed_loop(const OrcIcemCfg* c, const float* params9, const float* x0, const uint32_t key_in[2],
                          const float* best_seq_in, int T, float* states /*[T,3]*/, float* rewards /*[T]*/,
                          float* actions /*[T]*/, float* best_seq_out, uint32_t key_out[2]) {
  Pend p;
  memcpy(&p, params9, sizeof(p));
  NoiseTables* nt = (NoiseTables*)malloc(sizeof(NoiseTables));
  noise_tables(nt, c->horizon, c->exponent);
  Work w;
  work_alloc(&w, c);
  float x[3] = {x0[0], x0[1], x0[2]}, xn[3], seq[ORC_MAX_H], nseq[ORC_MAX_H], bv;
  uint32_t key[2] = {key_in[0], key_in[1]}, nkey[2];
  memcpy(seq, best_seq_in, sizeof(float) * (size_t)c->horizon);
  for (int t = 0; t < T; ++t) {
    icem_optimize_one(c, &p, nt, x, key, seq, nseq, &bv, nkey, w.acts, w.vals, w.items, w.skeys);
    memcpy(seq, nseq, sizeof(float) * (size_t)c->horizon);
    key[0] = nkey[0]; key[1] = nkey[1];
    rewards[t] = pend_step(&p, x, seq[0], xn);
    actions[t] = seq[0];
    x[0] = xn[0]; x[1] = xn[1]; x[2] = xn[2];
    states[3 * t] = x[0]; states[3 * t + 1] = x[1]; states[3 * t + 2] = x[2];
  }
  memcpy(best_seq_out, seq, sizeof(float) * (size_t)c->horizon);
  key_out[0] = key[0]; key_out[1] = key[1];
  work_free(&w);
  free(nt);
}

/* ---------------------------------------------------------------- wrapped env rollouts ------ */
/* obs/steps/done in-out [E,...]; actions [T,E]; outputs time-major (any may be NULL). */
int orc_env_rollout(const float* params9, int episode_length, int action_repeat, float* obs, float* steps,
                    float* done, const float* first_obs, const float* actions, int E, int T, float* observation_out,
                    float* reward_out, float* discount_out, float* next_observation_out, float* truncation_out,
                    int num_threads) {
  Pend p;
  memcpy(&p, params9, sizeof(p));
  int used = 1;
#ifdef _OPENMP
  if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
#pragma omp parallel
  {
#ifdef _OPENMP
#pragma omp single
    used = omp_get_num_threads();
#endif
#pragma omp for schedule(static)
    for (int e = 0; e < E; ++e) {
      float x[3] = {obs[3 * e], obs[3 * e + 1], obs[3 * e + 2]}, xn[3];
      float st = steps[e], dn = done[e];
      for (int t = 0; t < T; ++t) {
        const size_t o = (size_t)t * E + e;
        st = (dn != 0.0f) ? 0.0f : st;
        dn = 0.0f;
        if (observation_out) memcpy(observation_out + 3 * o, x, 12);
        float rew = 0.0f;
        for (int r = 0; r < action_repeat; ++r) {
          rew += pend_step(&p, x, actions[o], xn);
          x[0] = xn[0]; x[1] = xn[1]; x[2] = xn[2];
        }
        st += (float)action_repeat;
        const int over = st >= (float)episode_length;
        const float trunc = over ? (1.0f - dn) : 0.0f;
        dn = over ? 1.0f : dn;
        if (dn != 0.0f) memcpy(x, first_obs + 3 * e, 12);
        if (next_observation_out) memcpy(next_observation_out + 3 * o, x, 12);
        if (reward_out) reward_out[o] = rew;
        if (discount_out) discount_out[o] = 1.0f - dn;
        if (truncation_out) truncation_out[o] = trunc;
      }
      memcpy(obs + 3 * e, x, 12);
      steps[e] = st;
      done[e] = dn;
    }
  }
  return used;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
