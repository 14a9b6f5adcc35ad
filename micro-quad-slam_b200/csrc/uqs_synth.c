/*
 * uqs_synth.c -- seeded synthetic flight logs (host C, pthreads).  libuqs_synth.so
 *
 * The reference has no log generator, simulator or recorded flight
 * (SURVEY.md section 4); the post-flight replay is fed by the fields the
 * reference logs per ToF frame -- pose, yaw, flow rates, flow quality and the
 * ToF ranges (scanrec_t, uav_local_nav.c:1522-1547).  This file produces those
 * fields, as SoA arrays, for the BASELINE.json configurations.
 *
 *  world      : lattice of axis-aligned rectangular rooms (one room for C1-C3/C5,
 *               a 10 m lattice for the C4 building sweep); the true range of a
 *               beam is the exact exit distance from the room that contains the
 *               pose, computed in binary64.
 *  trajectory : closed ellipse-with-harmonic curve (kind 0) or lawn-mower sweep
 *               (kind 1), always inside 60 % of the grid half-extent so that the
 *               reference would never recenter (uav_local_nav.c:328-332);
 *               yaw sweeps at yaw_rate_dps, wrapped to [-180, 180).
 *  sensors    : range = true + sigma_r * g, clipped to [0.02, max_range], with
 *               p_dropout NaNs; flow rate = body velocity / h + sigma_f * g + bias
 *               (bias ~ sigma_b * g per flight); of_q = 200, p_lowq of samples 30.
 *               g is an Irwin-Hall(4) approximate Gaussian built from integer
 *               PRNG output, so a seed reproduces the same bytes on any libm.
 *  PRNG       : xorshift64*, seed = 0x5EED0000 + config_id*65536 + flight_id.
 *
 * Everything here is input fabrication; no mapping arithmetic lives in this file.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct uqs_synth_cfg {
  int32_t  config_id;       /* goes into the seed                                   */
  int32_t  n_samples;       /* flow/pose samples per flight                         */
  int32_t  frames_per_sample; /* 1: 32 beams per sample; 2: 64 beams (second frame at yaw+45) */
  int32_t  traj_kind;       /* 0 closed curve, 1 lawn-mower                         */
  float    rate_hz;         /* sample rate                                          */
  float    room_w, room_h;  /* room size (m); lattice period                        */
  float    room_x0, room_y0;/* lower-left corner of the room containing the origin  */
  float    traj_ax, traj_ay;/* kind 0: semi-axes (m); kind 1: half-extent of sweep   */
  float    traj_period_s;   /* kind 0: loop period; kind 1: unused                   */
  float    speed_mps;       /* kind 1: sweep speed                                   */
  float    line_spacing_m;  /* kind 1: distance between sweep lines                  */
  float    yaw_rate_dps;    /* 20 deg/s                                              */
  float    max_range_m;     /* 4.0                                                   */
  float    sigma_r, sigma_f, sigma_b;
  float    p_dropout, p_lowq;
  float    h_m;             /* flight height, 0.5 m                                  */
  int32_t  shared_truth;    /* 1: all flights fly the same true trajectory (drift ensemble) */
  int32_t  range_mm;        /* 1: ranges lie on the sensor's 1 mm lattice: (float)mm * 0.001f, the value the
                               reference derives from the raw u16 reading (uav_local_nav.c:1328) */
} uqs_synth_cfg;

typedef struct { uint64_t s; } rng_t;

static inline uint64_t rng_next(rng_t* r) {
  uint64_t x = r->s;
  x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
  r->s = x;
  return x * 0x2545F4914F6CDD1DULL;
}
static inline float rng_uniform(rng_t* r) { return (float)(rng_next(r) >> 40) * (1.0f / 16777216.0f); }
/* Irwin-Hall(4): sum of four U(0,1), centred and scaled to unit variance */
static inline float rng_gauss(rng_t* r) {
  uint64_t a = rng_next(r), b = rng_next(r);
  float s = (float)(uint32_t)(a >> 40) + (float)(uint32_t)((a >> 16) & 0xFFFFFF) +
            (float)(uint32_t)(b >> 40) + (float)(uint32_t)((b >> 16) & 0xFFFFFF);
  return (s * (1.0f / 16777216.0f) - 2.0f) * 1.7320508f;
}
static void rng_seed(rng_t* r, uint64_t seed) {
  r->s = seed * 0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL;
  if (!r->s) r->s = 0x5EED;
  for (int i = 0; i < 8; i++) rng_next(r);
}

/* true pose and world velocity at time t (s) */
static void truth_at(const uqs_synth_cfg* c, double t, double* x, double* y, double* vx, double* vy,
                     double* yaw_deg) {
  if (c->traj_kind == 0) {
    const double w = 2.0 * M_PI / c->traj_period_s;
    /* ellipse plus a third harmonic: smooth, closed, |x|<=ax, |y|<=ay */
    *x = c->traj_ax * (0.85 * cos(w * t) + 0.15 * cos(3.0 * w * t));
    *y = c->traj_ay * (0.85 * sin(w * t) - 0.15 * sin(3.0 * w * t));
    *vx = c->traj_ax * w * (-0.85 * sin(w * t) - 0.45 * sin(3.0 * w * t));
    *vy = c->traj_ay * w * (0.85 * cos(w * t) - 0.45 * cos(3.0 * w * t));
  } else {
    /* boustrophedon: lines along x at y = -ay + j*spacing, alternating direction */
    const double line_len = 2.0 * c->traj_ax;
    const double seg = line_len + c->line_spacing_m;          /* one line + one step up   */
    const int n_lines = (int)floor(2.0 * c->traj_ay / c->line_spacing_m) + 1;
    double s = c->speed_mps * t;
    const double total = seg * n_lines;
    int lap = (int)floor(s / total);
    s -= total * lap;
    int j = (int)floor(s / seg);
    double u = s - seg * j;
    int up = (lap & 1);                                        /* odd laps sweep back down */
    int jj = up ? (n_lines - 1 - j) : j;
    double ydir = up ? -1.0 : 1.0;
    double y0 = -c->traj_ay + jj * c->line_spacing_m;
    int fwd = ((j & 1) == 0);
    if (u <= line_len) {
      *x = fwd ? (-c->traj_ax + u) : (c->traj_ax - u);
      *y = y0;
      *vx = fwd ? c->speed_mps : -c->speed_mps;
      *vy = 0.0;
    } else {
      *x = fwd ? c->traj_ax : -c->traj_ax;
      *y = y0 + ydir * (u - line_len);
      *vx = 0.0;
      *vy = ydir * c->speed_mps;
    }
    if (*y > c->traj_ay) *y = c->traj_ay;
    if (*y < -c->traj_ay) *y = -c->traj_ay;
  }
  double yaw = fmod(c->yaw_rate_dps * t + 180.0, 360.0);
  if (yaw < 0) yaw += 360.0;
  *yaw_deg = yaw - 180.0;
}

/* exit distance from the lattice room containing (px,py) along angle th (rad) */
static double room_range(const uqs_synth_cfg* c, double px, double py, double th) {
  double ix = floor((px - c->room_x0) / c->room_w), iy = floor((py - c->room_y0) / c->room_h);
  double xl = c->room_x0 + ix * c->room_w, xh = xl + c->room_w;
  double yl = c->room_y0 + iy * c->room_h, yh = yl + c->room_h;
  double dx = cos(th), dy = sin(th), best = 1e30;
  if (dx > 1e-12) { double t = (xh - px) / dx; if (t < best) best = t; }
  if (dx < -1e-12) { double t = (xl - px) / dx; if (t < best) best = t; }
  if (dy > 1e-12) { double t = (yh - py) / dy; if (t < best) best = t; }
  if (dy < -1e-12) { double t = (yl - py) / dy; if (t < best) best = t; }
  return best;
}

typedef struct {
  const uqs_synth_cfg* cfg;
  int first_flight, n_flights, tid, n_threads;
  int flight_id0;
  const float* true_ranges;  /* shared truth: [n_frames][32], else NULL */
  uint32_t* t_ms; float *rx, *ry, *h, *yaw; uint8_t* q; float* ranges; float *xt, *yt;
} synth_job;

static const double CENTRE_DEG[4] = { 0.0, 90.0, 180.0, -90.0 };

static void true_ranges_for_sample(const uqs_synth_cfg* c, double px, double py, double yaw_deg,
                                   float* out /* frames_per_sample*32 */) {
  for (int f = 0; f < c->frames_per_sample; f++) {
    double yawf = yaw_deg + 45.0 * f;
    for (int b = 0; b < 32; b++) {
      int d = b >> 3, col = b & 7;
      double ang = (yawf + CENTRE_DEG[d] + ((col - 3.5) / 3.5) * 31.5) * (M_PI / 180.0);
      out[f * 32 + b] = (float)room_range(c, px, py, ang);
    }
  }
}

static void gen_flight(const synth_job* J, int fl) {
  const uqs_synth_cfg* c = J->cfg;
  const int n = c->n_samples, fps = c->frames_per_sample;
  const size_t so = (size_t)fl * n, fo = (size_t)fl * n * fps;
  rng_t r;
  rng_seed(&r, 0x5EED0000ULL + (uint64_t)c->config_id * 65536ULL + (uint64_t)(J->flight_id0 + fl));
  const float bias_x = c->sigma_b * rng_gauss(&r), bias_y = c->sigma_b * rng_gauss(&r);
  /* per-flight phase so that non-shared flights differ */
  const double t0 = c->shared_truth ? 0.0 : (double)rng_uniform(&r) * c->traj_period_s;
  const double dt = 1.0 / c->rate_hz;
  float tr[64];
  for (int i = 0; i < n; i++) {
    double t = t0 + i * dt, px, py, vx, vy, yaw;
    truth_at(c, t, &px, &py, &vx, &vy, &yaw);
    J->t_ms[so + i] = (uint32_t)llround((i * dt) * 1000.0);
    J->xt[so + i] = (float)px;
    J->yt[so + i] = (float)py;
    J->yaw[so + i] = (float)yaw;
    J->h[so + i] = c->h_m;
    double a = yaw * (M_PI / 180.0), ca = cos(a), sa = sin(a);
    double vbx = vx * ca + vy * sa, vby = -vx * sa + vy * ca;     /* world -> body */
    J->rx[so + i] = (float)(vbx / c->h_m) + c->sigma_f * rng_gauss(&r) + bias_x;
    J->ry[so + i] = (float)(vby / c->h_m) + c->sigma_f * rng_gauss(&r) + bias_y;
    J->q[so + i] = (rng_uniform(&r) < c->p_lowq) ? 30 : 200;
    const float* truth;
    if (J->true_ranges) truth = J->true_ranges + (size_t)i * fps * 32;
    else { true_ranges_for_sample(c, px, py, yaw, tr); truth = tr; }
    float* out = J->ranges + (fo + (size_t)i * fps) * 32;
    for (int k = 0; k < fps * 32; k++) {
      float v = truth[k] + c->sigma_r * rng_gauss(&r);
      if (v < 0.02f) v = 0.02f;
      if (v > c->max_range_m) v = c->max_range_m;
      if (c->range_mm) {
        long mm = lrintf(v * 1000.0f);
        if (mm < 0) mm = 0;
        if (mm > 65534) mm = 65534;
        v = (float)mm * 0.001f;
      }
      if (rng_uniform(&r) < c->p_dropout) v = NAN;
      out[k] = v;
    }
  }
}

static void* synth_worker(void* arg) {
  const synth_job* J = (const synth_job*)arg;
  for (int fl = J->tid; fl < J->n_flights; fl += J->n_threads) gen_flight(J, fl);
  return NULL;
}

/*
 * Fill SoA logs for flights [flight_id0, flight_id0 + n_flights).
 *   t_ms, of_rate_x, of_rate_y, h_m, yaw_deg, x_true, y_true : [n_flights][n_samples]
 *   of_q : [n_flights][n_samples] u8
 *   ranges : [n_flights][n_samples*frames_per_sample][32]
 * yaw_deg holds the per-SAMPLE yaw; with frames_per_sample == 2 the caller
 * expands poses to frames (second frame at yaw + 45.0f).
 */
int uqs_synth_generate(const uqs_synth_cfg* cfg, int flight_id0, int n_flights, int n_threads,
                       uint32_t* t_ms, float* of_rate_x, float* of_rate_y, float* h_m,
                       float* yaw_deg, uint8_t* of_q, float* ranges, float* x_true, float* y_true) {
  if (!cfg || n_flights <= 0 || cfg->n_samples <= 0) return 3;
  if (cfg->frames_per_sample != 1 && cfg->frames_per_sample != 2) return 3;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  if (n_threads > n_flights) n_threads = n_flights;
  float* shared = NULL;
  if (cfg->shared_truth) {
    const size_t per = (size_t)cfg->frames_per_sample * 32;
    shared = (float*)malloc(sizeof(float) * per * cfg->n_samples);
    if (!shared) return 6;
    const double dt = 1.0 / cfg->rate_hz;
    for (int i = 0; i < cfg->n_samples; i++) {
      double px, py, vx, vy, yaw;
      truth_at(cfg, i * dt, &px, &py, &vx, &vy, &yaw);
      true_ranges_for_sample(cfg, px, py, yaw, shared + per * i);
    }
  }
  pthread_t th[256];
  synth_job jobs[256];
  for (int t = 0; t < n_threads; t++) {
    jobs[t] = (synth_job){ cfg, 0, n_flights, t, n_threads, flight_id0, shared,
                           t_ms, of_rate_x, of_rate_y, h_m, yaw_deg, of_q, ranges, x_true, y_true };
    pthread_create(&th[t], NULL, synth_worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  free(shared);
  return 0;
}
