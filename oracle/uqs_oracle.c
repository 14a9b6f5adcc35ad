/*
 * uqs_oracle.c -- TEST INFRASTRUCTURE.  A plain-C, single-threaded CPU restatement
 * of the reference's mapping path with run-time geometry, used ONLY as the checker
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product
 * (libuqs_mapping.so) never links, loads or calls anything in this directory.
 *
 * PARITY PIN: the reference ships no tests or golden vectors for this path
 * (SURVEY.md section 4).  This restatement is pinned against the reference's OWN
 * code, extracted and compiled by oracle/build_ref.sh into oracle/_ref/ -- see
 * tests/test_oracle_vs_reference.py (byte-equal grids on every geometry, the
 * SURVEY Appendix-D known answer, hypothesis-generated poses/ranges).  The pose
 * stage P0 has NO counterpart in the reference (poses come from the flight
 * controller's EKF, uav_local_nav.c:1168-1195): for P0 this file is the
 * specification and its parity is "unpinned" (DESIGN.md section P0).
 *
 * Third-party arithmetic on the path: glibc 2.39 libm sincosf / lrintf, called
 * exactly where the reference calls them (uav_local_nav.c:209-210, :300-301).
 * orc_sincosf_restated() re-states glibc's published algorithm
 * (sysdeps/ieee754/flt-32/s_sincosf.c, FMA build) so that the device code can be
 * checked against the same operation sequence; it is itself checked against
 * libm exhaustively (tests/test_sincosf.py, oracle/sweep_sincosf).
 *
 * Build: gcc -O2 -ffp-contract=off (no -march=native, no -ffast-math).
 */
#define _GNU_SOURCE   /* sincosf */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/uqs_mapping.h"

/* ------------------------------------------------------------------ */
/* grid model: uav_local_nav.c:194-216                                 */
/* ------------------------------------------------------------------ */

/* clamp_lo(), uav_local_nav.c:199-203 */
static inline int8_t sat_lo(const uqs_params* p, int v) {
  if (v < p->lo_min) return (int8_t)p->lo_min;
  if (v > p->lo_max) return (int8_t)p->lo_max;
  return (int8_t)v;
}

/* world_to_grid(), uav_local_nav.c:205-214 (map_inited is the caller's business) */
int orc_world_to_grid(const uqs_params* p, float wx, float wy, int* gx, int* gy) {
  float ddx = wx - p->origin_x;
  float ddy = wy - p->origin_y;
  int cx = (int)lrintf(ddx / p->res_m) + (p->W / 2);
  int cy = (int)lrintf(ddy / p->res_m) + (p->H / 2);
  if (cx < 0 || cx >= p->W || cy < 0 || cy >= p->H) return 0;
  *gx = cx;
  *gy = cy;
  return 1;
}

/* raycast_update(), uav_local_nav.c:241-278.  Returns the number of cells visited
 * (iterations of the while(1) at :254-277), 0 if the ray is dropped (:243-244). */
long orc_raycast(const uqs_params* p, int8_t* grid, float ax, float ay, float bx, float by,
                 int hit) {
  int cx, cy, ex, ey;
  if (!orc_world_to_grid(p, ax, ay, &cx, &cy)) return 0;
  if (!orc_world_to_grid(p, bx, by, &ex, &ey)) return 0;

  const int adx = abs(ex - cx), ady = abs(ey - cy);
  const int stepx = (cx < ex) ? 1 : -1;
  const int stepy = (cy < ey) ? 1 : -1;
  int e = adx - ady;              /* err = dx + dy with dy = -|.| (:246-250) */
  long visited = 0;

  for (;;) {
    int8_t* cell = &grid[(size_t)cy * (size_t)p->W + (size_t)cx];   /* idx(), :216 */
    visited++;
    if (cx == ex && cy == ey) {
      /* end cell: +occ on a hit, -(free/2) (integer division, = 0 for free=1) otherwise (:262-268) */
      int delta = hit ? p->lo_occ : -(p->lo_free / 2);
      *cell = sat_lo(p, (int)*cell + delta);
      break;
    }
    *cell = sat_lo(p, (int)*cell - p->lo_free);                      /* :258-260 */
    const int twice = 2 * e;                                        /* :272-274 */
    if (twice >= -ady) { e -= ady; cx += stepx; }
    if (twice <= adx)  { e += adx; cy += stepy; }
    if (cx < 0 || cx >= p->W || cy < 0 || cy >= p->H) break;          /* :276, unreachable */
  }
  return visited;
}

/* map_update_from_beams(), uav_local_nav.c:280-306; beams = tof_beams_m[4][8]. */
long orc_frame(const uqs_params* p, int8_t* grid, float px, float py, float yaw_deg,
               const float* beams) {
  static const float centre_deg[4] = { 0.0f, 90.0f, 180.0f, -90.0f };   /* :283 */
  const float half_fov = p->fov_deg * 0.5f;                             /* :284 */
  const float hit_below = p->max_range_m - p->hit_margin_m;             /* :292 */
  long visited = 0;
  for (int b = 0; b < UQS_BEAMS_PER_FRAME; b++) {
    const int d = b >> 3, c = b & 7;
    float r = beams[b];
    if (isnan(r)) continue;                                             /* :289 */
    if (r <= p->min_range_m) continue;                                  /* :290 */
    const int hit = (r < hit_below);
    if (r > p->max_range_m) r = p->max_range_m;                         /* :293 */
    float u = ((float)c - 3.5f) / 3.5f;                                 /* :295 */
    float off = u * half_fov;                                           /* :296 */
    float a_deg = yaw_deg + centre_deg[d] + off;                        /* :298 */
    float a = a_deg * ((float)M_PI / 180.0f);                           /* :299 */
    float sn, cs;
    sincosf(a, &sn, &cs);                                               /* :300-301 */
    float qx = px + r * cs;
    float qy = py + r * sn;
    visited += orc_raycast(p, grid, px, py, qx, qy, hit);               /* :303 */
  }
  return visited;
}

/* the replay loop = body of log_tick(), uav_local_nav.c:1633-1635, gate forced true */
long orc_replay(const uqs_params* p, int8_t* grid, long n_frames, const float* x,
                const float* y, const float* yaw_deg, const float* ranges) {
  long visited = 0;
  for (long i = 0; i < n_frames; i++)
    visited += orc_frame(p, grid, x[i], y[i], yaw_deg[i], ranges + i * UQS_BEAMS_PER_FRAME);
  return visited;
}

/* end cell of every beam of one frame (for index-by-index parity of A3/A6). */
void orc_beam_cells(const uqs_params* p, float px, float py, float yaw_deg, const float* beams,
                    int32_t* cells /* [32][2] */, int32_t* origin /* [2] */) {
  static const float centre_deg[4] = { 0.0f, 90.0f, 180.0f, -90.0f };
  const float half_fov = p->fov_deg * 0.5f;
  int ox = -1, oy = -1;
  int have_o = orc_world_to_grid(p, px, py, &ox, &oy);
  origin[0] = have_o ? ox : -1;
  origin[1] = have_o ? oy : -1;
  for (int b = 0; b < UQS_BEAMS_PER_FRAME; b++) {
    const int d = b >> 3, c = b & 7;
    cells[2 * b] = cells[2 * b + 1] = -1;
    float r = beams[b];
    if (isnan(r) || r <= p->min_range_m || !have_o) continue;
    if (r > p->max_range_m) r = p->max_range_m;
    float u = ((float)c - 3.5f) / 3.5f;
    float a = (yaw_deg + centre_deg[d] + u * half_fov) * ((float)M_PI / 180.0f);
    float sn, cs;
    sincosf(a, &sn, &cs);
    int ex, ey;
    if (orc_world_to_grid(p, px + r * cs, py + r * sn, &ex, &ey)) {
      cells[2 * b] = ex;
      cells[2 * b + 1] = ey;
    }
  }
}

/* ------------------------------------------------------------------ */
/* rows either side of the path               */
/* ------------------------------------------------------------------ */

/* N1: robust_col_dist_m() + compute_beams_and_minima(), uav_local_nav.c:1320-1359.
 * raw = 512 bytes: 4 sensors x 64 cells x u16 LE mm (the frame payload at &frame[5], :1345). */
void orc_beams_from_scan(const uint8_t* raw, float max_range, float* beams /* [32] */, float* dir_min /* [4] or NULL */) {
  for (int d = 0; d < 4; d++) {
    const uint8_t* sensor = raw + d * 128;
    float lowest_dir = NAN;
    for (int c = 0; c < 8; c++) {
      float first = NAN, runner_up = NAN;
      for (int row = 0; row < 8; row++) {
        const uint8_t* q = sensor + (row * 8 + c) * 2;
        const unsigned mm = (unsigned)q[0] | ((unsigned)q[1] << 8);
        if (mm == 0xFFFFu || mm == 0u) continue;
        float metres = (float)mm * 0.001f;
        if (metres <= 0.02f) continue;
        if (metres > max_range) metres = max_range;
        if (isnan(first) || metres < first) { runner_up = first; first = metres; }
        else if (isnan(runner_up) || metres < runner_up) runner_up = metres;
      }
      const float chosen = !isnan(runner_up) ? runner_up : first;
      beams[d * 8 + c] = chosen;
      if (!isnan(chosen) && (isnan(lowest_dir) || chosen < lowest_dir)) lowest_dir = chosen;
    }
    if (dir_min) dir_min[d] = lowest_dir;
  }
}

/* N2: map_recenter_shift(), uav_local_nav.c:308-322 */
void orc_recenter_shift(const uqs_params* p, int8_t* grid, int sx, int sy) {
  const size_t cells = (size_t)p->W * p->H;
  int8_t* tmp = (int8_t*)calloc(cells, 1);
  for (int yy = 0; yy < p->H; yy++) {
    const int v = yy + sy;
    if (v < 0 || v >= p->H) continue;
    for (int xx = 0; xx < p->W; xx++) {
      const int u = xx + sx;
      if (u < 0 || u >= p->W) continue;
      tmp[(size_t)yy * p->W + xx] = grid[(size_t)v * p->W + u];
    }
  }
  memcpy(grid, tmp, cells);
  free(tmp);
}

/* N2: map_recentre_if_needed(), uav_local_nav.c:324-353; moves p->origin and shifts the grid; returns 1 if it did */
int orc_recentre_if_needed(uqs_params* p, int8_t* grid, float px, float py, int* sx_out, int* sy_out) {
  const float half = p->size_m * 0.5f;
  const float thresh = half * 0.60f;
  const float ddx = px - p->origin_x, ddy = py - p->origin_y;
  if (fabsf(ddx) < thresh && fabsf(ddy) < thresh) return 0;
  int sx = (int)lrintf(ddx / p->res_m), sy = (int)lrintf(ddy / p->res_m);
  const int cap = (int)(half / p->res_m * 0.5f);
  if (sx > cap) sx = cap;
  if (sx < -cap) sx = -cap;
  if (sy > cap) sy = cap;
  if (sy < -cap) sy = -cap;
  if (sx == 0 && sy == 0) return 0;
  orc_recenter_shift(p, grid, sx, sy);
  p->origin_x += (float)sx * p->res_m;
  p->origin_y += (float)sy * p->res_m;
  if (sx_out) *sx_out = sx;
  if (sy_out) *sy_out = sy;
  return 1;
}

/* log_tick() with recentering, uav_local_nav.c:1629-1635; returns the number of recenter events */
int orc_replay_recentering(uqs_params* p, int8_t* grid, long n_frames, const float* x, const float* y,
                           const float* yaw_deg, const float* ranges, long* updates) {
  int events = 0;
  long visited = 0;
  for (long i = 0; i < n_frames; i++) {
    events += orc_recentre_if_needed(p, grid, x[i], y[i], NULL, NULL);
    visited += orc_frame(p, grid, x[i], y[i], yaw_deg[i], ranges + i * UQS_BEAMS_PER_FRAME);
  }
  if (updates) *updates = visited;
  return events;
}

/* N3: frontier_score_dir(), uav_local_nav.c:356-385 */
int orc_frontier_score_dir(const uqs_params* p, const int8_t* grid, float px, float py, float yaw_deg, float offset_deg) {
  static const float fan_deg[3] = { 0.0f, 15.0f, -15.0f };
  const float reach = 2.5f;
  const float stride = p->res_m * 2.0f;
  int n_unknown = 0, n_free = 0, n_occ = 0;
  for (int r = 0; r < 3; r++) {
    const float a = (yaw_deg + offset_deg + fan_deg[r]) * ((float)M_PI / 180.0f);
    float sn, cs;
    sincosf(a, &sn, &cs);
    for (float t = stride; t <= reach; t += stride) {
      int gx, gy;
      if (!orc_world_to_grid(p, px + t * cs, py + t * sn, &gx, &gy)) break;
      const int8_t v = grid[(size_t)gy * p->W + gx];
      if (v >= -1 && v <= 1) n_unknown++;
      else if (v > 10) n_occ++;
      else if (v < -10) n_free++;
    }
  }
  return n_unknown * 3 + n_free - n_occ * 4;
}

/* ------------------------------------------------------------------ */
/* P0 -- builder-defined dead reckoning  */
/* ------------------------------------------------------------------ */
void orc_pose_integrate(long n, const uint32_t* t_ms, const float* rate_x, const float* rate_y,
                        const float* h_m, const float* yaw_deg, const uint8_t* q,
                        float* xo, float* yo) {
  if (n <= 0) return;
  float px = 0.0f, py = 0.0f;
  xo[0] = 0.0f;
  yo[0] = 0.0f;
  for (long i = 1; i < n; i++) {
    float dt = (float)(t_ms[i] - t_ms[i - 1]) * 0.001f;
    float inc_n = 0.0f, inc_e = 0.0f;
    /* quality gate mirrors uav_local_nav.c:943; NaN inputs contribute nothing */
    if (q[i] >= 50 && !isnan(rate_x[i]) && !isnan(rate_y[i]) && !isnan(h_m[i]) &&
        !isnan(yaw_deg[i])) {
      float vbx = rate_x[i] * h_m[i];            /* angular rate x height, :1160-1161 */
      float vby = rate_y[i] * h_m[i];
      float a = yaw_deg[i] * ((float)M_PI / 180.0f);
      float sn, cs;
      sincosf(a, &sn, &cs);
      float vn = vbx * cs - vby * sn;
      float ve = vbx * sn + vby * cs;
      inc_n = vn * dt;
      inc_e = ve * dt;
    }
    px = px + inc_n;
    py = py + inc_e;
    xo[i] = px;
    yo[i] = py;
  }
}

/* The same binary32 increments summed in binary64: the yardstick for the look-back scan variant (mode 1),
 * which cannot be bit-identical to the binary32 serial sum. */
void orc_pose_integrate_f64(long n, const uint32_t* t_ms, const float* rate_x, const float* rate_y,
                            const float* h_m, const float* yaw_deg, const uint8_t* q,
                            double* xo, double* yo) {
  if (n <= 0) return;
  double px = 0.0, py = 0.0;
  xo[0] = 0.0;
  yo[0] = 0.0;
  for (long i = 1; i < n; i++) {
    float dt = (float)(t_ms[i] - t_ms[i - 1]) * 0.001f;
    float inc_n = 0.0f, inc_e = 0.0f;
    if (q[i] >= 50 && !isnan(rate_x[i]) && !isnan(rate_y[i]) && !isnan(h_m[i]) &&
        !isnan(yaw_deg[i])) {
      float vbx = rate_x[i] * h_m[i];
      float vby = rate_y[i] * h_m[i];
      float a = yaw_deg[i] * ((float)M_PI / 180.0f);
      float sn, cs;
      sincosf(a, &sn, &cs);
      float vn = vbx * cs - vby * sn;
      float ve = vbx * sn + vby * cs;
      inc_n = vn * dt;
      inc_e = ve * dt;
    }
    px += (double)inc_n;
    py += (double)inc_e;
    xo[i] = px;
    yo[i] = py;
  }
}

/* ------------------------------------------------------------------ */
/* glibc 2.39 sincosf (FMA build) restated -- SURVEY.md Appendix B       */
/* ------------------------------------------------------------------ */
typedef struct { double c0, c1, c2, c3, c4, s1, s2, s3; } orc_sc_poly;
static const orc_sc_poly SC_POS = {
  1.0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16,
  -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13 };
static const orc_sc_poly SC_NEG = {
  -1.0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16,
  -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13 };
static const double SC_HPI_INV = 0x1.45f306dc9c883p+23; /* 2/pi * 2^24 */
static const double SC_HPI = 0x1.921fb54442d18p+0;      /* pi/2 */

static inline uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* 2/pi as a bit string, 24 overlapping 32-bit windows (glibc __inv_pio4) and pi/2 * 2^-62: the
 * |y| >= 120 branch (reduce_large) multiplies the 24-bit mantissa by 96 bits of 2/pi in integers. */
static const uint32_t SC_INV_PIO4[24] = {
  0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529, 0x441529fc, 0x1529fc27,
  0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, 0x34ddc0db, 0xddc0db62, 0xc0db6295,
  0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041 };
static const double SC_PI63 = 0x1.921FB54442D18p-62;

/* returns 1 for every finite y (all three branches of glibc's sincosf are restated); Inf/NaN give NaN, NaN
 * like libm and return 1 as well -- the return value is kept for callers written against the earlier form */
int orc_sincosf_restated(float y, float* sinp, float* cosp) {
  const uint32_t top12 = (f32_bits(y) >> 20) & 0x7ff;
  double x = (double)y, xs, x2;
  const orc_sc_poly* p = &SC_POS;
  int n = 0;
  if (top12 >= 0x42f) {
    if (top12 >= 0x7f8) {              /* Inf, NaN: y - y */
      *sinp = *cosp = y - y;
      return 1;
    }
    /* |y| >= 120: reduce_large */
    uint32_t xi = f32_bits(y);
    const int sign = (int)(xi >> 31);
    const uint32_t* arr = &SC_INV_PIO4[(xi >> 26) & 15];
    const int shift = (xi >> 23) & 7;
    uint64_t res0, res1, res2, nn;
    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;
    res0 = (uint32_t)(xi * arr[0]);
    res1 = (uint64_t)xi * arr[4];
    res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    nn = (res0 + (1ULL << 61)) >> 62;
    res0 -= nn << 62;
    x = (double)(int64_t)res0 * SC_PI63;
    n = (int)nn;
    static const double sgn4[4] = { 1.0, -1.0, -1.0, 1.0 };
    xs = x * sgn4[(n + sign) & 3];
    x2 = x * x;
    if ((n + sign) & 2) p = &SC_NEG;
    goto poly;
  }
  if (top12 < 0x3f4) {                 /* |y| < pi/4 */
    if (top12 < 0x398) {               /* |y| < 2^-12 */
      *sinp = y;
      *cosp = 1.0f;
      return 1;
    }
    x2 = x * x;
    xs = x;
  } else {                             /* |y| < 120 */
    double r = x * SC_HPI_INV;
    n = ((int32_t)r + 0x800000) >> 24;
    x = fma(-(double)n, SC_HPI, x);    /* one rounding */
    static const double sgn[4] = { 1.0, -1.0, -1.0, 1.0 };
    xs = x * sgn[n & 3];
    x2 = x * x;
    if (n & 2) p = &SC_NEG;
  }
poly:;
  double x3 = x2 * xs, x4 = x2 * x2;
  double s1 = fma(x2, p->s3, p->s2);
  double c2 = fma(x2, p->c4, p->c3);
  double c1 = fma(x2, p->c1, p->c0);
  double x5 = x2 * x3, x6 = x2 * x4;
  double S = fma(x3, p->s1, xs);
  double C = fma(x4, p->c2, c1);
  S = fma(x5, s1, S);
  C = fma(x6, c2, C);
  if (n & 1) { *cosp = (float)S; *sinp = (float)C; }
  else       { *sinp = (float)S; *cosp = (float)C; }
  return 1;
}

/* Compare the restatement with libm over the float bit patterns [lo_bits, hi_bits)
 * (positive floats) and their negatives; returns the number of mismatches.
 * n_threads > 1 splits the range over pthreads (the full |y| < 120 sweep is 2.2 G floats). */
typedef struct { uint32_t lo, hi, stride, first; long bad; } sweep_job;

static void* sweep_worker(void* arg) {
  sweep_job* j = (sweep_job*)arg;
  for (uint64_t b = j->lo; b < j->hi; b += j->stride) {
    for (int sgn = 0; sgn < 2; sgn++) {
      uint32_t u = (uint32_t)b | (sgn ? 0x80000000u : 0u);
      float y, s0, c0, s1, c1;
      memcpy(&y, &u, 4);
      sincosf(y, &s0, &c0);
      if (!orc_sincosf_restated(y, &s1, &c1)) continue;
      if (f32_bits(s0) != f32_bits(s1) || f32_bits(c0) != f32_bits(c1)) {
        if (!j->bad) j->first = u;
        j->bad++;
      }
    }
  }
  return NULL;
}

long orc_sincosf_sweep(uint32_t lo_bits, uint32_t hi_bits, uint32_t stride, int n_threads,
                       uint32_t* first_bad) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  sweep_job jobs[256];
  const uint64_t span = (uint64_t)hi_bits - lo_bits;
  const uint64_t per = (span / n_threads / stride + 1) * stride;
  for (int t = 0; t < n_threads; t++) {
    uint64_t a = lo_bits + per * t, b = a + per;
    if (a > hi_bits) a = hi_bits;
    if (b > hi_bits) b = hi_bits;
    jobs[t] = (sweep_job){ (uint32_t)a, (uint32_t)b, stride, 0, 0 };
    pthread_create(&th[t], NULL, sweep_worker, &jobs[t]);
  }
  long bad = 0;
  uint32_t first = 0;
  for (int t = 0; t < n_threads; t++) {
    pthread_join(th[t], NULL);
    if (jobs[t].bad && !bad) first = jobs[t].first;
    bad += jobs[t].bad;
  }
  if (first_bad) *first_bad = first;
  return bad;
}

void orc_libm_sincosf(long n, const float* a, float* s, float* c) {
  for (long i = 0; i < n; i++) sincosf(a[i], &s[i], &c[i]);
}

/* ------------------------------------------------------------------ */
/* arithmetic KATs used by the kernels' design */
/* ------------------------------------------------------------------ */

/* Closed form of the reference's Bresenham walk: with m = max(|dx|,|dy|),
 * n = min(|dx|,|dy|) the ray has m+1 cells and cell k is
 *   major = major0 + k*s_major,  minor = minor0 + s_minor * floor((k*n + m/2) / m)
 * (x is the major axis when |dx| >= |dy|).  Walks every end point with
 * |dx|,|dy| <= maxd through the step rule of uav_local_nav.c:246-274 and returns
 * the number of cells that disagree with the closed form (0 expected). */
long orc_bresenham_closed_form_check(int maxd) {
  long bad = 0;
  for (int ddy = -maxd; ddy <= maxd; ddy++) {
    for (int ddx = -maxd; ddx <= maxd; ddx++) {
      const int adx = abs(ddx), ady = abs(ddy);
      const int stepx = (0 < ddx) ? 1 : -1, stepy = (0 < ddy) ? 1 : -1;
      const int m = adx > ady ? adx : ady, n = adx > ady ? ady : adx;
      const int xmajor = adx >= ady;
      int e = adx - ady, cx = 0, cy = 0, k = 0;
      for (;;) {
        int q = m ? (k * n + (m >> 1)) / m : 0;
        int fx = xmajor ? k * stepx : q * stepx;
        int fy = xmajor ? q * stepy : k * stepy;
        if (fx != cx || fy != cy || k > m) bad++;
        if (cx == ddx && cy == ddy) { if (k != m) bad++; break; }
        const int twice = 2 * e;
        if (twice >= -ady) { e -= ady; cx += stepx; }
        if (twice <= adx)  { e += adx; cy += stepy; }
        k++;
      }
    }
  }
  return bad;
}

/* The kernels replace floor((k*n + m/2) / m) by one multiply-high:
 *   q = (2*(k*n + m/2) * ceil(2^31 / m)) >> 32.
 * Exhaustive check for every 1 <= m <= 1024, 0 <= n <= m, 0 <= k <= m (the cap the
 * library enforces on ray length).  Returns the number of mismatches. */
long orc_magic_division_check(void) {
  long bad = 0;
  for (uint32_t m = 1; m <= 1024; m++) {
    const uint32_t inv = (uint32_t)(((1ull << 31) + m - 1) / m);
    const uint32_t h2 = 2 * (m >> 1);
    for (uint32_t n = 0; n <= m; n++) {
      const uint32_t n2 = 2 * n;
      for (uint32_t k = 0; k <= m; k++) {
        const uint32_t t2 = k * n2 + h2;
        const uint32_t q = (uint32_t)(((uint64_t)t2 * inv) >> 32);
        if (q != (k * n + (m >> 1)) / m) bad++;
      }
    }
  }
  return bad;
}

/* Clamp-add monoid: the map x -> clamp(x + a, lo, hi) composed over a
 * sequence of {-1, +6} steps, each clamped to [-80, 80], equals the step-by-step result
 * for every start value, under any re-association.  n_seq random sequences. */
typedef struct { int a, lo, hi; } orc_caf;
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline orc_caf caf_then(orc_caf f, orc_caf g) {   /* apply f first, then g */
  orc_caf r;
  r.a = clampi(f.a + g.a, -400, 400);
  r.lo = clampi(f.lo + g.a, g.lo, g.hi);
  r.hi = clampi(f.hi + g.a, g.lo, g.hi);
  return r;
}
long orc_clamp_monoid_check(long n_seq, uint64_t seed) {
  long bad = 0;
  uint64_t s = seed ? seed : 88172645463325252ull;
  for (long it = 0; it < n_seq; it++) {
    int len, ops[64];
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    len = 1 + (int)(s % 64);
    for (int i = 0; i < len; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; ops[i] = (s & 3) ? -1 : 6; }
    /* left fold and a split-in-the-middle association */
    orc_caf left = { 0, -80, 80 }, a = { 0, -80, 80 }, b = { 0, -80, 80 };
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    const int cut = (int)(s % (uint64_t)(len + 1));
    for (int i = 0; i < len; i++) {
      orc_caf op = { ops[i], -80, 80 };
      left = caf_then(left, op);
      if (i < cut) a = caf_then(a, op); else b = caf_then(b, op);
    }
    orc_caf split = caf_then(a, b);
    for (int x0 = -80; x0 <= 80; x0++) {
      int v = x0;
      for (int i = 0; i < len; i++) v = clampi(v + ops[i], -80, 80);
      if (clampi(x0 + left.a, left.lo, left.hi) != v) bad++;
      if (clampi(x0 + split.a, split.lo, split.hi) != v) bad++;
    }
  }
  return bad;
}

/* Time-slice map used by the sub-tile kernel: a slice's effect on a cell is kept as
 * (f(lo_min), f(lo_max), a) with a the plain sum of deltas saturated at +-30000, and applied as
 * f(x) = min(max(x + a, f(lo_min)), f(lo_max)).  Random op sequences (incl. very long ones, so that
 * a saturates), every start value, and a two-slice composition; returns the number of mismatches. */
long orc_slice_map_check(long n_seq, uint64_t seed) {
  long bad = 0;
  uint64_t s = seed ? seed : 0x9E3779B97F4A7C15ull;
  for (long it = 0; it < n_seq; it++) {
    int lens[2];
    for (int h = 0; h < 2; h++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      lens[h] = (s & 7) == 0 ? (int)(s % 70000) : (int)(s % 300);
    }
    int flo[2], fhi[2], a[2];
    static int8_t ops[2][70000];
    for (int h = 0; h < 2; h++) {
      flo[h] = -80; fhi[h] = 80; a[h] = 0;
      for (int i = 0; i < lens[h]; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const int d = (s % 10) < 8 ? -1 : ((s % 10) == 8 ? 6 : 0);
        ops[h][i] = (int8_t)d;
        flo[h] = clampi(flo[h] + d, -80, 80);
        fhi[h] = clampi(fhi[h] + d, -80, 80);
        a[h] = clampi(a[h] + d, -30000, 30000);
      }
    }
    for (int x0 = -80; x0 <= 80; x0++) {
      int v = x0;
      for (int h = 0; h < 2; h++)
        for (int i = 0; i < lens[h]; i++) v = clampi(v + ops[h][i], -80, 80);
      int w = x0;
      for (int h = 0; h < 2; h++) {
        w = w + a[h];
        if (w < flo[h]) w = flo[h];
        if (w > fhi[h]) w = fhi[h];
      }
      if (w != v) bad++;
    }
  }
  return bad;
}

/* FNV-1a over a byte buffer (grid fingerprints in tests and golden files). */
/* 64-bit digest of a grid, the formula of include/uqs_mapping.h (uqs_grid_hash): sum over cells i of
 * splitmix64(i << 8 | byte).  Restated here so that the checker does not call the product. */
uint64_t orc_grid_hash64(const uint8_t* g, size_t n) {
  uint64_t h = 0;
  for (size_t i = 0; i < n; i++) {
    uint64_t z = (((uint64_t)i << 8) | g[i]) + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    h += z ^ (z >> 31);
  }
  return h;
}

uint32_t orc_fnv1a32(const uint8_t* p, size_t n) {
  uint32_t h = 0x811c9dc5u;
  for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x01000193u; }
  return h;
}
