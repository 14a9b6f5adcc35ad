/*
 * uqs_scanlog.c -- N4: reader of the reference's binary scan log (host C, file I/O only).
 *
 * Format written by uav_local_nav.c: the 7-byte text header "SCLOG2\n" when the file is created
 * (log_init, :1499-1508; the file is opened in append mode, so flights concatenate without further
 * headers), then packed 569-byte records `scanrec_t` (:1522-1547) with magic 'SCN2' = 0x324E4353:
 *   u32 magic, u32 host_ms, u32 scan_ms, f32 x_m, y_m, yaw_deg, alt_m, roll_rad, pitch_rad, rf_m,
 *   of_rate_x, of_rate_y, u8 of_q, u8 state, u8 kf_flags, u16 pad, u32 sys_health, u8 grid_raw[512].
 * x_m / y_m / yaw_deg are NaN when the FC had no position / attitude yet (:1559-1561); such records
 * cannot be mapped and are skipped unless keep_nan_pose is set.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/uqs_mapping.h"

#define SCANREC_BYTES 569
#define SCANREC_MAGIC 0x324E4353u

static uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static float rdf(const uint8_t* p) { uint32_t u = rd32(p); float f; memcpy(&f, &u, 4); return f; }

/* Counts the mappable records of a scan log (returns < 0 on error: -1 open, -2 bad header/magic). */
long uqs_scanlog_count(const char* path, int keep_nan_pose) {
  return uqs_scanlog_read(path, keep_nan_pose, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
}

/* Reads up to max_records records into SoA arrays (any output may be NULL).  Returns the number of
 * records that qualify (which may exceed max_records: call with max_records = 0 to size buffers). */
long uqs_scanlog_read(const char* path, int keep_nan_pose, long max_records, uint32_t* host_ms, uint32_t* scan_ms,
                      float* x_m, float* y_m, float* yaw_deg, float* alt_m, float* of_rate_x, float* of_rate_y,
                      uint8_t* of_q, uint8_t* kf_flags, uint8_t* grid_raw /* [n][512] */) {
  FILE* fp = fopen(path, "rb");
  if (!fp) return -1;
  uint8_t rec[SCANREC_BYTES];
  char hdr[7];
  if (fread(hdr, 1, 7, fp) != 7 || memcmp(hdr, "SCLOG2\n", 7) != 0) { fclose(fp); return -2; }
  long n = 0;
  for (;;) {
    size_t got = fread(rec, 1, SCANREC_BYTES, fp);
    if (got == 0) break;
    if (got >= 7 && memcmp(rec, "SCLOG2\n", 7) == 0) {        /* tolerate a repeated header (concatenated files) */
      memmove(rec, rec + 7, got - 7);
      size_t more = fread(rec + got - 7, 1, 7, fp);
      got = got - 7 + more;
    }
    if (got < SCANREC_BYTES) break;                            /* truncated tail: power was cut mid-record */
    if (rd32(rec) != SCANREC_MAGIC) { fclose(fp); return -2; }
    const float x = rdf(rec + 12), y = rdf(rec + 16), yaw = rdf(rec + 20);
    if (!keep_nan_pose && (isnan(x) || isnan(y) || isnan(yaw))) continue;
    if (n < max_records) {
      if (host_ms) host_ms[n] = rd32(rec + 4);
      if (scan_ms) scan_ms[n] = rd32(rec + 8);
      if (x_m) x_m[n] = x;
      if (y_m) y_m[n] = y;
      if (yaw_deg) yaw_deg[n] = yaw;
      if (alt_m) alt_m[n] = rdf(rec + 24);
      if (of_rate_x) of_rate_x[n] = rdf(rec + 40);
      if (of_rate_y) of_rate_y[n] = rdf(rec + 44);
      if (of_q) of_q[n] = rec[48];
      if (kf_flags) kf_flags[n] = rec[50];
      if (grid_raw) memcpy(grid_raw + (size_t)n * 512, rec + 57, 512);
    }
    n++;
  }
  fclose(fp);
  return n;
}

/* ---- navlog.csv (uav_local_nav.c:1489-1494 header, :1586-1625 rows) -------------------------------------------
 * One text row per log_tick at LOG_HZ, 22 comma-separated columns:
 *   t_ms,state,want_arm,armed,mode,yaw_deg,alt_m,alt_src,x_m,y_m,vx_mps,vy_mps,rf_m,of_q,of_rate_x,of_rate_y,
 *   tof_f,tof_r,tof_b,tof_l,batt_v,batt_cells
 * Missing values are the text "nan" (:1596-1616); of_q is 0 when the flow sample is stale (:1612).  The file is
 * opened in append mode and the header line is written only into an empty file, so flights concatenate; a row cut
 * short by a power loss is dropped.  These columns are exactly what P0 consumes (t_ms, of_rate_x/y, height =
 * rf_m or alt_m, yaw_deg, of_q) plus the FC pose the reference maps with (x_m, y_m). */
#define NAVLOG_COLS 22

static float field_f32(const char* s) { return strtof(s, NULL); }          /* "nan" -> NaN, like the writer's text */

/* Reads up to max_rows rows into SoA arrays (any output may be NULL).  Returns the number of well-formed data
 * rows in the file (which may exceed max_rows: call with max_rows = 0 to size buffers), -1 if it cannot be opened. */
long uqs_navlog_read(const char* path, long max_rows, uint32_t* t_ms, float* yaw_deg, float* alt_m, float* x_m,
                     float* y_m, float* vx_mps, float* vy_mps, float* rf_m, uint8_t* of_q, float* of_rate_x,
                     float* of_rate_y, float* tof4 /* [n][4]: front, right, back, left (filtered, m) */) {
  FILE* fp = fopen(path, "r");
  if (!fp) return -1;
  char line[1024];
  long n = 0;
  while (fgets(line, sizeof line, fp)) {
    size_t len = strlen(line);
    if (len == 0 || line[len - 1] != '\n') {                       /* over-long or unterminated (truncated) row */
      if (len == sizeof line - 1) { int c; while ((c = fgetc(fp)) != EOF && c != '\n') { } }
      continue;
    }
    if (line[0] < '0' || line[0] > '9') continue;                  /* header line(s), blank lines */
    char* col[NAVLOG_COLS];
    int nc = 0;
    char* p = line;
    col[nc++] = p;
    for (; *p; p++) {
      if (*p == ',') {
        *p = 0;
        if (nc < NAVLOG_COLS) col[nc] = p + 1;
        nc++;
      } else if (*p == '\n' || *p == '\r') {
        *p = 0;
      }
    }
    if (nc != NAVLOG_COLS) continue;
    if (n < max_rows) {
      if (t_ms) t_ms[n] = (uint32_t)strtoull(col[0], NULL, 10);     /* low 32 bits: P0 only uses differences */
      if (yaw_deg) yaw_deg[n] = field_f32(col[5]);
      if (alt_m) alt_m[n] = field_f32(col[6]);
      if (x_m) x_m[n] = field_f32(col[8]);
      if (y_m) y_m[n] = field_f32(col[9]);
      if (vx_mps) vx_mps[n] = field_f32(col[10]);
      if (vy_mps) vy_mps[n] = field_f32(col[11]);
      if (rf_m) rf_m[n] = field_f32(col[12]);
      if (of_q) { unsigned long q = strtoul(col[13], NULL, 10); of_q[n] = (uint8_t)(q > 255 ? 255 : q); }
      if (of_rate_x) of_rate_x[n] = field_f32(col[14]);
      if (of_rate_y) of_rate_y[n] = field_f32(col[15]);
      if (tof4) for (int d = 0; d < 4; d++) tof4[(size_t)n * 4 + d] = field_f32(col[16 + d]);
    }
    n++;
  }
  fclose(fp);
  return n;
}
