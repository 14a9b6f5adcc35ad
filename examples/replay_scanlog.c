/*
 * replay_scanlog.c -- post-flight replay harness in plain C (the host side the north star asks for):
 * scanlog.bin (uav_local_nav.c:1522-1581) -> beams (N1) -> occupancy grid (hot path), all through
 * include/uqs_mapping.h.  Also drives the drop-in symbols the way log_tick() does and checks that both
 * routes give the same grid.
 *
 *   gcc -O2 -Iinclude examples/replay_scanlog.c -L micro-quad-slam_b200 -luqs_mapping -lm -o replay_scanlog
 *   LD_LIBRARY_PATH=micro-quad-slam_b200 ./replay_scanlog /mnt/sdcard/scanlog.bin 400 0.05
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "uqs_mapping.h"

static unsigned fnv1a(const unsigned char* p, size_t n) {
  unsigned h = 0x811c9dc5u;
  for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x01000193u; }
  return h;
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s scanlog.bin W res_m\n", argv[0]); return 2; }
  const char* path = argv[1];
  uqs_params p;
  uqs_params_default(&p);
  p.W = p.H = atoi(argv[2]);
  p.res_m = (float)atof(argv[3]);
  p.size_m = p.W * p.res_m;

  long n = uqs_scanlog_count(path, 0);
  if (n <= 0) { fprintf(stderr, "cannot read %s (%ld)\n", path, n); return 1; }
  float *x = malloc(n * 4), *y = malloc(n * 4), *yaw = malloc(n * 4), *beams = malloc((size_t)n * 128);
  unsigned char* raw = malloc((size_t)n * 512);
  int8_t* grid = malloc((size_t)p.W * p.H);
  if (uqs_scanlog_read(path, 0, n, NULL, NULL, x, y, yaw, NULL, NULL, NULL, NULL, NULL, raw) != n) return 1;

  if (uqs_init(0)) { fprintf(stderr, "%s\n", uqs_last_error()); return 1; }
  p.origin_x = x[0];                                   /* map centred on the first pose, like :2188-2189 */
  p.origin_y = y[0];
  uqs_stats st;
  if (uqs_beams_from_scans(n, raw, p.max_range_m, beams, NULL) ||
      uqs_replay(&p, 1, (int)n, x, y, yaw, beams, grid, &st)) {
    fprintf(stderr, "%s\n", uqs_last_error());
    return 1;
  }
  const unsigned h_batch = fnv1a((unsigned char*)grid, (size_t)p.W * p.H);

  /* the same log through the reference's own symbols (HOVER init :2187-2194, log_tick :1633-1635) */
  if (uqs_dropin_configure(&p)) { fprintf(stderr, "%s\n", uqs_last_error()); return 1; }
  map_origin_x = x[0];
  map_origin_y = y[0];
  map_reset();
  map_inited = true;
  for (long i = 0; i < n; i++) {
    memcpy(tof_beams_m, beams + i * 32, sizeof(tof_beams_m));
    map_update_from_beams(x[i], y[i], yaw[i]);
  }
  uqs_dropin_flush();
  const unsigned h_dropin = fnv1a((unsigned char*)occ_grid, (size_t)p.W * p.H);

  printf("records=%ld updates=%llu fnv_batch=%08x fnv_dropin=%08x %s\n", n, (unsigned long long)st.ray_cell_updates, h_batch,
         h_dropin, h_batch == h_dropin ? "MATCH" : "MISMATCH");
  uqs_shutdown();
  return h_batch == h_dropin ? 0 : 1;
}
