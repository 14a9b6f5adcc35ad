/*
 * replay_building_multi.c -- BASELINE config 4 in plain C: ONE large occupancy grid tiled across every GPU of
 * the box from one host thread: scanlog.bin (uav_local_nav.c:1522-1581) -> beams (N1)
 * -> uqs_multi_replay_banded(): log slices up on N PCIe links, all-gathered over NVLink, each GPU replays the
 * row band of occ_grid (uav_local_nav.c:188) it owns, one NCCL all-gather assembles the grid.  The grid is
 * replayed a second time on GPU 0 alone; both must be byte-identical (ownership, unlike a sum of partial grids,
 * is exact under the per-update clamp at uav_local_nav.c:259-260).
 *
 *   gcc -O2 -Iinclude examples/replay_building_multi.c -L micro-quad-slam_b200 -luqs_mapping -lm -o replay_building_multi
 *   LD_LIBRARY_PATH=micro-quad-slam_b200 ./replay_building_multi scanlog.bin 16384 0.01 8
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
  if (argc < 5) { fprintf(stderr, "usage: %s scanlog.bin W res_m n_gpus\n", argv[0]); return 2; }
  const char* path = argv[1];
  uqs_params p;
  uqs_params_default(&p);
  p.W = p.H = atoi(argv[2]);
  p.res_m = (float)atof(argv[3]);
  p.size_m = p.W * p.res_m;
  const int n_gpus = atoi(argv[4]);

  long n = uqs_scanlog_count(path, 0);
  if (n <= 0) { fprintf(stderr, "cannot read %s (%ld)\n", path, n); return 1; }
  float *x = malloc(n * 4), *y = malloc(n * 4), *yaw = malloc(n * 4), *beams = malloc((size_t)n * 128);
  unsigned char* raw = malloc((size_t)n * 512);
  const size_t cells = (size_t)p.W * p.H;
  int8_t *grid_multi = malloc(cells), *grid_one = malloc(cells);
  if (!x || !y || !yaw || !beams || !raw || !grid_multi || !grid_one) return 1;
  if (uqs_scanlog_read(path, 0, n, NULL, NULL, x, y, yaw, NULL, NULL, NULL, NULL, NULL, raw) != n) return 1;
  p.origin_x = x[0];
  p.origin_y = y[0];

  /* every GPU of the box, one host thread */
  if (uqs_multi_init(n_gpus, NULL)) { fprintf(stderr, "%s\n", uqs_last_error()); return 1; }
  uqs_stats st, st1;
  if (uqs_beams_from_scans(n, raw, p.max_range_m, beams, NULL) ||              /* on device 0 (the current one) */
      uqs_multi_replay_banded(&p, (int)n, x, y, yaw, beams, grid_multi, &st)) {
    fprintf(stderr, "%s\n", uqs_last_error());
    return 1;
  }
  int edges[17];
  const int nb = uqs_band_edges(edges);                 /* cuts that give every GPU the same share of the log */
  for (int r = 0; r < nb; r++) printf("gpu %d owns rows [%d, %d)\n", r, edges[r], edges[r + 1]);
  /* the same log on one GPU */
  if (uqs_multi_select(0) || uqs_replay(&p, 1, (int)n, x, y, yaw, beams, grid_one, &st1)) {
    fprintf(stderr, "%s\n", uqs_last_error());
    return 1;
  }
  const unsigned h_multi = fnv1a((unsigned char*)grid_multi, cells), h_one = fnv1a((unsigned char*)grid_one, cells);
  const int same = h_multi == h_one && memcmp(grid_multi, grid_one, cells) == 0 && st.ray_cell_updates == st1.ray_cell_updates;
  printf("records=%ld gpus=%d nccl=%d updates=%llu fnv_multi=%08x fnv_one=%08x %s\n", n, uqs_multi_count(), uqs_nccl_version(),
         (unsigned long long)st.ray_cell_updates, h_multi, h_one, same ? "MATCH" : "MISMATCH");
  uqs_multi_shutdown();
  return same ? 0 : 1;
}
