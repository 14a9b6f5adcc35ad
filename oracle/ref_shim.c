/*
 * ref_shim.c -- TEST INFRASTRUCTURE (oracle/_ref).  Not part of the product.
 *
 * This file is appended, by oracle/build_ref.sh, to the mapping block that the
 * script extracts by line range from /root/reference/uav_local_nav.c.  The
 * reference keeps every mapping symbol `static` inside one translation unit
 * (uav_local_nav.c:188-192, :205, :241, :280), so the only way to call the
 * reference's own code is to live in the same TU.  Nothing here re-states the
 * algorithm: every function forwards to the reference's functions unchanged.
 *
 * The wrappers mirror how the reference itself drives the path:
 *   - ref_reset()        = the map-init site, uav_local_nav.c:2187-2194
 *   - ref_frame()        = the body of log_tick(), uav_local_nav.c:1629-1635,
 *                          with pose_good_for_mapping() forced true
 *   - ref_replay()       = that body in a loop over a whole log (timed by bench)
 */

int ref_map_w(void) { return MAP_W; }
int ref_map_h(void) { return MAP_H; }
float ref_map_res(void) { return MAP_RES_M; }
float ref_map_size(void) { return MAP_SIZE_M; }

void ref_reset(float ox, float oy) {
  map_origin_x = ox;
  map_origin_y = oy;
  memset(occ_grid, 0, sizeof(occ_grid));
  map_inited = true;
  pending_kf_flags = 0;
}

int8_t* ref_grid(void) { return occ_grid; }
float ref_origin_x(void) { return map_origin_x; }
float ref_origin_y(void) { return map_origin_y; }
int ref_recentered(void) { return (pending_kf_flags & KF_MAP_RECENTER) ? 1 : 0; }

void ref_set_beams(const float* b32) { memcpy(tof_beams_m, b32, sizeof(tof_beams_m)); }

/* uav_local_nav.c:1629-1635 with the gate true; recenter optional so the
 * harness can assert that bounded synthetic logs never trigger it. */
void ref_frame(float x, float y, float yaw_deg, const float* b32, int allow_recenter) {
  memcpy(tof_beams_m, b32, sizeof(tof_beams_m));
  if (allow_recenter) map_recentre_if_needed(x, y);
  map_update_from_beams(x, y, yaw_deg);
}

void ref_replay(long n_frames, const float* x, const float* y, const float* yaw_deg,
                const float* ranges /* n_frames*32 */, int allow_recenter) {
  for (long i = 0; i < n_frames; i++) {
    memcpy(tof_beams_m, ranges + i * 32, sizeof(tof_beams_m));
    if (allow_recenter) map_recentre_if_needed(x[i], y[i]);
    map_update_from_beams(x[i], y[i], yaw_deg[i]);
  }
}

int ref_world_to_grid(float x, float y, int* gx, int* gy) {
  int a = -1, b = -1;
  bool ok = world_to_grid(x, y, &a, &b);
  *gx = a; *gy = b;
  return ok ? 1 : 0;
}

void ref_raycast_update(float x0, float y0, float x1, float y1, int hit) {
  raycast_update(x0, y0, x1, y1, hit != 0);
}

void ref_recenter_shift(int sx, int sy) { map_recenter_shift(sx, sy); }
void ref_recentre_if_needed(float x, float y) { map_recentre_if_needed(x, y); }

int ref_frontier_score_dir(float x, float y, float yaw_deg, float off_deg) {
  return frontier_score_dir(x, y, yaw_deg, off_deg);
}

/* N1: raw 518-byte scan frame -> tof_beams_m (uav_local_nav.c:1344-1359). */
void ref_beams_from_frame(const uint8_t* frame518, float* out32) {
  compute_beams_and_minima(frame518);
  memcpy(out32, tof_beams_m, sizeof(tof_beams_m));
}

uint8_t ref_xor8(const uint8_t* p, int len) { return xor8(p, len); }
