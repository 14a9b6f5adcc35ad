// uqs_dropin.cu -- the reference's own mapping symbols, backed by the CUDA path.
//
// uav_local_nav.c keeps the map and its functions file-static (uav_local_nav.c:188-192,
// :205, :241, :280); built with lines 181-385 removed and include/uqs_mapping.h included
// instead, the rest of that file links against the symbols below unchanged
// (INTEGRATION.md).  Semantics kept from the reference: silent no-op when !map_inited
// (:206, :281), whole ray dropped when an end is off-grid (:243-244), updates applied in
// call order.  Calls are queued in pinned host memory and replayed on the device in
// order by the same kernels as the batch API (accumulate mode); occ_grid is a pinned
// host mirror refreshed by uqs_dropin_flush() / before any read the library can see.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "uqs_host.h"

extern "C" {
int8_t* occ_grid = nullptr;
bool map_inited = false;
float map_origin_x = NAN;
float map_origin_y = NAN;
float tof_beams_m[4][8];
uint8_t pending_kf_flags = 0;
}

namespace uqs {

namespace {
constexpr int kQueueCap = 8192;

struct DropIn {
  bool configured = false;
  uqs_params cfg;
  int8_t* d_grid = nullptr;
  // pinned queue, SoA like the batch API
  float *qx = nullptr, *qy = nullptr, *qthird = nullptr, *qranges = nullptr;
  uint8_t* qkind = nullptr;
  int* w2g = nullptr;            // pinned, result of k_world_to_grid_one
  int count = 0;
  float snap_ox = 0.f, snap_oy = 0.f;   // origin the queued entries were issued under
  bool mirror_stale = false;
  Context* owner = nullptr;      // the device context the drop-in map lives on (the one current at configure time)
} D;

[[noreturn]] void die(const char* what) {
  fprintf(stderr, "libuqs_mapping: %s -- %s (there is no CPU fallback)\n", what, uqs_last_error());
  abort();
}

int flush_queue() {
  if (!D.count) return UQS_OK;
  uqs_params p = D.cfg;
  p.origin_x = D.snap_ox;
  p.origin_y = D.snap_oy;
  DevParams dp;
  int rc = make_dev_params(&p, &dp);
  if (rc) return rc;
  cudaStream_t st = g_ctx.stream();
  const size_t n = (size_t)D.count;
  if ((rc = g_ctx.in_x.ensure(kQueueCap * 4)) || (rc = g_ctx.in_y.ensure(kQueueCap * 4)) ||
      (rc = g_ctx.in_yaw.ensure(kQueueCap * 4)) || (rc = g_ctx.in_ranges.ensure(kQueueCap * 128)) ||
      (rc = g_ctx.in_kind.ensure(kQueueCap)))
    return rc;
  cudaError_t e = cudaMemcpyAsync(g_ctx.in_x.p, D.qx, n * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(g_ctx.in_y.p, D.qy, n * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(g_ctx.in_yaw.p, D.qthird, n * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(g_ctx.in_ranges.p, D.qranges, n * 128, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(g_ctx.in_kind.p, D.qkind, n, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return cuda_fail(e, "drop-in queue H2D");
  rc = replay_device(dp, 1, D.count, (float*)g_ctx.in_x.p, (float*)g_ctx.in_y.p, (float*)g_ctx.in_yaw.p,
                     (float*)g_ctx.in_ranges.p, (uint8_t*)g_ctx.in_kind.p, D.d_grid, 1, 0, p.H, true);
  if (rc) return rc;
  uqs_stats qs;
  if ((rc = fetch_stats(&qs, n))) return rc;             // synchronises; a dropped ray is an error, never silent
  D.count = 0;
  D.mirror_stale = true;
  return UQS_OK;
}

void enqueue(uint8_t kind, float a, float b, float c, const float* r32) {
  if (!D.configured) die("map symbol used before uqs_dropin_configure() succeeded");
  if (D.count && (map_origin_x != D.snap_ox || map_origin_y != D.snap_oy)) {
    if (flush_queue()) die("replay failed");
  }
  if (!D.count) { D.snap_ox = map_origin_x; D.snap_oy = map_origin_y; }
  const int i = D.count++;
  D.qx[i] = a; D.qy[i] = b; D.qthird[i] = c; D.qkind[i] = kind;
  memcpy(D.qranges + (size_t)i * 32, r32, 128);
  if (D.count == kQueueCap && flush_queue()) die("replay failed");
}
}  // namespace

void dropin_release() {
  if (D.owner && D.owner != g_cur) return;      // another device's context is shutting down
  if (D.d_grid) cudaFree(D.d_grid);
  if (occ_grid) cudaFreeHost(occ_grid);
  if (D.qx) cudaFreeHost(D.qx);
  if (D.qy) cudaFreeHost(D.qy);
  if (D.qthird) cudaFreeHost(D.qthird);
  if (D.qranges) cudaFreeHost(D.qranges);
  if (D.qkind) cudaFreeHost(D.qkind);
  if (D.w2g) cudaFreeHost(D.w2g);
  D = DropIn();
  occ_grid = nullptr;
  map_inited = false;
}

}  // namespace uqs

using namespace uqs;

extern "C" {

int uqs_dropin_configure(const uqs_params* p) {
  int rc = check_ready();
  if (rc == UQS_ERR_NOT_INIT && (rc = uqs_init(0))) return rc;
  DevParams dp;
  if ((rc = make_dev_params(p, &dp))) return rc;
  D.owner = nullptr;               // reconfiguring: whatever device held the old map, free it
  dropin_release();
  const size_t cells = (size_t)p->W * p->H;
  cudaError_t e = cudaMalloc(&D.d_grid, cells);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&occ_grid, cells, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&D.qx, kQueueCap * 4, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&D.qy, kQueueCap * 4, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&D.qthird, kQueueCap * 4, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&D.qranges, kQueueCap * 128, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&D.qkind, kQueueCap, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&D.w2g, 32, cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaMemsetAsync(D.d_grid, 0, cells, g_ctx.stream());
  if (e == cudaSuccess) e = cudaStreamSynchronize(g_ctx.stream());
  if (e != cudaSuccess) {
    dropin_release();
    return cuda_fail(e, "uqs_dropin_configure allocation");
  }
  memset(occ_grid, 0, cells);
  D.cfg = *p;
  D.owner = g_cur;
  D.configured = true;
  map_origin_x = p->origin_x;
  map_origin_y = p->origin_y;
  map_inited = false;
  pending_kf_flags = 0;
  return UQS_OK;
}

void map_reset(void) {
  if (!D.configured) die("map_reset() before uqs_dropin_configure()");
  D.count = 0;   /* queued updates would be wiped by the memset anyway */
  cudaError_t e = cudaMemsetAsync(D.d_grid, 0, (size_t)D.cfg.W * D.cfg.H, g_ctx.stream());
  if (e == cudaSuccess) e = cudaStreamSynchronize(g_ctx.stream());
  if (e != cudaSuccess) { cuda_fail(e, "map_reset"); die("map_reset failed"); }
  memset(occ_grid, 0, (size_t)D.cfg.W * D.cfg.H);
  D.mirror_stale = false;
}

void uqs_dropin_flush(void) {
  if (!D.configured) die("uqs_dropin_flush() before uqs_dropin_configure()");
  if (flush_queue()) die("replay failed");
  if (D.mirror_stale) {
    cudaError_t e = cudaMemcpyAsync(occ_grid, D.d_grid, (size_t)D.cfg.W * D.cfg.H, cudaMemcpyDeviceToHost, g_ctx.stream());
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_ctx.stream());
    if (e != cudaSuccess) { cuda_fail(e, "occ_grid D2H"); die("grid read-back failed"); }
    D.mirror_stale = false;
  }
}

/* Push the host mirror to the device (after the caller edited occ_grid[] directly). */
int uqs_dropin_upload(void) {
  if (!D.configured) { set_error("drop-in not configured"); return UQS_ERR_NOT_INIT; }
  int rc = flush_queue();
  if (rc) return rc;
  cudaError_t e = cudaMemcpyAsync(D.d_grid, occ_grid, (size_t)D.cfg.W * D.cfg.H, cudaMemcpyHostToDevice, g_ctx.stream());
  if (e == cudaSuccess) e = cudaStreamSynchronize(g_ctx.stream());
  if (e != cudaSuccess) return cuda_fail(e, "occ_grid H2D");
  D.mirror_stale = false;
  return UQS_OK;
}

bool world_to_grid(float x, float y, int* gx, int* gy) {
  if (!map_inited) return false;                               /* :206 */
  if (!D.configured) die("world_to_grid() before uqs_dropin_configure()");
  uqs_params p = D.cfg;
  p.origin_x = map_origin_x;
  p.origin_y = map_origin_y;
  DevParams dp;
  if (make_dev_params(&p, &dp)) die("bad map parameters");
  int* dptr = nullptr;
  cudaError_t e = cudaHostGetDevicePointer((void**)&dptr, D.w2g, 0);
  if (e == cudaSuccess) {
    k_world_to_grid_one<<<1, 1, 0, g_ctx.stream()>>>(dp, x, y, dptr);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(g_ctx.stream());
  if (e != cudaSuccess) { cuda_fail(e, "world_to_grid"); die("world_to_grid kernel failed"); }
  g_ctx.launches += 1;
  if (!D.w2g[0]) return false;
  *gx = D.w2g[1];
  *gy = D.w2g[2];
  return true;
}

void raycast_update(float x0, float y0, float x1, float y1, bool hit_occ) {
  if (!map_inited) return;      /* world_to_grid() fails first thing at :243 when !map_inited */
  float r[32];
  memset(r, 0, sizeof(r));
  r[0] = y1;
  r[1] = hit_occ ? 1.0f : 0.0f;
  enqueue(1, x0, y0, x1, r);
}

void map_update_from_beams(float x_m, float y_m, float yaw_deg) {
  if (!map_inited) return;      /* :281 */
  enqueue(0, x_m, y_m, yaw_deg, &tof_beams_m[0][0]);
}

/* ---- N2 / N3 drop-ins (uav_local_nav.c:308-385) ------------------------------------------------ */

void map_recenter_shift(int sx_cells, int sy_cells) {
  if (!D.configured) die("map_recenter_shift() before uqs_dropin_configure()");
  if (flush_queue()) die("replay failed");
  const size_t cells = (size_t)D.cfg.W * D.cfg.H;
  int rc = g_ctx.out_grids.ensure(cells);
  if (rc) die("allocation failed");
  cudaStream_t st = g_ctx.stream();
  k_recenter_shift<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(D.d_grid, (int8_t*)g_ctx.out_grids.p, D.cfg.W, D.cfg.H,
                                                                    sx_cells, sy_cells);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(D.d_grid, g_ctx.out_grids.p, cells, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { cuda_fail(e, "map_recenter_shift"); die("recenter shift failed"); }
  g_ctx.launches += 1;
  D.mirror_stale = true;
}

void map_recentre_if_needed(float x_m, float y_m) {
  if (!map_inited) return;                                     /* :325 */
  if (!D.configured) die("map_recentre_if_needed() before uqs_dropin_configure()");
  int* dptr = nullptr;
  cudaError_t e = cudaHostGetDevicePointer((void**)&dptr, D.w2g, 0);
  if (e == cudaSuccess) {
    k_recenter_decide_one<<<1, 1, 0, g_ctx.stream()>>>(D.cfg.res_m, D.cfg.size_m, map_origin_x, map_origin_y, x_m, y_m, dptr);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(g_ctx.stream());
  if (e != cudaSuccess) { cuda_fail(e, "map_recentre_if_needed"); die("recenter decision failed"); }
  g_ctx.launches += 1;
  if (!D.w2g[0]) return;
  const int sx = D.w2g[1], sy = D.w2g[2];
  map_recenter_shift(sx, sy);                                  /* flushes updates queued under the old origin */
  memcpy(&map_origin_x, &D.w2g[3], 4);
  memcpy(&map_origin_y, &D.w2g[4], 4);
  pending_kf_flags |= (1u << 5);                               /* KF_MAP_RECENTER, :225,:350 */
  printf("Map recenter: shift (%d,%d) cells => new origin (%.2f,%.2f)\n", sx, sy, map_origin_x, map_origin_y);
}

int frontier_score_dir(float x_m, float y_m, float yaw_deg, float offset_deg) {
  if (!map_inited) return 0;                                   /* :357 */
  if (!D.configured) die("frontier_score_dir() before uqs_dropin_configure()");
  if (flush_queue()) die("replay failed");                     /* reads must see every queued update */
  uqs_params p = D.cfg;
  p.origin_x = map_origin_x;
  p.origin_y = map_origin_y;
  DevParams dp;
  if (make_dev_params(&p, &dp)) die("bad map parameters");
  int rc = g_ctx.in_x.ensure(64);
  if (!rc) rc = g_ctx.w->counters.ensure(64 * 8);
  if (rc) die("allocation failed");
  cudaStream_t st = g_ctx.stream();
  const float q[4] = { x_m, y_m, yaw_deg, offset_deg };
  float* dq = (float*)g_ctx.in_x.p;
  int* dptr = nullptr;
  cudaError_t e = cudaMemcpyAsync(dq, q, 16, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&dptr, D.w2g, 0);
  if (e == cudaSuccess) {
    k_frontier_scores<<<1, 32, 0, st>>>(dp, D.d_grid, 1, dq, dq + 1, dq + 2, dq + 3, dptr,
                                         (unsigned long long*)g_ctx.w->counters.p + 24);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { cuda_fail(e, "frontier_score_dir"); die("frontier kernel failed"); }
  g_ctx.launches += 1;
  return D.w2g[0];
}

}  // extern "C"
