// uqs_multi.cu -- the multi-GPU part of the C ABI.
//
//   * many flights (configs 3, 5): uqs_flight_shard() -- contiguous blocks of flights per GPU, NO data-path
//     collective; each GPU runs the ordinary single-device calls on its block.
//   * one very large grid (config 4): uqs_row_band() -- every GPU OWNS a band of rows of occ_grid
//     (uav_local_nav.c:188) and replays, in log order, every frame that can reach it.  The reference clamps
//     after every update (:259-260), so partial grids cannot be summed (an ncclSum of halos is exact only while
//     no clamp engages); ownership keeps every cell's updates on one GPU in reference order, and the single
//     exchange is a gather of disjoint bands over NCCL (NVLink / NVSwitch).
//
// Two host models:
//   (i)  one process per GPU (torchrun / MPI): uqs_init(local_rank), uqs_comm_unique_id() on rank 0, the id
//        broadcast by the launcher's own means, uqs_comm_init_rank() -> ncclCommInitRank;
//   (ii) one host thread driving N devices (a plain C harness): uqs_multi_init() -> ncclCommInitAll, one device
//        context per GPU, uqs_multi_select() to address one of them with the single-device calls.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, preferring a copy the process has already loaded, e.g.
// PyTorch's) so that the library itself has no link-time dependency a single-GPU user would have to satisfy.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "uqs_host.h"

namespace uqs {

namespace {

struct Nccl {
  void* handle = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
} N;

int nccl_load() {
  if (N.handle) return UQS_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // the copy this process already uses, if any
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("NCCL is not available (dlopen libnccl.so.2: %s)", dlerror());
    return UQS_ERR_NO_NCCL;
  }
  bool ok = true;
  auto sym = [&](const char* name) { void* p = dlsym(h, name); ok = ok && p != nullptr; return p; };
  N.GetVersion = (decltype(N.GetVersion))sym("ncclGetVersion");
  N.GetUniqueId = (decltype(N.GetUniqueId))sym("ncclGetUniqueId");
  N.CommInitRank = (decltype(N.CommInitRank))sym("ncclCommInitRank");
  N.CommInitAll = (decltype(N.CommInitAll))sym("ncclCommInitAll");
  N.CommDestroy = (decltype(N.CommDestroy))sym("ncclCommDestroy");
  N.AllGather = (decltype(N.AllGather))sym("ncclAllGather");
  N.Broadcast = (decltype(N.Broadcast))sym("ncclBroadcast");
  N.Send = (decltype(N.Send))sym("ncclSend");
  N.Recv = (decltype(N.Recv))sym("ncclRecv");
  N.GroupStart = (decltype(N.GroupStart))sym("ncclGroupStart");
  N.GroupEnd = (decltype(N.GroupEnd))sym("ncclGroupEnd");
  N.GetErrorString = (decltype(N.GetErrorString))sym("ncclGetErrorString");
  if (!ok) {
    set_error("libnccl.so.2 lacks a required symbol");
    return UQS_ERR_NO_NCCL;
  }
  N.handle = h;
  return UQS_OK;
}

int nccl_fail(ncclResult_t r, const char* what) {
  set_error("NCCL error in %s: %s", what, N.GetErrorString ? N.GetErrorString(r) : "?");
  return UQS_ERR_NCCL;
}

// the contexts of host model (ii)
constexpr int kMaxDevices = 16;
Context g_multi[kMaxDevices];
int g_n_multi = 0;
// device-side log and grid of the banded replay, one set per device context
struct BandBufs { DevBuf x, y, yaw, ranges, grid; };
BandBufs g_bands[kMaxDevices];

}  // namespace

// ---- balanced ownership --------------------------------------------------------------------------------------------
// Equal row bands are only balanced when the flight covers the grid evenly; a building sweep covers the middle of a
// 16384-row grid and leaves the outer bands idle.  The owned bands are therefore cut so that every rank gets the
// same share of the log: a histogram of the frames' origin rows (device), widened by the sensor's reach (a frame
// works on every band its rays can touch), and a prefix sum on the host.  Any partition gives the same bytes --
// ownership is what makes the result exact, not where the cuts are -- and every rank derives the same cuts from
// the same log.
__global__ void k_row_histogram(const __grid_constant__ DevParams p, int n_frames, const float* __restrict__ x,
                                const float* __restrict__ y, int align, unsigned* __restrict__ hist) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  int gx, gy;
  if (world_to_grid(p, x[f], y[f], gx, gy)) atomicAdd(&hist[gy / align], 1u);
}

static int g_balance = 1;                          // uqs_set_band_balance

// edges[0..world]: rank r owns rows [edges[r], edges[r+1]); multiples of `align` (except the last = H)
static int band_edges(const DevParams& dp, int n_frames, const float* x_dev, const float* y_dev, int world, int align, int* edges) {
  for (int r = 0; r <= world; r++) {
    int r0, rows;
    uqs_row_band(dp.H, std::min(r, world - 1), world, align, &r0, &rows);
    edges[r] = r < world ? r0 : dp.H;
  }
  if (!g_balance || world <= 1) return UQS_OK;
  const int units = (dp.H + align - 1) / align;
  int rc = g_ctx.in_kind.ensure((size_t)units * sizeof(unsigned));
  if (rc) return rc;
  cudaStream_t st = g_ctx.stream();
  unsigned* d_hist = (unsigned*)g_ctx.in_kind.p;
  std::vector<unsigned> h(units);
  cudaError_t e = zero_async(d_hist, (size_t)units * sizeof(unsigned), st);
  if (e == cudaSuccess) {
    k_row_histogram<<<(unsigned)((n_frames + 255) / 256), 256, 0, st>>>(dp, n_frames, x_dev, y_dev, align, d_hist);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), d_hist, (size_t)units * sizeof(unsigned), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "row histogram");
  g_ctx.launches += 2;
  // a frame at row y works on rows y +- reach: box filter of that width through prefix sums
  const int reach = std::max(1, (int)std::ceil(dp.max_range / dp.res / (float)align) + 1);
  std::vector<unsigned long long> pre(units + 1, 0), work(units + 1, 0);
  for (int u = 0; u < units; u++) pre[u + 1] = pre[u] + h[u];
  if (pre[units] == 0) return UQS_OK;                               // no frame on the grid: equal bands
  for (int u = 0; u < units; u++) {
    const int a = std::max(0, u - reach), b = std::min(units, u + reach + 1);
    work[u + 1] = work[u] + (pre[b] - pre[a]);
  }
  const unsigned long long total = work[units];
  int u = 0;
  for (int r = 1; r < world; r++) {
    const unsigned long long want = total / world * r + total % world * r / world;
    while (u < units && work[u] < want) u++;
    // every rank keeps at least one unit of rows (its ray set-up counts the log like everyone else's)
    const int lo = edges[r - 1] + align, hi = dp.H - (world - r) * align;
    edges[r] = hi >= lo ? std::min(std::max(u * align, lo), hi) : std::min(lo, dp.H);
  }
  edges[world] = dp.H;
  return UQS_OK;
}

void comm_release() {
  if (g_ctx.comm && N.CommDestroy) N.CommDestroy((ncclComm_t)g_ctx.comm);
  g_ctx.comm = nullptr;
  g_ctx.comm_rank = 0;
  g_ctx.comm_nranks = 1;
}

// Gather the owned row bands of `grid` (full W*H on every rank, band r valid on rank r) so that every rank holds
// the whole grid.  In place; equal bands: one ncclAllGather, ragged: grouped ncclSend/ncclRecv between all pairs.
static int gather_bands(Context& c, const int* edges, int8_t* grid, int W, int H, cudaStream_t st) {
  const int n = c.comm_nranks;
  if (n <= 1) return UQS_OK;
  (void)H;
  ncclComm_t comm = (ncclComm_t)c.comm;
  int r0, rows;
  const int first_rows = edges[1] - edges[0];
  bool equal = true;
  for (int r = 0; r < n; r++) equal = equal && edges[r + 1] - edges[r] == first_rows;
  ncclResult_t q;
  if (equal) {
    r0 = edges[c.comm_rank];
    q = N.AllGather(grid + (size_t)r0 * W, grid, (size_t)first_rows * W, ncclInt8, comm, st);
    return q == ncclSuccess ? UQS_OK : nccl_fail(q, "ncclAllGather(bands)");
  }
  // ragged bands: every rank sends its band to every peer and receives theirs, all in one group (the transfers run
  // concurrently over NVSwitch; eight back-to-back broadcasts of the same bytes took 2.5 ms for 268 MB on 8 GPUs)
  if ((q = N.GroupStart()) != ncclSuccess) return nccl_fail(q, "ncclGroupStart");
  const int me = c.comm_rank;
  const int my0 = edges[me], my_rows = edges[me + 1] - edges[me];
  for (int r = 0; r < n && q == ncclSuccess; r++) {
    if (r == me) continue;
    r0 = edges[r];
    rows = edges[r + 1] - edges[r];
    if (my_rows > 0) q = N.Send(grid + (size_t)my0 * W, (size_t)my_rows * W, ncclInt8, r, comm, st);
    if (q == ncclSuccess && rows > 0) q = N.Recv(grid + (size_t)r0 * W, (size_t)rows * W, ncclInt8, r, comm, st);
  }
  if (q != ncclSuccess) {
    N.GroupEnd();
    return nccl_fail(q, "ncclSend/ncclRecv(band)");
  }
  q = N.GroupEnd();
  return q == ncclSuccess ? UQS_OK : nccl_fail(q, "ncclGroupEnd");
}

}  // namespace uqs

using namespace uqs;

extern "C" {

/* ---- partitions (host arithmetic; usable without a device) -------------------------------------------------- */

void uqs_flight_shard(int n_flights, int rank, int world, int* first, int* count) {
  if (world <= 0 || rank < 0 || rank >= world || n_flights < 0) { if (first) *first = 0; if (count) *count = 0; return; }
  const int base = n_flights / world, extra = n_flights % world;
  if (first) *first = rank * base + std::min(rank, extra);
  if (count) *count = base + (rank < extra ? 1 : 0);
}

void uqs_row_band(int H, int rank, int world, int align, int* row0, int* rows) {
  if (world <= 0 || rank < 0 || rank >= world || H < 0) { if (row0) *row0 = 0; if (rows) *rows = 0; return; }
  if (align < 1) align = 1;
  const int units = (H + align - 1) / align;
  const int base = units / world, extra = units % world;
  const int u0 = rank * base + std::min(rank, extra);
  const int u1 = u0 + base + (rank < extra ? 1 : 0);
  const int r0 = std::min(u0 * align, H), r1 = std::min(u1 * align, H);
  if (row0) *row0 = r0;
  if (rows) *rows = r1 - r0;
}

/* ---- (i) one process per GPU ------------------------------------------------------------------------------------ */

int uqs_comm_unique_id(void* id128) {
  int rc = nccl_load();
  if (rc) return rc;
  if (!id128) { set_error("uqs_comm_unique_id: NULL"); return UQS_ERR_BAD_ARG; }
  ncclUniqueId id;
  ncclResult_t q = N.GetUniqueId(&id);
  if (q != ncclSuccess) return nccl_fail(q, "ncclGetUniqueId");
  static_assert(sizeof(id) == UQS_COMM_ID_BYTES, "uqs_mapping.h states the id size");
  memcpy(id128, &id, sizeof(id));
  return UQS_OK;
}

int uqs_comm_init_rank(const void* id128, int nranks, int rank) {
  int rc = check_ready();
  if (rc) return rc;
  if ((rc = nccl_load())) return rc;
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) { set_error("uqs_comm_init_rank: bad argument"); return UQS_ERR_BAD_ARG; }
  comm_release();
  cudaSetDevice(g_ctx.device);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  ncclResult_t q = N.CommInitRank(&comm, nranks, id, rank);
  if (q != ncclSuccess) return nccl_fail(q, "ncclCommInitRank");
  g_ctx.comm = comm;
  g_ctx.comm_rank = rank;
  g_ctx.comm_nranks = nranks;
  return UQS_OK;
}

int uqs_comm_destroy(void) {
  if (!g_ctx.ready) return UQS_OK;
  cudaStreamSynchronize(g_ctx.stream());
  comm_release();
  return UQS_OK;
}

int uqs_comm_nranks(void) { return g_ctx.ready ? g_ctx.comm_nranks : 0; }

/* The cuts uqs_replay_banded_dev would use for `world` ranks on this log (device pointers): edges_out[0..world]. */
int uqs_balanced_row_bands_dev(const uqs_params* p, int n_frames, const float* x_dev, const float* y_dev, int world,
                               int* edges_out) {
  int rc = check_ready();
  if (rc) return rc;
  DevParams dp;
  if ((rc = make_dev_params(p, &dp))) return rc;
  if (n_frames <= 0 || !x_dev || !y_dev || world < 1 || world > kMaxDevices || !edges_out) { set_error("uqs_balanced_row_bands_dev: bad argument"); return UQS_ERR_BAD_ARG; }
  return band_edges(dp, n_frames, x_dev, y_dev, world, 4, edges_out);
}

int uqs_set_band_balance(int on) {
  g_balance = on != 0;
  return UQS_OK;
}

/* edges_out[0..nranks]: the row cuts of the last banded replay on the current context (rank r owned rows
 * [edges_out[r], edges_out[r+1])).  Returns the number of ranks. */
int uqs_band_edges(int* edges_out) {
  if (!g_ctx.ready || !edges_out) return 0;
  for (int r = 0; r <= g_ctx.comm_nranks; r++) edges_out[r] = g_ctx.band_edges[r];
  return g_ctx.comm_nranks;
}
int uqs_comm_rank(void) { return g_ctx.ready ? g_ctx.comm_rank : 0; }

int uqs_nccl_version(void) {
  if (nccl_load()) return 0;
  int v = 0;
  return N.GetVersion(&v) == ncclSuccess ? v : 0;
}

/* Config 4 on this rank: replay the rank's owned row band of ONE grid from a log every rank holds on its device,
 * then (gather != 0) the single exchange -- every rank ends up with the whole grid.  Asynchronous on the current
 * stream unless stats is given. */
int uqs_replay_banded_dev(const uqs_params* p, int n_frames, const float* x, const float* y, const float* yaw,
                          const float* ranges, int8_t* grid, int gather, uqs_stats* stats) {
  int rc = check_ready();
  if (rc) return rc;
  DevParams dp;
  if ((rc = make_dev_params(p, &dp))) return rc;
  if (n_frames <= 0 || !x || !y || !yaw || !ranges || !grid) { set_error("uqs_replay_banded_dev: NULL pointer or non-positive size"); return UQS_ERR_BAD_ARG; }
  if (g_ctx.comm_nranks > 1 && !g_ctx.comm) { set_error("uqs_replay_banded_dev: no communicator (uqs_comm_init_rank)"); return UQS_ERR_NOT_INIT; }
  int* edges = g_ctx.band_edges;
  if ((rc = band_edges(dp, n_frames, x, y, g_ctx.comm_nranks, 4, edges))) return rc;
  const int r0 = edges[g_ctx.comm_rank], rows = edges[g_ctx.comm_rank + 1] - r0;
  if (rows > 0 && (rc = replay_device(dp, 1, n_frames, x, y, yaw, ranges, nullptr, grid, 0, r0, rows, true))) return rc;
  if (gather && (rc = gather_bands(g_ctx, edges, grid, p->W, p->H, g_ctx.stream()))) return rc;
  // every rank's ray set-up sees the whole log, so the counters are the whole log's on every rank
  if (rows > 0) return fetch_stats(stats, (uint64_t)n_frames);
  if (stats) memset(stats, 0, sizeof(*stats));       // more ranks than 4-row units: this rank owns nothing
  return UQS_OK;
}

/* Host-buffer form for model (i): every rank passes the same log.  Rank r uploads only its 1/N slice over PCIe;
 * the slices are all-gathered over NVLink; then as uqs_replay_banded_dev with the gather; grid_out (may be NULL
 * on ranks that do not want the result) receives the whole grid. */
int uqs_replay_banded(const uqs_params* p, int n_frames, const float* x, const float* y, const float* yaw,
                      const float* ranges, int8_t* grid_out, uqs_stats* stats) {
  int rc = check_ready();
  if (rc) return rc;
  DevParams dp;
  if ((rc = make_dev_params(p, &dp))) return rc;
  if (n_frames <= 0 || !x || !y || !yaw || !ranges) { set_error("uqs_replay_banded: NULL pointer or non-positive size"); return UQS_ERR_BAD_ARG; }
  const int n = g_ctx.comm_nranks, me = g_ctx.comm_rank;
  if (n > 1 && !g_ctx.comm) { set_error("uqs_replay_banded: no communicator (uqs_comm_init_rank)"); return UQS_ERR_NOT_INIT; }
  cudaStream_t st = g_ctx.stream();
  const size_t per = ((size_t)n_frames + n - 1) / n;               // frames per slice (the last one may be short)
  const size_t padded = per * n;
  if ((rc = g_ctx.in_x.ensure(padded * 4)) || (rc = g_ctx.in_y.ensure(padded * 4)) || (rc = g_ctx.in_yaw.ensure(padded * 4)) ||
      (rc = g_ctx.in_ranges.ensure(padded * 128)) || (rc = g_ctx.out_grids.ensure((size_t)p->W * p->H)))
    return rc;
  const size_t f0 = std::min(per * me, (size_t)n_frames), nf = std::min(per, (size_t)n_frames - f0);
  struct Arr { DevBuf* b; const float* src; size_t elems; } arrs[4] = {
    { &g_ctx.in_x, x, 1 }, { &g_ctx.in_y, y, 1 }, { &g_ctx.in_yaw, yaw, 1 }, { &g_ctx.in_ranges, ranges, 32 } };
  cudaError_t e = cudaSuccess;
  for (auto& a : arrs)
    if (e == cudaSuccess && nf)
      e = cudaMemcpyAsync((float*)a.b->p + f0 * a.elems, a.src + f0 * a.elems, nf * a.elems * 4, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return cuda_fail(e, "log slice H2D");
  if (n > 1) {
    ncclResult_t q = N.GroupStart();
    for (auto& a : arrs)
      if (q == ncclSuccess)
        q = N.AllGather((float*)a.b->p + per * me * a.elems, a.b->p, per * a.elems, ncclFloat32, (ncclComm_t)g_ctx.comm, st);
    if (q != ncclSuccess) { N.GroupEnd(); return nccl_fail(q, "ncclAllGather(log)"); }
    if ((q = N.GroupEnd()) != ncclSuccess) return nccl_fail(q, "ncclGroupEnd");
  }
  int8_t* grid = (int8_t*)g_ctx.out_grids.p;
  uqs_stats local;
  if ((rc = uqs_replay_banded_dev(p, n_frames, (float*)g_ctx.in_x.p, (float*)g_ctx.in_y.p, (float*)g_ctx.in_yaw.p,
                                  (float*)g_ctx.in_ranges.p, grid, 1, stats ? stats : &local)))
    return rc;
  if (grid_out) {
    e = cudaMemcpyAsync(grid_out, grid, (size_t)p->W * p->H, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return cuda_fail(e, "grid D2H");
  }
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "uqs_replay_banded sync");
  return UQS_OK;
}

/* ---- (ii) one host thread, N devices -------------------------------------------------------------------------- */

int uqs_multi_init(int n_devices, const int* devices) {
  int rc = nccl_load();
  if (n_devices > 1 && rc) return rc;
  if (n_devices < 1 || n_devices > kMaxDevices) { set_error("uqs_multi_init: 1..%d devices", kMaxDevices); return UQS_ERR_BAD_ARG; }
  uqs_multi_shutdown();
  int devs[kMaxDevices];
  for (int i = 0; i < n_devices; i++) devs[i] = devices ? devices[i] : i;
  for (int i = 0; i < n_devices; i++) {
    if ((rc = context_init(g_multi[i], devs[i]))) {
      g_n_multi = i;
      uqs_multi_shutdown();
      return rc;
    }
  }
  g_n_multi = n_devices;
  if (n_devices > 1) {
    ncclComm_t comms[kMaxDevices];
    ncclResult_t q = N.CommInitAll(comms, n_devices, devs);
    if (q != ncclSuccess) { uqs_multi_shutdown(); return nccl_fail(q, "ncclCommInitAll"); }
    for (int i = 0; i < n_devices; i++) {
      g_multi[i].comm = comms[i];
      g_multi[i].comm_rank = i;
      g_multi[i].comm_nranks = n_devices;
    }
  }
  return uqs_multi_select(0);
}

int uqs_multi_count(void) { return g_n_multi; }

int uqs_multi_select(int i) {
  if (i < 0 || i >= g_n_multi) { set_error("uqs_multi_select: %d outside 0..%d", i, g_n_multi - 1); return UQS_ERR_BAD_ARG; }
  g_cur = &g_multi[i];
  cudaError_t e = cudaSetDevice(g_cur->device);
  return e == cudaSuccess ? UQS_OK : cuda_fail(e, "cudaSetDevice");
}

void uqs_multi_shutdown(void) {
  for (int i = 0; i < g_n_multi; i++) {
    g_cur = &g_multi[i];
    BandBufs& B = g_bands[i];
    if (g_multi[i].ready) {
      cudaSetDevice(g_multi[i].device);
      DevBuf* all[] = { &B.x, &B.y, &B.yaw, &B.ranges, &B.grid };
      for (DevBuf* b : all) b->release();
    }
    context_shutdown(g_multi[i]);
    g_multi[i] = Context();
  }
  g_n_multi = 0;
  context_select_single();
}

/* Config 4 from one host thread: the log is cut into N slices, device i uploads slice i (N PCIe links in
 * parallel), the slices are all-gathered over NVLink, every device replays its owned row band, the bands are
 * all-gathered (the path's single grid exchange), and device i copies band i back into grid_out (again N links).
 * Nothing blocks between devices: every step is enqueued on all devices before the first wait. */
int uqs_multi_replay_banded(const uqs_params* p, int n_frames, const float* x, const float* y, const float* yaw,
                            const float* ranges, int8_t* grid_out, uqs_stats* stats) {
  const int n = g_n_multi;
  if (n < 1) { set_error("uqs_multi_replay_banded: uqs_multi_init() first"); return UQS_ERR_NOT_INIT; }
  DevParams dp;
  int rc = make_dev_params(p, &dp);
  if (rc) return rc;
  if (n_frames <= 0 || !x || !y || !yaw || !ranges || !grid_out) { set_error("uqs_multi_replay_banded: NULL pointer or non-positive size"); return UQS_ERR_BAD_ARG; }
  const size_t per = ((size_t)n_frames + n - 1) / n, padded = per * n;
  const size_t cells = (size_t)p->W * p->H;
  cudaError_t e = cudaSuccess;
  // 1. slices up
  for (int i = 0; i < n; i++) {
    if ((rc = uqs_multi_select(i))) return rc;
    BandBufs& B = g_bands[i];
    if ((rc = B.x.ensure(padded * 4)) || (rc = B.y.ensure(padded * 4)) || (rc = B.yaw.ensure(padded * 4)) ||
        (rc = B.ranges.ensure(padded * 128)) || (rc = B.grid.ensure(cells)))
      return rc;
    const size_t f0 = std::min(per * i, (size_t)n_frames), nf = std::min(per, (size_t)n_frames - f0);
    cudaStream_t st = g_ctx.stream();
    if (nf) {
      e = cudaMemcpyAsync((float*)B.x.p + f0, x + f0, nf * 4, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync((float*)B.y.p + f0, y + f0, nf * 4, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync((float*)B.yaw.p + f0, yaw + f0, nf * 4, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync((float*)B.ranges.p + f0 * 32, ranges + f0 * 32, nf * 128, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) return cuda_fail(e, "log slice H2D");
    }
  }
  // 2. log all-gather over NVLink
  if (n > 1) {
    ncclResult_t q = N.GroupStart();
    for (int i = 0; i < n && q == ncclSuccess; i++) {
      BandBufs& B = g_bands[i];
      cudaStream_t st = g_multi[i].stream();
      ncclComm_t comm = (ncclComm_t)g_multi[i].comm;
      struct { DevBuf* b; size_t elems; } arrs[4] = { { &B.x, 1 }, { &B.y, 1 }, { &B.yaw, 1 }, { &B.ranges, 32 } };
      for (auto& a : arrs)
        if (q == ncclSuccess) q = N.AllGather((float*)a.b->p + per * i * a.elems, a.b->p, per * a.elems, ncclFloat32, comm, st);
    }
    if (q != ncclSuccess) { N.GroupEnd(); return nccl_fail(q, "ncclAllGather(log)"); }
    if ((q = N.GroupEnd()) != ncclSuccess) return nccl_fail(q, "ncclGroupEnd");
  }
  // 3. owned bands (cuts from the log on device 0; the same on every device)
  int edges[kMaxDevices + 1];
  if ((rc = uqs_multi_select(0))) return rc;
  if ((rc = band_edges(dp, n_frames, (float*)g_bands[0].x.p, (float*)g_bands[0].y.p, n, 4, edges))) return rc;
  for (int i = 0; i < n; i++) memcpy(g_multi[i].band_edges, edges, sizeof(int) * (n + 1));
  for (int i = 0; i < n; i++) {
    if ((rc = uqs_multi_select(i))) return rc;
    BandBufs& B = g_bands[i];
    const int r0 = edges[i], rows = edges[i + 1] - edges[i];
    if (rows > 0 && (rc = replay_device(dp, 1, n_frames, (float*)B.x.p, (float*)B.y.p, (float*)B.yaw.p, (float*)B.ranges.p, nullptr,
                                        (int8_t*)B.grid.p, 0, r0, rows, true)))
      return rc;
  }
  // 4. the grid exchange
  if (n > 1) {
    ncclResult_t q = N.GroupStart();
    if (q != ncclSuccess) return nccl_fail(q, "ncclGroupStart");
    for (int i = 0; i < n; i++) {
      if ((rc = gather_bands(g_multi[i], edges, (int8_t*)g_bands[i].grid.p, p->W, p->H, g_multi[i].stream()))) { N.GroupEnd(); return rc; }
    }
    if ((q = N.GroupEnd()) != ncclSuccess) return nccl_fail(q, "ncclGroupEnd");
  }
  // 5. bands down, one link each
  for (int i = 0; i < n; i++) {
    if ((rc = uqs_multi_select(i))) return rc;
    const int r0 = edges[i], rows = edges[i + 1] - edges[i];
    if (rows <= 0) continue;
    e = cudaMemcpyAsync(grid_out + (size_t)r0 * p->W, (int8_t*)g_bands[i].grid.p + (size_t)r0 * p->W, (size_t)rows * p->W,
                        cudaMemcpyDeviceToHost, g_ctx.stream());
    if (e != cudaSuccess) return cuda_fail(e, "band D2H");
  }
  for (int i = 0; i < n; i++) {
    if ((rc = uqs_multi_select(i))) return rc;
    if ((e = cudaStreamSynchronize(g_ctx.stream())) != cudaSuccess) return cuda_fail(e, "uqs_multi_replay_banded sync");
  }
  rc = uqs_multi_select(0);
  if (rc) return rc;
  return fetch_stats(stats, (uint64_t)n_frames);
}

/* device pointer of device i's copy of the whole grid after uqs_multi_replay_banded (for device-side consumers) */
const int8_t* uqs_multi_grid_dev(int i) { return (i >= 0 && i < g_n_multi) ? (const int8_t*)g_bands[i].grid.p : nullptr; }

}  // extern "C"
