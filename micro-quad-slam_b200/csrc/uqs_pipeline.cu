// uqs_pipeline.cu -- host-buffer entry points (uqs_replay, uqs_replay_flow).
//
// The caller's logs live in host memory; the work is cut into chunks of flights whose three steps run on
// separate streams so that PCIe traffic hides behind the kernels:
//     H2D(c+1..c+3)  ||  P0 + ray set-up + replay (c, c+1 on two alternating streams)  ||  D2H(c-1)
// Four staging buffers rotate: the copy engine runs up to three chunks ahead of the kernels (with two, the
// H2D of chunk c+1 had to wait for the kernels of chunk c-1 and the copy engine idled a quarter of the time).
// Copies overlap only from page-locked host memory (cudaHostRegister / cudaHostAlloc / torch
// pin_memory); pageable buffers still work, serialised by the driver.
#include <algorithm>
#include <cstring>
#include <vector>

#include "uqs_host.h"

namespace uqs {

namespace {

constexpr int kStages = 4;

struct Stage {                 // device staging of one chunk (kStages of them rotate)
  DevBuf t, rx, ry, h, yaw, q, x, y, ranges, grids, packed, boxes;
  cudaEvent_t in_ready = nullptr, computed = nullptr, out_done = nullptr;
};

struct Pipeline {
  bool made = false;
  cudaStream_t s_in = nullptr, s_out = nullptr, s_cmp[2] = { nullptr, nullptr };
  Stage st[kStages];
};

// one pipeline per device context, created on first use
Pipeline& pipe() {
  if (!g_ctx.pipeline) g_ctx.pipeline = new Pipeline();
  return *static_cast<Pipeline*>(g_ctx.pipeline);
}
#define P pipe()

int pipeline_init() {
  if (P.made) return UQS_OK;
  cudaError_t e = cudaStreamCreateWithFlags(&P.s_in, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&P.s_out, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&P.s_cmp[0], cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&P.s_cmp[1], cudaStreamNonBlocking);
  for (int i = 0; i < kStages && e == cudaSuccess; i++) {
    e = cudaEventCreateWithFlags(&P.st[i].in_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P.st[i].computed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P.st[i].out_done, cudaEventDisableTiming);
  }
  if (e != cudaSuccess) return cuda_fail(e, "pipeline_init");
  P.made = true;
  return UQS_OK;
}

}  // namespace

void pipeline_release() {
  if (!g_ctx.pipeline) return;
  struct Drop { ~Drop() { delete static_cast<Pipeline*>(g_ctx.pipeline); g_ctx.pipeline = nullptr; } } drop;
  if (!P.made) return;
  for (auto& s : P.st) {
    DevBuf* all[] = { &s.t, &s.rx, &s.ry, &s.h, &s.yaw, &s.q, &s.x, &s.y, &s.ranges, &s.grids, &s.packed, &s.boxes };
    for (DevBuf* b : all) b->release();
    cudaEventDestroy(s.in_ready);
    cudaEventDestroy(s.computed);
    cudaEventDestroy(s.out_done);
  }
  cudaStreamDestroy(P.s_in);
  cudaStreamDestroy(P.s_out);
  cudaStreamDestroy(P.s_cmp[0]);
  cudaStreamDestroy(P.s_cmp[1]);
}

// flow != nullptr: P0 from flow samples (t_ms, rate_x, rate_y, h, q) then replay; else poses x,y given.
struct HostLogs {
  const uint32_t* t_ms; const float *rx, *ry, *h; const uint8_t* q;     // flow form
  const float *x, *y;                                                   // pose form
  const float* yaw; const float* ranges;
  float *pox, *poy;                                                     // optional pose output (flow form)
  const uint16_t* ranges_mm = nullptr;                                  // ranges as u16 millimetres instead of float metres
  // boxed output (instead of grids_out): per flight its touched box and the box's cells
  int32_t* boxes_out = nullptr; uint64_t* offsets_out = nullptr; int8_t* packed_out = nullptr;
  size_t packed_cap = 0; size_t* packed_bytes = nullptr;
};

// Copies each flight's box [x0,x1) x [y0,y1) of its W x H grid into a slot of `slot` bytes (row pitch x1 - x0).
__global__ void k_pack_boxes(const int8_t* __restrict__ grids, const int4* __restrict__ boxes, int W, int H, size_t slot,
                             int8_t* __restrict__ packed) {
  const int f = blockIdx.x;
  const int4 b = boxes[f];
  const int bw = b.z - b.x, bh = b.w - b.y;
  const int8_t* g = grids + (size_t)f * W * H + (size_t)b.y * W + b.x;
  int8_t* out = packed + (size_t)f * slot;
  if (((bw | b.x | W) & 3) == 0) {
    const int wpr = bw >> 2;
    for (int i = threadIdx.x; i < wpr * bh; i += blockDim.x) {
      const int r = i / wpr, c = i - r * wpr;
      reinterpret_cast<uint32_t*>(out)[i] = reinterpret_cast<const uint32_t*>(g + (size_t)r * W)[c];
    }
  } else {
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) {
      const int r = i / bw, c = i - r * bw;
      out[i] = g[(size_t)r * W + c];
    }
  }
}

__global__ void k_whole_boxes(int n, int W, int H, int4* boxes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) boxes[i] = make_int4(0, 0, W, H);
}

static int host_pipeline_run(const uqs_params* p, const DevParams& dp, int n_flights, int n_frames, const HostLogs& L,
                             int8_t* grids_out, uqs_stats* stats);

int host_pipeline(const uqs_params* p, const DevParams& dp, int n_flights, int n_frames, const HostLogs& L,
                  int8_t* grids_out, uqs_stats* stats) {
  const int rc = host_pipeline_run(p, dp, n_flights, n_frames, L, grids_out, stats);
  g_ctx.chip_shared = false;
  if (rc && P.made) {
    // an early return must not leave copies from / into the caller's buffers in flight
    for (cudaStream_t s : { P.s_in, P.s_cmp[0], P.s_cmp[1], P.s_out }) cudaStreamSynchronize(s);
    cudaGetLastError();
  }
  return rc;
}

static int host_pipeline_run(const uqs_params* p, const DevParams& dp_in, int n_flights, int n_frames, const HostLogs& L,
                             int8_t* grids_out, uqs_stats* stats) {
  int rc = pipeline_init();
  if (rc) return rc;
  DevParams dp = dp_in;
  dp.ranges_u16 = L.ranges_mm ? 1 : 0;
  const bool boxed = L.boxes_out != nullptr;
  const size_t rbytes = L.ranges_mm ? 64 : 128;                      // range bytes per frame
  size_t packed_used = 0;
  if (L.packed_bytes) *L.packed_bytes = 0;
  const bool flow = L.t_ms != nullptr;
  const size_t cells = (size_t)p->W * p->H;
  // Chunk schedule: chunks of 2 flights per SM between a first and a last chunk of 1 per SM (the first chunk's H2D and
  // the last chunk's kernels and D2H hide behind nothing).  The copies, not the kernels, bound this pipeline on the
  // 400x400 ensemble (1.83 GB in at ~51 GB/s while grids flow out), so small chunks -- a short tail after the last
  // byte has arrived -- win: measured 39.8 ms at 2 per SM, 42 ms at 3, 43 ms at 4.
  std::vector<int> starts;                      // flight index where chunk c begins; starts.back() = n_flights
  {
    const int sm = g_ctx.sm_count;
    int chunk_cap = g_ctx.host_chunk > 0 ? g_ctx.host_chunk : 2 * sm;
    const size_t stage_budget = (size_t)6 << 30;                          // bound the staging buffers (kStages x chunk)
    const size_t per_flight = (size_t)n_frames * (24 + rbytes) + cells * (boxed ? 2 : 1);     // log bytes + the flight's grid(s)
    if ((size_t)chunk_cap * per_flight * kStages > stage_budget)
      chunk_cap = std::max<int>(1, (int)(stage_budget / (per_flight * kStages)));
    if (g_ctx.host_chunk == 0 && chunk_cap == 2 * sm && n_flights >= 8 * sm) {
      starts.push_back(0);
      const int body = n_flights - 2 * sm, n_mid = (body + chunk_cap - 1) / chunk_cap;
      for (int i = 0; i < n_mid; i++) starts.push_back(sm + (int)((long long)body * i / n_mid));
      starts.push_back(n_flights - sm);
    } else {
      const int n_want = (n_flights + chunk_cap - 1) / chunk_cap;
      for (int i = 0; i < n_want; i++) starts.push_back((int)((long long)n_flights * i / n_want));
    }
    starts.push_back(n_flights);
  }
  const int n_chunks = (int)starts.size() - 1;
  g_ctx.chip_shared = n_chunks > 1;             // engine and warps-per-CTA choice: neighbours fill the chip
  int chunk = 0;                                // largest chunk: size of the staging buffers
  for (int c = 0; c < n_chunks; c++) chunk = std::max(chunk, starts[c + 1] - starts[c]);
  cudaError_t e = cudaSuccess;
  // everything the caller queued on its stream happens before the pipeline starts
  cudaEvent_t ev_start = P.st[0].out_done;            // any event will do before the first D2H uses it
  if ((e = cudaEventRecord(ev_start, g_ctx.user_stream())) != cudaSuccess) return cuda_fail(e, "record start");
  for (cudaStream_t s : { P.s_in, P.s_cmp[0], P.s_cmp[1] })
    if ((e = cudaStreamWaitEvent(s, ev_start, 0)) != cudaSuccess) return cuda_fail(e, "wait start");
  struct WorkScope {                                   // kernels of chunk c run on s_cmp[c&1] with works[1 + (c&1)]
    explicit WorkScope(int i) { g_ctx.works[1 + i].stream = P.s_cmp[i]; g_ctx.w = &g_ctx.works[1 + i]; }
    ~WorkScope() { g_ctx.w = &g_ctx.works[0]; }
  };

  for (int i = 0; i < std::min(kStages, n_chunks); i++) {
    Stage& S = P.st[i];
    const size_t n = (size_t)chunk * n_frames;
    if (flow && ((rc = S.t.ensure(n * 4)) || (rc = S.rx.ensure(n * 4)) || (rc = S.ry.ensure(n * 4)) ||
                 (rc = S.h.ensure(n * 4)) || (rc = S.q.ensure(n))))
      return rc;
    if ((rc = S.yaw.ensure(n * 4)) || (rc = S.x.ensure(n * 4)) || (rc = S.y.ensure(n * 4)) ||
        (rc = S.ranges.ensure(n * 128)) || (rc = S.grids.ensure((size_t)chunk * cells)))
      return rc;
  }

  auto upload = [&](int c) -> cudaError_t {
    Stage& S = P.st[c % kStages];
    const int f0 = starts[c], nf = starts[c + 1] - f0;
    const size_t o = (size_t)f0 * n_frames, n = (size_t)nf * n_frames;
    cudaError_t r = cudaSuccess;
    if (c >= kStages) r = cudaStreamWaitEvent(P.s_in, S.computed, 0);         // inputs of chunk c-kStages consumed
    KernelTimer t_in(3, P.s_in);
    auto cp = [&](DevBuf& b, const void* src, size_t bytes) {
      if (r == cudaSuccess) r = cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, P.s_in);
    };
    if (flow) {
      cp(S.t, L.t_ms + o, n * 4); cp(S.rx, L.rx + o, n * 4); cp(S.ry, L.ry + o, n * 4);
      cp(S.h, L.h + o, n * 4); cp(S.q, L.q + o, n);
    } else {
      cp(S.x, L.x + o, n * 4); cp(S.y, L.y + o, n * 4);
    }
    cp(S.yaw, L.yaw + o, n * 4);
    if (L.ranges_mm) cp(S.ranges, L.ranges_mm + o * 32, n * 64);
    else cp(S.ranges, L.ranges + o * 32, n * 128);
    t_in.stop();
    if (r == cudaSuccess) r = cudaEventRecord(S.in_ready, P.s_in);
    return r;
  };

  // upload(c') waits for the kernels of chunk c'-kStages, so it can only be queued once those are: the copies
  // of chunks 0..kStages-2 go first, then iteration c queues the copy of chunk c+kStages-1
  for (int c = 0; c < std::min(kStages - 1, n_chunks); c++)
    if ((e = upload(c)) != cudaSuccess) return cuda_fail(e, "H2D");
  for (int c = 0; c < n_chunks; c++) {
    Stage& S = P.st[c % kStages];
    const int f0 = starts[c], nf = starts[c + 1] - f0;
    const size_t o = (size_t)f0 * n_frames, n = (size_t)nf * n_frames;
    if (c + kStages - 1 < n_chunks && (e = upload(c + kStages - 1)) != cudaSuccess) return cuda_fail(e, "H2D");
    WorkScope scope(c & 1);
    cudaStream_t sc = P.s_cmp[c & 1];
    if ((e = cudaStreamWaitEvent(sc, S.in_ready, 0)) != cudaSuccess) return cuda_fail(e, "wait H2D");
    if (c >= kStages && (e = cudaStreamWaitEvent(sc, S.out_done, 0)) != cudaSuccess) return cuda_fail(e, "wait D2H");   // grid buffer free
    if (g_ctx.copy_only) {
      // copies and their ordering only
    } else if (flow) {
      if ((rc = pose_device(nf, n_frames, (uint32_t*)S.t.p, (float*)S.rx.p, (float*)S.ry.p, (float*)S.h.p, (float*)S.yaw.p,
                            (uint8_t*)S.q.p, (float*)S.x.p, (float*)S.y.p, 0)))
        return rc;
    }
    if (!g_ctx.copy_only &&
        (rc = replay_device(dp, nf, n_frames, (float*)S.x.p, (float*)S.y.p, (float*)S.yaw.p, (float*)S.ranges.p, nullptr,
                            (int8_t*)S.grids.p, 0, 0, p->H, c < 2)))
      return rc;
    size_t slot = 0, out_at = 0;
    if (boxed && !g_ctx.copy_only) {
      // The chunk's touched boxes (k_flight_boxes ran inside replay_device when the resident engine was a candidate; its
      // maxima came back through mapped memory before the replay was launched) -> one slot of max_w x max_h bytes per
      // flight, so the size of the D2H is known here without another synchronisation.
      const Work& wk = *g_ctx.w;
      if ((rc = S.boxes.ensure((size_t)nf * sizeof(int4)))) return rc;
      int4* bx = (int4*)S.boxes.p;
      if (wk.boxes_valid && wk.boxes_n == nf) {
        slot = (size_t)std::max(wk.box_w, 1) * std::max(wk.box_h, 1);
        e = cudaMemcpyAsync(bx, wk.boxes.p, (size_t)nf * sizeof(int4), cudaMemcpyDeviceToDevice, sc);
      } else {
        slot = cells;
        k_whole_boxes<<<(unsigned)((nf + 127) / 128), 128, 0, sc>>>(nf, p->W, p->H, bx);
        e = cudaGetLastError();
        g_ctx.launches += 1;
      }
      slot = (slot + 15) & ~(size_t)15;
      if (e == cudaSuccess && (rc = S.packed.ensure((size_t)nf * slot))) return rc;
      if (e == cudaSuccess) {
        k_pack_boxes<<<(unsigned)nf, 256, 0, sc>>>((const int8_t*)S.grids.p, bx, p->W, p->H, slot, (int8_t*)S.packed.p);
        e = cudaGetLastError();
        g_ctx.launches += 1;
      }
      if (e != cudaSuccess) return cuda_fail(e, "k_pack_boxes");
      out_at = packed_used;
      packed_used += (size_t)nf * slot;
      if (L.packed_bytes) *L.packed_bytes = packed_used;
      if (packed_used > L.packed_cap) {
        set_error("boxed output needs more than the %zu bytes given (%zu so far at flight %d of %d)", L.packed_cap, packed_used, f0 + nf, n_flights);
        return UQS_ERR_NOMEM;
      }
      for (int i = 0; i < nf; i++) L.offsets_out[f0 + i] = out_at + (size_t)i * slot;
    }
    if ((e = cudaEventRecord(S.computed, sc)) != cudaSuccess) return cuda_fail(e, "record");
    if ((e = cudaStreamWaitEvent(P.s_out, S.computed, 0)) != cudaSuccess) return cuda_fail(e, "wait compute");
    KernelTimer t_out(4, P.s_out);
    if (boxed) {
      e = cudaSuccess;
      if (!g_ctx.copy_only) {
        e = cudaMemcpyAsync(L.packed_out + out_at, S.packed.p, (size_t)nf * slot, cudaMemcpyDeviceToHost, P.s_out);
        if (e == cudaSuccess) e = cudaMemcpyAsync(L.boxes_out + (size_t)f0 * 4, S.boxes.p, (size_t)nf * sizeof(int4), cudaMemcpyDeviceToHost, P.s_out);
      }
    } else {
      e = cudaMemcpyAsync(grids_out + (size_t)f0 * cells, S.grids.p, (size_t)nf * cells, cudaMemcpyDeviceToHost, P.s_out);
    }
    if (e == cudaSuccess && flow && L.pox && L.poy) {
      e = cudaMemcpyAsync(L.pox + o, S.x.p, n * 4, cudaMemcpyDeviceToHost, P.s_out);
      if (e == cudaSuccess) e = cudaMemcpyAsync(L.poy + o, S.y.p, n * 4, cudaMemcpyDeviceToHost, P.s_out);
    }
    t_out.stop();
    if (e == cudaSuccess) e = cudaEventRecord(S.out_done, P.s_out);
    if (e != cudaSuccess) return cuda_fail(e, "D2H");
  }
  // the call is synchronous: results are in the caller's buffers when it returns
  if ((e = cudaStreamSynchronize(P.s_out)) != cudaSuccess) return cuda_fail(e, "D2H sync");
  if ((e = cudaStreamSynchronize(P.s_in)) != cudaSuccess) return cuda_fail(e, "H2D sync");
  if (g_ctx.copy_only) return UQS_OK;
  uqs_stats local;
  return fetch_stats_mask(stats ? stats : &local, (uint64_t)n_flights * n_frames, n_chunks > 1 ? 6u : 2u);
}

}  // namespace uqs

using namespace uqs;

extern "C" {

int uqs_replay(const uqs_params* p, int n_flights, int n_frames, const float* x, const float* y,
               const float* yaw, const float* ranges, int8_t* grids_out, uqs_stats* stats) {
  int rc = check_ready();
  if (rc) return rc;
  DevParams dp;
  if ((rc = make_dev_params(p, &dp))) return rc;
  if (n_flights <= 0 || n_frames <= 0 || !x || !y || !yaw || !ranges || !grids_out) {
    set_error("uqs_replay: NULL pointer or non-positive size");
    return UQS_ERR_BAD_ARG;
  }
  HostLogs L = { nullptr, nullptr, nullptr, nullptr, nullptr, x, y, yaw, ranges, nullptr, nullptr };
  return host_pipeline(p, dp, n_flights, n_frames, L, grids_out, stats);
}

int uqs_replay_flow(const uqs_params* p, int n_flights, int n_samples, const uint32_t* t_ms,
                    const float* rx, const float* ry, const float* h, const float* yaw,
                    const uint8_t* q, const float* ranges, int8_t* grids_out, float* pox, float* poy,
                    uqs_stats* stats) {
  int rc = check_ready();
  if (rc) return rc;
  DevParams dp;
  if ((rc = make_dev_params(p, &dp))) return rc;
  if (n_flights <= 0 || n_samples <= 0 || !t_ms || !rx || !ry || !h || !yaw || !q || !ranges || !grids_out) {
    set_error("uqs_replay_flow: NULL pointer or non-positive size");
    return UQS_ERR_BAD_ARG;
  }
  HostLogs L = { t_ms, rx, ry, h, q, nullptr, nullptr, yaw, ranges, pox, poy };
  return host_pipeline(p, dp, n_flights, n_samples, L, grids_out, stats);
}

/* uqs_replay_flow with the ranges as u16 millimetres (0xFFFF = no return), the unit the ToF sensors deliver
 * (uav_local_nav.c:1327-1328): converted on the device with the reference's own (float)mm * 0.001f.  Halves the
 * host-to-device traffic of a log. */
int uqs_replay_flow_mm(const uqs_params* p, int n_flights, int n_samples, const uint32_t* t_ms,
                       const float* rx, const float* ry, const float* h, const float* yaw,
                       const uint8_t* q, const uint16_t* ranges_mm, int8_t* grids_out, float* pox, float* poy,
                       uqs_stats* stats) {
  int rc = check_ready();
  if (rc) return rc;
  DevParams dp;
  if ((rc = make_dev_params(p, &dp))) return rc;
  if (n_flights <= 0 || n_samples <= 0 || !t_ms || !rx || !ry || !h || !yaw || !q || !ranges_mm || !grids_out) {
    set_error("uqs_replay_flow_mm: NULL pointer or non-positive size");
    return UQS_ERR_BAD_ARG;
  }
  HostLogs L = { t_ms, rx, ry, h, q, nullptr, nullptr, yaw, nullptr, pox, poy };
  L.ranges_mm = ranges_mm;
  return host_pipeline(p, dp, n_flights, n_samples, L, grids_out, stats);
}

/* uqs_replay_flow with BOXED output: instead of n_flights dense W x H grids the call returns, per flight, the
 * bounding box of every cell the flight touched (boxes_out[f] = x0, y0, x1, y1, exclusive upper corner) and the
 * box's cells, row-major with row pitch x1 - x0, at packed_out + offsets_out[f].  Every cell outside the box is 0
 * (uav_local_nav.c:2190 and no update).  ranges (float metres) or ranges_mm (u16 millimetres): exactly one non-NULL.
 * packed_cap = bytes available at packed_out (n_flights * W * H always suffices); *packed_bytes = bytes used
 * (UQS_ERR_NOMEM if more were needed).  uqs_unpack_boxed() expands the result to dense grids. */
int uqs_replay_flow_boxed(const uqs_params* p, int n_flights, int n_samples, const uint32_t* t_ms,
                          const float* rx, const float* ry, const float* h, const float* yaw, const uint8_t* q,
                          const float* ranges, const uint16_t* ranges_mm, int32_t* boxes_out, uint64_t* offsets_out,
                          int8_t* packed_out, size_t packed_cap, size_t* packed_bytes, float* pox, float* poy,
                          uqs_stats* stats) {
  int rc = check_ready();
  if (rc) return rc;
  DevParams dp;
  if ((rc = make_dev_params(p, &dp))) return rc;
  if (n_flights <= 0 || n_samples <= 0 || !t_ms || !rx || !ry || !h || !yaw || !q || (!ranges == !ranges_mm) || !boxes_out ||
      !offsets_out || !packed_out) {
    set_error("uqs_replay_flow_boxed: NULL pointer, non-positive size, or not exactly one of ranges / ranges_mm");
    return UQS_ERR_BAD_ARG;
  }
  HostLogs L = { t_ms, rx, ry, h, q, nullptr, nullptr, yaw, ranges, pox, poy };
  L.ranges_mm = ranges_mm;
  L.boxes_out = boxes_out; L.offsets_out = offsets_out; L.packed_out = packed_out;
  L.packed_cap = packed_cap; L.packed_bytes = packed_bytes;
  return host_pipeline(p, dp, n_flights, n_samples, L, nullptr, stats);
}

/* Host helper: dense grids [n_flights][H][W] from the boxed form (plain host code). */
int uqs_unpack_boxed(const uqs_params* p, int n_flights, const int32_t* boxes, const uint64_t* offsets,
                     const int8_t* packed, int8_t* grids_out) {
  if (!p || n_flights <= 0 || !boxes || !offsets || !packed || !grids_out) { set_error("uqs_unpack_boxed: bad argument"); return UQS_ERR_BAD_ARG; }
  const size_t cells = (size_t)p->W * p->H;
  memset(grids_out, 0, cells * (size_t)n_flights);
  for (int f = 0; f < n_flights; f++) {
    const int32_t* b = boxes + (size_t)f * 4;
    const int bw = b[2] - b[0], bh = b[3] - b[1];
    if (bw <= 0 || bh <= 0) continue;
    if (b[0] < 0 || b[1] < 0 || b[2] > p->W || b[3] > p->H) { set_error("uqs_unpack_boxed: box %d outside the grid", f); return UQS_ERR_BAD_ARG; }
    for (int r = 0; r < bh; r++)
      memcpy(grids_out + (size_t)f * cells + (size_t)(b[1] + r) * p->W + b[0], packed + offsets[f] + (size_t)r * bw, (size_t)bw);
  }
  return UQS_OK;
}

int uqs_set_copy_only(int on) {
  int rc = check_ready();
  if (rc) return rc;
  g_ctx.copy_only = on != 0;
  return UQS_OK;
}

/* Flights per chunk of the host-buffer pipeline (0 = automatic). */
int uqs_set_host_chunk(int flights) {
  if (flights < -1) { set_error("negative chunk"); return UQS_ERR_BAD_ARG; }
  g_ctx.host_chunk = flights;
  return UQS_OK;
}

}  // extern "C"
