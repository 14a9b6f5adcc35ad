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
#include <vector>

#include "uqs_host.h"

namespace uqs {

namespace {

constexpr int kStages = 4;

struct Stage {                 // device staging of one chunk (kStages of them rotate)
  DevBuf t, rx, ry, h, yaw, q, x, y, ranges, grids;
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
    DevBuf* all[] = { &s.t, &s.rx, &s.ry, &s.h, &s.yaw, &s.q, &s.x, &s.y, &s.ranges, &s.grids };
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
};

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

static int host_pipeline_run(const uqs_params* p, const DevParams& dp, int n_flights, int n_frames, const HostLogs& L,
                             int8_t* grids_out, uqs_stats* stats) {
  int rc = pipeline_init();
  if (rc) return rc;
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
    const size_t per_flight = (size_t)n_frames * 152 + cells;             // log bytes + the flight's grid
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
    cp(S.ranges, L.ranges + o * 32, n * 128);
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
    if (flow) {
      if ((rc = pose_device(nf, n_frames, (uint32_t*)S.t.p, (float*)S.rx.p, (float*)S.ry.p, (float*)S.h.p, (float*)S.yaw.p,
                            (uint8_t*)S.q.p, (float*)S.x.p, (float*)S.y.p, 0)))
        return rc;
    }
    if ((rc = replay_device(dp, nf, n_frames, (float*)S.x.p, (float*)S.y.p, (float*)S.yaw.p, (float*)S.ranges.p, nullptr,
                            (int8_t*)S.grids.p, 0, 0, p->H, c < 2)))
      return rc;
    if ((e = cudaEventRecord(S.computed, sc)) != cudaSuccess) return cuda_fail(e, "record");
    if ((e = cudaStreamWaitEvent(P.s_out, S.computed, 0)) != cudaSuccess) return cuda_fail(e, "wait compute");
    KernelTimer t_out(4, P.s_out);
    e = cudaMemcpyAsync(grids_out + (size_t)f0 * cells, S.grids.p, (size_t)nf * cells, cudaMemcpyDeviceToHost, P.s_out);
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

/* Flights per chunk of the host-buffer pipeline (0 = automatic). */
int uqs_set_host_chunk(int flights) {
  if (flights < -1) { set_error("negative chunk"); return UQS_ERR_BAD_ARG; }
  g_ctx.host_chunk = flights;
  return UQS_OK;
}

}  // extern "C"
