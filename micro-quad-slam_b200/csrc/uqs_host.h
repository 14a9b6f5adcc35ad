// uqs_host.h -- state shared by the host-side translation units of libuqs_mapping.so.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>

#include "../../include/uqs_mapping.h"
#include "uqs_kernels.cuh"

namespace uqs {

// grow-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes);
  void release();
};

struct Context {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t ext_stream = nullptr;
  bool use_ext = false;
  int tune_sw = 0, tune_sh = 0, tune_slices = 0;
  int engine = 0;                               // 0 auto, 1 warp-owned sub-tiles, 2 grid resident per CTA
  int flight_warps = 0;                         // warps per CTA of the resident engine (0 = 16)
  size_t scratch_budget = (size_t)12 << 30;     // ray/frame records held at once
  unsigned long long launches = 0;              // kernels launched by this library
  // scratch
  DevBuf ws_rays, ws_frames, ws_groups, ws_counters, ws_inc, ws_scan;
  // staging for the host-buffer entry points
  DevBuf in_t, in_rx, in_ry, in_h, in_yaw, in_q, in_x, in_y, in_ranges, in_kind, out_grids;

  // optional per-kernel timing (uqs_set_profiling): event pairs on the launching stream
  bool profiling = false;
  struct Span { cudaEvent_t a, b; int kind; };   // kind 0 pose, 1 ray set-up, 2 replay
  std::vector<Span> spans;

  cudaStream_t stream() const { return use_ext ? ext_stream : own_stream; }
  void release_all() {
    DevBuf* all[] = { &ws_rays, &ws_frames, &ws_groups, &ws_counters, &ws_inc, &ws_scan, &in_t, &in_rx,
                      &in_ry, &in_h, &in_yaw, &in_q, &in_x, &in_y, &in_ranges, &in_kind, &out_grids };
    for (DevBuf* b : all) b->release();
  }
};

constexpr size_t kFlightSmemMax = 227u * 1024u - 64u;   // dynamic + the kernel's few static bytes

extern Context g_ctx;

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int check_ready();
int make_dev_params(const uqs_params* p, DevParams* d);
int replay_device(const DevParams& dp, int n_flights, int n_frames, const float* x, const float* y,
                  const float* yaw, const float* ranges, const uint8_t* kind, int8_t* grids,
                  int accumulate, int row0, int rows, bool reset_stats);
int fetch_stats(uqs_stats* stats, uint64_t frames);
void dropin_release();

// RAII-free helper: records an event pair around a kernel launch when profiling is on
struct KernelTimer {
  int kind;
  cudaEvent_t a = nullptr, b = nullptr;
  explicit KernelTimer(int k) : kind(k) {
    if (g_ctx.profiling && cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess)
      cudaEventRecord(a, g_ctx.stream());
  }
  void stop() {
    if (a && b) {
      cudaEventRecord(b, g_ctx.stream());
      g_ctx.spans.push_back({a, b, kind});
    }
  }
};

}  // namespace uqs
