// uqs_host.h -- state shared by the host-side translation units of libuqs_mapping.so.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

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
  size_t scratch_budget = (size_t)12 << 30;     // ray/frame records held at once
  unsigned long long launches = 0;              // kernels launched by this library
  // scratch
  DevBuf ws_rays, ws_frames, ws_groups, ws_counters, ws_inc, ws_scan;
  // staging for the host-buffer entry points
  DevBuf in_t, in_rx, in_ry, in_h, in_yaw, in_q, in_x, in_y, in_ranges, in_kind, out_grids;

  cudaStream_t stream() const { return use_ext ? ext_stream : own_stream; }
  void release_all() {
    DevBuf* all[] = { &ws_rays, &ws_frames, &ws_groups, &ws_counters, &ws_inc, &ws_scan, &in_t, &in_rx,
                      &in_ry, &in_h, &in_yaw, &in_q, &in_x, &in_y, &in_ranges, &in_kind, &out_grids };
    for (DevBuf* b : all) b->release();
  }
};

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

}  // namespace uqs
