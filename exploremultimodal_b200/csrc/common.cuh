// Shared host-side plumbing of libmome: error reporting, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/mome.h"

namespace mome {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return MOME_ERR_CUDA;
  }
  return MOME_OK;
}

#define MOME_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      mome::set_error(__VA_ARGS__);  \
      return MOME_ERR_ARG;           \
    }                                \
  } while (0)

// rowwise.cu: column sums of the q and v thirds of dqkv (used by block.cu)
int colsum_qv(const void* dqkv, int dtype, int64_t rows, int64_t d, float* dq_bias, float* dv_bias, void* ws, size_t ws_bytes,
              cudaStream_t s);

bool gemm_prof_active();  // gemm_tcgen05.cu: per-launch event profiling (bench.py roofline) is on

int droppath_scales2(const int32_t* row_sample, int64_t rows, const uint32_t* seed, uint32_t salt1, uint32_t salt2, float p, float* out1,
                     float* out2, cudaStream_t stream);

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

}  // namespace mome
