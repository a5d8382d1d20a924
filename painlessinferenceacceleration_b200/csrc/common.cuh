// Shared host/device helpers of libpia_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/pia_b200.h"

namespace pia {

void set_error(const char *fmt, ...);
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

#define PIA_CUDA_CHECK(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      pia::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PIA_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define PIA_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      pia::set_error(__VA_ARGS__);      \
      return PIA_ERR_INVALID;           \
    }                                   \
  } while (0)

// checks the launch itself (legal during stream capture: no sync)
#define PIA_LAUNCH_CHECK()                                                                  \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      pia::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PIA_ERR_CUDA;                                                                  \
    }                                                                                       \
    pia::count_launch();                                                                    \
  } while (0)

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

}  // namespace pia
