// Shared host/device helpers of libpia_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <utility>

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

// Programmatic dependent launch (PDL): every kernel of the decode step is launched with the
// programmatic-stream-serialization attribute, announces its dependents at once and waits for its predecessor
// (griddepcontrol.wait = predecessor complete + memory visible) before its first global access.  The next kernel's
// CTAs are then resident, with barriers/TMEM set up, when the predecessor's last CTA retires, which hides the
// launch + prologue latency between the ~330 small kernels of a step.  PIA_PDL=0 disables the attribute (the
// device-side instructions are no-ops for a normally launched kernel).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

}  // namespace pia
