// common.cuh -- shared helpers of libocc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "occ_b200.h"

namespace occb200 {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

#define OCC_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess) {                                                                       \
      ::occb200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return 1;                                                                                     \
    }                                                                                               \
  } while (0)

#define OCC_KERNEL_OK(name)                                                                     \
  do {                                                                                          \
    ::occb200::count_launch();                                                                  \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) {                                                                   \
      ::occb200::set_error("%s:%d: launch of %s -> %s", __FILE__, __LINE__, name,               \
                           cudaGetErrorString(e__));                                            \
      return 1;                                                                                 \
    }                                                                                           \
  } while (0)

#define OCC_REQUIRE(cond, msg)                                          \
  do {                                                                  \
    if (!(cond)) {                                                      \
      ::occb200::set_error("%s:%d: %s", __FILE__, __LINE__, msg);       \
      return 2;                                                         \
    }                                                                   \
  } while (0)

// Streaming (read-once) global loads: keep them out of L1 so the small hot tables stay there.
__device__ __forceinline__ float2 ld_stream2(const float2 *p) {   // 8-byte streaming load (read once)
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream4(const float4 *p) {   // 16-byte streaming load (read once)
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

}  // namespace occb200
