// Shared helpers for the acr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/acr_b200.h"

namespace acr {

void set_error(const char* fmt, ...);

#define ACR_CHECK_ARG(cond, ...)                      \
  do {                                                \
    if (!(cond)) {                                    \
      acr::set_error(__VA_ARGS__);                    \
      return ACR_B200_EINVAL;                         \
    }                                                 \
  } while (0)

#define ACR_CHECK_CUDA(expr)                                                           \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      acr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                     __LINE__);                                                        \
      return ACR_B200_ECUDA;                                                           \
    }                                                                                  \
  } while (0)

#define ACR_CHECK_LAUNCH() ACR_CHECK_CUDA(cudaGetLastError())

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Opt a kernel into more than 48 KB of dynamic shared memory, once per (kernel instance, device): the attribute is
// per device, and one process may drive several (`done` is the caller's static per-instance device bit mask).
template <typename Kernel>
static inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, unsigned long long* done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (__atomic_load_n(done, __ATOMIC_ACQUIRE) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) __atomic_fetch_or(done, bit, __ATOMIC_RELEASE);
  return e;
}

// ---- 16-bit activation type helpers ------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }

// 8 x 16-bit values <-> 8 floats through one 128-bit access
template <typename T>
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const T* p = reinterpret_cast<const T*>(&u);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = to_f32<T>(p[i]);
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  T* p = reinterpret_cast<T*>(&u);
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = from_f32<T>(f[i]);
  return u;
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one whole 32-byte sector per thread and instruction
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

}  // namespace acr
