// Shared helpers for the slotvps_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <vector>

#include "../../include/slotvps_b200.h"

namespace slotvps {

constexpr int C = SLOTVPS_C;          // 256
constexpr int CIN = SLOTVPS_CIN;      // 128
constexpr float LN_EPS = 1e-5f;
constexpr float BN_EPS = 1e-5f;

extern thread_local char g_err[512];
extern thread_local int64_t g_launches;

// Optional per-launch timing (bench.py's roofline leg): when enabled every launch records one CUDA
// event on its stream right after the kernel; durations are differences of consecutive events.
void prof_mark(const char* name, cudaStream_t s);
extern thread_local bool g_prof_on;
extern thread_local int g_prof_grid;      // grid size (CTAs) of the next marked launch, 0 = not recorded

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

#define SV_CHECK_CUDA(expr)                                                            \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) return ::slotvps::fail(SLOTVPS_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

// requires a `cudaStream_t s` in scope (the stream the kernel was launched on)
#define SV_CHECK_LAUNCH(name)                                                          \
  do {                                                                                 \
    ++::slotvps::g_launches;                                                           \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) return ::slotvps::fail(SLOTVPS_ECUDA, "launch %s: %s", name, cudaGetErrorString(_e)); \
    if (::slotvps::g_prof_on) ::slotvps::prof_mark(name, s);                           \
  } while (0)

// First statement of an entry point (after `cudaStream_t s`): while profiling, the time the stream sat idle waiting
// for the host to call us is booked under "(host gap)" instead of inflating the entry point's first kernel.
#define SV_PROF_ENTRY()                                                                \
  do {                                                                                 \
    if (::slotvps::g_prof_on) ::slotvps::prof_mark("(host gap)", s);                   \
  } while (0)

#define SV_REQUIRE(cond, msg)                                                          \
  do {                                                                                 \
    if (!(cond)) return ::slotvps::fail(SLOTVPS_EINVAL, "%s (%s)", msg, #cond);        \
  } while (0)

#define SV_TRY(expr)                                                                   \
  do {                                                                                 \
    int _r = (expr);                                                                   \
    if (_r != SLOTVPS_OK) return _r;                                                   \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (context) attribute: remember (function, device, bytes)
// per host thread instead of a process-wide flag, so a second GPU in the same process gets its own call.
inline int ensure_dyn_smem(const void* fn, size_t bytes) {
  struct Done { const void* fn; int dev; size_t bytes; };
  static thread_local std::vector<Done> done;
  int dev = 0;
  SV_CHECK_CUDA(cudaGetDevice(&dev));
  for (auto& d : done)
    if (d.fn == fn && d.dev == dev) {
      if (d.bytes >= bytes) return SLOTVPS_OK;
      SV_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      d.bytes = bytes;
      return SLOTVPS_OK;
    }
  SV_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  done.push_back({fn, dev, bytes});
  return SLOTVPS_OK;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Bump allocator over the caller-provided workspace.
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
  template <typename T>
  T* take(size_t n) {
    size_t o = align_up(off);
    off = o + n * sizeof(T);
    return (T*)(base ? base + o : nullptr);
  }
  bool ok() const { return off <= cap; }
};

}  // namespace slotvps
