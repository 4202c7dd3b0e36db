// common.cuh — shared helpers for the sm_100a kernels of libgmlm_b200.so.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/gmlm_b200.h"

namespace gmlm {

// ---------------------------------------------------------------- error plumbing
char* err_buf();  // thread-local, 512 bytes (defined in graph_build.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define GMLM_CUDA_TRY(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::gmlm::fail(GMLM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                          __FILE__, __LINE__);                                           \
  } while (0)

#define GMLM_LAUNCH_CHECK()                                                              \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess)                                                               \
      return ::gmlm::fail(GMLM_ERR_CUDA, "kernel launch failed: %s (%s:%d)",              \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                   \
  } while (0)

#define GMLM_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return ::gmlm::fail(GMLM_ERR_INVALID, __VA_ARGS__); \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kMaxDevices = 64;

inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

// cached PER DEVICE (a process may drive several GPUs); the benign race writes the same value
inline int num_sms() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

// workspace carving (256-byte aligned sub-buffers)
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
  size_t used() const { return (off + 255) & ~size_t(255); }
};

// ------------------------------------------------------------- element access
// A "pack" is what one lane moves with one memory instruction: VEC elements.
template <typename T, int VEC>
struct Pack;

template <>
struct Pack<float, 4> {
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = __ldg(reinterpret_cast<const float4*>(p)); }
  // coherent load (plain ld.global) for operands the same kernel also writes
  __device__ __forceinline__ void load_rw(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { __stcs(reinterpret_cast<float4*>(p), v); }
  __device__ __forceinline__ void unpack(float* f) const { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
  __device__ __forceinline__ void pack(const float* f) { v = make_float4(f[0], f[1], f[2], f[3]); }
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
};

template <>
struct Pack<float, 1> {
  float v;
  __device__ __forceinline__ void load(const float* p) { v = __ldg(p); }
  __device__ __forceinline__ void load_rw(const float* p) { v = *p; }
  __device__ __forceinline__ void store(float* p) const { *p = v; }
  __device__ __forceinline__ void unpack(float* f) const { f[0] = v; }
  __device__ __forceinline__ void pack(const float* f) { v = f[0]; }
  __device__ __forceinline__ void zero() { v = 0.f; }
};

template <>
struct Pack<__nv_bfloat16, 8> {
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void load_rw(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { __stcs(reinterpret_cast<uint4*>(p), v); }
  __device__ __forceinline__ void unpack(float* f) const {
    // bf16 -> fp32 is a 16-bit left shift
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
    f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
    f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
  }
  __device__ __forceinline__ void pack(const float* f) {
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]);
    __nv_bfloat162 d = __floats2bfloat162_rn(f[6], f[7]);
    v.x = *reinterpret_cast<uint32_t*>(&a); v.y = *reinterpret_cast<uint32_t*>(&b);
    v.z = *reinterpret_cast<uint32_t*>(&c); v.w = *reinterpret_cast<uint32_t*>(&d);
  }
  __device__ __forceinline__ void zero() { v = make_uint4(0u, 0u, 0u, 0u); }
};

template <>
struct Pack<__nv_bfloat16, 1> {
  __nv_bfloat16 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *p; }
  __device__ __forceinline__ void load_rw(const __nv_bfloat16* p) { v = *p; }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *p = v; }
  __device__ __forceinline__ void unpack(float* f) const { f[0] = __bfloat162float(v); }
  __device__ __forceinline__ void pack(const float* f) { v = __float2bfloat16_rn(f[0]); }
  __device__ __forceinline__ void zero() { v = __float2bfloat16_rn(0.f); }
};

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_float(float v);
template <>
__device__ __forceinline__ float from_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// streaming (read-once) index loads: keep them out of L1 and first-to-evict in L2 so the
// gathered feature rows keep the cache.
__device__ __forceinline__ int32_t ld_stream(const int32_t* p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }

}  // namespace gmlm
