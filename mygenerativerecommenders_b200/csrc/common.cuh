// common.cuh — shared host/device helpers for libgrb200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <atomic>
#include "../../include/grb200.h"

namespace grb {

// ---- host: error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define GRB_REQUIRE(cond, code, ...)                      \
  do {                                                    \
    if (!(cond)) {                                        \
      grb::set_error(__VA_ARGS__);                        \
      return (code);                                      \
    }                                                     \
  } while (0)

#define GRB_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      grb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                     __LINE__);                                                        \
      return GRB_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

#define GRB_LAUNCH_OK()                                                                 \
  do {                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      grb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),        \
                     __FILE__, __LINE__);                                               \
      return GRB_ERR_CUDA;                                                              \
    }                                                                                   \
    grb::count_launch();                                                                \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
int num_sms();

// ---- device helpers ---------------------------------------------------------------------------
template <int BITS>
struct IndexT;
template <>
struct IndexT<32> { using type = int32_t; };
template <>
struct IndexT<64> { using type = int64_t; };

__device__ __forceinline__ int64_t load_index(const void* p, int64_t i, int bits) {
  return bits == 32 ? (int64_t) reinterpret_cast<const int32_t*>(p)[i]
                    : reinterpret_cast<const int64_t*>(p)[i];
}

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

// exact-ish SiLU in fp32: x * sigmoid(x) with IEEE division (matches ATen's x / (1 + exp(-x)))
__device__ __forceinline__ float silu_f32(float x) { return x / (1.0f + expf(-x)); }
// d/dx silu = s * (1 + x * (1 - s)),  s = sigmoid(x)
__device__ __forceinline__ float dsilu_f32(float x) {
  float s = 1.0f / (1.0f + expf(-x));
  return s * (1.0f + x * (1.0f - s));
}

// Philox4x32-10 (Salmon et al., the generator behind curand / torch's CUDA RNG): counter (c0, c1, 0, 0)
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
  uint32_t c2 = 0u, c3 = 0u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// number of table entries <= d  (table ascending, n entries) — the reference's bucketization_fn
// (hstu.py:579-581) tabulated on the host; result in [0, n].
__device__ __forceinline__ int bucket_of(const int64_t* __restrict__ thr, int n, int64_t d) {
  int lo = 0, hi = n;  // first index with thr[idx] > d
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (thr[mid] <= d) lo = mid + 1; else hi = mid;
  }
  return lo;
}

}  // namespace grb
