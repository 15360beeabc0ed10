// hstu_attn_sm100.cuh — pieces shared by the tcgen05 attention forward and backward kernels:
// integer-only time bucketing (octave table), the per-sequence timestamp range scan that
// enables 32-bit deltas, and small math / barrier helpers.
#pragma once
#include "common.cuh"
#include "sm100_ptx.cuh"
#include <cuda_fp16.h>

namespace grb {

constexpr int AT_BM = 128;          // query rows per tile
constexpr int AT_BN = 128;          // key rows per tile
constexpr int AT_D = 64;            // head dim (dqk = dv)
constexpr int AT_TILE_BYTES = AT_BN * AT_D * 2;  // 16 KiB: one 128 x 64 bf16 tile

struct alignas(16) OctRec { uint32_t base, t1, t2, t3; };

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t tanh_approx_f16x2(uint32_t x) {
  uint32_t y;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

// mbarrier wait for the single-thread roles (TMA / MMA issuers): a long suspend-time hint lets
// the hardware park the thread instead of re-issuing try_wait, which steals issue slots from the
// epilogue warps sharing the scheduler.
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(200000u)
        : "memory");
  } while (!ok);
}

// ext_ts[b, idx] with the reference's (B, N+1) extension: idx == N reads ts[b, N-1] (hstu.py:113-115)
__device__ __forceinline__ int64_t ext_ts_at(const int64_t* ts, int64_t b, int64_t N, int64_t idx) {
  if (idx >= N) idx = N - 1;
  return ts[b * N + idx];
}

// Octave table: for d in [2^e, 2^(e+1)) bucket(d) = base + (d>=t1) + (d>=t2) + (d>=t3).
// Built by one warp (lane = e) from the host-tabulated thresholds.  flags[0] = 1 when the table
// cannot express the thresholds (4+ in one octave, or bucket(0) != bucket(1)): then every lookup
// takes the exact binary search.  flags[1] = bucket(0).
__device__ __forceinline__ void build_octave_table(OctRec* oct, int* flags, const int64_t* thr,
                                                   int nb, int lane) {
  const int e = lane;
  const int64_t lo = 1ll << e, hi = (1ll << (e + 1)) - 1;
  const int base = bucket_of(thr, nb, lo);
  OctRec r;
  r.base = (uint32_t) base;
  uint32_t t[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int idx = base + i;
    t[i] = (idx < nb && thr[idx] <= hi) ? (uint32_t) thr[idx] : 0xffffffffu;
  }
  r.t1 = t[0]; r.t2 = t[1]; r.t3 = t[2];
  bool bad = (base + 3 < nb && thr[base + 3] <= hi);
  const int b0 = bucket_of(thr, nb, 0);
  if (e == 0 && b0 != base) bad = true;   // octave 0 also serves d == 0
  oct[e] = r;
  const unsigned any_bad = __ballot_sync(0xffffffffu, bad);
  if (lane == 0) { flags[0] = any_bad != 0; flags[1] = b0; }
}

// same table, precomputed on the host (grb_bucket_octaves): 130 words, copied by one warp
__device__ __forceinline__ void load_octave_table(OctRec* oct, int* flags, const uint32_t* pre,
                                                  int lane) {
  const uint4 v = reinterpret_cast<const uint4*>(pre)[lane];
  OctRec r; r.base = v.x; r.t1 = v.y; r.t2 = v.z; r.t3 = v.w;
  oct[lane] = r;
  if (lane == 0) { flags[0] = (int) pre[128]; flags[1] = (int) pre[129]; }
}

// out-of-line exact search: keeps the (rare) 64-bit path from bloating the unrolled hot loops
static __device__ __noinline__ int bucket_search_noinline(const int64_t* thr, int n, int64_t d) {
  return bucket_of(thr, n, d);
}

// exact bucket of a 64-bit |delta|; table fast path below 2^32 - 1
__device__ __forceinline__ int bucket_wide(const OctRec* __restrict__ oct,
                                           const int64_t* __restrict__ thr_g, int nb, bool slow,
                                           int64_t d) {
  // 0xffffffff doubles as the "no threshold" marker of the table, so it takes the slow path too
  if (slow || (uint64_t) d >= 0xffffffffull) return bucket_search_noinline(thr_g, nb, d);
  const uint32_t u = (uint32_t) d;
  const OctRec r = oct[31 - __clz(u | 1u)];
  return (int) r.base + (u >= r.t1) + (u >= r.t2) + (u >= r.t3);
}
// bucket of a 32-bit |delta| < 0xffffffff, table known to be valid: no branches
__device__ __forceinline__ int bucket_narrow(const OctRec* __restrict__ oct, uint32_t u) {
  const OctRec r = oct[31 - __clz(u | 1u)];
  return (int) r.base + (u >= r.t1) + (u >= r.t2) + (u >= r.t3);
}

// min / max of ts[b, 0..count) over the 256 epilogue threads (tid256 in [0,256)).  If the span
// fits in 32 bits every |delta| of the sequence does, and the kernels use 32-bit arithmetic on
// (ts - tmin).  red: 2 * NT/32 x int64 of shared memory.  All NT threads must call it.
struct TsRange { int64_t tmin; bool narrow; };
template <int NT = 256>
__device__ __forceinline__ TsRange scan_ts_range(const int64_t* __restrict__ ts_row, int count,
                                                 int tid256, int64_t* red, int bar_id) {
  int64_t mn = INT64_MAX, mx = INT64_MIN;
  for (int i = tid256; i < count; i += NT) {
    const int64_t v = ts_row[i];
    mn = v < mn ? v : mn;
    mx = v > mx ? v : mx;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int64_t a = __shfl_xor_sync(0xffffffffu, mn, o);
    const int64_t c = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn;
    mx = c > mx ? c : mx;
  }
  if ((tid256 & 31) == 0) { red[(tid256 >> 5) * 2] = mn; red[(tid256 >> 5) * 2 + 1] = mx; }
  named_bar_sync(bar_id, NT);
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) {
    const int64_t a = red[w * 2], c = red[w * 2 + 1];
    mn = a < mn ? a : mn;
    mx = c > mx ? c : mx;
  }
  TsRange r;
  r.tmin = mn;
  r.narrow = (count > 0) && ((uint64_t) mx - (uint64_t) mn) <= 0xfffffffeull;
  return r;
}

}  // namespace grb
