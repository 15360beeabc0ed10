// AdamW over a list of fp32 tensors in one launch (the optimizer half of the train step:
// torch.optim.AdamW as configured by the reference, configs/model/hstu.yaml optimizer block and
// src/generative_recommenders_pl/models/generative_recommenders.py:254-322).
//
// The step is pure streaming: per element 4 loads (p, g, m, v) and 3 stores, 28 B.  At C2 the
// 131 263 x 256 embedding table alone is 941 MB per step, so the kernel is written for HBM: every
// block owns one 8192-element chunk of one tensor, all loads of a chunk half are issued before the
// first use (8 x LDG.128 in flight per thread), and the tensor list travels in the kernel parameters
// (no device-side table to keep in sync with autograd's freshly allocated gradients).
#include <cmath>
#include "common.cuh"

namespace grb {
namespace {

constexpr int ADAMW_MAX_TENSORS = 64;
constexpr int ADAMW_THREADS = 256;
constexpr int ADAMW_CHUNK = 8192;     // elements per block: 256 threads x 8 float4

struct AdamwList {
  float* p[ADAMW_MAX_TENSORS];
  const float* g[ADAMW_MAX_TENSORS];
  float* m[ADAMW_MAX_TENSORS];
  float* v[ADAMW_MAX_TENSORS];
  int64_t numel[ADAMW_MAX_TENSORS];
  int32_t first_block[ADAMW_MAX_TENSORS + 1];
  int32_t n;
};

struct AdamwHyper {
  float decay;        // 1 - lr * weight_decay
  float beta1, beta2;
  float one_m_beta1, one_m_beta2;
  float step_size;    // lr / (1 - beta1^t)
  float inv_bc2_sqrt; // 1 / sqrt(1 - beta2^t)
  float eps;
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamwHyper& h) {
  p *= h.decay;
  m = m + (g - m) * h.one_m_beta1;
  v = h.beta2 * v + h.one_m_beta2 * g * g;
  const float denom = sqrtf(v) * h.inv_bc2_sqrt + h.eps;
  p -= h.step_size * (m / denom);
}

__global__ void __launch_bounds__(ADAMW_THREADS)
adamw_kernel(const __grid_constant__ AdamwList L, const AdamwHyper h) {
  // which tensor does this block belong to?  (n <= 64: a short binary search over the block prefix)
  int lo = 0, hi = L.n;
  const int b = (int)blockIdx.x;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (L.first_block[mid] <= b) lo = mid; else hi = mid;
  }
  const int t = lo;
  const int64_t base = (int64_t)(b - L.first_block[t]) * ADAMW_CHUNK;
  const int64_t left = L.numel[t] - base;
  float* __restrict__ p = L.p[t] + base;
  const float* __restrict__ g = L.g[t] + base;
  float* __restrict__ m = L.m[t] + base;
  float* __restrict__ v = L.v[t] + base;
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15u) == 0;
  if (vec && left >= ADAMW_CHUNK) {
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 rp[4], rg[4], rm[4], rv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = (half * 4 + j) * ADAMW_THREADS + threadIdx.x;
        // evict-first both ways: the working set is 7x the L2, nothing is re-read before it is gone
        // (measured on the C2 table: 198 us with default policies, 182 us with .cs loads and stores)
        rp[j] = __ldcs(p4 + i);
        rg[j] = __ldcs(g4 + i);
        rm[j] = __ldcs(m4 + i);
        rv[j] = __ldcs(v4 + i);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = (half * 4 + j) * ADAMW_THREADS + threadIdx.x;
        adamw_one(rp[j].x, rg[j].x, rm[j].x, rv[j].x, h);
        adamw_one(rp[j].y, rg[j].y, rm[j].y, rv[j].y, h);
        adamw_one(rp[j].z, rg[j].z, rm[j].z, rv[j].z, h);
        adamw_one(rp[j].w, rg[j].w, rm[j].w, rv[j].w, h);
        __stcs(p4 + i, rp[j]);
        __stcs(m4 + i, rm[j]);
        __stcs(v4 + i, rv[j]);
      }
    }
  } else {
    const int64_t n = left < ADAMW_CHUNK ? left : ADAMW_CHUNK;
    for (int64_t i = threadIdx.x; i < n; i += ADAMW_THREADS) {
      float pp = p[i], mm = m[i], vv = v[i];
      adamw_one(pp, g[i], mm, vv, h);
      p[i] = pp; m[i] = mm; v[i] = vv;
    }
  }
}

}  // namespace
}  // namespace grb

extern "C" int grb_adamw_step(int n, float* const* p, const float* const* g, float* const* m,
                              float* const* v, const int64_t* numel, double lr, double beta1,
                              double beta2, double eps, double weight_decay, double bias_correction1,
                              double bias_correction2, grb_stream_t stream) {
  using namespace grb;
  GRB_REQUIRE(n >= 0 && (n == 0 || (p && g && m && v && numel)), GRB_ERR_INVALID_ARG,
              "grb_adamw_step: null tensor list");
  GRB_REQUIRE(bias_correction1 > 0.0 && bias_correction2 > 0.0, GRB_ERR_INVALID_ARG,
              "grb_adamw_step: bias corrections must be positive (step >= 1)");
  AdamwHyper h;
  h.decay = (float)(1.0 - lr * weight_decay);
  h.beta1 = (float)beta1;
  h.beta2 = (float)beta2;
  h.one_m_beta1 = (float)(1.0 - beta1);
  h.one_m_beta2 = (float)(1.0 - beta2);
  h.step_size = (float)(lr / bias_correction1);
  h.inv_bc2_sqrt = (float)(1.0 / sqrt(bias_correction2));
  h.eps = (float)eps;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int i = 0;
  while (i < n) {
    AdamwList L;
    L.n = 0;
    int64_t blocks = 0;
    while (i < n && L.n < ADAMW_MAX_TENSORS) {
      GRB_REQUIRE(numel[i] >= 0 && (numel[i] == 0 || (p[i] && g[i] && m[i] && v[i])), GRB_ERR_INVALID_ARG,
                  "grb_adamw_step: tensor %d has a null pointer or a negative size", i);
      const int64_t nb = (numel[i] + ADAMW_CHUNK - 1) / ADAMW_CHUNK;
      if (nb == 0) { ++i; continue; }
      if (blocks + nb > 0x7fffffff) break;
      L.p[L.n] = p[i]; L.g[L.n] = g[i]; L.m[L.n] = m[i]; L.v[L.n] = v[i];
      L.numel[L.n] = numel[i];
      L.first_block[L.n] = (int32_t)blocks;
      blocks += nb;
      ++L.n; ++i;
    }
    if (L.n == 0) {
      GRB_REQUIRE(i >= n, GRB_ERR_UNSUPPORTED, "grb_adamw_step: tensor %d is too large", i);
      break;
    }
    L.first_block[L.n] = (int32_t)blocks;
    adamw_kernel<<<(unsigned)blocks, ADAMW_THREADS, 0, st>>>(L, h);
    GRB_LAUNCH_OK();
  }
  return GRB_OK;
}


// ---------------------------------------------------------------------------------------------------
// fp32 -> bf16 copies of a list of tensors in one launch: the compute-dtype shadows of the fp32 master
// weights of all STU layers (hstu.py:300-305, :404-413 run under bf16 autocast in the reference; here the
// masters stay fp32 and every projection used to cast its own weight, one launch each, every step).
// ---------------------------------------------------------------------------------------------------
namespace grb {
namespace {

struct CastList {
  const float* src[ADAMW_MAX_TENSORS];
  __nv_bfloat16* dst[ADAMW_MAX_TENSORS];
  int64_t numel[ADAMW_MAX_TENSORS];
  int32_t first_block[ADAMW_MAX_TENSORS + 1];
  int32_t n;
};

__global__ void __launch_bounds__(ADAMW_THREADS) cast_many_kernel(const __grid_constant__ CastList L) {
  int lo = 0, hi = L.n;
  const int b = (int) blockIdx.x;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (L.first_block[mid] <= b) lo = mid; else hi = mid;
  }
  const int64_t base = (int64_t) (b - L.first_block[lo]) * ADAMW_CHUNK;
  const int64_t left = L.numel[lo] - base;
  const float* __restrict__ src = L.src[lo] + base;
  __nv_bfloat16* __restrict__ dst = L.dst[lo] + base;
  const bool vec = ((((uintptr_t) src) & 15u) | (((uintptr_t) dst) & 7u)) == 0;
#pragma unroll
  for (int j = 0; j < ADAMW_CHUNK / (4 * ADAMW_THREADS); ++j) {
    const int64_t e = ((int64_t) j * ADAMW_THREADS + threadIdx.x) * 4;
    if (vec && e + 4 <= left) {
      const float4 v = *reinterpret_cast<const float4*>(src + e);
      const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), c = __floats2bfloat162_rn(v.z, v.w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&a);
      o.y = *reinterpret_cast<const uint32_t*>(&c);
      *reinterpret_cast<uint2*>(dst + e) = o;
    } else {
      for (int64_t i = e; i < e + 4 && i < left; ++i) dst[i] = __float2bfloat16_rn(src[i]);
    }
  }
}

}  // namespace
}  // namespace grb

extern "C" int grb_cast_f32_bf16_many(int n, const float* const* src, void* const* dst, const int64_t* numel,
                                      grb_stream_t stream) {
  using namespace grb;
  GRB_REQUIRE(n >= 0 && (n == 0 || (src && dst && numel)), GRB_ERR_INVALID_ARG, "cast_many: bad arguments");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  for (int t0 = 0; t0 < n; t0 += ADAMW_MAX_TENSORS) {
    CastList L{};
    int64_t blocks = 0;
    const int cnt = n - t0 < ADAMW_MAX_TENSORS ? n - t0 : ADAMW_MAX_TENSORS;
    for (int i = 0; i < cnt; ++i) {
      GRB_REQUIRE(src[t0 + i] && dst[t0 + i] && numel[t0 + i] >= 0, GRB_ERR_INVALID_ARG, "cast_many: tensor %d", t0 + i);
      L.src[i] = src[t0 + i];
      L.dst[i] = reinterpret_cast<__nv_bfloat16*>(dst[t0 + i]);
      L.numel[i] = numel[t0 + i];
      L.first_block[i] = (int32_t) blocks;
      blocks += ceil_div(numel[t0 + i], ADAMW_CHUNK);
      GRB_REQUIRE(blocks < (1ll << 31), GRB_ERR_UNSUPPORTED, "cast_many: too many elements");
    }
    L.first_block[cnt] = (int32_t) blocks;
    L.n = cnt;
    if (blocks == 0) continue;
    cast_many_kernel<<<(unsigned) blocks, ADAMW_THREADS, 0, st>>>(L);
    GRB_LAUNCH_OK();
  }
  return GRB_OK;
}
