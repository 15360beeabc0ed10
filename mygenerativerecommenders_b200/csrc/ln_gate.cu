// ln_gate.cu — y = gate * LayerNorm(x) (no affine), one warp per jagged row, warp-shuffle
// reductions.  Reference: hstu.py:258-264 (_norm_input / _norm_attn_output) and :402
// (o_input = u * norm(attn_output)).  gate == NULL gives the plain LayerNorm of :300.
// HBM-bound: fwd reads x (+gate) once from HBM (re-reads hit L1), writes y.
#include "common.cuh"
#include <initializer_list>

namespace grb {

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

constexpr int LN_WARPS = 8;

template <typename T>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_gate_fwd_kernel(
    const T* __restrict__ x, int64_t ldx, const T* __restrict__ gate, int64_t ldg,
    T* __restrict__ y, int64_t ldy, float* __restrict__ mean, float* __restrict__ rstd,
    int64_t rows, int W, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t) blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * ldx;
  float s = 0.f;
  for (int c = lane; c < W; c += 32) s += ldf<T>(xr + c);
  const float mu = warp_sum(s) / (float) W;
  float q = 0.f;
  for (int c = lane; c < W; c += 32) {
    const float d = ldf<T>(xr + c) - mu;
    q = fmaf(d, d, q);
  }
  const float var = warp_sum(q) / (float) W;
  const float rs = rsqrtf(var + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  T* yr = y + row * ldy;
  if (gate) {
    const T* gr = gate + row * ldg;
    for (int c = lane; c < W; c += 32)
      stf<T>(yr + c, ldf<T>(gr + c) * ((ldf<T>(xr + c) - mu) * rs));
  } else {
    for (int c = lane; c < W; c += 32) stf<T>(yr + c, (ldf<T>(xr + c) - mu) * rs);
  }
}

template <typename T>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_gate_bwd_kernel(
    const T* __restrict__ x, int64_t ldx, const T* __restrict__ gate, int64_t ldg,
    const T* __restrict__ dy, int64_t lddy, const float* __restrict__ mean,
    const float* __restrict__ rstd, T* __restrict__ dx, int64_t lddx, T* __restrict__ dgate,
    int64_t lddg, int64_t rows, int W, const T* __restrict__ res, int64_t ldres) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t) blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * ldx;
  const T* dyr = dy + row * lddy;
  const T* gr = gate ? gate + row * ldg : nullptr;
  const float mu = mean[row], rs = rstd[row];
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < W; c += 32) {
    const float xh = (ldf<T>(xr + c) - mu) * rs;
    const float dxh = ldf<T>(dyr + c) * (gr ? ldf<T>(gr + c) : 1.0f);
    s1 += dxh;
    s2 = fmaf(dxh, xh, s2);
  }
  const float m1 = warp_sum(s1) / (float) W;
  const float m2 = warp_sum(s2) / (float) W;
  T* dxr = dx + row * lddx;
  T* dgr = dgate ? dgate + row * lddg : nullptr;
  for (int c = lane; c < W; c += 32) {
    const float xh = (ldf<T>(xr + c) - mu) * rs;
    const float dyv = ldf<T>(dyr + c);
    const float dxh = dyv * (gr ? ldf<T>(gr + c) : 1.0f);
    stf<T>(dxr + c, rs * (dxh - m1 - xh * m2) + (res ? ldf<T>(res + row * ldres + c) : 0.f));
    if (dgr) stf<T>(dgr + c, dyv * xh);
  }
}

// ---- bf16 rows of 256 / 512 / 1024 elements: every lane owns NCH chunks of 8 consecutive
// elements (one 16-byte access each), the row lives in registers, each tensor is touched once ----
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(f[2 * i + 1]), "f"(f[2 * i]));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Dropout on the output (hstu.py:404-408, dropout(u * norm(a)) in front of the output projection) is
// drawn inside the kernels: Philox4x32-10 keyed by the caller's per-step device seed, counter = (row,
// 8-element chunk, salt of the call site); one call yields the eight 16-bit draws of a chunk; keep iff
// draw >= thr16 = round(p * 2^16), kept values scaled by 2^16 / (2^16 - thr16).  The backward kernel
// regenerates the mask: no mask tensor, no separate dropout / masked-scale passes.  Same distribution as
// torch's dropout, not the same stream (parity tests run with p = 0 / eval, as for the input path).
struct LnDrop {
  const int64_t* seed;
  uint32_t salt, thr16;
  float keep_scale;
};
__device__ __forceinline__ void ln_drop_mask8(const LnDrop& d, int64_t row, int chunk, float (&m)[8]) {
  const uint64_t sd = (uint64_t) d.seed[0];
  const uint4 u = philox4x32_10((uint32_t) row, (uint32_t) chunk | (d.salt << 8), (uint32_t) sd, (uint32_t) (sd >> 32));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[2 * i] = (w[i] & 0xffffu) >= d.thr16 ? d.keep_scale : 0.f;
    m[2 * i + 1] = (w[i] >> 16) >= d.thr16 ? d.keep_scale : 0.f;
  }
}

template <int NCH, bool GATE, bool DROP>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_gate_fwd_bf16v_kernel(
    const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ gate,
    int64_t ldg, __nv_bfloat16* __restrict__ y, int64_t ldy, float* __restrict__ mean,
    float* __restrict__ rstd, int64_t rows, float eps, LnDrop drop) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t) blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  constexpr int W = 256 * NCH;
  float xv[NCH][8], gv[GATE ? NCH : 1][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    unpack8(*reinterpret_cast<const uint4*>(x + row * ldx + 256 * k + 8 * lane), xv[k]);
    if (GATE) unpack8(*reinterpret_cast<const uint4*>(gate + row * ldg + 256 * k + 8 * lane), gv[k]);
#pragma unroll
    for (int e = 0; e < 8; ++e) s += xv[k][e];
  }
  const float mu = warp_sum(s) / (float) W;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NCH; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float d = xv[k][e] - mu; q = fmaf(d, d, q); }
  const float rs = rsqrtf(warp_sum(q) / (float) W + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (GATE ? gv[k][e] : 1.0f) * ((xv[k][e] - mu) * rs);
    if (DROP) {
      float m[8];
      ln_drop_mask8(drop, row, 32 * k + lane, m);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] *= m[e];
    }
    *reinterpret_cast<uint4*>(y + row * ldy + 256 * k + 8 * lane) = pack8(o);
  }
}

// res != NULL: dx = (LayerNorm backward) + res — the gradient that reaches x through the residual branch
// (hstu.py:413, new_outputs = o(...) + x), which autograd would add with one more elementwise pass.
template <int NCH, bool GATE, bool DROP>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_gate_bwd_bf16v_kernel(
    const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ gate,
    int64_t ldg, const __nv_bfloat16* __restrict__ dy, int64_t lddy, const float* __restrict__ mean,
    const float* __restrict__ rstd, __nv_bfloat16* __restrict__ dx, int64_t lddx,
    __nv_bfloat16* __restrict__ dgate, int64_t lddg, int64_t rows,
    const __nv_bfloat16* __restrict__ res, int64_t ldres, LnDrop drop) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t) blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  constexpr int W = 256 * NCH;
  const float mu = mean[row], rs = rstd[row];
  float xh[NCH][8], dyv[NCH][8], dxh[NCH][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    float xv[8], gv[8];
    unpack8(*reinterpret_cast<const uint4*>(x + row * ldx + 256 * k + 8 * lane), xv);
    unpack8(*reinterpret_cast<const uint4*>(dy + row * lddy + 256 * k + 8 * lane), dyv[k]);
    if (GATE) unpack8(*reinterpret_cast<const uint4*>(gate + row * ldg + 256 * k + 8 * lane), gv);
    if (DROP) {   // dy arrives for the dropped output: the same mask and scale on the way back
      float m[8];
      ln_drop_mask8(drop, row, 32 * k + lane, m);
#pragma unroll
      for (int e = 0; e < 8; ++e) dyv[k][e] *= m[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      xh[k][e] = (xv[e] - mu) * rs;
      dxh[k][e] = dyv[k][e] * (GATE ? gv[e] : 1.0f);
      s1 += dxh[k][e];
      s2 = fmaf(dxh[k][e], xh[k][e], s2);
    }
  }
  const float m1 = warp_sum(s1) / (float) W;
  const float m2 = warp_sum(s2) / (float) W;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    float o[8], og[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o[e] = rs * (dxh[k][e] - m1 - xh[k][e] * m2);
      og[e] = dyv[k][e] * xh[k][e];
    }
    if (res) {
      float rv[8];
      unpack8(*reinterpret_cast<const uint4*>(res + row * ldres + 256 * k + 8 * lane), rv);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += rv[e];
    }
    *reinterpret_cast<uint4*>(dx + row * lddx + 256 * k + 8 * lane) = pack8(o);
    if (GATE) *reinterpret_cast<uint4*>(dgate + row * lddg + 256 * k + 8 * lane) = pack8(og);
  }
}

static bool ln_vec_ok(int64_t W, std::initializer_list<const void*> ptrs, std::initializer_list<int64_t> lds) {
  if (W != 256 && W != 512 && W != 1024) return false;
  for (const void* p : ptrs) if (p && (reinterpret_cast<uintptr_t>(p) & 15)) return false;
  for (int64_t l : lds) if (l % 8) return false;
  return true;
}

}  // namespace grb

using namespace grb;

// y = x / max(||x||_2, eps) per row (negative_sampler.py:31-37 _maybe_l2_norm; postprocessors.py:47-55).
// inv[row] = 1 / max(||x||, eps), negated when the clamp is active (||x|| < eps) so that the
// backward knows which branch of clamp(min=eps) the row took.
__global__ void __launch_bounds__(LN_WARPS * 32) l2norm_fwd_kernel(
    const float* __restrict__ x, int64_t ldx, float* __restrict__ y, int64_t ldy,
    float* __restrict__ inv, int64_t rows, int W, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t) blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  float q = 0.f;
  for (int c = lane; c < W; c += 32) { const float v = xr[c]; q = fmaf(v, v, q); }
  const float nrm = sqrtf(warp_sum(q));
  const float iv = 1.0f / fmaxf(nrm, eps);
  if (lane == 0) inv[row] = nrm >= eps ? iv : -iv;
  float* yr = y + row * ldy;
  for (int c = lane; c < W; c += 32) yr[c] = xr[c] * iv;
}
// dx = inv * (dy - y * <y, dy>) when the norm passed the clamp, inv * dy otherwise
__global__ void __launch_bounds__(LN_WARPS * 32) l2norm_bwd_kernel(
    const float* __restrict__ y, int64_t ldy, const float* __restrict__ dy, int64_t lddy,
    const float* __restrict__ inv, float* __restrict__ dx, int64_t lddx, int64_t rows, int W) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t) blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* yr = y + row * ldy;
  const float* gr = dy + row * lddy;
  float* dr = dx + row * lddx;
  const float iv = inv[row];
  if (iv < 0.f) {
    for (int c = lane; c < W; c += 32) dr[c] = -iv * gr[c];
    return;
  }
  float d = 0.f;
  for (int c = lane; c < W; c += 32) d = fmaf(yr[c], gr[c], d);
  d = warp_sum(d);
  for (int c = lane; c < W; c += 32) dr[c] = iv * (gr[c] - yr[c] * d);
}

extern "C" {

int grb_l2norm_fwd(const float* x, int64_t ldx, float* y, int64_t ldy, float* inv, int64_t rows,
                   int64_t W, float eps, grb_stream_t stream) {
  GRB_REQUIRE(x && y && inv && rows >= 0 && W > 0 && W < (1 << 30) && eps > 0.f, GRB_ERR_INVALID_ARG,
              "l2norm_fwd: bad arguments");
  if (rows == 0) return GRB_OK;
  l2norm_fwd_kernel<<<(unsigned) ceil_div(rows, LN_WARPS), LN_WARPS * 32, 0,
                      reinterpret_cast<cudaStream_t>(stream)>>>(x, ldx, y, ldy, inv, rows, (int) W, eps);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_l2norm_bwd(const float* y, int64_t ldy, const float* dy, int64_t lddy, const float* inv,
                   float* dx, int64_t lddx, int64_t rows, int64_t W, grb_stream_t stream) {
  GRB_REQUIRE(y && dy && inv && dx && rows >= 0 && W > 0, GRB_ERR_INVALID_ARG, "l2norm_bwd: bad arguments");
  if (rows == 0) return GRB_OK;
  l2norm_bwd_kernel<<<(unsigned) ceil_div(rows, LN_WARPS), LN_WARPS * 32, 0,
                      reinterpret_cast<cudaStream_t>(stream)>>>(y, ldy, dy, lddy, inv, dx, lddx, rows, (int) W);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

static bool ln_drop_of(const grb_ln_gate_args* a, LnDrop* d, const char* who) {
  *d = LnDrop{};
  if (a->p_drop <= 0.f) return true;
  if (!(a->p_drop < 1.f) || !a->seed) { set_error("%s: dropout needs 0 <= p < 1 and a device seed", who); return false; }
  uint32_t thr = (uint32_t) lrintf(a->p_drop * 65536.f);
  if (thr > 65535u) thr = 65535u;
  d->seed = a->seed;
  d->salt = (uint32_t) a->salt & 0xffffffu;
  d->thr16 = thr;
  d->keep_scale = 65536.f / (float) (65536u - thr);
  return true;
}

int grb_ln_gate_fwd_ex(const grb_ln_gate_args* a, grb_stream_t stream) {
  GRB_REQUIRE(a && a->x && a->y && a->mean && a->rstd && a->rows >= 0 && a->W > 0 && a->W < (1 << 30),
              GRB_ERR_INVALID_ARG, "ln_gate_fwd: bad arguments");
  GRB_REQUIRE(a->dtype == GRB_F32 || a->dtype == GRB_BF16, GRB_ERR_INVALID_ARG, "ln_gate_fwd: dtype");
  LnDrop drop;
  if (!ln_drop_of(a, &drop, "ln_gate_fwd")) return GRB_ERR_INVALID_ARG;
  if (a->rows == 0) return GRB_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned) ceil_div(a->rows, LN_WARPS);
  const int64_t W = a->W, rows = a->rows;
  if (a->dtype == GRB_BF16 && ln_vec_ok(W, {a->x, a->gate, a->y}, {a->ldx, a->gate ? a->ldg : 0, a->ldy})) {
    auto X = (const __nv_bfloat16*) a->x; auto G = (const __nv_bfloat16*) a->gate; auto Y = (__nv_bfloat16*) a->y;
#define GRB_LN_FWD2(NCH, GATE, DROP)                                                                  \
    ln_gate_fwd_bf16v_kernel<NCH, GATE, DROP><<<grid, LN_WARPS * 32, 0, st>>>(X, a->ldx, G, a->ldg, Y, a->ldy, a->mean, a->rstd, rows, a->eps, drop)
#define GRB_LN_FWD(NCH)                                                                              \
    if (a->gate) { if (drop.seed) GRB_LN_FWD2(NCH, true, true); else GRB_LN_FWD2(NCH, true, false); } \
    else { if (drop.seed) GRB_LN_FWD2(NCH, false, true); else GRB_LN_FWD2(NCH, false, false); }
    if (W == 256) { GRB_LN_FWD(1) } else if (W == 512) { GRB_LN_FWD(2) } else { GRB_LN_FWD(4) }
#undef GRB_LN_FWD
#undef GRB_LN_FWD2
    GRB_LAUNCH_OK();
    return GRB_OK;
  }
  GRB_REQUIRE(!drop.seed, GRB_ERR_UNSUPPORTED,
              "ln_gate_fwd: fused dropout needs bf16 rows of 256 / 512 / 1024 elements, 16-byte aligned");
  if (a->dtype == GRB_F32)
    ln_gate_fwd_kernel<float><<<grid, LN_WARPS * 32, 0, st>>>(
        (const float*) a->x, a->ldx, (const float*) a->gate, a->ldg, (float*) a->y, a->ldy, a->mean, a->rstd,
        rows, (int) W, a->eps);
  else
    ln_gate_fwd_kernel<__nv_bfloat16><<<grid, LN_WARPS * 32, 0, st>>>(
        (const __nv_bfloat16*) a->x, a->ldx, (const __nv_bfloat16*) a->gate, a->ldg, (__nv_bfloat16*) a->y,
        a->ldy, a->mean, a->rstd, rows, (int) W, a->eps);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

// a->y is dy here
int grb_ln_gate_bwd_ex(const grb_ln_gate_args* a, grb_stream_t stream) {
  GRB_REQUIRE(a && a->x && a->y && a->dx && a->mean && a->rstd && a->rows >= 0 && a->W > 0, GRB_ERR_INVALID_ARG,
              "ln_gate_bwd: bad arguments");
  GRB_REQUIRE((a->gate == nullptr) == (a->dgate == nullptr), GRB_ERR_INVALID_ARG,
              "ln_gate_bwd: gate and dgate must both be null or both non-null");
  GRB_REQUIRE(a->dtype == GRB_F32 || a->dtype == GRB_BF16, GRB_ERR_INVALID_ARG, "ln_gate_bwd: dtype");
  LnDrop drop;
  if (!ln_drop_of(a, &drop, "ln_gate_bwd")) return GRB_ERR_INVALID_ARG;
  if (a->rows == 0) return GRB_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned) ceil_div(a->rows, LN_WARPS);
  const int64_t W = a->W, rows = a->rows;
  if (a->dtype == GRB_BF16 &&
      ln_vec_ok(W, {a->x, a->gate, a->y, a->dx, a->dgate, a->res},
                {a->ldx, a->gate ? a->ldg : 0, a->ldy, a->lddx, a->dgate ? a->lddg : 0, a->res ? a->ldres : 0})) {
    auto X = (const __nv_bfloat16*) a->x; auto G = (const __nv_bfloat16*) a->gate; auto DY = (const __nv_bfloat16*) a->y;
    auto DX = (__nv_bfloat16*) a->dx; auto DG = (__nv_bfloat16*) a->dgate; auto RS = (const __nv_bfloat16*) a->res;
#define GRB_LN_BWD2(NCH, GATE, DROP)                                                                  \
    ln_gate_bwd_bf16v_kernel<NCH, GATE, DROP><<<grid, LN_WARPS * 32, 0, st>>>(X, a->ldx, G, a->ldg, DY, a->ldy, a->mean, a->rstd, DX, a->lddx, DG, a->lddg, rows, RS, a->ldres, drop)
#define GRB_LN_BWD(NCH)                                                                              \
    if (a->gate) { if (drop.seed) GRB_LN_BWD2(NCH, true, true); else GRB_LN_BWD2(NCH, true, false); } \
    else { if (drop.seed) GRB_LN_BWD2(NCH, false, true); else GRB_LN_BWD2(NCH, false, false); }
    if (W == 256) { GRB_LN_BWD(1) } else if (W == 512) { GRB_LN_BWD(2) } else { GRB_LN_BWD(4) }
#undef GRB_LN_BWD
#undef GRB_LN_BWD2
    GRB_LAUNCH_OK();
    return GRB_OK;
  }
  GRB_REQUIRE(!drop.seed, GRB_ERR_UNSUPPORTED,
              "ln_gate_bwd: fused dropout needs bf16 rows of 256 / 512 / 1024 elements, 16-byte aligned");
  if (a->dtype == GRB_F32)
    ln_gate_bwd_kernel<float><<<grid, LN_WARPS * 32, 0, st>>>(
        (const float*) a->x, a->ldx, (const float*) a->gate, a->ldg, (const float*) a->y, a->ldy, a->mean, a->rstd,
        (float*) a->dx, a->lddx, (float*) a->dgate, a->lddg, rows, (int) W, (const float*) a->res, a->ldres);
  else
    ln_gate_bwd_kernel<__nv_bfloat16><<<grid, LN_WARPS * 32, 0, st>>>(
        (const __nv_bfloat16*) a->x, a->ldx, (const __nv_bfloat16*) a->gate, a->ldg,
        (const __nv_bfloat16*) a->y, a->ldy, a->mean, a->rstd, (__nv_bfloat16*) a->dx, a->lddx,
        (__nv_bfloat16*) a->dgate, a->lddg, rows, (int) W, (const __nv_bfloat16*) a->res, a->ldres);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_ln_gate_fwd(const void* x, int64_t ldx, const void* gate, int64_t ldg, void* y,
                    int64_t ldy, float* mean, float* rstd, int64_t rows, int64_t W, float eps,
                    int dtype, grb_stream_t stream) {
  grb_ln_gate_args a{};
  a.x = x; a.ldx = ldx; a.gate = gate; a.ldg = ldg; a.y = y; a.ldy = ldy; a.mean = mean; a.rstd = rstd;
  a.rows = rows; a.W = W; a.eps = eps; a.dtype = dtype;
  return grb_ln_gate_fwd_ex(&a, stream);
}

int grb_ln_gate_bwd(const void* x, int64_t ldx, const void* gate, int64_t ldg, const void* dy,
                    int64_t lddy, const float* mean, const float* rstd, void* dx, int64_t lddx,
                    void* dgate, int64_t lddg, int64_t rows, int64_t W, int dtype,
                    grb_stream_t stream) {
  grb_ln_gate_args a{};
  a.x = x; a.ldx = ldx; a.gate = gate; a.ldg = ldg; a.y = const_cast<void*>(dy); a.ldy = lddy;
  a.mean = const_cast<float*>(mean); a.rstd = const_cast<float*>(rstd);
  a.dx = dx; a.lddx = lddx; a.dgate = dgate; a.lddg = lddg; a.rows = rows; a.W = W; a.dtype = dtype;
  return grb_ln_gate_bwd_ex(&a, stream);
}

}
