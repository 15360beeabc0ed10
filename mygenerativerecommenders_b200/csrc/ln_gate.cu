// ln_gate.cu — y = gate * LayerNorm(x) (no affine), one warp per jagged row, warp-shuffle
// reductions.  Reference: hstu.py:258-264 (_norm_input / _norm_attn_output) and :402
// (o_input = u * norm(attn_output)).  gate == NULL gives the plain LayerNorm of :300.
// HBM-bound: fwd reads x (+gate) once from HBM (re-reads hit L1), writes y.
#include "common.cuh"

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
    int64_t lddg, int64_t rows, int W) {
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
    stf<T>(dxr + c, rs * (dxh - m1 - xh * m2));
    if (dgr) stf<T>(dgr + c, dyv * xh);
  }
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

int grb_ln_gate_fwd(const void* x, int64_t ldx, const void* gate, int64_t ldg, void* y,
                    int64_t ldy, float* mean, float* rstd, int64_t rows, int64_t W, float eps,
                    int dtype, grb_stream_t stream) {
  GRB_REQUIRE(x && y && mean && rstd && rows >= 0 && W > 0 && W < (1 << 30), GRB_ERR_INVALID_ARG,
              "ln_gate_fwd: bad arguments");
  GRB_REQUIRE(dtype == GRB_F32 || dtype == GRB_BF16, GRB_ERR_INVALID_ARG, "ln_gate_fwd: dtype");
  if (rows == 0) return GRB_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned) ceil_div(rows, LN_WARPS);
  if (dtype == GRB_F32)
    ln_gate_fwd_kernel<float><<<grid, LN_WARPS * 32, 0, st>>>(
        (const float*) x, ldx, (const float*) gate, ldg, (float*) y, ldy, mean, rstd, rows,
        (int) W, eps);
  else
    ln_gate_fwd_kernel<__nv_bfloat16><<<grid, LN_WARPS * 32, 0, st>>>(
        (const __nv_bfloat16*) x, ldx, (const __nv_bfloat16*) gate, ldg, (__nv_bfloat16*) y, ldy,
        mean, rstd, rows, (int) W, eps);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_ln_gate_bwd(const void* x, int64_t ldx, const void* gate, int64_t ldg, const void* dy,
                    int64_t lddy, const float* mean, const float* rstd, void* dx, int64_t lddx,
                    void* dgate, int64_t lddg, int64_t rows, int64_t W, int dtype,
                    grb_stream_t stream) {
  GRB_REQUIRE(x && dy && dx && mean && rstd && rows >= 0 && W > 0, GRB_ERR_INVALID_ARG,
              "ln_gate_bwd: bad arguments");
  GRB_REQUIRE((gate == nullptr) == (dgate == nullptr), GRB_ERR_INVALID_ARG,
              "ln_gate_bwd: gate and dgate must both be null or both non-null");
  GRB_REQUIRE(dtype == GRB_F32 || dtype == GRB_BF16, GRB_ERR_INVALID_ARG, "ln_gate_bwd: dtype");
  if (rows == 0) return GRB_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned) ceil_div(rows, LN_WARPS);
  if (dtype == GRB_F32)
    ln_gate_bwd_kernel<float><<<grid, LN_WARPS * 32, 0, st>>>(
        (const float*) x, ldx, (const float*) gate, ldg, (const float*) dy, lddy, mean, rstd,
        (float*) dx, lddx, (float*) dgate, lddg, rows, (int) W);
  else
    ln_gate_bwd_kernel<__nv_bfloat16><<<grid, LN_WARPS * 32, 0, st>>>(
        (const __nv_bfloat16*) x, ldx, (const __nv_bfloat16*) gate, ldg,
        (const __nv_bfloat16*) dy, lddy, mean, rstd, (__nv_bfloat16*) dx, lddx,
        (__nv_bfloat16*) dgate, lddg, rows, (int) W);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}
