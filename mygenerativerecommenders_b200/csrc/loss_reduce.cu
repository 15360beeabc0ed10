// loss_reduce.cu — the last line of the sampled-softmax loss, autoregressive_losses.py:306 (and :172):
//     (jagged_loss * supervision_weights).sum() / supervision_weights.sum()
// as one launch each way.  On torch ops it is a multiply, two reductions and a division forward, and a
// fill, two multiplies and an expand backward: seven launches of 2-7 us around 14 k numbers.
#include "common.cuh"

namespace grb {
namespace {

constexpr int WM_THREADS = 1024;

// out[0] = sum(x * w) / sum(w), out[1] = sum(w).  One CTA: fixed summation order (deterministic).
__global__ void __launch_bounds__(WM_THREADS) weighted_mean_fwd_kernel(const float* __restrict__ x,
                                                                       const float* __restrict__ w, int64_t n,
                                                                       float* __restrict__ out) {
  __shared__ float s_xw[WM_THREADS / 32], s_w[WM_THREADS / 32];
  float a = 0.f, b = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += WM_THREADS) {
    const float wi = w[i];
    a = fmaf(x[i], wi, a);
    b += wi;
  }
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_xw[warp] = a; s_w[warp] = b; }
  __syncthreads();
  if (warp == 0) {
    a = warp_sum(s_xw[lane]);
    b = warp_sum(s_w[lane]);
    if (lane == 0) { out[0] = a / b; out[1] = b; }
  }
}

// gx[i] = gout * w[i] / sum(w)
__global__ void __launch_bounds__(256) weighted_mean_bwd_kernel(const float* __restrict__ w,
                                                                const float* __restrict__ out,
                                                                const float* __restrict__ gout, int64_t n,
                                                                float* __restrict__ gx) {
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gx[i] = gout[0] * w[i] / out[1];
}

}  // namespace
}  // namespace grb

using namespace grb;

extern "C" {

int grb_weighted_mean_fwd(const float* x, const float* w, int64_t n, float* out2, grb_stream_t stream) {
  GRB_REQUIRE(x && w && out2 && n >= 0, GRB_ERR_INVALID_ARG, "weighted_mean_fwd: bad arguments");
  weighted_mean_fwd_kernel<<<1, WM_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, w, n, out2);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_weighted_mean_bwd(const float* w, const float* out2, const float* gout, int64_t n, float* gx,
                          grb_stream_t stream) {
  GRB_REQUIRE(w && out2 && gout && gx && n >= 0, GRB_ERR_INVALID_ARG, "weighted_mean_bwd: bad arguments");
  if (n == 0) return GRB_OK;
  weighted_mean_bwd_kernel<<<(unsigned) ceil_div(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w, out2, gout, n, gx);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}
