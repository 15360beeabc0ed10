// silu_split.cu — the activation between the UVQK projection and its consumers
// (hstu.py:304-320: batched_mm_output = F.silu(torch.mm(normed_x, _uvqk)); u, v, q, k = split).
//
// Forward: y = x / (1 + exp(-x)) over a (rows, W) matrix, fp32 math.  Backward: the four consumers
// (gate, attention V / Q / K) hand back four separate (rows, w_i) gradients; autograd would first
// concatenate them (one full pass) and then run silu_backward (another): here one pass reads the four
// pieces in place and writes d pre-activation:  dx = dy * s * (1 + x * (1 - s)),  s = sigmoid(x).
// HBM-bound: forward 2, backward 3 passes over rows * W elements, 16-byte accesses.
#include "common.cuh"

namespace grb {
namespace {

template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int E = 4;
  static __device__ __forceinline__ void load(const float* p, float (&f)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int E = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 v = __bfloat1622float2(h[e]);
      f[2 * e] = v.x; f[2 * e + 1] = v.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 o;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
    o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
    o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(p) = o;
  }
};

constexpr int SS_UNROLL = 2;

// sigmoid in the precision the output can hold: exact expf + IEEE division for fp32 tensors (parity
// with ATen at 1e-6); for bf16 tensors ex2.approx + rcp.approx (two MUFU ops, ~1e-6 relative error at
// every x, far below bf16's 2^-8 rounding).  With the exact form the kernel is ALU-bound (~25
// instructions per element: 19 us forward / 43 us backward at 14k x 1024 against 14 / 27 us for ATen's
// silu and cat + silu_backward); with the fast form it is 9 / 17 us.
template <typename T> __device__ __forceinline__ float sigmoid_of(float x);
template <> __device__ __forceinline__ float sigmoid_of<float>(float x) { return 1.0f / (1.0f + expf(-x)); }
template <> __device__ __forceinline__ float sigmoid_of<__nv_bfloat16>(float x) {
  return __fdividef(1.0f, 1.0f + __expf(-x));
}

struct SplitGrads {          // up to four column blocks [begin_i, begin_{i+1}) of the (rows, W) matrix
  const void* g[4];          // (rows, w_i), row stride ld[i] elements; NULL = zero gradient
  int64_t ld[4];
  int begin[5];
  int n;
};

// (row, first element) of 16-byte chunk i of a (rows, cpr chunks) matrix.  Chunk counts below 2^31 (every
// real shape) take a 32-bit division: the 64-bit one is ~100 instructions, paid per chunk, and made these
// streaming kernels instruction-bound.
__device__ __forceinline__ void chunk_row_col(int64_t i, int cpr, bool small, int64_t& r, int& c) {
  if (small) {
    const uint32_t q = (uint32_t) i / (uint32_t) cpr;
    r = q;
    c = (int) ((uint32_t) i - q * (uint32_t) cpr);
  } else {
    r = i / cpr;
    c = (int) (i - r * cpr);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) silu_fwd_kernel(const T* __restrict__ x, int64_t ldx, T* __restrict__ y,
                                                       int64_t ldy, int64_t rows, int W) {
  constexpr int E = Vec16<T>::E;
  const int cpr = W / E;                                   // 16-byte chunks per row
  const int64_t total = rows * cpr;
  // SS_UNROLL independent 16-byte loads per thread before the first use
  const int64_t base = (int64_t) blockIdx.x * (256 * SS_UNROLL) + threadIdx.x;
  float f[SS_UNROLL][E];
  int64_t off_y[SS_UNROLL];
#pragma unroll
  for (int u = 0; u < SS_UNROLL; ++u) {
    const int64_t i = base + u * 256;
    off_y[u] = -1;
    if (i < total) {
      int64_t r;
      int c;
      chunk_row_col(i, cpr, total < (1ll << 31), r, c);
      c *= E;
      Vec16<T>::load(x + r * ldx + c, f[u]);
      off_y[u] = r * ldy + c;
    }
  }
#pragma unroll
  for (int u = 0; u < SS_UNROLL; ++u) {
    if (off_y[u] < 0) continue;
#pragma unroll
    for (int e = 0; e < E; ++e)
      f[u][e] = sizeof(T) == 4 ? silu_f32(f[u][e]) : f[u][e] * sigmoid_of<T>(f[u][e]);
    Vec16<T>::store(y + off_y[u], f[u]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) silu_split_bwd_kernel(const T* __restrict__ x, int64_t ldx,
                                                             const __grid_constant__ SplitGrads sg,
                                                             T* __restrict__ dx, int64_t lddx, int64_t rows, int W) {
  constexpr int E = Vec16<T>::E;
  const int cpr = W / E;
  const int64_t total = rows * cpr;
  const int64_t base = (int64_t) blockIdx.x * (256 * SS_UNROLL) + threadIdx.x;
  float xf[SS_UNROLL][E], gf[SS_UNROLL][E];
  int64_t off_d[SS_UNROLL];
  bool has[SS_UNROLL];
#pragma unroll
  for (int u = 0; u < SS_UNROLL; ++u) {
    const int64_t i = base + u * 256;
    off_d[u] = -1;
    has[u] = false;
    if (i < total) {
      int64_t r;
      int c;
      chunk_row_col(i, cpr, total < (1ll << 31), r, c);
      c *= E;
      int blk = 0;
#pragma unroll
      for (int b = 1; b < 4; ++b) blk += (b < sg.n && c >= sg.begin[b]) ? 1 : 0;
      Vec16<T>::load(x + r * ldx + c, xf[u]);
      const T* g = reinterpret_cast<const T*>(sg.g[blk]);
      if (g) {
        Vec16<T>::load(g + r * sg.ld[blk] + (c - sg.begin[blk]), gf[u]);
        has[u] = true;
      }
      off_d[u] = r * lddx + c;
    }
  }
#pragma unroll
  for (int u = 0; u < SS_UNROLL; ++u) {
    if (off_d[u] < 0) continue;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float sg_ = sigmoid_of<T>(xf[u][e]);
      gf[u][e] = has[u] ? gf[u][e] * (sg_ * (1.0f + xf[u][e] * (1.0f - sg_))) : 0.f;
    }
    Vec16<T>::store(dx + off_d[u], gf[u]);
  }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace grb

extern "C" int grb_silu_fwd(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int32_t W,
                            int32_t dtype, grb_stream_t stream) {
  using namespace grb;
  GRB_REQUIRE(rows >= 0 && W > 0, GRB_ERR_INVALID_ARG, "silu_fwd: bad sizes");
  GRB_REQUIRE(dtype == GRB_F32 || dtype == GRB_BF16, GRB_ERR_INVALID_ARG, "silu_fwd: dtype");
  if (rows == 0) return GRB_OK;
  GRB_REQUIRE(x && y, GRB_ERR_INVALID_ARG, "silu_fwd: null tensor");
  const int es = dtype == GRB_F32 ? 4 : 2, E = 16 / es;
  GRB_REQUIRE(W % E == 0 && (ldx * es) % 16 == 0 && (ldy * es) % 16 == 0 && al16(x) && al16(y),
              GRB_ERR_UNSUPPORTED, "silu_fwd: rows must be 16-byte aligned and W a multiple of %d", E);
  const int64_t total = rows * (W / E);
  const unsigned grid = (unsigned) ceil_div(total, 256 * SS_UNROLL);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == GRB_F32)
    silu_fwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), ldx, static_cast<float*>(y), ldy, rows, W);
  else
    silu_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ldx,
                                                         static_cast<__nv_bfloat16*>(y), ldy, rows, W);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

extern "C" int grb_silu_split_bwd(const void* x, int64_t ldx, int32_t n_blocks, const void* const* grads,
                                  const int64_t* ld_grads, const int32_t* widths, void* dx, int64_t lddx,
                                  int64_t rows, int32_t dtype, grb_stream_t stream) {
  using namespace grb;
  GRB_REQUIRE(rows >= 0 && n_blocks >= 1 && n_blocks <= 4 && grads && ld_grads && widths, GRB_ERR_INVALID_ARG,
              "silu_split_bwd: 1..4 column blocks");
  GRB_REQUIRE(dtype == GRB_F32 || dtype == GRB_BF16, GRB_ERR_INVALID_ARG, "silu_split_bwd: dtype");
  if (rows == 0) return GRB_OK;
  GRB_REQUIRE(x && dx, GRB_ERR_INVALID_ARG, "silu_split_bwd: null tensor");
  const int es = dtype == GRB_F32 ? 4 : 2, E = 16 / es;
  SplitGrads sg{};
  sg.n = n_blocks;
  int W = 0;
  for (int i = 0; i < n_blocks; ++i) {
    GRB_REQUIRE(widths[i] > 0 && widths[i] % E == 0, GRB_ERR_UNSUPPORTED,
                "silu_split_bwd: block widths must be multiples of %d", E);
    GRB_REQUIRE(grads[i] == nullptr || (al16(grads[i]) && (ld_grads[i] * es) % 16 == 0), GRB_ERR_UNSUPPORTED,
                "silu_split_bwd: gradient %d is not 16-byte aligned", i);
    sg.g[i] = grads[i]; sg.ld[i] = ld_grads[i]; sg.begin[i] = W;
    W += widths[i];
  }
  sg.begin[n_blocks] = W;
  GRB_REQUIRE((ldx * es) % 16 == 0 && (lddx * es) % 16 == 0 && al16(x) && al16(dx), GRB_ERR_UNSUPPORTED,
              "silu_split_bwd: rows must be 16-byte aligned");
  const int64_t total = rows * (W / E);
  const unsigned grid = (unsigned) ceil_div(total, 256 * SS_UNROLL);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == GRB_F32)
    silu_split_bwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), ldx, sg,
                                                       static_cast<float*>(dx), lddx, rows, W);
  else
    silu_split_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ldx, sg,
                                                               static_cast<__nv_bfloat16*>(dx), lddx, rows, W);
  GRB_LAUNCH_OK();
  return GRB_OK;
}
