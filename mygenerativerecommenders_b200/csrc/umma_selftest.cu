// umma_selftest.cu — building-block self test: every tcgen05 operand mode the attention and
// retrieval kernels rely on (K-major / MN-major shared-memory operands under the 128-byte TMA
// swizzle, A operand from TMEM) is run as one small GEMM and compared with a host reference.
// Also home of the host-side tensor-map helper.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include <cudaTypedefs.h>
#include <vector>
#include <cmath>
#include <mutex>

namespace grb {

static PFN_cuTensorMapEncodeTiled_v12000 tmap_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  });
  return encode;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols) {
  auto encode = tmap_encoder();
  GRB_REQUIRE(encode != nullptr, GRB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
  GRB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld_elems * 2) % 16 == 0,
              GRB_ERR_INVALID_ARG, "TMA needs a 16-byte aligned base and row stride");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                      gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GRB_REQUIRE(r == CUDA_SUCCESS, GRB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int) r);
  return GRB_OK;
}

// fp32 matrix, box of box_rows x 16 floats (64-byte rows, 64-byte swizzle): the shape the
// attention backward uses for its bulk reduce-add of dQ tiles.
int make_tmap_f32_2d_sw64(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                          uint64_t ld_elems, uint32_t box_rows) {
  auto encode = tmap_encoder();
  GRB_REQUIRE(encode != nullptr, GRB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
  GRB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld_elems * 4) % 16 == 0,
              GRB_ERR_INVALID_ARG, "TMA needs a 16-byte aligned base and row stride");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 4};
  cuuint32_t box[2] = {16, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim,
                      gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GRB_REQUIRE(r == CUDA_SUCCESS, GRB_ERR_CUDA, "cuTensorMapEncodeTiled (f32) failed (%d)", (int) r);
  return GRB_OK;
}

enum { OP_K_MAJOR = 0, OP_MN_MAJOR = 1, OP_TMEM = 2, OP_TMEM_F16 = 3 };

// D[128][N] = A[128][K] * B[N][K]^T with the operands staged as the mode says.
__global__ void __launch_bounds__(128) umma_probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmB,
                                                         const __nv_bfloat16* __restrict__ a_rowmajor,
                                                         float* __restrict__ D, int N, int K,
                                                         int a_mode, int b_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  using namespace ptx;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sA = smem;                       // up to 128 x 256 bf16 = 64 KiB
  uint8_t* sB = smem + 65536;               // up to 256 x 256 bf16 = 128 KiB
  const uint32_t bl = smem_u32(&bar_load), bm = smem_u32(&bar_mma);
  if (tid == 0) {
    mbar_init(bl, 1);
    mbar_init(bm, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t a_tmem_col = 256;

  if (tid == 0) {
    uint32_t bytes = 0;
    if (a_mode == OP_K_MAJOR) {
      for (int kc = 0; kc < K / 64; ++kc)
        tma_load_2d(smem_u32(sA + kc * 128 * 128), &tmA, kc * 64, 0, bl);
      bytes += 128 * K * 2;
    } else if (a_mode == OP_MN_MAJOR) {     // global [K][128]; blocks of [K rows][64 cols]
      for (int mc = 0; mc < 2; ++mc) tma_load_2d(smem_u32(sA + mc * K * 128), &tmA, mc * 64, 0, bl);
      bytes += 128 * K * 2;
    }
    if (b_mode == OP_K_MAJOR) {             // global [N][K]; blocks of [N rows][64 cols]
      for (int kc = 0; kc < K / 64; ++kc)
        tma_load_2d(smem_u32(sB + kc * N * 128), &tmB, kc * 64, 0, bl);
    } else {                                // global [K][N]; blocks of [K rows][64 cols]
      for (int nc = 0; nc < N / 64; ++nc)
        tma_load_2d(smem_u32(sB + nc * K * 128), &tmB, nc * 64, 0, bl);
    }
    bytes += N * K * 2;
    mbar_arrive_expect_tx(bl, bytes);
  }
  if (a_mode == OP_TMEM || a_mode == OP_TMEM_F16) {
    // thread = row; 16 bf16 (8 packed columns) per store
    const uint32_t lane_base = (uint32_t) (warp * 32) << 16;
    for (int ks = 0; ks < K / 16; ++ks) {
      uint32_t v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float lo = __bfloat162float(a_rowmajor[tid * K + ks * 16 + 2 * i]);
        const float hi = __bfloat162float(a_rowmajor[tid * K + ks * 16 + 2 * i + 1]);
        v[i] = a_mode == OP_TMEM_F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
      }
      tmem_st8(tmem + lane_base + a_tmem_col + ks * 8, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    mbar_wait(bl, 0);
    tc_fence_after();
    const uint32_t idesc = a_mode == OP_TMEM_F16
        ? make_idesc_f16a_bf16b(128, N, false, b_mode == OP_MN_MAJOR)
        : make_idesc_bf16(128, N, a_mode == OP_MN_MAJOR, b_mode == OP_MN_MAJOR);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t bdesc;
      if (b_mode == OP_K_MAJOR)
        bdesc = make_smem_desc_sw128(smem_u32(sB + (ks / 4) * N * 128 + (ks % 4) * 32), 0, 1024);
      else
        bdesc = make_smem_desc_sw128(smem_u32(sB + ks * 2048), K * 128, 1024);
      if (a_mode == OP_TMEM || a_mode == OP_TMEM_F16) {
        umma_ts(tmem, tmem + a_tmem_col + ks * 8, bdesc, idesc, ks > 0);
      } else {
        uint64_t adesc;
        if (a_mode == OP_K_MAJOR)
          adesc = make_smem_desc_sw128(smem_u32(sA + (ks / 4) * 128 * 128 + (ks % 4) * 32), 0, 1024);
        else
          adesc = make_smem_desc_sw128(smem_u32(sA + ks * 2048), K * 128, 1024);
        umma_ss(tmem, adesc, bdesc, idesc, ks > 0);
      }
    }
    umma_commit(bm);
  }
  mbar_wait(bm, 0);
  tc_fence_after();
  const uint32_t lane_base = (uint32_t) (warp * 32) << 16;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + lane_base + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) D[tid * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

struct ProbeMode { int a_mode, b_mode, N, K; };

}  // namespace grb

using namespace grb;

extern "C" int grb_selftest_umma(float* errs, int max_modes, grb_stream_t stream) {
  static const ProbeMode modes[] = {
      {OP_K_MAJOR, OP_K_MAJOR, 128, 128},   // 0: S = Q K^T
      {OP_K_MAJOR, OP_MN_MAJOR, 64, 128},   // 1: O = P V, P in smem
      {OP_TMEM, OP_MN_MAJOR, 64, 128},      // 2: O = P V, P in TMEM
      {OP_TMEM, OP_K_MAJOR, 128, 64},       // 3: A from TMEM, K-major B
      {OP_MN_MAJOR, OP_MN_MAJOR, 64, 128},  // 4: dV = P^T dO (both MN-major)
      {OP_K_MAJOR, OP_K_MAJOR, 256, 256},   // 5: retrieval score tile
      {OP_MN_MAJOR, OP_K_MAJOR, 128, 64},   // 6: MN-major A with K-major B
      // (probed once: an fp16 A operand against a bf16 B operand raises "illegal instruction" on
      //  sm_100a — kind::f16 wants both operands in the same 16-bit format — so P stays bf16)
  };
  const int n_modes = (int) (sizeof(modes) / sizeof(modes[0]));
  GRB_REQUIRE(errs != nullptr && max_modes > 0, GRB_ERR_INVALID_ARG, "selftest: bad arguments");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = 65536 + 131072;
  GRB_CUDA_OK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int) smem));
  int run = 0;
  for (int m = 0; m < n_modes && m < max_modes; ++m, ++run) {
    const ProbeMode pm = modes[m];
    const int M = 128, N = pm.N, K = pm.K;
    std::vector<float> A(M * K), B(N * K);
    uint32_t s = 12345u + m;
    auto rnd = [&s]() { s = s * 1664525u + 1013904223u; return (float) ((int) ((s >> 16) % 33) - 16) / 8.0f; };
    for (auto& v : A) v = rnd();
    for (auto& v : B) v = rnd();
    // device layouts
    std::vector<__nv_bfloat16> Ad(M * K), Bd(N * K), Arow(M * K);
    for (int i = 0; i < M; ++i)
      for (int k = 0; k < K; ++k) {
        Arow[i * K + k] = __float2bfloat16(A[i * K + k]);
        if (pm.a_mode == OP_MN_MAJOR) Ad[k * M + i] = Arow[i * K + k];
        else Ad[i * K + k] = Arow[i * K + k];
      }
    for (int j = 0; j < N; ++j)
      for (int k = 0; k < K; ++k) {
        if (pm.b_mode == OP_MN_MAJOR) Bd[k * N + j] = __float2bfloat16(B[j * K + k]);
        else Bd[j * K + k] = __float2bfloat16(B[j * K + k]);
      }
    __nv_bfloat16 *dA = nullptr, *dB = nullptr, *dArow = nullptr;
    float* dD = nullptr;
    GRB_CUDA_OK(cudaMalloc(&dA, Ad.size() * 2));
    GRB_CUDA_OK(cudaMalloc(&dB, Bd.size() * 2));
    GRB_CUDA_OK(cudaMalloc(&dArow, Arow.size() * 2));
    GRB_CUDA_OK(cudaMalloc(&dD, (size_t) M * N * 4));
    GRB_CUDA_OK(cudaMemcpyAsync(dA, Ad.data(), Ad.size() * 2, cudaMemcpyHostToDevice, st));
    GRB_CUDA_OK(cudaMemcpyAsync(dB, Bd.data(), Bd.size() * 2, cudaMemcpyHostToDevice, st));
    GRB_CUDA_OK(cudaMemcpyAsync(dArow, Arow.data(), Arow.size() * 2, cudaMemcpyHostToDevice, st));
    GRB_CUDA_OK(cudaMemsetAsync(dD, 0xff, (size_t) M * N * 4, st));
    CUtensorMap tmA, tmB;
    int rc;
    if (pm.a_mode == OP_MN_MAJOR) rc = make_tmap_bf16_2d(&tmA, dA, K, M, M, K);
    else rc = make_tmap_bf16_2d(&tmA, dA, M, K, K, M);
    if (rc != GRB_OK) return rc;
    if (pm.b_mode == OP_MN_MAJOR) rc = make_tmap_bf16_2d(&tmB, dB, K, N, N, K);
    else rc = make_tmap_bf16_2d(&tmB, dB, N, K, K, N);
    if (rc != GRB_OK) return rc;
    umma_probe_kernel<<<1, 128, smem, st>>>(tmA, tmB, dArow, dD, N, K, pm.a_mode, pm.b_mode);
    GRB_LAUNCH_OK();
    std::vector<float> Dh((size_t) M * N);
    GRB_CUDA_OK(cudaMemcpyAsync(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost, st));
    GRB_CUDA_OK(cudaStreamSynchronize(st));
    float maxerr = 0.f;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        float ref = 0.f;
        for (int k = 0; k < K; ++k) ref += A[i * K + k] * B[j * K + k];
        float e = std::fabs(ref - Dh[(size_t) i * N + j]);
        if (!(e == e)) e = 1e30f;  // NaN
        if (e > maxerr) maxerr = e;
      }
    errs[m] = maxerr;
    cudaFree(dA); cudaFree(dB); cudaFree(dArow); cudaFree(dD);
  }
  return run;
}
