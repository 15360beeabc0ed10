// proj_gemm.cu — the HSTU layer's projections on Blackwell tensor cores, with the elementwise work
// around them fused into the epilogue.
//
// Reference: /root/reference/src/generative_recommenders_pl/models/sequential_encoders/hstu.py
//   :302-320  batched_mm_output = mm(normed_x, _uvqk) ; F.silu ; torch.split        (UVQK projection)
//   :404-413  _o(dropout(u * norm(attn))) + x                                        (output projection)
// and their backward (dgrad + wgrad).  Six GEMM shapes per layer, one kernel:
//
//   C[M, N] = epilogue( A[M, K] * B[K, N] ),  bf16 operands, fp32 accumulation in TMEM.
//     A: "K-major" = stored (M, K) row-major, or "MN-major" = stored (K, M) row-major (a transposed
//        operand read in place: the weight-gradient GEMMs contract over the token dimension);
//     B: "K-major" = stored (N, K) row-major (nn.Linear weights), or "MN-major" = stored (K, N)
//        row-major (_uvqk, activations).
//   Epilogues: PLAIN  bf16 C                                   (both dgrads)
//              SILU2  bf16 C and bf16 SiLU(C)                  (UVQK forward: the backward needs C)
//              BIAS_RES  bf16 (C + bias[n] + residual[m, n])   (output projection forward)
//              F32_ADD  fp32 C added into the output with vector reds (weight gradients, split-K)
//
// Persistent, one CTA per SM, tile 128 x 256, K step 64:
//   warp 0 (TMA) : A and B slabs into a 4-stage ring (128-byte swizzle; 1 box for a K-major operand,
//                  one 64x64 box per 64-wide chunk for an MN-major one)
//   warp 1 (MMA) : tcgen05.mma M128 N256 K16 x 4 per stage into one of two 256-column TMEM
//                  accumulators, so the epilogue of tile i overlaps the main loop of tile i + 1
//   warps 2..17  : epilogue, thread = output row, warpgroup g = columns [64 g, 64 g + 64) (16 warps:
//                  with 8 the epilogue ran at 27 % issue utilisation, two warps per scheduler)
// Tiles are ordered n-fastest, so the CTAs running concurrently share A slabs through L2.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "hstu_attn_sm100.cuh"

namespace grb {

using namespace ptx;

namespace {

constexpr int PG_BM = 128, PG_BN = 256, PG_BK = 64;
constexpr int PG_STAGES = 4;
constexpr int PG_A_BYTES = PG_BM * PG_BK * 2;          // 16 KiB
constexpr int PG_B_BYTES = PG_BN * PG_BK * 2;          // 32 KiB
constexpr int PG_STAGE_BYTES = PG_A_BYTES + PG_B_BYTES;
constexpr int PG_EPI_WARPS = 16;
constexpr int PG_THREADS = 64 + 32 * PG_EPI_WARPS;

struct PgSmem {
  static constexpr int ring = 0;
  static constexpr int stage = ring + PG_STAGES * PG_STAGE_BYTES;   // per epilogue warp [32 rows][64 bytes]
  static constexpr int bars = stage + PG_EPI_WARPS * 2048;
  static constexpr int total = bars + 256;
};
static_assert(PgSmem::total + 1024 <= 232448, "shared memory budget");

struct PgParams {
  int64_t M, N, K;
  int a_mn, b_mn, epi;
  int n_mt, n_nt, splits;          // tiles along M, along N, K splits (F32_ADD only)
  int64_t k_per_split;             // multiple of PG_BK
  void* out0; int64_t ldo0;
  void* out1; int64_t ldo1;
  const float* bias;
  const __nv_bfloat16* res; int64_t ldres;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__device__ __forceinline__ float silu_fast(float x) {
  // x * sigmoid(x), sigmoid through ex2.approx + rcp.approx (as csrc/silu_split.cu for bf16 tensors)
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

__global__ void __launch_bounds__(PG_THREADS, 1) proj_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   PgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  using L = PgSmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  const uint32_t bar_full = smem_u32(bars);                         // [STAGES]
  const uint32_t bar_empty = smem_u32(bars + PG_STAGES);            // [STAGES]
  const uint32_t bar_acc_full = smem_u32(bars + 2 * PG_STAGES);     // [2]
  const uint32_t bar_acc_empty = smem_u32(bars + 2 * PG_STAGES + 2);   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * PG_STAGES + 4);

  if (tid == 0) {
    for (int s = 0; s < PG_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_acc_full + 8 * s, 1); mbar_init(bar_acc_empty + 8 * s, PG_EPI_WARPS); }
    fence_barrier_init();
    prefetch_tensormap(&tmA); prefetch_tensormap(&tmB);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int64_t n_tiles = (int64_t) p.n_mt * p.n_nt * p.splits;
  // tile w -> (m tile, split, n tile), n fastest
  auto decode = [&](int64_t w, int& mt, int& nt, int& sp) {
    nt = (int) (w % p.n_nt);
    const int64_t rest = w / p.n_nt;
    sp = (int) (rest % p.splits);
    mt = (int) (rest / p.splits);
  };
  auto k_range = [&](int sp, int64_t& k0, int& nkb) {
    k0 = (int64_t) sp * p.k_per_split;
    int64_t k1 = k0 + p.k_per_split;
    if (k1 > p.K) k1 = p.K;
    nkb = k1 > k0 ? (int) ((k1 - k0 + PG_BK - 1) / PG_BK) : 0;
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t w = blockIdx.x; w < n_tiles; w += gridDim.x) {
        int mt, nt, sp, nkb;
        int64_t k0;
        decode(w, mt, nt, sp);
        k_range(sp, k0, nkb);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t st = it % PG_STAGES;
          mbar_wait_parked(bar_empty + 8 * st, ((it / PG_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(bar_full + 8 * st, PG_STAGE_BYTES);
          const uint32_t sa = smem_u32(smem + L::ring + st * PG_STAGE_BYTES);
          const uint32_t sb = sa + PG_A_BYTES;
          const int kk = (int) (k0 + (int64_t) kb * PG_BK);
          if (!p.a_mn) {
            tma_load_2d(sa, &tmA, kk, mt * PG_BM, bar_full + 8 * st);                 // box 128 rows x 64 k
          } else {
            for (int mc = 0; mc < PG_BM / 64; ++mc)                                      // boxes 64 k x 64 m
              tma_load_2d(sa + mc * 8192, &tmA, mt * PG_BM + mc * 64, kk, bar_full + 8 * st);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &tmB, kk, nt * PG_BN, bar_full + 8 * st);                 // box 256 rows x 64 k
          } else {
            for (int nc = 0; nc < PG_BN / 64; ++nc)                                      // boxes 64 k x 64 n
              tma_load_2d(sb + nc * 8192, &tmB, nt * PG_BN + nc * 64, kk, bar_full + 8 * st);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp, one elected lane issues) =================
    const uint32_t idesc = make_idesc_bf16(PG_BM, PG_BN, p.a_mn != 0, p.b_mn != 0);
    // K-major operand: rows of 128 bytes, 8-row groups 1024 bytes apart, 16 k = 32 bytes
    // MN-major operand: 64-wide chunks 8192 bytes apart (LBO), 8 k-rows 1024 bytes, 16 k = 2048 bytes
    const uint32_t a_lbo = p.a_mn ? 8192u : 0u, b_lbo = p.b_mn ? 8192u : 0u;
    const uint32_t a_step = p.a_mn ? 2048u : 32u, b_step = p.b_mn ? 2048u : 32u;
    const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smem + L::ring), a_lbo, 1024);
    const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smem + L::ring) + PG_A_BYTES, b_lbo, 1024);
    auto adv = [](uint64_t d, uint32_t bytes) { return d + (uint64_t) (bytes >> 4); };
    uint32_t it = 0, tile = 0;
    for (int64_t w = blockIdx.x; w < n_tiles; w += gridDim.x, ++tile) {
      int mt, nt, sp, nkb;
      int64_t k0;
      decode(w, mt, nt, sp);
      k_range(sp, k0, nkb);
      const uint32_t ab = tile & 1;
      mbar_wait_parked(bar_acc_empty + 8 * ab, ((tile >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t st = it % PG_STAGES;
        mbar_wait_parked(bar_full + 8 * st, (it / PG_STAGES) & 1);
        tc_fence_after();
        const uint64_t ad = adv(a_desc0, st * PG_STAGE_BYTES), bd = adv(b_desc0, st * PG_STAGE_BYTES);
#pragma unroll
        for (int ks = 0; ks < PG_BK / 16; ++ks)
          umma_ss_warp(tmem + ab * PG_BN, adv(ad, ks * a_step), adv(bd, ks * b_step), idesc, (kb > 0) || (ks > 0));
        umma_commit_warp(bar_empty + 8 * st);
      }
      umma_commit_warp(bar_acc_full + 8 * ab);
    }
  } else {
    // ================= epilogue =================
    const int wq = warp & 3;
    const int g = (warp - 2) >> 2;                    // columns [64 g, 64 g + 64) of the tile
    const int r = (wq << 5) | lane;
    const uint32_t lane_base = (uint32_t) (wq * 32) << 16;
    uint32_t tile = 0;
    for (int64_t w = blockIdx.x; w < n_tiles; w += gridDim.x, ++tile) {
      int mt, nt, sp, nkb;
      int64_t k0;
      decode(w, mt, nt, sp);
      k_range(sp, k0, nkb);
      const uint32_t ab = tile & 1;
      mbar_wait(bar_acc_full + 8 * ab, (tile >> 1) & 1);
      tc_fence_after();
      // The accumulator arrives with thread = row: stored as it is, one instruction would touch 32
      // different 128-byte lines (32 LSU wavefronts; the first version of this epilogue spent 16 k
      // cycles per tile on its stores).  Each 32-column chunk is therefore transposed through a
      // per-warp staging block (16-byte chunk c of row rr at chunk c ^ ((rr >> 1) & mask): conflict
      // free both ways) and leaves as row segments of 64 (bf16) or 128 (fp32) contiguous bytes.
      const int64_t row0 = (int64_t) mt * PG_BM + 32 * wq;          // first row of this warp
      const int64_t col0 = (int64_t) nt * PG_BN + 64 * g;
      uint8_t* stg = smem + L::stage + (warp - 2) * 2048;
      // bf16 chunk [32 rows][64 B]: lane -> row 8 i + lane / 4, 16-byte chunk lane % 4   (i = 0..3)
      auto flush_bf16 = [&](const uint32_t (&w)[16], __nv_bfloat16* base, int64_t ld, int64_t col) {
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
              make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = 8 * i + (lane >> 2), c = lane & 3;
          const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((c ^ ((rr >> 1) & 3)) << 4));
          if (row0 + rr < p.M) *reinterpret_cast<uint4*>(base + (row0 + rr) * ld + col + 8 * c) = v;
        }
      };
#pragma unroll 1
      for (int c32 = 0; c32 < 2; ++c32) {
        uint32_t acc[32];
        if (nkb > 0) {
          tmem_ld32(tmem + lane_base + ab * PG_BN + 64 * g + 32 * c32, acc);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] = 0u;
        }
        const int64_t col = col0 + 32 * c32;
        if (p.epi == GRB_GEMM_EPI_F32_ADD) {
          // fp32: two passes of 16 columns = 64 bytes per row through the same staging block
          float* o = reinterpret_cast<float*>(p.out0);
#pragma unroll
          for (int h16 = 0; h16 < 2; ++h16) {
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
                  make_uint4(acc[16 * h16 + 4 * c], acc[16 * h16 + 4 * c + 1], acc[16 * h16 + 4 * c + 2],
                             acc[16 * h16 + 4 * c + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = 8 * i + (lane >> 2), c = lane & 3;
              const float4 v = *reinterpret_cast<const float4*>(stg + rr * 64 + ((c ^ ((rr >> 1) & 3)) << 4));
              if (row0 + rr < p.M) red_add_v4(o + (row0 + rr) * p.ldo0 + col + 16 * h16 + 4 * c, v.x, v.y, v.z, v.w);
            }
          }
        } else if (p.epi == GRB_GEMM_EPI_SILU2) {
          uint32_t a16[16];
#pragma unroll
          for (int e2 = 0; e2 < 16; ++e2)
            a16[e2] = pack_bf16x2(__uint_as_float(acc[2 * e2]), __uint_as_float(acc[2 * e2 + 1]));
          flush_bf16(a16, reinterpret_cast<__nv_bfloat16*>(p.out0), p.ldo0, col);
#pragma unroll
          for (int e2 = 0; e2 < 16; ++e2) {
            // SiLU of the bf16-rounded pre-activation: what a separate silu pass over the stored
            // tensor computes, so forward and backward see the same x
            const float xr0 = __uint_as_float(a16[e2] << 16), xr1 = __uint_as_float(a16[e2] & 0xffff0000u);
            a16[e2] = pack_bf16x2(silu_fast(xr0), silu_fast(xr1));
          }
          flush_bf16(a16, reinterpret_cast<__nv_bfloat16*>(p.out1), p.ldo1, col);
        } else {
          const bool br = p.epi == GRB_GEMM_EPI_BIAS_RES;
          uint32_t o16[16];
          if (br && p.res) {
            // the residual chunk comes in through the same staging block, coalesced
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = 8 * i + (lane >> 2), c = lane & 3;
              uint4 v = make_uint4(0u, 0u, 0u, 0u);
              if (row0 + rr < p.M) v = *reinterpret_cast<const uint4*>(p.res + (row0 + rr) * p.ldres + col + 8 * c);
              *reinterpret_cast<uint4*>(stg + rr * 64 + ((c ^ ((rr >> 1) & 3)) << 4)) = v;
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 v = *reinterpret_cast<const uint4*>(stg + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4));
              const uint32_t rw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e2 = 0; e2 < 4; ++e2) {
                acc[8 * c + 2 * e2] = __float_as_uint(__uint_as_float(acc[8 * c + 2 * e2]) + __uint_as_float(rw[e2] << 16));
                acc[8 * c + 2 * e2 + 1] =
                    __float_as_uint(__uint_as_float(acc[8 * c + 2 * e2 + 1]) + __uint_as_float(rw[e2] & 0xffff0000u));
              }
            }
          }
          if (br && p.bias) {
#pragma unroll
            for (int v4 = 0; v4 < 8; ++v4) {
              const float4 bv = *reinterpret_cast<const float4*>(p.bias + col + 4 * v4);
              acc[4 * v4] = __float_as_uint(__uint_as_float(acc[4 * v4]) + bv.x);
              acc[4 * v4 + 1] = __float_as_uint(__uint_as_float(acc[4 * v4 + 1]) + bv.y);
              acc[4 * v4 + 2] = __float_as_uint(__uint_as_float(acc[4 * v4 + 2]) + bv.z);
              acc[4 * v4 + 3] = __float_as_uint(__uint_as_float(acc[4 * v4 + 3]) + bv.w);
            }
          }
#pragma unroll
          for (int e2 = 0; e2 < 16; ++e2)
            o16[e2] = pack_bf16x2(__uint_as_float(acc[2 * e2]), __uint_as_float(acc[2 * e2 + 1]));
          flush_bf16(o16, reinterpret_cast<__nv_bfloat16*>(p.out0), p.ldo0, col);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * ab);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// column sums of a bf16 (rows, W) matrix into fp32 out[W] (+=): the bias gradient of nn.Linear
// (hstu.py:404-413 backward).  Vector path (W % 8 == 0, aligned): lane = 8 columns (one 16-byte load
// per row), warp w of the block takes rows r0 + w, r0 + w + 8, ... with all its loads independent, the
// 8 warps meet in shared memory, one red per (block, column).  One block per ~SM-count slice of rows.
constexpr int CS_ROWS = 104;
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                          int64_t rows, int W, float* __restrict__ out, int vec) {
  const int64_t r0 = (int64_t) blockIdx.x * CS_ROWS;
  const int64_t r1 = r0 + CS_ROWS < rows ? r0 + CS_ROWS : rows;
  if (!vec) {
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
      float acc = 0.f;
      for (int64_t r = r0; r < r1; ++r) acc += __bfloat162float(x[r * ldx + c]);
      if (acc != 0.f) atomicAdd(out + c, acc);
    }
    return;
  }
  __shared__ float part[8][264];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < W; c0 += 256) {
    const int c = c0 + 8 * lane;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c < W) {
#pragma unroll 4
      for (int64_t r = r0 + warp; r < r1; r += 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(x + r * ldx + c);
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[2 * k] += __uint_as_float(w4[k] << 16);
          acc[2 * k + 1] += __uint_as_float(w4[k] & 0xffff0000u);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) part[warp][8 * lane + k] = acc[k];
    __syncthreads();
    const int cc = c0 + threadIdx.x;
    if (cc < W) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
      if (t != 0.f) atomicAdd(out + cc, t);
    }
    __syncthreads();
  }
}

// out_a[wa] = column sums of a (rows, wa), out_b[wb] = column sums of b (rows, wb), fp32, one launch:
// the privatised copies of d ts_w / d pos_w the attention backward fills (functional._HstuAttention).
// Block = 32 columns of the concatenated [a | b]; warp w sums rows w, w + 8, ...; shared-memory finish.
__global__ void __launch_bounds__(256) colsum_f32_pair_kernel(const float* __restrict__ a, int wa,
                                                              float* __restrict__ out_a,
                                                              const float* __restrict__ b, int wb,
                                                              float* __restrict__ out_b, int rows) {
  __shared__ float part[8][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane;
  const bool in_a = c < wa;
  const float* src = in_a ? a + c : b + (c - wa);
  const int ld = in_a ? wa : wb;
  float acc = 0.f;
  if (c < wa + wb) {
#pragma unroll 4
    for (int r = warp; r < rows; r += 8) acc += src[(int64_t) r * ld];
  }
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < wa + wb) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][lane];
    if (in_a) out_a[c] = t; else out_b[c - wa] = t;
  }
}

}  // namespace
}  // namespace grb

using namespace grb;

extern "C" {

int grb_proj_gemm(const grb_proj_gemm_args* a, grb_stream_t stream) {
  GRB_REQUIRE(a != nullptr, GRB_ERR_INVALID_ARG, "proj_gemm: null args");
  GRB_REQUIRE(a->M >= 0 && a->N > 0 && a->K >= 0 && a->A && a->B && a->out0, GRB_ERR_INVALID_ARG,
              "proj_gemm: bad arguments");
  GRB_REQUIRE(a->N % PG_BN == 0, GRB_ERR_UNSUPPORTED, "proj_gemm: N must be a multiple of %d (N=%lld)", PG_BN,
              (long long) a->N);
  GRB_REQUIRE(a->epi >= GRB_GEMM_EPI_PLAIN && a->epi <= GRB_GEMM_EPI_F32_ADD, GRB_ERR_INVALID_ARG,
              "proj_gemm: unknown epilogue");
  GRB_REQUIRE(a->epi != GRB_GEMM_EPI_SILU2 || a->out1, GRB_ERR_INVALID_ARG, "proj_gemm: SILU2 needs out1");
  GRB_REQUIRE(a->epi == GRB_GEMM_EPI_F32_ADD || a->K % PG_BK == 0 || a->a_mn, GRB_ERR_UNSUPPORTED,
              "proj_gemm: K must be a multiple of %d for K-major A", PG_BK);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int esz = a->epi == GRB_GEMM_EPI_F32_ADD ? 4 : 2;
  GRB_REQUIRE(al16(a->A) && al16(a->B) && al16(a->out0) && al16(a->out1) && al16(a->res) && al16(a->bias) &&
                  (a->lda * 2) % 16 == 0 && (a->ldb * 2) % 16 == 0 && (a->ldo0 * esz) % 16 == 0 &&
                  (a->ldo1 * 2) % 16 == 0 && (a->ldres * 2) % 16 == 0,
              GRB_ERR_INVALID_ARG, "proj_gemm: pointers and row strides must be 16-byte aligned");
  if (a->M == 0) return GRB_OK;
  CUtensorMap tmA, tmB;
  int rc;
  // A: K-major = (M, K) row-major, box 128 x 64 ; MN-major = (K, M) row-major, box 64 k x 64 m
  if (!a->a_mn) rc = make_tmap_bf16_2d(&tmA, a->A, a->M, a->K, a->lda, PG_BM);
  else rc = make_tmap_bf16_2d(&tmA, a->A, a->K, a->M, a->lda, 64);
  if (rc != GRB_OK) return rc;
  // B: K-major = (N, K) row-major, box 256 x 64 ; MN-major = (K, N) row-major, box 64 k x 64 n
  if (!a->b_mn) rc = make_tmap_bf16_2d(&tmB, a->B, a->N, a->K, a->ldb, PG_BN);
  else rc = make_tmap_bf16_2d(&tmB, a->B, a->K, a->N, a->ldb, 64);
  if (rc != GRB_OK) return rc;
  PgParams p{};
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_mn; p.b_mn = a->b_mn; p.epi = a->epi;
  p.n_mt = (int) ceil_div(a->M, PG_BM);
  p.n_nt = (int) (a->N / PG_BN);
  const int sms = num_sms();
  p.splits = 1;
  p.k_per_split = ceil_div(a->K > 0 ? a->K : 1, PG_BK) * PG_BK;
  if (a->epi == GRB_GEMM_EPI_F32_ADD) {
    // split K so that every SM has a work item: the output tiles alone are far fewer than SMs
    const int64_t tiles = (int64_t) p.n_mt * p.n_nt;
    const int64_t kblocks = ceil_div(a->K, PG_BK);
    int64_t s = tiles >= sms ? 1 : sms / tiles;
    if (s > kblocks) s = kblocks;
    if (s < 1) s = 1;
    p.k_per_split = ceil_div(kblocks, s) * PG_BK;
    p.splits = (int) ceil_div(a->K > 0 ? a->K : 1, p.k_per_split);
  }
  p.out0 = a->out0; p.ldo0 = a->ldo0; p.out1 = a->out1; p.ldo1 = a->ldo1;
  p.bias = a->bias; p.res = reinterpret_cast<const __nv_bfloat16*>(a->res); p.ldres = a->ldres;
  const int64_t n_tiles = (int64_t) p.n_mt * p.n_nt * p.splits;
  const unsigned grid = (unsigned) (n_tiles < sms ? n_tiles : sms);
  const size_t smem = PgSmem::total + 1024;
  GRB_CUDA_OK(cudaFuncSetAttribute(proj_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  proj_gemm_kernel<<<grid, PG_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_colsum_bf16(const void* x, int64_t ldx, int64_t rows, int32_t W, float* out, grb_stream_t stream) {
  GRB_REQUIRE(x && out && rows >= 0 && W > 0, GRB_ERR_INVALID_ARG, "colsum_bf16: bad arguments");
  if (rows == 0) return GRB_OK;
  const int vec = (W % 8 == 0) && (ldx % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  colsum_bf16_kernel<<<(unsigned) ceil_div(rows, (int64_t) CS_ROWS), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, rows, W, out, vec);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_colsum_f32_pair(const float* a, int32_t wa, float* out_a, const float* b, int32_t wb, float* out_b,
                        int32_t rows, grb_stream_t stream) {
  GRB_REQUIRE(a && out_a && wa > 0 && wb >= 0 && (wb == 0 || (b && out_b)) && rows >= 0, GRB_ERR_INVALID_ARG,
              "colsum_f32_pair: bad arguments");
  colsum_f32_pair_kernel<<<(unsigned) ceil_div((int64_t) wa + wb, (int64_t) 32), 256, 0,
                           reinterpret_cast<cudaStream_t>(stream)>>>(a, wa, out_a, b, wb, out_b, rows);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}
