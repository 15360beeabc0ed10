// mips_epilogue.cuh — the epilogue contract shared by the CUDA-core and tcgen05 score kernels.
#pragma once
#include "common.cuh"

namespace grb {

constexpr int MIPS_TILE_N = 128;  // items per tile: the sampling granule

// GMAX / PRIVATE: the two passes of the small-batch plan (B <= 128, one query block; mips_small.cu)
enum { MIPS_EPI_STORE = 0, MIPS_EPI_FILTER = 1, MIPS_EPI_GMAX = 2, MIPS_EPI_PRIVATE = 3 };

// candidate of a private sub-list: (score bits, item index)
struct __align__(8) MipsCand { uint32_t score; int32_t item; };
constexpr int MIPS_SUB_SPARE = 32;   // spare entries behind a sub-list's capacity (mips_sm100.cu)
constexpr int MIPS_SUB_PER_CTA = 4;  // sub-lists per (CTA, row): one per 32-column quarter of a tile

struct ScoreEpi {
  int mode;
  // launch tile u covers item tile  tile_stride * (grp * (u / per) + (u % per) + first):
  // (grp, per, first) = (1, 1, 0) is a plain stride; (4, 3, 1) skips every 4th multiple, which
  // is how refinement phase p visits exactly the tiles the coarser phases have not seen.
  int64_t tile_stride;
  int32_t grp, per, first;
  // STORE: out[row * Xs + u * MIPS_TILE_N + c] = score (items >= X -> -inf)
  int64_t Xs;
  float* out;
  // FILTER: append (score, item) with score >= tau[row]
  const float* tau;
  int32_t* counts;
  float* cscores;
  int32_t* cidx;
  int64_t cap;
  // GMAX: gmax[row * n_groups + u * (MIPS_TILE_N / group) + g] = max score of `group` consecutive
  // columns of launch tile u (group in {8, 16, 32}; items >= X count as -inf)
  float* gmax;
  int64_t n_groups;
  int32_t group;
  // PRIVATE: thread (row, column quarter q) of CTA c owns sub-list s = 4 c + q of row `row`:
  // sub_cand[(row * n_sub + s) * (sub_cap + MIPS_SUB_SPARE) + slot], slot counter in a register, no
  // atomics; hits past sub_cap land in the spare tail; sub_counts[row * n_sub + s] = ALL hits
  MipsCand* sub_cand;
  int32_t* sub_counts;
  int32_t n_sub, sub_cap;
};

// the small-batch plan (mips_small.cu): sizes and workspace offsets; ok == 0: not applicable
struct MipsSmallPlan {
  int ok;
  int64_t stride, n_sample_tiles, n_groups;
  int group, n_sub, sub_cap;
  int64_t off_tau, off_gmax, off_counts, off_cand, total;
};

// (launch tiles number < 2^24, since X < 2^31: the quotient is a 32-bit division — or none at all
//  for a plain stride — instead of the ~100-instruction 64-bit one, which every epilogue warp paid
//  per tile: ncu of the one-query-block pass, 246 instructions per warp and tile)
__device__ __forceinline__ int64_t epi_item_tile(const ScoreEpi& e, int64_t u) {
  const uint32_t uu = (uint32_t) u, per = (uint32_t) e.per;
  const uint32_t q = per == 1u ? uu : uu / per;
  return e.tile_stride * (int64_t) ((uint32_t) e.grp * q + (uu - q * per) + (uint32_t) e.first);
}

__device__ __forceinline__ void append_candidate(const ScoreEpi& e, int64_t row, float s,
                                                 int64_t item) {
  const int slot = atomicAdd(e.counts + row, 1);
  if (slot < e.cap) {
    e.cscores[row * e.cap + slot] = s;
    e.cidx[row * e.cap + slot] = (int32_t) item;
  }
}

// 8 consecutive scores of one query row: items item0..item0+7, tile-local columns col0..col0+7.
__device__ __forceinline__ void score_epilogue_row8(const ScoreEpi& e, int64_t row,
                                                    int64_t launch_tile, int64_t item0, int col0,
                                                    int64_t X, const float (&s)[8]) {
  if (e.mode == MIPS_EPI_STORE) {
    float* o = e.out + row * e.Xs + launch_tile * MIPS_TILE_N + col0;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = (item0 + c < X) ? s[c] : -INFINITY;
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    const float t = e.tau[row];
    int n = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) n += (s[c] >= t && item0 + c < X) ? 1 : 0;
    if (n) {  // one atomic per 8 scores, not one per candidate
      int slot = atomicAdd(e.counts + row, n);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (s[c] >= t && item0 + c < X) {
          if (slot < e.cap) {
            e.cscores[row * e.cap + slot] = s[c];
            e.cidx[row * e.cap + slot] = (int32_t) (item0 + c);
          }
          ++slot;
        }
    }
  }
}

}  // namespace grb
