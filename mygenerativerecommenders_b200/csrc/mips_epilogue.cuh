// mips_epilogue.cuh — the epilogue contract shared by the CUDA-core and tcgen05 score kernels.
#pragma once
#include "common.cuh"

namespace grb {

constexpr int MIPS_TILE_N = 128;  // items per tile: the sampling granule

enum { MIPS_EPI_STORE = 0, MIPS_EPI_FILTER = 1 };

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
};

__device__ __forceinline__ int64_t epi_item_tile(const ScoreEpi& e, int64_t u) {
  return e.tile_stride * ((int64_t) e.grp * (u / e.per) + (u % e.per) + e.first);
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
