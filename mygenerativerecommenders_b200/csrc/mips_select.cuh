// mips_select.cuh — what the two selection kernels (mips_topk.cu: candidate lists of the phased plan;
// mips_small.cu: private sub-lists of the small-batch plan) share: the order-preserving float key,
// and the last step of a row — the sorted best k' in shared memory -> (B, k) outputs, with the
// invalid-id filter and the target rank of candidate_index.py:125-158 / metrics/retrieval.py:40-68.
#pragma once
#include "common.cuh"
#include <climits>
#include <cmath>

namespace grb {

constexpr int SEL_THREADS = 256;
__device__ __forceinline__ uint32_t fkey(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

constexpr int SEL_KMAX = 2048;

constexpr int SEL_INVALID_MAX = 1024;

// f3 (candidate_index.py:125-158, metrics/retrieval.py:40-68): what the reference does with the
// (B, k') result of the top-k module — drop ids listed in invalid_ids[row], keep the first k, find
// the target's rank — happens here on the sorted shared-memory list, so that only (B, k) leaves.
struct SelectFilter {
  const int64_t* invalid;  // (B, n_invalid), row stride ld; NULL = none
  int64_t ld;
  int n_invalid;
  int k_out;               // entries written per row (k_out <= k; k - k_out <= n_invalid)
  const int64_t* target;   // (B) or NULL
  int32_t* ranks;          // (B) or NULL
};

// Row `row`: skey[r] / sid[r] (r < filled) hold the best entries, sorted (key descending, index
// ascending); sid is the item INDEX, mapped through id_map here.  Called by all THREADS threads of
// the CTA, after a __syncthreads() that made skey / sid visible.
template <int THREADS>
__device__ __forceinline__ void select_write_rows(int64_t row, int k, int filled, const uint32_t* skey,
                                                  long long* sid, const int64_t* __restrict__ id_map,
                                                  float* __restrict__ out_scores,
                                                  int64_t* __restrict__ out_ids, const SelectFilter& flt) {
  const int tid = threadIdx.x;
  if (!flt.invalid && !flt.ranks) {
    for (int r = tid; r < k; r += THREADS) {
      float s = -INFINITY;
      long long id = -1;
      if (r < filled) {
        s = fkey_inv(skey[r]);
        id = sid[r];
        if (id_map) id = id_map[id];
      }
      out_scores[row * k + r] = s;
      out_ids[row * k + r] = id;
    }
    return;
  }
  // ---- filtered tail: real ids, invalid flags, stable compaction to k_out, target rank --------
  __shared__ long long sinv[SEL_INVALID_MAX];
  __shared__ int wsum[THREADS / 32];
  __shared__ int rank_sh;
  const int ko = flt.k_out;
  const int ninv = flt.invalid ? flt.n_invalid : 0;
  for (int i = tid; i < ninv; i += THREADS) sinv[i] = flt.invalid[row * flt.ld + i];
  if (tid == 0) rank_sh = ko + 1;
  for (int r = tid; r < filled; r += THREADS)
    if (id_map) sid[r] = id_map[sid[r]];
  __syncthreads();
  const long long tgt = (flt.ranks && flt.target) ? flt.target[row] : 0;
  // each thread owns a contiguous run of the sorted list so that the compaction keeps its order
  const int per = (filled + THREADS - 1) / THREADS;
  const int r0 = tid * per, r1 = (r0 + per < filled) ? r0 + per : filled;
  unsigned long long keepmask = 0;   // per <= 2048 / 256 = 8
  int nkeep = 0;
  for (int r = r0; r < r1; ++r) {
    const long long id = sid[r];
    bool bad = false;
    for (int j = 0; j < ninv; ++j) bad |= (sinv[j] == id);
    if (!bad) { keepmask |= 1ull << (r - r0); ++nkeep; }
  }
  int incl = nkeep;
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  int base = incl - nkeep;
  for (int w = 0; w < wid; ++w) base += wsum[w];
  int total = 0;
  for (int w = 0; w < THREADS / 32; ++w) total += wsum[w];
  for (int r = r0; r < r1; ++r) {
    if (!((keepmask >> (r - r0)) & 1ull)) continue;
    if (base < ko) {
      out_scores[row * ko + base] = fkey_inv(skey[r]);
      out_ids[row * ko + base] = sid[r];
      if (flt.ranks && sid[r] == tgt) atomicMin(&rank_sh, base + 1);
    }
    ++base;
  }
  for (int r = (total < ko ? total : ko) + tid; r < ko; r += THREADS) {
    out_scores[row * ko + r] = -INFINITY;
    out_ids[row * ko + r] = -1;
  }
  __syncthreads();
  if (flt.ranks && tid == 0) flt.ranks[row] = rank_sh;
}

}  // namespace grb
