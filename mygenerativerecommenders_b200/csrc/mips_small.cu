// mips_small.cu — the small-batch plan of the fused MIPS top-k: one query block (B <= 128) against a
// corpus that has to stream from HBM once (the eval batches of C3, SURVEY 8d: 128 x 700 k x 64).
//
// Reference: /root/reference/src/generative_recommenders_pl/models/indexing/top_k.py:44-70 and the
// caller candidate_index.py:107-164 — same contract as mips_topk.cu (exact, sorted, ties -> lowest
// index, k' = k + #invalid ids, filter and target rank in the last kernel).
//
// The phased plan of mips_topk.cu (sample, then 4^-L ... 3/4 of the tiles with a threshold tightened in
// between) is built for thousands of queries, where candidate lists must stay short.  With one query
// block it is nine dependent launches, an atomic round trip per (warp, tile) in the filter epilogue and
// four one-CTA-per-row radix selects whose histograms serialise on two or three exponent bins: 0.27 ms
// for 90 MB of items at C3.  With <= 128 rows the candidate lists can be 10x longer and private:
//
//   1. GMAX     score every s-th item tile (s = 8, 16 ...) and keep, per row, the maximum of every
//               group of G consecutive columns (mips_sm100.cu, MIPS_EPI_GMAX)          ~1/s of the corpus
//   2. tau      tau[b] = k'-th largest group maximum of row b.  Every group maximum is the score of
//               a distinct item, so >= k' items score >= tau[b]; with n_groups >> k' few of the best
//               sampled items share a group and tau is close to the k'-th largest sampled score.
//   3. PRIVATE  score ALL tiles; thread (row, 32-column quarter) of CTA c appends its hits >= tau[row]
//               to its own sub-list (slot counter in a register: no atomics, one pass over TMEM).
//               Expected k' s hits per row, spread over 4 x #CTAs sub-lists.
//   4. select   one 1024-thread CTA per row: gather the sub-lists into registers as 64-bit keys
//               (score key << 32 | ~index: unique, so "ties -> lowest index" is plain descending order),
//               bitwise binary search for the k'-th largest, bitonic sort of the winners, then the same
//               output step as topk_select_kernel (mips_select.cuh).
//
// The binary searches count with one warp reduction + one barrier per bit: no shared-memory atomics, so
// equal exponents cost nothing.  Sampled tiles are scored twice (1/s extra traffic) by the same MMA
// sequence, so their scores are bit-identical in both passes.
//
// Overflow (a sub-list or the 8192-key register file of the select is too small: adversarial item order,
// massive ties) sets status[0] like the phased plan does; the host wrapper's exact re-run then arrives
// with an explicit candidate capacity, which selects the phased plan.
#include "common.cuh"
#include "mips_epilogue.cuh"
#include "mips_select.cuh"
#include <cstdlib>

namespace grb {

constexpr int MSM_THREADS = 1024;
constexpr int MSM_VPT = 8;                           // candidate keys per thread of the select
constexpr int MSM_CMAX = MSM_THREADS * MSM_VPT;      // 8192 candidates per row
constexpr int MSM_SUBMAX = 1024;                     // sub-lists per row (4 x CTAs of the PRIVATE pass)
constexpr int MSM_GROUPS_MAX = MSM_THREADS * 32;     // group maxima per row: 8 per thread of the tau kernel, up to 4 folded on load

// Sum of `local` over the 1024 threads of the CTA, returned to every thread: warp reduction, one
// barrier, warp reduction.  `wred` is int[2][64]; `phase` alternates between its halves, so that a
// fast warp's next call cannot overwrite partial results a slow warp still reads.
__device__ __forceinline__ int block_sum_1024(int local, int (*wred)[64], int& phase) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w = __reduce_add_sync(0xffffffffu, local);
  if (lane == 0) wred[phase][warp] = w;
  __syncthreads();
  const int t = __reduce_add_sync(0xffffffffu, wred[phase][lane]);
  phase ^= 1;
  return t;
}

// A T with #{keys >= T} >= kk, as large as possible up to its low `stop_bit` bits (stop_bit = 0: the
// kk-th largest key).  Keys == 0 are empty slots; 1 <= kk <= number of non-empty keys.  One count per
// bit, but only below the highest bit on which the keys differ at all (scores of one query's best
// candidates share sign, exponent and the first mantissa bits: ~20 of 32 rounds remain).
template <int VPT>
__device__ __forceinline__ uint32_t kth_largest_u32(const uint32_t (&key)[VPT], int kk, int stop_bit,
                                                    int (*wred)[64], int& phase) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t all_and = 0xffffffffu, all_or = 0u;
#pragma unroll
  for (int i = 0; i < VPT; ++i)
    if (key[i]) { all_and &= key[i]; all_or |= key[i]; }
  all_and = __reduce_and_sync(0xffffffffu, all_and);
  all_or = __reduce_or_sync(0xffffffffu, all_or);
  if (lane == 0) { wred[phase][warp] = (int) all_and; wred[phase][32 + warp] = (int) all_or; }
  __syncthreads();
  all_and = __reduce_and_sync(0xffffffffu, (uint32_t) wred[phase][lane]);
  all_or = __reduce_or_sync(0xffffffffu, (uint32_t) wred[phase][32 + lane]);
  phase ^= 1;
  const uint32_t diff = all_and ^ all_or;
  if (diff == 0u) return all_and;            // every key equal
  const int top = 31 - __clz(diff);
  uint32_t T = all_and & ~((2u << top) - 1u);   // the common prefix
#pragma unroll 1
  for (int bit = top; bit >= stop_bit; --bit) {
    const uint32_t cand = T | (1u << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < VPT; ++i) c += key[i] >= cand ? 1 : 0;
    if (block_sum_1024(c, wred, phase) >= kk) T = cand;
  }
  return T;
}

// ---------------------------------------------------------------------------------------------
// 2. tau[row] = k-th largest of the row's group maxima (MERGE adjacent groups folded on load: a
//    maximum over more columns, still the score of one item)
// ---------------------------------------------------------------------------------------------
constexpr int MSM_TAU_VPT = 8;
constexpr int MSM_TAU_STOP_BIT = 10;   // tau rounded down by < 2^-13 relative: a handful of extra candidates

template <int MERGE>
__global__ void __launch_bounds__(MSM_THREADS) mips_small_tau_kernel(const float* __restrict__ gmax,
                                                                     int64_t n_groups, int k,
                                                                     float* __restrict__ tau) {
  __shared__ int wred[2][64];
  const float* x = gmax + (int64_t) blockIdx.x * n_groups;
  const int64_t n = n_groups / MERGE;
  uint32_t key[MSM_TAU_VPT];
#pragma unroll
  for (int i = 0; i < MSM_TAU_VPT; ++i) {
    const int64_t e = (int64_t) threadIdx.x + (int64_t) MSM_THREADS * i;
    float v = 0.f;
    if (e < n) {
      if (MERGE == 1) v = x[e];
      else if (MERGE == 2) { const float2 f = reinterpret_cast<const float2*>(x)[e]; v = fmaxf(f.x, f.y); }
      else { const float4 f = reinterpret_cast<const float4*>(x)[e]; v = fmaxf(fmaxf(f.x, f.y), fmaxf(f.z, f.w)); }
    }
    key[i] = e < n ? fkey(v) : 0u;
  }
  int phase = 0;
  const uint32_t T = kth_largest_u32<MSM_TAU_VPT>(key, k, MSM_TAU_STOP_BIT, wred, phase);
  if (threadIdx.x == 0) tau[blockIdx.x] = fkey_inv(T);
}

// ---------------------------------------------------------------------------------------------
// 4. exact top-k of a row's private sub-lists
// ---------------------------------------------------------------------------------------------
constexpr int MSM_SELECT_SMEM = MSM_CMAX * 8;   // dynamic: the gathered keys, later the sort buffer

__global__ void __launch_bounds__(MSM_THREADS) mips_small_select_kernel(
    const MipsCand* __restrict__ sub_cand, const int32_t* __restrict__ sub_counts, int n_sub, int sub_cap,
    int k, const int64_t* __restrict__ id_map, float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
    int32_t* __restrict__ status, int32_t overflow_floor, SelectFilter flt) {
  extern __shared__ __align__(16) unsigned long long stage[];   // [MSM_CMAX]
  __shared__ int wred[2][64];
  __shared__ uint32_t skey[SEL_KMAX];
  __shared__ int wscan[32];
  __shared__ int sh_slot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row = blockIdx.x;
  int phase = 0;

  // sub-list lengths -> prefix offsets; the true hit count (what an exact re-run must hold)
  int cnt = 0;
  if (tid < n_sub) cnt = sub_counts[row * n_sub + tid];
  const int len = cnt < sub_cap ? cnt : sub_cap;
  int incl = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wscan[warp] = incl;
  if (tid == 0) sh_slot = 0;
  const int hits = block_sum_1024(cnt, wred, phase);           // <= X < 2^31; barrier: wscan visible
  const int clipped = __syncthreads_or(cnt > sub_cap ? 1 : 0);
  int base = 0, total = 0;
  for (int w = 0; w < 32; ++w) {
    const int v = wscan[w];
    base += w < warp ? v : 0;
    total += v;
  }
  if (clipped || total > MSM_CMAX) {   // uniform over the CTA
    if (tid == 0) atomicMax(status, hits > overflow_floor ? hits : overflow_floor);
    return;
  }
  // thread t copies sub-list t behind its prefix offset, as unique 64-bit keys
  // (score key << 32 | ~index: descending order = score descending, index ascending)
  if (tid < n_sub) {
    const MipsCand* src = sub_cand + (row * n_sub + tid) * (int64_t) (sub_cap + MIPS_SUB_SPARE);
    unsigned long long* dst = stage + (base + incl - len);
#pragma unroll 4
    for (int j = 0; j < len; ++j) {
      const MipsCand c = src[j];
      dst[j] = ((unsigned long long) fkey(__uint_as_float(c.score)) << 32) | (uint32_t) ~(uint32_t) c.item;
    }
  }
  __syncthreads();
  uint32_t hi[MSM_VPT], lo[MSM_VPT];
#pragma unroll
  for (int i = 0; i < MSM_VPT; ++i) {
    const int e = tid + MSM_THREADS * i;
    const unsigned long long v = e < total ? stage[e] : 0ull;
    hi[i] = (uint32_t) (v >> 32);
    lo[i] = (uint32_t) v;
  }

  // threshold: the kk-th largest 64-bit key, high word first
  const int kk = total < k ? total : k;
  uint32_t t_hi = 1u, t_lo = 0u;   // total <= k: every non-empty key
  if (total > k) {
    t_hi = kth_largest_u32<MSM_VPT>(hi, kk, 0, wred, phase);
    int c2 = 0;   // (#hi > t_hi) | (#hi == t_hi) << 16
#pragma unroll
    for (int i = 0; i < MSM_VPT; ++i) c2 += (hi[i] > t_hi ? 1 : 0) + (hi[i] == t_hi ? 65536 : 0);
    c2 = block_sum_1024(c2, wred, phase);
    const int above = c2 & 0xffff, ties = c2 >> 16, need = kk - above;
    if (ties > need) {   // equal scores straddle the cut: the `need` lowest indices = largest ~index
      uint32_t tl[MSM_VPT];
#pragma unroll
      for (int i = 0; i < MSM_VPT; ++i) tl[i] = hi[i] == t_hi ? lo[i] : 0u;   // ~index >= 2^31: never empty
      t_lo = kth_largest_u32<MSM_VPT>(tl, need, 0, wred, phase);
    }
  } else {
    __syncthreads();   // `stage` is read; it becomes the sort buffer below
  }

  // winners -> the sort buffer (exactly kk of them), padded to a power of two with empty keys.
  // (every path above ends in a barrier after the last read of `stage`)
  unsigned long long* const sbuf = stage;
  int np2 = 1;
  while (np2 < k) np2 <<= 1;
  for (int i = tid; i < np2; i += MSM_THREADS) sbuf[i] = 0ull;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MSM_VPT; ++i) {
    const bool win = hi[i] != 0u && (hi[i] > t_hi || (hi[i] == t_hi && lo[i] >= t_lo));
    const unsigned m = __ballot_sync(0xffffffffu, win);
    if (m) {
      int b0 = 0;
      if (lane == 0) b0 = atomicAdd(&sh_slot, __popc(m));
      b0 = __shfl_sync(0xffffffffu, b0, 0);
      const int s = b0 + __popc(m & ((1u << lane) - 1u));
      if (win && s < np2) sbuf[s] = ((unsigned long long) hi[i] << 32) | lo[i];
    }
  }
  __syncthreads();
  // bitonic sort, descending: one comparator per thread, only the warps that hold one take part
  // (named barrier 1), the others wait at the CTA barrier behind it
  const int n_sort = np2 / 2 < 32 ? 32 : np2 / 2;
  if (tid < n_sort) {
    for (int size = 2; size <= np2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        if (tid < np2 / 2) {
          const int i = ((tid & ~(stride - 1)) << 1) | (tid & (stride - 1));
          const int j = i | stride;
          const bool up = (i & size) == 0;
          const unsigned long long a = sbuf[i], b = sbuf[j];
          if ((a > b) != up && a != b) { sbuf[i] = b; sbuf[j] = a; }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(n_sort) : "memory");
      }
    }
  }
  __syncthreads();
  // unpack in place: skey[r] = score key, sid[r] = item index
  unsigned long long v[SEL_KMAX / MSM_THREADS];
#pragma unroll
  for (int j = 0; j < SEL_KMAX / MSM_THREADS; ++j) {
    const int r = tid + MSM_THREADS * j;
    v[j] = r < np2 ? sbuf[r] : 0ull;
  }
  __syncthreads();
  long long* sid = reinterpret_cast<long long*>(sbuf);
#pragma unroll
  for (int j = 0; j < SEL_KMAX / MSM_THREADS; ++j) {
    const int r = tid + MSM_THREADS * j;
    if (r < np2) {
      skey[r] = (uint32_t) (v[j] >> 32);
      sid[r] = v[j] ? (long long) (int32_t) ~(uint32_t) v[j] : LLONG_MAX;
    }
  }
  __syncthreads();
  select_write_rows<MSM_THREADS>(row, k, kk, skey, sid, id_map, out_scores, out_ids, flt);
}

// ---------------------------------------------------------------------------------------------
// Plan and orchestration
// ---------------------------------------------------------------------------------------------
bool mips_sm100_supported(const grb_mips_topk_args* a);
int mips_scores_sm100(const grb_mips_topk_args* a, const ScoreEpi& epi, int64_t n_launch_tiles,
                      cudaStream_t st);
void mips_sm100_work_split(int64_t n_launch_tiles, int64_t n_qb, int64_t* chunk_out, int64_t* n_chunks_out,
                           unsigned* grid_out);

static int64_t align256s(int64_t x) { return (x + 255) & ~255ll; }

static int small_env(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}

int plan_mips_small(const grb_mips_topk_args* a, int64_t ksel, MipsSmallPlan* S) {
  *S = MipsSmallPlan{};
  static const int enabled = small_env("GRB_MIPS_SMALL", 1);           // developer switches, read once
  static const int stride_env = small_env("GRB_MIPS_SMALL_STRIDE", 0);
  if (!enabled || a->B > 128 || a->B <= 0 || !mips_sm100_supported(a)) return GRB_OK;
  const int64_t n_tiles = ceil_div(a->X, MIPS_TILE_N);
  int64_t s = stride_env > 0 ? stride_env : (n_tiles < 32768 ? 8 : 16);
  // the select holds 8192 candidates per row; expected k' s (+ what the group-maximum estimate of the
  // k'-th sampled score lets through): keep 1.5x of room, else this plan would only ever overflow
  while (s > 2 && ksel * s * 3 / 2 > MSM_CMAX) s /= 2;
  if (ksel * s * 3 / 2 > MSM_CMAX) return GRB_OK;
  const int64_t n_sample = ceil_div(n_tiles, s);
  int G = 8;
  while (G < 32 && n_sample * (MIPS_TILE_N / G) > 8192 && n_sample * (MIPS_TILE_N / (2 * G)) >= 8 * ksel) G *= 2;
  const int64_t n_groups = n_sample * (MIPS_TILE_N / G);
  // >= k' real groups even if the last sampled tile is the ragged one, and enough of them for the estimate
  const int merge = n_groups <= MSM_THREADS * MSM_TAU_VPT ? 1 : n_groups <= 2 * MSM_THREADS * MSM_TAU_VPT ? 2 : 4;
  if (n_groups > MSM_GROUPS_MAX || n_groups / merge < 4 * ksel + MIPS_TILE_N / G) return GRB_OK;
  int64_t chunk, n_chunks;
  unsigned grid;
  mips_sm100_work_split(n_tiles, 1, &chunk, &n_chunks, &grid);
  const int n_sub = MIPS_SUB_PER_CTA * (int) grid;
  if (n_sub > MSM_SUBMAX) return GRB_OK;
  const int64_t mean = ceil_div(ksel * s, n_sub);
  int64_t sub_cap = 6 * mean;
  if (sub_cap < 24) sub_cap = 24;
  S->ok = 1;
  S->stride = s; S->n_sample_tiles = n_sample; S->n_groups = n_groups; S->group = G;
  S->n_sub = n_sub; S->sub_cap = (int) sub_cap;
  int64_t o = 0;
  S->off_tau = o;    o = align256s(o + a->B * 4);
  S->off_gmax = o;   o = align256s(o + a->B * n_groups * 4);
  S->off_counts = o; o = align256s(o + a->B * (int64_t) n_sub * 4);
  S->off_cand = o;   o = align256s(o + a->B * (int64_t) n_sub * (sub_cap + MIPS_SUB_SPARE) * (int64_t) sizeof(MipsCand));
  S->total = o;
  return GRB_OK;
}

int run_mips_small(const grb_mips_topk_args* a, const MipsSmallPlan& S, int ksel, int32_t overflow_floor,
                   cudaStream_t st) {
  auto ws = reinterpret_cast<unsigned char*>(a->workspace);
  float* tau = reinterpret_cast<float*>(ws + S.off_tau);
  float* gmax = reinterpret_cast<float*>(ws + S.off_gmax);
  int32_t* counts = reinterpret_cast<int32_t*>(ws + S.off_counts);
  MipsCand* cand = reinterpret_cast<MipsCand*>(ws + S.off_cand);
  const int64_t n_tiles = ceil_div(a->X, MIPS_TILE_N);

  ScoreEpi epi{};
  epi.mode = MIPS_EPI_GMAX; epi.tile_stride = S.stride; epi.grp = 1; epi.per = 1; epi.first = 0;
  epi.gmax = gmax; epi.n_groups = S.n_groups; epi.group = S.group;
  epi.tau = tau; epi.sub_cand = cand; epi.sub_counts = counts; epi.n_sub = S.n_sub; epi.sub_cap = S.sub_cap;
  int rc = mips_scores_sm100(a, epi, S.n_sample_tiles, st);
  if (rc != GRB_OK) return rc;

  if (S.n_groups <= MSM_THREADS * MSM_TAU_VPT)
    mips_small_tau_kernel<1><<<(unsigned) a->B, MSM_THREADS, 0, st>>>(gmax, S.n_groups, ksel, tau);
  else if (S.n_groups <= 2 * MSM_THREADS * MSM_TAU_VPT)
    mips_small_tau_kernel<2><<<(unsigned) a->B, MSM_THREADS, 0, st>>>(gmax, S.n_groups, ksel, tau);
  else
    mips_small_tau_kernel<4><<<(unsigned) a->B, MSM_THREADS, 0, st>>>(gmax, S.n_groups, ksel, tau);
  GRB_LAUNCH_OK();

  epi.mode = MIPS_EPI_PRIVATE; epi.tile_stride = 1;
  rc = mips_scores_sm100(a, epi, n_tiles, st);
  if (rc != GRB_OK) return rc;

  SelectFilter flt{};
  flt.invalid = a->n_invalid > 0 ? a->invalid_ids : nullptr;
  flt.ld = a->ld_invalid; flt.n_invalid = a->n_invalid; flt.k_out = a->k;
  flt.target = a->target_ids; flt.ranks = a->target_ids ? a->out_ranks : nullptr;
  GRB_CUDA_OK(cudaFuncSetAttribute(mips_small_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   MSM_SELECT_SMEM));
  mips_small_select_kernel<<<(unsigned) a->B, MSM_THREADS, MSM_SELECT_SMEM, st>>>(
      cand, counts, S.n_sub, S.sub_cap, ksel, a->item_ids, a->out_scores, a->out_ids, a->status,
      overflow_floor, flt);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}  // namespace grb
