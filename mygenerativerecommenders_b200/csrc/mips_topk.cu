// mips_topk.cu — brute-force MIPS top-k with the score matrix kept out of HBM.
//
// Reference: /root/reference/src/generative_recommenders_pl/models/indexing/top_k.py:44-70
//   all_logits = mm(q, items_t) ; topk(all_logits, k, sorted) ; item_ids[idx]
// and the caller models/indexing/candidate_index.py:107-164 (k' = k + #invalid ids).
//
// Exact, four stream-ordered phases (see include/grb200.h):
//   1. score a strided sample of 128-item tiles                    -> (B, Xs) fp32
//   2. tau[b] = k-th largest sample score (radix select)           -> >= k items score >= tau[b]
//   3. score ALL tiles with the same code path (bit-identical scores for sampled items) and
//      append (score, index) with score >= tau[b] to row b's candidate list
//   4. exact select + bitonic sort of each candidate list (ties -> lowest index), id gather.
// Expected candidates per row ~ k * stride; with Xs ~ sqrt(2 k X) the extra traffic is
// O(B sqrt(k X)) instead of the reference's O(B X).
//
// This file holds the CUDA-core score kernel (fp32 tables, any D — the reference's dtype) and
// the selection kernels.  The tcgen05 score kernel for bf16 tables is in mips_sm100.cu and
// plugs into the same epilogue contract (ScoreEpi).
#include "common.cuh"
#include "mips_epilogue.cuh"
#include "mips_select.cuh"
#include <cmath>
#include <cfloat>

namespace grb {

// ---------------------------------------------------------------------------------------------
// CUDA-core score kernel: 128 queries x 128 items per CTA, BK = 16, 8x8 per thread.
// ---------------------------------------------------------------------------------------------
constexpr int SG_BM = 128, SG_BN = MIPS_TILE_N, SG_BK = 16, SG_LD = 132, SG_THREADS = 256;

template <typename T> __device__ __forceinline__ float ld_as_f32(const T* p);
template <> __device__ __forceinline__ float ld_as_f32<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as_f32<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

template <typename T>
__global__ void __launch_bounds__(SG_THREADS) mips_scores_simt(
    const T* __restrict__ Q, int64_t ldq, const T* __restrict__ I, int64_t ldi, int64_t B,
    int64_t X, int D, ScoreEpi epi) {
  __shared__ __align__(16) float As[SG_BK][SG_LD];
  __shared__ __align__(16) float Bs[SG_BK][SG_LD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t tile = epi_item_tile(epi, (int64_t) blockIdx.x);
  const int64_t n0 = tile * SG_BN;
  const int64_t m0 = (int64_t) blockIdx.y * SG_BM;

  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[a][c] = 0.f;

  const int lr = tid >> 1, lk = (tid & 1) * 8;
  for (int k0 = 0; k0 < D; k0 += SG_BK) {
    {
      const int64_t row = m0 + lr;
      const T* src = Q + row * ldq + k0 + lk;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        As[lk + i][lr] = (row < B && k0 + lk + i < D) ? ld_as_f32<T>(src + i) : 0.f;
      const int64_t it = n0 + lr;
      const T* srcb = I + it * ldi + k0 + lk;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        Bs[lk + i][lr] = (it < X && k0 + lk + i < D) ? ld_as_f32<T>(srcb + i) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int64_t row = m0 + ty * 8 + a;
    if (row >= B) continue;
    score_epilogue_row8(epi, row, (int64_t) blockIdx.x, n0 + tx * 8, tx * 8, X, acc[a]);
  }
}

// ---------------------------------------------------------------------------------------------
// Phase 2: k-th largest of each row (radix select, 4 x 8 bits).  One CTA per row.
// ---------------------------------------------------------------------------------------------

// Finds the kk-th largest (1-based) key among keys matching (key & mask) == prefix restricted to
// digit `shift`.  hist must be 256 ints of smem; result broadcast through sh[0..2] = (digit, rank
// inside the digit's bin, size of the bin).  Warp 0 walks the bins from 255 down, 8 per lane.
__device__ __forceinline__ void pick_digit_desc(int* hist, int* sh, int kk) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int h[8], s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { h[j] = hist[255 - 8 * lane - j]; s += h[j]; }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int excl = incl - s;
    if (excl < kk && kk <= incl) {   // the lane whose bins hold the kk-th largest
      int cum = excl;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (cum < kk && cum + h[j] >= kk) { sh[0] = 255 - 8 * lane - j; sh[1] = kk - cum; sh[2] = h[j]; }
        cum += h[j];
      }
    } else if (lane == 31 && incl < kk) {   // fewer than kk keys: callers rule this out
      sh[0] = 0; sh[1] = kk; sh[2] = 0;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(SEL_THREADS) row_kth_largest_kernel(
    const float* __restrict__ scores, int64_t ld, int64_t L, const int32_t* __restrict__ counts,
    int k, float* __restrict__ tau) {
  __shared__ int hist[256];
  __shared__ int sh[3];
  const float* x = scores + (int64_t) blockIdx.x * ld;
  if (counts) {  // candidate lists: only the filled prefix (capped) is valid
    const int64_t c = counts[blockIdx.x];
    L = c < L ? c : L;
    if (L < k) return;  // cannot happen once phase 0 has contributed k entries; keep tau
  }
  uint32_t prefix = 0, mask = 0;
  int kk = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += SEL_THREADS) hist[i] = 0;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < L; i += SEL_THREADS) {
      const uint32_t key = fkey(x[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
    }
    __syncthreads();
    pick_digit_desc(hist, sh, kk);
    prefix |= (uint32_t) sh[0] << shift;
    mask |= 255u << shift;
    kk = sh[1];
    __syncthreads();
  }
  if (threadIdx.x == 0) tau[blockIdx.x] = fkey_inv(prefix);
}

// Phase 3 when every tile was sampled (small corpora): filter the stored scores.
// Sampled scores -> candidates.  Column c of the sample is item
// (c / 128) * sample_stride * 128 + c % 128.
// CTA = (2048-column chunk, query row): eight independent coalesced loads per thread, and a warp
// appends its hits with ONE atomic (ballot + prefix popcount) instead of one per candidate.
constexpr int FD_COLS = 2048;
__global__ void __launch_bounds__(256) filter_dense_kernel(const float* __restrict__ scores, int64_t ld,
                                                           int64_t ncols, int64_t sample_stride,
                                                           int64_t X, ScoreEpi epi) {
  const int64_t row = blockIdx.y;
  const float tau = epi.tau[row];
  const float* x = scores + row * ld;
  const int lane = threadIdx.x & 31;
  const int64_t c0 = (int64_t) blockIdx.x * FD_COLS + threadIdx.x;
  float s[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) s[u] = (c0 + 256 * u < ncols) ? x[c0 + 256 * u] : -INFINITY;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int64_t col = c0 + 256 * u;
    const int64_t item = (col / MIPS_TILE_N) * sample_stride * MIPS_TILE_N + col % MIPS_TILE_N;
    const bool hit = col < ncols && item < X && s[u] >= tau;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) {
      int base = 0;
      if (lane == 0) base = atomicAdd(epi.counts + row, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (hit) {
        const int slot = base + __popc(m & ((1u << lane) - 1u));
        if (slot < epi.cap) {
          epi.cscores[row * epi.cap + slot] = s[u];
          epi.cidx[row * epi.cap + slot] = (int32_t) item;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Phase 4 / shard merge: exact top-k of a candidate list, sorted, ties -> lowest id.
// ---------------------------------------------------------------------------------------------
template <typename IdT>
__global__ void __launch_bounds__(SEL_THREADS) topk_select_kernel(
    const float* __restrict__ cs, const IdT* __restrict__ cid, const int32_t* __restrict__ counts,
    int64_t cap, int k, const int64_t* __restrict__ id_map, float* __restrict__ out_scores,
    int64_t* __restrict__ out_ids, int32_t* __restrict__ status, SelectFilter flt) {
  __shared__ int hist[256];
  __shared__ int sh[3];
  __shared__ uint32_t skey[SEL_KMAX];
  __shared__ long long sid[SEL_KMAX];
  __shared__ int slot;
  const int64_t row = blockIdx.x;
  int64_t c = counts ? (int64_t) counts[row] : cap;
  if (c > cap) {
    if (threadIdx.x == 0 && status) atomicMax(status, (int) (c > 0x7fffffff ? 0x7fffffff : c));
    c = cap;
  }
  const float* x = cs + row * cap;
  const IdT* ids = cid + row * cap;
  const int tid = threadIdx.x;

  uint32_t key_k = 0;        // every key >= key_k qualifies when c <= k
  long long id_thr = LLONG_MAX;
  if (c > k) {
    uint32_t prefix = 0, mask = 0;
    int kk = k, ties = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
      __syncthreads();
      for (int64_t i = tid; i < c; i += SEL_THREADS) {
        const uint32_t key = fkey(x[i]);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
      }
      __syncthreads();
      pick_digit_desc(hist, sh, kk);
      prefix |= (uint32_t) sh[0] << shift;
      mask |= 255u << shift;
      kk = sh[1];
      ties = sh[2];
      __syncthreads();
    }
    key_k = prefix;
    if (ties > kk) {
      // need the kk smallest ids among the `ties` entries with key == key_k:
      // radix select (ascending) over the 64-bit ids, 8 x 8 bits.
      unsigned long long ipre = 0, imask = 0;
      int need = kk;
      for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
        __syncthreads();
        for (int64_t i = tid; i < c; i += SEL_THREADS) {
          if (fkey(x[i]) != key_k) continue;
          const unsigned long long u = (unsigned long long) (long long) ids[i];
          if ((u & imask) == ipre) atomicAdd(&hist[(int) ((u >> shift) & 255ull)], 1);
        }
        __syncthreads();
        if (tid == 0) {
          int cum = 0, digit = 255, rem = need;
          for (int b = 0; b < 256; ++b) {
            const int h = hist[b];
            if (cum + h >= need) { digit = b; rem = need - cum; break; }
            cum += h;
          }
          sh[0] = digit; sh[1] = rem;
        }
        __syncthreads();
        ipre |= (unsigned long long) sh[0] << shift;
        imask |= 255ull << shift;
        need = sh[1];
        __syncthreads();
      }
      id_thr = (long long) ipre;
    }
  }
  // collect
  if (tid == 0) slot = 0;
  int np2 = 1;
  while (np2 < k) np2 <<= 1;
  for (int i = tid; i < np2; i += SEL_THREADS) { skey[i] = 0u; sid[i] = LLONG_MAX; }
  __syncthreads();
  for (int64_t i = tid; i < c; i += SEL_THREADS) {
    const uint32_t key = fkey(x[i]);
    const long long id = (long long) ids[i];
    if (key > key_k || (key == key_k && id <= id_thr)) {
      const int s = atomicAdd(&slot, 1);
      if (s < k) { skey[s] = key; sid[s] = id; }
    }
  }
  __syncthreads();
  // bitonic sort, order: key descending, id ascending
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < np2; i += SEL_THREADS) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;  // "up" block: best first
          const uint32_t ki = skey[i], kj = skey[j];
          const long long ii = sid[i], ij = sid[j];
          const bool i_before_j = (ki > kj) || (ki == kj && ii < ij);
          if (i_before_j != up) { skey[i] = kj; skey[j] = ki; sid[i] = ij; sid[j] = ii; }
        }
      }
      __syncthreads();
    }
  }
  select_write_rows<SEL_THREADS>(row, k, (int) (c < k ? c : k), skey, sid, id_map, out_scores, out_ids, flt);
}

// ---------------------------------------------------------------------------------------------
// Orchestration
// ---------------------------------------------------------------------------------------------
// Refinement radix: phase p visits the tiles that are multiples of R^(levels-p) but not of
// R^(levels-p+1) and adds ~(R-1) k candidates per row.  Measured at C4 (4096 x 10 M): R = 4 (five
// phases, 3/4 of the corpus in the last one) 18.8 ms, R = 8 (three phases, 7/8 in the last one)
// 20.4 ms: the longer candidate lists cost more in the selection kernels than the saved phases.
constexpr int MIPS_RADIX = 4;

struct MipsPlan {
  int64_t n_tiles, stride, n_sample_tiles, Xs, cap;
  int64_t auto_cap;     // the capacity the plan picks on its own (cand_cap = 0)
  MipsSmallPlan small;  // B <= 128, bf16: mips_small.cu; taken unless the caller sized the lists itself
  int levels;  // stride = MIPS_RADIX^levels; refinement phases 1..levels
  int64_t off_tau, off_counts, off_sample, off_cscores, off_cidx, total;
};

static int64_t align256(int64_t x) { return (x + 255) & ~255ll; }

// implemented in mips_small.cu
int plan_mips_small(const grb_mips_topk_args* a, int64_t ksel, MipsSmallPlan* S);
int run_mips_small(const grb_mips_topk_args* a, const MipsSmallPlan& S, int ksel, int32_t overflow_floor,
                   cudaStream_t st);

// the selection size: the reference over-selects k' = min(k + #invalid ids, X) (candidate_index.py:132)
static int64_t mips_ksel(const grb_mips_topk_args* a) {
  const int64_t ks = (int64_t) a->k + (a->invalid_ids ? a->n_invalid : 0);
  return ks < a->X ? ks : a->X;
}

static int plan_mips(const grb_mips_topk_args* a, MipsPlan* P) {
  GRB_REQUIRE(a && a->B >= 0 && a->X > 0 && a->D > 0 && a->k > 0, GRB_ERR_INVALID_ARG,
              "mips_topk: bad sizes");
  GRB_REQUIRE(a->k <= a->X, GRB_ERR_INVALID_ARG, "mips_topk: k=%d exceeds corpus size %lld",
              a->k, (long long) a->X);
  GRB_REQUIRE(a->n_invalid >= 0 && a->n_invalid <= SEL_INVALID_MAX && (a->n_invalid == 0 || a->invalid_ids),
              GRB_ERR_UNSUPPORTED, "mips_topk: n_invalid=%d (at most %d, with a list)", a->n_invalid,
              SEL_INVALID_MAX);
  const int64_t ksel = mips_ksel(a);
  GRB_REQUIRE(ksel <= SEL_KMAX, GRB_ERR_UNSUPPORTED, "mips_topk: k + n_invalid = %lld exceeds %d",
              (long long) ksel, SEL_KMAX);
  GRB_REQUIRE(a->X < (1ll << 31), GRB_ERR_UNSUPPORTED, "mips_topk: corpus too large");
  P->n_tiles = ceil_div(a->X, MIPS_TILE_N);
  // Phase 0 scores every (R^levels)-th tile; phase p = 1..levels scores the tiles that are
  // multiples of R^(levels-p) but not of R^(levels-p+1), tightening tau[b] in between.  Each
  // refinement phase is expected to add ~(R-1) k candidates per row ((R-1)x the items seen so far).
  auto ipow = [](int l) { int64_t v = 1; for (int i = 0; i < l; ++i) v *= MIPS_RADIX; return v; };
  int64_t min_tiles = ceil_div(8 * ksel, MIPS_TILE_N);
  if (min_tiles < 32) min_tiles = 32;
  int levels = 0;
  if (a->sample_stride > 0) {
    int64_t s4 = 1;
    while (s4 * MIPS_RADIX <= a->sample_stride && levels < 8) { s4 *= MIPS_RADIX; ++levels; }
  } else {
    while (levels < 6 && P->n_tiles / ipow(levels + 1) >= min_tiles) ++levels;
  }
  int64_t stride = ipow(levels);
  P->n_sample_tiles = ceil_div(P->n_tiles, stride);
  // the sample must hold at least k real items (its last tile may be partial)
  while (levels > 0) {
    const int64_t last = (P->n_sample_tiles - 1) * stride;
    const int64_t real = (P->n_sample_tiles - 1) * MIPS_TILE_N +
        (last == P->n_tiles - 1 ? a->X - last * MIPS_TILE_N : MIPS_TILE_N);
    if (real >= ksel) break;
    --levels; stride = ipow(levels); P->n_sample_tiles = ceil_div(P->n_tiles, stride);
  }
  P->levels = levels;
  P->stride = stride;
  P->Xs = P->n_sample_tiles * MIPS_TILE_N;
  auto clamp_cap = [&](int64_t c) { c = c > a->X ? a->X : c; return c < ksel ? ksel : c; };
  P->auto_cap = clamp_cap(ksel * (4 + 2 * (MIPS_RADIX - 1) * levels) + 1024);
  P->cap = a->cand_cap > 0 ? clamp_cap(a->cand_cap) : P->auto_cap;
  int64_t o = 0;
  P->off_tau = o;     o = align256(o + a->B * 4);
  P->off_counts = o;  o = align256(o + a->B * 4);
  P->off_sample = o;  o = align256(o + a->B * P->Xs * 4);
  P->off_cscores = o; o = align256(o + a->B * P->cap * 4);
  P->off_cidx = o;    o = align256(o + a->B * P->cap * 4);
  P->total = o;
  // The small-batch plan runs when the capacity is the automatic one; an explicit capacity is the host
  // wrapper's exact re-run after an overflow (status[0] + 1024 > auto_cap) and takes the phased plan.
  // Either way the workspace holds both layouts, so the choice never depends on who sized it.
  int rc = plan_mips_small(a, ksel, &P->small);
  if (rc != GRB_OK) return rc;
  if (P->small.ok && (P->cap != P->auto_cap || 4 * P->auto_cap > a->X)) P->small.ok = 0;
  if (P->small.ok && P->small.total > P->total) P->total = P->small.total;
  return GRB_OK;
}

// implemented in mips_sm100.cu (tcgen05 score kernel, bf16 tables)
int mips_scores_sm100(const grb_mips_topk_args* a, const ScoreEpi& epi, int64_t n_launch_tiles,
                      cudaStream_t st);
bool mips_sm100_supported(const grb_mips_topk_args* a);

static int launch_scores(const grb_mips_topk_args* a, const ScoreEpi& epi, int64_t n_launch_tiles,
                         cudaStream_t st) {
  if (a->dtype == GRB_BF16 && mips_sm100_supported(a))
    return mips_scores_sm100(a, epi, n_launch_tiles, st);
  dim3 grid((unsigned) n_launch_tiles, (unsigned) ceil_div(a->B, SG_BM));
  if (a->dtype == GRB_F32)
    mips_scores_simt<float><<<grid, SG_THREADS, 0, st>>>(
        (const float*) a->queries, a->ldq, (const float*) a->items, a->ldi, a->B, a->X,
        (int) a->D, epi);
  else
    mips_scores_simt<__nv_bfloat16><<<grid, SG_THREADS, 0, st>>>(
        (const __nv_bfloat16*) a->queries, a->ldq, (const __nv_bfloat16*) a->items, a->ldi, a->B,
        a->X, (int) a->D, epi);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}  // namespace grb

using namespace grb;

extern "C" {

int64_t grb_mips_topk_workspace_bytes(grb_mips_topk_args* a) {
  MipsPlan P{};
  int rc = plan_mips(a, &P);
  if (rc != GRB_OK) return rc;
  a->sample_stride = P.stride;
  a->cand_cap = P.cap;
  return P.total;
}

int grb_mips_topk(const grb_mips_topk_args* a, grb_stream_t stream) {
  MipsPlan P{};
  int rc = plan_mips(a, &P);
  if (rc != GRB_OK) return rc;
  GRB_REQUIRE(a->dtype == GRB_F32 || a->dtype == GRB_BF16, GRB_ERR_INVALID_ARG,
              "mips_topk: dtype");
  GRB_REQUIRE(a->queries && a->items && a->out_scores && a->out_ids && a->workspace &&
                  a->status,
              GRB_ERR_INVALID_ARG, "mips_topk: null pointer");
  GRB_REQUIRE(a->workspace_bytes >= P.total, GRB_ERR_WORKSPACE,
              "mips_topk: workspace %lld < required %lld", (long long) a->workspace_bytes,
              (long long) P.total);
  GRB_REQUIRE(a->B <= 65535ll * SG_BM, GRB_ERR_UNSUPPORTED, "mips_topk: too many queries");
  if (a->B == 0) return GRB_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  auto ws = reinterpret_cast<unsigned char*>(a->workspace);
  float* tau = reinterpret_cast<float*>(ws + P.off_tau);
  int32_t* counts = reinterpret_cast<int32_t*>(ws + P.off_counts);
  float* sample = reinterpret_cast<float*>(ws + P.off_sample);
  float* cscores = reinterpret_cast<float*>(ws + P.off_cscores);
  int32_t* cidx = reinterpret_cast<int32_t*>(ws + P.off_cidx);
  if (P.small.ok) return run_mips_small(a, P.small, (int) mips_ksel(a), (int32_t) P.auto_cap, st);
  GRB_CUDA_OK(cudaMemsetAsync(counts, 0, a->B * 4, st));

  ScoreEpi epi{};
  epi.mode = MIPS_EPI_STORE; epi.tile_stride = P.stride; epi.grp = 1; epi.per = 1; epi.first = 0;
  epi.Xs = P.Xs; epi.out = sample;
  epi.tau = tau; epi.counts = counts; epi.cscores = cscores; epi.cidx = cidx; epi.cap = P.cap;
  rc = launch_scores(a, epi, P.n_sample_tiles, st);
  if (rc != GRB_OK) return rc;

  const int ksel = (int) mips_ksel(a);
  row_kth_largest_kernel<<<(unsigned) a->B, SEL_THREADS, 0, st>>>(sample, P.Xs, P.Xs, nullptr,
                                                                  ksel, tau);
  GRB_LAUNCH_OK();
  {
    GRB_REQUIRE(a->B <= 65535, GRB_ERR_UNSUPPORTED, "mips_topk: more than 65535 queries per call");
    epi.mode = MIPS_EPI_FILTER;
    dim3 grid((unsigned) ceil_div(P.Xs, FD_COLS), (unsigned) a->B);
    filter_dense_kernel<<<grid, 256, 0, st>>>(sample, P.Xs, P.Xs, P.stride, a->X, epi);
    GRB_LAUNCH_OK();
  }
  for (int ph = 1; ph <= P.levels; ++ph) {
    int64_t S = P.stride;
    for (int i = 0; i < ph; ++i) S /= MIPS_RADIX;
    const int64_t nS = ceil_div(P.n_tiles, S);
    const int64_t n_ph = nS - ceil_div(nS, MIPS_RADIX);
    if (n_ph > 0) {
      epi.mode = MIPS_EPI_FILTER; epi.tile_stride = S; epi.grp = MIPS_RADIX; epi.per = MIPS_RADIX - 1; epi.first = 1;
      rc = launch_scores(a, epi, n_ph, st);
      if (rc != GRB_OK) return rc;
    }
    if (ph < P.levels) {
      row_kth_largest_kernel<<<(unsigned) a->B, SEL_THREADS, 0, st>>>(cscores, P.cap, P.cap,
                                                                      counts, ksel, tau);
      GRB_LAUNCH_OK();
    }
  }
  SelectFilter flt{};
  flt.invalid = a->n_invalid > 0 ? a->invalid_ids : nullptr;
  flt.ld = a->ld_invalid; flt.n_invalid = a->n_invalid; flt.k_out = a->k;
  flt.target = a->target_ids; flt.ranks = a->target_ids ? a->out_ranks : nullptr;
  topk_select_kernel<int32_t><<<(unsigned) a->B, SEL_THREADS, 0, st>>>(
      cscores, cidx, counts, P.cap, ksel, a->item_ids, a->out_scores, a->out_ids, a->status, flt);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_topk_select(const float* cand_scores, const int64_t* cand_ids, const int32_t* counts,
                    int64_t B, int64_t cap, int32_t k, const int64_t* id_map, float* out_scores,
                    int64_t* out_ids, grb_stream_t stream) {
  GRB_REQUIRE(cand_scores && cand_ids && out_scores && out_ids && B >= 0 && cap > 0 && k > 0,
              GRB_ERR_INVALID_ARG, "topk_select: bad arguments");
  GRB_REQUIRE(k <= SEL_KMAX, GRB_ERR_UNSUPPORTED, "topk_select: k=%d exceeds %d", k, SEL_KMAX);
  if (B == 0) return GRB_OK;
  topk_select_kernel<int64_t><<<(unsigned) B, SEL_THREADS, 0,
                                reinterpret_cast<cudaStream_t>(stream)>>>(
      cand_scores, cand_ids, counts, cap, k, id_map, out_scores, out_ids, nullptr, SelectFilter{});
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}
