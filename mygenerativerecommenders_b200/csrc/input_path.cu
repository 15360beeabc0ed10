// input_path.cu — the steps either side of the HSTU stack as jagged kernels.
//
// Reference (paths under /root/reference/src/generative_recommenders_pl/models/):
//   embeddings/embeddings.py:94-97            item_emb[past_ids]                       (B, N, D) fp32
//   preprocessors/learnable_positional_embedding.py:42-58
//       x = emb * sqrt(D) + pos_emb[0..N) ; dropout ; x *= (past_ids != 0)
//   sequential_encoders/hstu.py:502           dense_to_jagged(x) -> (T, D)
//   postprocessors/postprocessors.py:47-55    x / clamp(||x||, eps)
// The reference materialises three padded (B, N, D) fp32 tensors (plus the dropout mask) before
// it drops the padding; here one kernel gathers the table rows of the VALID positions only,
// applies scale + positional embedding + dropout and writes the jagged rows in the compute dtype.
// Backward: one kernel scatter-adds into the table gradient and the positional-embedding gradient
// (16-byte vector reds), regenerating the dropout mask from the same counter-based generator.
//
// Dropout mask: Philox4x32-10 keyed by a 64-bit seed the caller draws per step (torch's generator,
// so it is CUDA-graph safe), counter = element index / 4; keep iff u32 >= p * 2^32, scaled 1/(1-p).
// Same distribution as torch's dropout, not the same stream (parity tests run with p = 0 / eval).
#include "common.cuh"

namespace grb {
namespace {

// ---------------------------------------------------------------------------------------------
// b5  in-batch cache: the distinct ids of a batch, ascending, in two launches and without a sync.
// negative_sampler.py:187-196 keeps `torch.unique(ids[presences])` (sorted ascending; the draw of :208-211
// indexes it, so the order is part of the contract).  With a small id space that is a membership table
// and a compaction.  The table is persistent and stamped with an epoch instead of being cleared:
//   mark     for the first cnt_b = offsets[b+1] - offsets[b] + rows_extra ids of row b (id 0 = padding, ids
//            past the table are skipped): old = exch(flags[id], epoch); the first marker of an id also
//            counts it in chunk_cnt[id / 1024]
//   compact  one CTA: exclusive scan of the chunk counts; every warp walks its chunks (32 coalesced loads,
//            32 ballots: lane l holds the membership mask of ids [1024 c + 32 l, + 32)) and writes the ids
//            in ascending order to uniq[0 .. count); uniq[count .. n_out) = 0, count stored, chunk counts
//            cleared for the next call, epoch += 1.
// (Replaces dense_to_jagged + arange + compare + where + zeros + scatter_ + zero_ + sum + the four kernels
//  of nonzero_static: 12 launches of the train step.)
// ---------------------------------------------------------------------------------------------
constexpr int IBC_CHUNK = 1024;
constexpr int IBC_MAX_CHUNKS = 4096;      // ids < 2^22
__global__ void __launch_bounds__(256) inbatch_mark_kernel(const int64_t* __restrict__ ids, int64_t N,
                                                           const void* __restrict__ offsets, int index_bits,
                                                           int rows_extra, int64_t n_flags,
                                                           int32_t* __restrict__ flags, int32_t* __restrict__ chunk_cnt,
                                                           const int32_t* __restrict__ epoch) {
  const int64_t b = blockIdx.y;
  int64_t cnt = load_index(offsets, b + 1, index_bits) - load_index(offsets, b, index_bits) + rows_extra;
  if (cnt > N) cnt = N;
  const int32_t e = epoch[0];
  for (int64_t j = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += (int64_t) gridDim.x * blockDim.x) {
    const int64_t id = ids[b * N + j];
    if (id > 0 && id < n_flags && atomicExch(flags + id, e) != e) atomicAdd(chunk_cnt + (id / IBC_CHUNK), 1);
  }
}

constexpr int IBC_THREADS = 1024;
__global__ void __launch_bounds__(IBC_THREADS) inbatch_compact_kernel(const int32_t* __restrict__ flags,
                                                                      int64_t n_flags, int32_t* __restrict__ chunk_cnt,
                                                                      int32_t* __restrict__ epoch,
                                                                      int64_t* __restrict__ uniq, int64_t n_out,
                                                                      int64_t* __restrict__ count) {
  __shared__ int chunk_base[IBC_MAX_CHUNKS];
  __shared__ int wtot[32];
  __shared__ int total_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int32_t e = epoch[0];
  const int n_chunks = (int) ((n_flags + IBC_CHUNK - 1) / IBC_CHUNK);
  // exclusive scan of the chunk counts: thread t owns chunks [t * per, (t + 1) * per)
  const int per = (n_chunks + IBC_THREADS - 1) / IBC_THREADS;      // <= 4
  int cc[4], c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ch = tid * per + k;
    cc[k] = (k < per && ch < n_chunks) ? chunk_cnt[ch] : 0;
    c += cc[k];
  }
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wtot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = wtot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += v;
    }
    wtot[lane] = w;                       // inclusive warp totals
    if (lane == 31) total_s = w;
  }
  __syncthreads();
  int run = (wid > 0 ? wtot[wid - 1] : 0) + incl - c;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ch = tid * per + k;
    if (k < per && ch < n_chunks) {
      chunk_base[ch] = run;
      run += cc[k];
      chunk_cnt[ch] = 0;                  // clean for the next call
    }
  }
  __syncthreads();
  // chunks with members, round-robin over the warps
  for (int ch = wid; ch < n_chunks; ch += 32) {
    const int have = (ch + 1 < n_chunks ? chunk_base[ch + 1] : total_s) - chunk_base[ch];
    if (have == 0) continue;              // warp-uniform
    int32_t v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) {        // all loads before the first ballot
      const int64_t i = (int64_t) ch * IBC_CHUNK + 32 * k + lane;
      v[k] = i < n_flags ? flags[i] : 0;  // e >= 1: 0 is never a member
    }
    uint32_t m = 0u;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const uint32_t word = __ballot_sync(0xffffffffu, v[k] == e);
      if (lane == k) m = word;
    }
    if (ch == 0) m &= lane == 0 ? ~1u : ~0u;     // id 0 is padding (never marked; belt and braces)
    const int cw = __popc(m);
    int inw = cw;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inw, o);
      if (lane >= o) inw += t;
    }
    int64_t pos = chunk_base[ch] + inw - cw;
    const int64_t id0 = (int64_t) ch * IBC_CHUNK + 32 * lane;
    while (m) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      if (pos < n_out) uniq[pos] = id0 + bit;
      ++pos;
    }
  }
  const int total = total_s;
  for (int64_t i = total + tid; i < n_out; i += IBC_THREADS) uniq[i] = 0;
  if (tid == 0) {
    count[0] = total < n_out ? total : n_out;
    epoch[0] = e == 0x7fffffff ? 1 : e + 1;   // (a wrap would need a cleared table: 2^31 steps away)
  }
}

// jagged row t -> (sequence b, position i); offsets ascending, off[B] = valid rows.  Warp-uniform.
__device__ __forceinline__ bool locate_row(const void* __restrict__ offsets, int index_bits, int B, int64_t t,
                                           int& b, int64_t& i) {
  if (t >= load_index(offsets, B, index_bits)) return false;
  int lo = 0, hi = B;                       // largest b with off[b] <= t
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (load_index(offsets, mid, index_bits) <= t) lo = mid; else hi = mid;
  }
  b = lo;
  i = t - load_index(offsets, lo, index_bits);
  return true;
}

struct InP {
  const float* table; int64_t ldt; int64_t V;
  const int64_t* ids; int64_t N;              // (B, N)
  const void* offsets; int index_bits; int B;
  const float* pos; int64_t ldp;              // (>= N, D)
  float scale, p_drop, keep_scale;
  const int64_t* seed;                        // device scalar, or nullptr when p_drop == 0
  int64_t rows; int D;
  void* out; int64_t ldo; int out_bf16;       // forward output / backward input (dy)
  float* d_table; float* d_pos;               // backward
};

template <bool BWD>
__global__ void __launch_bounds__(256) jagged_input_kernel(InP P) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t) blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= P.rows) return;
  int b = 0;
  int64_t i = 0;
  int64_t id = 0;
  if (locate_row(P.offsets, P.index_bits, P.B, t, b, i) && i < P.N) id = P.ids[(int64_t) b * P.N + i];
  const bool live = id > 0 && id < P.V;
  uint32_t k0 = 0u, k1 = 0u;
  const uint32_t thr = P.p_drop > 0.f ? (uint32_t) fminf(P.p_drop * 4294967296.0f, 4294967295.0f) : 0u;
  if (P.p_drop > 0.f) { const uint64_t s = (uint64_t) *P.seed; k0 = (uint32_t) s; k1 = (uint32_t) (s >> 32); }
  for (int c = 4 * lane; c < P.D; c += 128) {
    float m[4] = {1.f, 1.f, 1.f, 1.f};
    if (P.p_drop > 0.f) {
      const uint64_t ctr = ((uint64_t) t * (uint64_t) P.D + (uint64_t) c) >> 2;
      const uint4 u = philox4x32_10((uint32_t) ctr, (uint32_t) (ctr >> 32), k0, k1);
      m[0] = u.x >= thr ? P.keep_scale : 0.f; m[1] = u.y >= thr ? P.keep_scale : 0.f;
      m[2] = u.z >= thr ? P.keep_scale : 0.f; m[3] = u.w >= thr ? P.keep_scale : 0.f;
    }
    if (!BWD) {
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        const float4 e = *reinterpret_cast<const float4*>(P.table + id * P.ldt + c);
        const float4 pe = *reinterpret_cast<const float4*>(P.pos + i * P.ldp + c);
        y.x = fmaf(e.x, P.scale, pe.x) * m[0]; y.y = fmaf(e.y, P.scale, pe.y) * m[1];
        y.z = fmaf(e.z, P.scale, pe.z) * m[2]; y.w = fmaf(e.w, P.scale, pe.w) * m[3];
      }
      if (P.out_bf16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo); o.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(P.out) + t * P.ldo + c) = o;
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(P.out) + t * P.ldo + c) = y;
      }
    } else if (live) {
      float4 g;
      if (P.out_bf16) {
        const uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(P.out) + t * P.ldo + c);
        g.x = __uint_as_float(raw.x << 16); g.y = __uint_as_float(raw.x & 0xffff0000u);
        g.z = __uint_as_float(raw.y << 16); g.w = __uint_as_float(raw.y & 0xffff0000u);
      } else {
        g = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(P.out) + t * P.ldo + c);
      }
      g.x *= m[0]; g.y *= m[1]; g.z *= m[2]; g.w *= m[3];
      if (P.d_pos)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(P.d_pos + i * (int64_t) P.D + c),
                     "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w) : "memory");
      if (P.d_table)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(P.d_table + id * (int64_t) P.D + c),
                     "f"(g.x * P.scale), "f"(g.y * P.scale), "f"(g.z * P.scale), "f"(g.w * P.scale) : "memory");
    }
  }
}

// y (rows, W) fp32 = x / max(||x||, eps), x bf16 or fp32 (the encoder's jagged output in the compute
// dtype: the cast to fp32 is folded in); inv as grb_l2norm_fwd.  Backward: dx in x's dtype.
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256) l2norm_cast_fwd_kernel(const T* __restrict__ x, int64_t ldx,
                                                              float* __restrict__ y, int64_t ldy,
                                                              float* __restrict__ inv, int64_t rows, int W, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t) blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * ldx;
  float q = 0.f;
  for (int c = lane; c < W; c += 32) { const float v = ldf(xr + c); q = fmaf(v, v, q); }
  const float nrm = sqrtf(warp_sum(q));
  const float iv = 1.0f / fmaxf(nrm, eps);
  if (lane == 0) inv[row] = nrm >= eps ? iv : -iv;
  float* yr = y + row * ldy;
  for (int c = lane; c < W; c += 32) yr[c] = ldf(xr + c) * iv;
}
template <typename T>
__global__ void __launch_bounds__(256) l2norm_cast_bwd_kernel(const float* __restrict__ y, int64_t ldy,
                                                              const float* __restrict__ dy, int64_t lddy,
                                                              const float* __restrict__ inv, T* __restrict__ dx,
                                                              int64_t lddx, int64_t rows, int W) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t) blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* yr = y + row * ldy;
  const float* gr = dy + row * lddy;
  T* dr = dx + row * lddx;
  const float iv = inv[row];
  if (iv < 0.f) {
    for (int c = lane; c < W; c += 32) stf(dr + c, -iv * gr[c]);
    return;
  }
  float d = 0.f;
  for (int c = lane; c < W; c += 32) d = fmaf(yr[c], gr[c], d);
  d = warp_sum(d);
  for (int c = lane; c < W; c += 32) stf(dr + c, iv * (gr[c] - yr[c] * d));
}

// rows >= offsets[B] of n_mats row-major matrices (same shape, mat_stride bytes apart) := 0.  Fixed row
// buckets (CUDA-graph mode) carry up to a bucket of padding rows the attention kernels never write;
// clearing only those replaces a memset of the whole (T_pad, W) output.
__global__ void __launch_bounds__(256) zero_tail_rows_kernel(unsigned char* __restrict__ base, int64_t ld_bytes,
                                                             int64_t rows, int64_t row_bytes, int64_t mat_stride,
                                                             const void* __restrict__ offsets, int index_bits,
                                                             int B) {
  const int64_t first = load_index(offsets, B, index_bits);
  const int64_t r = first + (int64_t) blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  unsigned char* p = base + (int64_t) blockIdx.y * mat_stride + r * ld_bytes;
  const int lane = threadIdx.x & 31;
  for (int64_t c = 16 * lane; c < row_bytes; c += 512) *reinterpret_cast<uint4*>(p + c) = make_uint4(0u, 0u, 0u, 0u);
}

// in-batch negatives draw (negative_sampler.py:208-211: randint(0, X_b) then a gather of the cached
// ids) as one pass: offset = Philox(seed, i) mod count with the count read on the device, id =
// cached_ids[offset].  62 random bits per draw: modulo bias < count / 2^62.
__global__ void __launch_bounds__(256) draw_negatives_kernel(const int64_t* __restrict__ seed,
                                                             const int64_t* __restrict__ count,
                                                             const int64_t* __restrict__ cached_ids, int64_t n,
                                                             int64_t* __restrict__ offsets,
                                                             int64_t* __restrict__ ids) {
  const int64_t i2 = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) * 2;   // two draws per Philox call
  if (i2 >= n) return;
  const uint64_t sd = (uint64_t) seed[0];
  const uint64_t ctr = (uint64_t) (i2 >> 1);
  const uint4 rnd = philox4x32_10((uint32_t) ctr, (uint32_t) (ctr >> 32), (uint32_t) sd, (uint32_t) (sd >> 32));
  uint64_t cnt = (uint64_t) count[0];
  if (cnt < 1) cnt = 1;
  const uint64_t d0 = ((((uint64_t) rnd.x << 32) | rnd.y) >> 2) % cnt;
  const uint64_t d1 = ((((uint64_t) rnd.z << 32) | rnd.w) >> 2) % cnt;
  offsets[i2] = (int64_t) d0;
  ids[i2] = cached_ids[d0];
  if (i2 + 1 < n) {
    offsets[i2 + 1] = (int64_t) d1;
    ids[i2 + 1] = cached_ids[d1];
  }
}

}  // namespace
}  // namespace grb

using namespace grb;

extern "C" {

static int jagged_input_launch(const grb_jagged_input_args* a, bool bwd, grb_stream_t stream) {
  GRB_REQUIRE(a != nullptr, GRB_ERR_INVALID_ARG, "jagged_input: null args");
  GRB_REQUIRE(a->index_bits == 32 || a->index_bits == 64, GRB_ERR_INVALID_ARG, "jagged_input: index_bits");
  GRB_REQUIRE(a->ids && a->offsets && a->io && a->B >= 0 && a->N > 0 && a->D > 0 && a->D % 4 == 0 && a->V > 0 &&
                  a->rows >= 0,
              GRB_ERR_INVALID_ARG, "jagged_input: bad arguments (D must be a multiple of 4)");
  GRB_REQUIRE(a->dtype == GRB_F32 || a->dtype == GRB_BF16, GRB_ERR_INVALID_ARG, "jagged_input: dtype");
  GRB_REQUIRE(a->p_drop >= 0.f && a->p_drop < 1.f && (a->p_drop == 0.f || a->seed), GRB_ERR_INVALID_ARG,
              "jagged_input: dropout needs 0 <= p < 1 and a seed");
  if (!bwd) GRB_REQUIRE(a->table && a->pos, GRB_ERR_INVALID_ARG, "jagged_input_fwd: table / pos are null");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int esz = a->dtype == GRB_BF16 ? 2 : 4;
  GRB_REQUIRE(al16(a->table) && al16(a->pos) && al16(a->d_table) && al16(a->d_pos) &&
                  (reinterpret_cast<uintptr_t>(a->io) & 7) == 0 && (a->ldt * 4) % 16 == 0 && (a->ldp * 4) % 16 == 0 &&
                  (a->ldio * esz) % 8 == 0,
              GRB_ERR_INVALID_ARG, "jagged_input: rows must be 16-byte aligned");
  if (a->rows == 0 || a->B == 0) return GRB_OK;
  InP P{};
  P.table = a->table; P.ldt = a->ldt; P.V = a->V;
  P.ids = a->ids; P.N = a->N;
  P.offsets = a->offsets; P.index_bits = a->index_bits; P.B = (int) a->B;
  P.pos = a->pos; P.ldp = a->ldp;
  P.scale = a->scale; P.p_drop = a->p_drop; P.keep_scale = 1.0f / (1.0f - a->p_drop);
  P.seed = a->seed;
  P.rows = a->rows; P.D = a->D;
  P.out = a->io; P.ldo = a->ldio; P.out_bf16 = a->dtype == GRB_BF16;
  P.d_table = a->d_table; P.d_pos = a->d_pos;
  const unsigned grid = (unsigned) ceil_div(a->rows, 8);
  auto st = reinterpret_cast<cudaStream_t>(stream);
  if (bwd) jagged_input_kernel<true><<<grid, 256, 0, st>>>(P);
  else jagged_input_kernel<false><<<grid, 256, 0, st>>>(P);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_zero_tail_rows(void* base, int64_t ld_bytes, int64_t rows, int64_t row_bytes, int32_t n_mats,
                       int64_t mat_stride_bytes, const void* offsets, int32_t index_bits, int64_t B,
                       int64_t max_tail_rows, grb_stream_t stream) {
  GRB_REQUIRE(base && offsets && rows >= 0 && row_bytes > 0 && row_bytes % 16 == 0 && ld_bytes % 16 == 0 &&
                  (reinterpret_cast<uintptr_t>(base) & 15) == 0 && n_mats > 0 && mat_stride_bytes % 16 == 0 &&
                  (index_bits == 32 || index_bits == 64) && B >= 0 && max_tail_rows >= 0,
              GRB_ERR_INVALID_ARG, "zero_tail_rows: bad arguments (16-byte aligned rows)");
  if (rows == 0 || max_tail_rows == 0) return GRB_OK;
  const int64_t span = max_tail_rows < rows ? max_tail_rows : rows;
  zero_tail_rows_kernel<<<dim3((unsigned) ceil_div(span, (int64_t) 8), (unsigned) n_mats), 256, 0,
                          reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<unsigned char*>(base), ld_bytes, rows, row_bytes, mat_stride_bytes, offsets, index_bits,
      (int) B);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_draw_negatives(const int64_t* seed, const int64_t* count, const int64_t* cached_ids, int64_t n,
                       int64_t* offsets, int64_t* ids, grb_stream_t stream) {
  GRB_REQUIRE(seed && count && cached_ids && offsets && ids && n >= 0, GRB_ERR_INVALID_ARG,
              "draw_negatives: bad arguments");
  if (n == 0) return GRB_OK;
  draw_negatives_kernel<<<(unsigned) ceil_div(ceil_div(n, (int64_t) 2), (int64_t) 256), 256, 0,
                          reinterpret_cast<cudaStream_t>(stream)>>>(seed, count, cached_ids, n, offsets, ids);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_inbatch_distinct_ids(const int64_t* ids, int64_t B, int64_t N, const void* offsets, int32_t index_bits,
                             int32_t rows_extra, int32_t* flags, int64_t n_flags, int32_t* epoch, int64_t* uniq,
                             int64_t n_out, int64_t* count, grb_stream_t stream) {
  GRB_REQUIRE(ids && offsets && flags && epoch && uniq && count && B >= 0 && N > 0 && n_flags > 1 && n_out >= 0 &&
                  B <= 65535 && n_flags <= (int64_t) IBC_MAX_CHUNKS * IBC_CHUNK && (index_bits == 32 || index_bits == 64),
              GRB_ERR_INVALID_ARG, "inbatch_distinct_ids: bad arguments");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  int32_t* chunk_cnt = flags + n_flags;       // the table's tail: one counter per 1024 ids
  if (B > 0) {
    inbatch_mark_kernel<<<dim3((unsigned) ceil_div(N, (int64_t) 256), (unsigned) B), 256, 0, st>>>(
        ids, N, offsets, index_bits, rows_extra, n_flags, flags, chunk_cnt, epoch);
    GRB_LAUNCH_OK();
  }
  inbatch_compact_kernel<<<1, IBC_THREADS, 0, st>>>(flags, n_flags, chunk_cnt, epoch, uniq, n_out, count);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_jagged_input_fwd(const grb_jagged_input_args* a, grb_stream_t stream) {
  return jagged_input_launch(a, false, stream);
}
int grb_jagged_input_bwd(const grb_jagged_input_args* a, grb_stream_t stream) {
  return jagged_input_launch(a, true, stream);
}

int grb_l2norm_cast_fwd(const void* x, int64_t ldx, int dtype, float* y, int64_t ldy, float* inv, int64_t rows,
                        int64_t W, float eps, grb_stream_t stream) {
  GRB_REQUIRE(x && y && inv && rows >= 0 && W > 0 && W < (1 << 30) && eps > 0.f, GRB_ERR_INVALID_ARG,
              "l2norm_cast_fwd: bad arguments");
  GRB_REQUIRE(dtype == GRB_F32 || dtype == GRB_BF16, GRB_ERR_INVALID_ARG, "l2norm_cast_fwd: dtype");
  if (rows == 0) return GRB_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned) ceil_div(rows, 8);
  if (dtype == GRB_BF16)
    l2norm_cast_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*) x, ldx, y, ldy, inv, rows, (int) W, eps);
  else
    l2norm_cast_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*) x, ldx, y, ldy, inv, rows, (int) W, eps);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_l2norm_cast_bwd(const float* y, int64_t ldy, const float* dy, int64_t lddy, const float* inv, void* dx,
                        int64_t lddx, int dtype, int64_t rows, int64_t W, grb_stream_t stream) {
  GRB_REQUIRE(y && dy && inv && dx && rows >= 0 && W > 0, GRB_ERR_INVALID_ARG, "l2norm_cast_bwd: bad arguments");
  GRB_REQUIRE(dtype == GRB_F32 || dtype == GRB_BF16, GRB_ERR_INVALID_ARG, "l2norm_cast_bwd: dtype");
  if (rows == 0) return GRB_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned) ceil_div(rows, 8);
  if (dtype == GRB_BF16)
    l2norm_cast_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(y, ldy, dy, lddy, inv, (__nv_bfloat16*) dx, lddx, rows, (int) W);
  else
    l2norm_cast_bwd_kernel<float><<<grid, 256, 0, st>>>(y, ldy, dy, lddy, inv, (float*) dx, lddx, rows, (int) W);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}
