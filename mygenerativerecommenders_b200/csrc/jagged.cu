// jagged.cu — offset cumsum and jagged <-> padded-dense conversions (HBM-bound byte movers).
//
// Reference call sites (under /root/reference/src/generative_recommenders_pl/models/utils/):
//   ops.py:18-38   asynchronous_complete_cumsum   -> grb_complete_cumsum
//   ops.py:41-64   dense_to_jagged                -> grb_dense_to_jagged
//   ops.py:67-114  jagged_to_padded_dense         -> grb_jagged_to_padded_dense
//   ops.py:171-187 get_current_embeddings         -> grb_gather_last_rows
//
// Layout insight: sequence b's jagged rows [off[b], off[b+1]) are ONE contiguous byte range that
// maps to the contiguous prefix of dense[b].  So both conversions are B independent contiguous
// copies (plus a pad fill), vectorised at the widest power of two that divides the row size and
// the base addresses.  One grid over (sequence, 16 KiB chunk); every thread keeps UNROLL
// independent 16-byte loads in flight before the first store.
#include "common.cuh"

namespace grb {

// ---------------------------------------------------------------------------------------------
// a1: complete cumsum.  One CTA, 1024 threads; a tile is 16 consecutive lengths per thread (all of a
// thread's loads issued before the first add: one round trip per 16 384 lengths, where a tile of one
// length per thread cost four barriers and a dependent load per 1024 — 16 us for the 11 k-entry offsets of
// the sampled-softmax backward), thread-local scan, block scan of the thread totals, running carry.
// ---------------------------------------------------------------------------------------------
constexpr int CS_ITEMS = 16;
template <typename IdxT>
__global__ void __launch_bounds__(1024) complete_cumsum_kernel(const IdxT* __restrict__ lengths,
                                                               IdxT* __restrict__ offsets,
                                                               int64_t B) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) { carry_s = 0; offsets[0] = (IdxT) 0; }
  __syncthreads();
  for (int64_t base = 0; base < B; base += 1024 * CS_ITEMS) {
    const int64_t i0 = base + (int64_t) tid * CS_ITEMS;
    long long x[CS_ITEMS];
#pragma unroll
    for (int k = 0; k < CS_ITEMS; ++k) x[k] = (i0 + k < B) ? (long long) lengths[i0 + k] : 0;
#pragma unroll
    for (int k = 1; k < CS_ITEMS; ++k) x[k] += x[k - 1];
    long long v = x[CS_ITEMS - 1];
    // inclusive warp scan of the thread totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    if (wid == 0) {
      long long w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      warp_tot[lane] = w;  // inclusive totals
    }
    __syncthreads();
    const long long before = carry_s + (wid > 0 ? warp_tot[wid - 1] : 0) + v - x[CS_ITEMS - 1];
#pragma unroll
    for (int k = 0; k < CS_ITEMS; ++k)
      if (i0 + k < B) offsets[i0 + k + 1] = (IdxT) (before + x[k]);
    __syncthreads();
    if (tid == 1023) carry_s = before + x[CS_ITEMS - 1];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// a2 / a3: per-sequence contiguous copies.
// ---------------------------------------------------------------------------------------------
template <int VEC> struct VecT;
template <> struct VecT<1>  { using type = unsigned char; };
template <> struct VecT<2>  { using type = unsigned short; };
template <> struct VecT<4>  { using type = unsigned int; };
template <> struct VecT<8>  { using type = uint2; };
template <> struct VecT<16> { using type = uint4; };

struct Pad16 { unsigned char b[16]; };

constexpr int kCopyThreads = 256;
constexpr int kCopyUnroll = 4;

// TO_DENSE = true : jagged -> padded dense (fills the tail with the pad pattern)
// TO_DENSE = false: dense -> jagged
template <int VEC, bool TO_DENSE>
__global__ void __launch_bounds__(kCopyThreads) jagged_copy_kernel(
    const unsigned char* __restrict__ src, unsigned char* __restrict__ dst,
    const void* __restrict__ offsets, int index_bits, int64_t N, int64_t row_bytes,
    int64_t batch_stride, int64_t chunks_per_seq, Pad16 pad) {
  using V = typename VecT<VEC>::type;
  const int64_t b = blockIdx.x / chunks_per_seq;
  const int64_t chunk = blockIdx.x % chunks_per_seq;
  const int64_t off0 = load_index(offsets, b, index_bits);
  const int64_t off1 = load_index(offsets, b + 1, index_bits);
  int64_t n = off1 - off0;
  if (n > N) n = N;
  if (n < 0) n = 0;
  const int64_t dense_vecs = N * row_bytes / VEC;   // vectors in dense[b]
  const int64_t valid_vecs = n * row_bytes / VEC;   // vectors that carry data
  const V* jag = reinterpret_cast<const V*>((TO_DENSE ? src : dst) + off0 * row_bytes);
  const V* den = reinterpret_cast<const V*>((TO_DENSE ? dst : src) + b * batch_stride);
  const V* s = TO_DENSE ? jag : den;
  V* d = const_cast<V*>(TO_DENSE ? den : jag);
  const int64_t limit = TO_DENSE ? dense_vecs : valid_vecs;
  const int64_t base = chunk * (int64_t) (kCopyThreads * kCopyUnroll) + threadIdx.x;
  if (base - threadIdx.x >= limit) return;
  V vals[kCopyUnroll];
  const V padv = *reinterpret_cast<const V*>(pad.b);
#pragma unroll
  for (int u = 0; u < kCopyUnroll; ++u) {
    int64_t i = base + (int64_t) u * kCopyThreads;
    if (i < valid_vecs) vals[u] = __ldcs(s + i);
    else vals[u] = padv;
  }
#pragma unroll
  for (int u = 0; u < kCopyUnroll; ++u) {
    int64_t i = base + (int64_t) u * kCopyThreads;
    if (i < limit) __stcs(d + i, vals[u]);
  }
}

static int pick_vec(const void* a, const void* b, int64_t row_bytes, int64_t batch_stride = 0) {
  uintptr_t x = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                (uintptr_t) row_bytes | (uintptr_t) batch_stride;
  int v = 16;
  while (v > 1 && (x & (uintptr_t) (v - 1))) v >>= 1;
  return v;
}

template <bool TO_DENSE>
static int launch_jagged_copy(const void* src, void* dst, const void* offsets, int64_t B,
                              int64_t N, int64_t row_bytes, int64_t batch_stride, int index_bits,
                              Pad16 pad, cudaStream_t st) {
  if (B == 0 || N == 0 || row_bytes == 0) return GRB_OK;
  if (batch_stride == 0) batch_stride = N * row_bytes;
  GRB_REQUIRE(batch_stride >= N * row_bytes, GRB_ERR_INVALID_ARG,
              "jagged copy: dense batch stride %lld < N*row_bytes %lld", (long long) batch_stride,
              (long long) (N * row_bytes));
  const int vec = pick_vec(src, dst, row_bytes, batch_stride);
  const int64_t dense_vecs = N * row_bytes / vec;
  const int64_t chunks = ceil_div(dense_vecs, (int64_t) kCopyThreads * kCopyUnroll);
  const int64_t grid = B * chunks;
  GRB_REQUIRE(grid < (1ll << 31), GRB_ERR_UNSUPPORTED, "jagged copy: grid too large (%lld)",
              (long long) grid);
  auto s = reinterpret_cast<const unsigned char*>(src);
  auto d = reinterpret_cast<unsigned char*>(dst);
#define GRB_JC(V)                                                                          \
  jagged_copy_kernel<V, TO_DENSE><<<(unsigned) grid, kCopyThreads, 0, st>>>(               \
      s, d, offsets, index_bits, N, row_bytes, batch_stride, chunks, pad)
  switch (vec) {
    case 16: GRB_JC(16); break;
    case 8:  GRB_JC(8); break;
    case 4:  GRB_JC(4); break;
    case 2:  GRB_JC(2); break;
    default: GRB_JC(1); break;
  }
#undef GRB_JC
  GRB_LAUNCH_OK();
  return GRB_OK;
}

// ---------------------------------------------------------------------------------------------
// a10: out[b] = dense[b, len[b]-1]   (flattened index like the reference, so len 0 wraps to the
// previous sequence's last row; b = 0, len = 0 would be row -1 and is clamped to a zero row).
// ---------------------------------------------------------------------------------------------
__global__ void gather_last_rows_kernel(const unsigned char* __restrict__ dense,
                                        const void* __restrict__ lengths,
                                        unsigned char* __restrict__ out, int64_t B, int64_t N,
                                        int64_t row_bytes, int index_bits, int vec, int scatter) {
  const int64_t b = blockIdx.x;
  const int64_t len = load_index(lengths, b, index_bits);
  int64_t flat = b * N + len - 1;
  const bool valid = flat >= 0 && flat < B * N;
  const unsigned char* s = dense + flat * row_bytes;
  unsigned char* o = out + b * row_bytes;
  if (scatter) {  // transpose: dense[flat] = out[b]
    if (!valid) return;
    unsigned char* dd = const_cast<unsigned char*>(dense) + flat * row_bytes;
    if (vec == 16) {
      for (int64_t i = threadIdx.x; i < row_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(dd)[i] = reinterpret_cast<const uint4*>(o)[i];
    } else {
      for (int64_t i = threadIdx.x; i < row_bytes; i += blockDim.x) dd[i] = o[i];
    }
    return;
  }
  if (vec == 16) {
    for (int64_t i = threadIdx.x; i < row_bytes / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(o)[i] =
          valid ? reinterpret_cast<const uint4*>(s)[i] : make_uint4(0, 0, 0, 0);
  } else {
    for (int64_t i = threadIdx.x; i < row_bytes; i += blockDim.x) o[i] = valid ? s[i] : 0;
  }
}

// every rank's block of local results -> the same column block of each peer's gather buffer
// (peer-to-peer stores; blockIdx.y = destination)
constexpr int P2P_MAX_DST = 16;
struct P2pDst { uint32_t* p[P2P_MAX_DST]; };
__global__ void __launch_bounds__(256) p2p_put_rows_kernel(const uint32_t* __restrict__ src, int64_t src_ld,
                                                           P2pDst dst, int64_t dst_ld, int64_t rows,
                                                           int64_t words) {
  uint32_t* out = dst.p[blockIdx.y];
  const int64_t total = rows * words;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t) gridDim.x * blockDim.x) {
    const int64_t r = i / words, c = i - r * words;
    out[r * dst_ld + c] = src[r * src_ld + c];
  }
}

// stream barrier across ranks: see grb_p2p_barrier
struct P2pSig { unsigned long long* p[P2P_MAX_DST]; };
__global__ void __launch_bounds__(32) p2p_barrier_kernel(P2pSig sig, int n_ranks, int rank, int slot,
                                                         unsigned long long target) {
  __threadfence_system();                       // earlier peer stores / reds of this stream first
  if ((int) threadIdx.x < n_ranks)
    atomicAdd_system(sig.p[threadIdx.x] + slot, 1ull);
  if (threadIdx.x == 0) {
    volatile unsigned long long* mine = sig.p[rank] + slot;
    while (*mine < target) __nanosleep(40);
    __threadfence_system();
  }
}

// two-shot all-reduce over peer memory, in place: rank r owns the float4 chunk range
// [r * per, (r + 1) * per): it reads that range from EVERY rank's buffer (peer loads over NVLink),
// sums in rank order (so every rank of a group computes bit-identical results), scales, and stores the
// result into the same range of EVERY rank's buffer (peer stores).  Nobody reads a range it does not
// own, so reading and writing the same buffers is safe between the two barriers the caller places
// around the launch.
struct P2pBufs { float4* p[P2P_MAX_DST]; };
__global__ void __launch_bounds__(256) p2p_allreduce_kernel(P2pBufs bufs, int n_ranks, int rank, int64_t n4,
                                                            float scale) {
  const int64_t per = (n4 + n_ranks - 1) / n_ranks;
  const int64_t lo = (int64_t) rank * per;
  int64_t hi = lo + per;
  if (hi > n4) hi = n4;
  for (int64_t i = lo + (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < hi;
       i += (int64_t) gridDim.x * blockDim.x) {
    float4 v[P2P_MAX_DST];
#pragma unroll
    for (int r = 0; r < P2P_MAX_DST; ++r)
      if (r < n_ranks) v[r] = bufs.p[r][i];          // all loads in flight before the first add
    float4 acc = v[0];
#pragma unroll
    for (int r = 1; r < P2P_MAX_DST; ++r)
      if (r < n_ranks) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
#pragma unroll
    for (int r = 0; r < P2P_MAX_DST; ++r)
      if (r < n_ranks) bufs.p[r][i] = acc;
  }
}

}  // namespace grb

using namespace grb;

extern "C" {

int grb_p2p_allreduce(void* const* bufs, int32_t n_ranks, int32_t rank, int64_t numel, float scale,
                      grb_stream_t stream) {
  GRB_REQUIRE(bufs && n_ranks > 0 && n_ranks <= P2P_MAX_DST && rank >= 0 && rank < n_ranks && numel >= 0 &&
                  numel % 4 == 0,
              GRB_ERR_INVALID_ARG, "p2p_allreduce: bad arguments (numel must be a multiple of 4)");
  if (numel == 0) return GRB_OK;
  P2pBufs b{};
  for (int i = 0; i < n_ranks; ++i) {
    GRB_REQUIRE(bufs[i] != nullptr && (reinterpret_cast<uintptr_t>(bufs[i]) & 15) == 0, GRB_ERR_INVALID_ARG,
                "p2p_allreduce: null or unaligned buffer");
    b.p[i] = reinterpret_cast<float4*>(bufs[i]);
  }
  const int64_t n4 = numel / 4, per = ceil_div(n4, (int64_t) n_ranks);
  int64_t blocks = ceil_div(per, (int64_t) 256);
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  p2p_allreduce_kernel<<<(unsigned) blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(b, n_ranks, rank,
                                                                                            n4, scale);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_complete_cumsum(const void* lengths, void* offsets, int64_t B, int index_bits,
                        grb_stream_t stream) {
  GRB_REQUIRE(index_bits == 32 || index_bits == 64, GRB_ERR_INVALID_ARG,
              "complete_cumsum: index_bits must be 32 or 64, got %d", index_bits);
  GRB_REQUIRE(B >= 0 && offsets != nullptr && (B == 0 || lengths != nullptr),
              GRB_ERR_INVALID_ARG, "complete_cumsum: bad arguments");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  if (index_bits == 32)
    complete_cumsum_kernel<int32_t><<<1, 1024, 0, st>>>(
        reinterpret_cast<const int32_t*>(lengths), reinterpret_cast<int32_t*>(offsets), B);
  else
    complete_cumsum_kernel<int64_t><<<1, 1024, 0, st>>>(
        reinterpret_cast<const int64_t*>(lengths), reinterpret_cast<int64_t*>(offsets), B);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_dense_to_jagged(const void* dense, const void* offsets, void* jagged, int64_t B,
                        int64_t N, int64_t row_bytes, int64_t dense_batch_stride_bytes,
                        int index_bits, grb_stream_t stream) {
  GRB_REQUIRE(index_bits == 32 || index_bits == 64, GRB_ERR_INVALID_ARG,
              "dense_to_jagged: index_bits must be 32 or 64");
  GRB_REQUIRE(B >= 0 && N >= 0 && row_bytes >= 0 && offsets, GRB_ERR_INVALID_ARG,
              "dense_to_jagged: bad arguments");
  Pad16 pad{};
  return launch_jagged_copy<false>(dense, jagged, offsets, B, N, row_bytes,
                                   dense_batch_stride_bytes, index_bits, pad,
                                   reinterpret_cast<cudaStream_t>(stream));
}

int grb_jagged_to_padded_dense(const void* jagged, const void* offsets, void* dense, int64_t B,
                               int64_t N, int64_t row_bytes, int64_t dense_batch_stride_bytes,
                               const void* pad_pattern, int elem_bytes, int index_bits,
                               grb_stream_t stream) {
  GRB_REQUIRE(index_bits == 32 || index_bits == 64, GRB_ERR_INVALID_ARG,
              "jagged_to_padded_dense: index_bits must be 32 or 64");
  GRB_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8,
              GRB_ERR_INVALID_ARG, "jagged_to_padded_dense: elem_bytes must be 1,2,4,8");
  GRB_REQUIRE(B >= 0 && N >= 0 && row_bytes >= 0 && offsets && row_bytes % elem_bytes == 0,
              GRB_ERR_INVALID_ARG, "jagged_to_padded_dense: bad arguments");
  Pad16 pad{};
  if (pad_pattern) {
    for (int i = 0; i < 16; ++i)
      pad.b[i] = reinterpret_cast<const unsigned char*>(pad_pattern)[i % elem_bytes];
  }
  return launch_jagged_copy<true>(jagged, dense, offsets, B, N, row_bytes,
                                  dense_batch_stride_bytes, index_bits, pad,
                                  reinterpret_cast<cudaStream_t>(stream));
}

int grb_gather_last_rows(const void* dense, const void* lengths, void* out, int64_t B, int64_t N,
                         int64_t row_bytes, int index_bits, int scatter, grb_stream_t stream) {
  GRB_REQUIRE(index_bits == 32 || index_bits == 64, GRB_ERR_INVALID_ARG,
              "gather_last_rows: index_bits must be 32 or 64");
  if (B == 0 || row_bytes == 0) return GRB_OK;
  const int vec = pick_vec(dense, out, row_bytes) == 16 ? 16 : 1;
  int threads = (int) (vec == 16 ? row_bytes / 16 : row_bytes);
  threads = threads < 32 ? 32 : (threads > 256 ? 256 : ((threads + 31) / 32) * 32);
  gather_last_rows_kernel<<<(unsigned) B, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const unsigned char*>(dense), lengths,
      reinterpret_cast<unsigned char*>(out), B, N, row_bytes, index_bits, vec, scatter);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_p2p_put_rows(const void* src, int64_t src_row_stride_bytes, void* const* dst, int32_t n_dst,
                     int64_t dst_row_stride_bytes, int64_t dst_col_offset_bytes, int64_t rows,
                     int64_t row_bytes, grb_stream_t stream) {
  using namespace grb;
  GRB_REQUIRE(src && dst && n_dst > 0 && n_dst <= P2P_MAX_DST && rows >= 0 && row_bytes > 0,
              GRB_ERR_INVALID_ARG, "p2p_put_rows: bad arguments");
  GRB_REQUIRE(row_bytes % 4 == 0 && src_row_stride_bytes % 4 == 0 && dst_row_stride_bytes % 4 == 0 &&
                  dst_col_offset_bytes % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0,
              GRB_ERR_INVALID_ARG, "p2p_put_rows: sizes and strides must be multiples of 4 bytes");
  if (rows == 0) return GRB_OK;
  P2pDst d{};
  for (int i = 0; i < n_dst; ++i) {
    GRB_REQUIRE(dst[i] != nullptr, GRB_ERR_INVALID_ARG, "p2p_put_rows: null destination");
    d.p[i] = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(dst[i]) + dst_col_offset_bytes);
  }
  const int64_t words = row_bytes / 4;
  const int64_t total = rows * words;
  const unsigned bx = (unsigned) (ceil_div(total, 256) < 4096 ? ceil_div(total, 256) : 4096);
  p2p_put_rows_kernel<<<dim3(bx, (unsigned) n_dst), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint32_t*>(src), src_row_stride_bytes / 4, d, dst_row_stride_bytes / 4, rows,
      words);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_p2p_barrier(void* const* signals, int32_t n_ranks, int32_t rank, int32_t slot, int64_t epoch,
                    grb_stream_t stream) {
  using namespace grb;
  GRB_REQUIRE(signals && n_ranks > 0 && n_ranks <= P2P_MAX_DST && rank >= 0 && rank < n_ranks &&
                  slot >= 0 && epoch > 0,
              GRB_ERR_INVALID_ARG, "p2p_barrier: bad arguments");
  P2pSig s{};
  for (int i = 0; i < n_ranks; ++i) {
    GRB_REQUIRE(signals[i] != nullptr, GRB_ERR_INVALID_ARG, "p2p_barrier: null signal array");
    s.p[i] = reinterpret_cast<unsigned long long*>(signals[i]);
  }
  p2p_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      s, n_ranks, rank, slot, (unsigned long long) n_ranks * (unsigned long long) epoch);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}
