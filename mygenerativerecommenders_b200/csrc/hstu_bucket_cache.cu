// hstu_bucket_cache.cu — time-bucket index tiles, computed once per batch.
//
// bucket(b, i, j) = #{t : thr[t] <= |ts[b, i+1] - ts[b, j]|}  (hstu.py:113-123) does not depend on
// the layer, the head or the direction (forward / backward): the reference recomputes it as a
// (B, N, N) int64 tensor in every layer; here it is tabulated once per batch as uint8 and read
// by every attention launch of the step (2 x layers x heads uses).
//
// Layout: sequence b owns TPS = NT (NT + 1) / 2 tile slots (NT = ceil(max_len / 128)), slot
// t(iq, jk) = iq (iq + 1) / 2 + jk for key tile jk <= query tile iq.  A slot is 32 KiB:
//   [0, 16K)   "Q orientation"  byte ((c / 16) * 128 + r) * 16 + c % 16   (r = query row, c = key)
//   [16K, 32K) "K orientation"  byte ((r / 16) * 128 + c) * 16 + r % 16
// so a forward thread (one query row) or a backward thread (one key row) reads its 128 entries as
// eight conflict-free 16-byte shared-memory loads after one 16 KiB bulk copy.
#include "hstu_attn_sm100.cuh"

namespace grb {

// MASKED (the short-sequence attention kernels): pairs outside the causal triangle or past the end
// of the sequence (j > i or i >= n) hold 255; those kernels map bucket 255 to a bias of -15000, where
// tanh.approx saturates to exactly -1, so P and dS come out as exact zeros without any compare.
// With MASKED and ts == nullptr (no relative bias) the valid pairs hold 0.
template <bool MASKED>
__global__ void __launch_bounds__(256) hstu_bucket_tiles_kernel(
    const void* __restrict__ offsets, int index_bits, const int64_t* __restrict__ ts, int64_t N,
    const int64_t* __restrict__ thr, int nb, const uint32_t* __restrict__ octaves, int NT,
    uint8_t* __restrict__ cache) {
  __shared__ OctRec oct[32];
  __shared__ int flags[2];
  __shared__ int64_t tsq[128], tsk[128];
  const int b = blockIdx.y;
  // slot -> (iq, jk)
  const int slot = blockIdx.x;
  int iq = (int) ((sqrtf(8.f * slot + 1.f) - 1.f) * 0.5f);
  while ((iq + 1) * (iq + 2) / 2 <= slot) ++iq;
  while (iq * (iq + 1) / 2 > slot) --iq;
  const int jk = slot - iq * (iq + 1) / 2;
  const int64_t off0 = load_index(offsets, b, index_bits);
  int64_t n = load_index(offsets, b + 1, index_bits) - off0;
  if (n > N) n = N;
  if ((int64_t) iq * 128 >= n) return;
  const int tid = threadIdx.x;
  const bool has_ts = ts != nullptr;
  if (has_ts) {
    if (tid < 32) {
      if (octaves) load_octave_table(oct, flags, octaves, tid);
      else build_octave_table(oct, flags, thr, nb, tid);
    }
    if (tid >= 128) tsq[tid - 128] = ext_ts_at(ts, b, N, (int64_t) iq * 128 + (tid - 128) + 1);
    else tsk[tid] = ext_ts_at(ts, b, N, (int64_t) jk * 128 + tid);
  }
  __syncthreads();
  const bool slow = has_ts && flags[0] != 0;
  const int TPS = NT * (NT + 1) / 2;
  uint8_t* tile = cache + ((int64_t) b * TPS + slot) * 32768;
  const int orient = tid >> 7;          // 0: thread = query row, 1: thread = key row
  const int rr = tid & 127;
  const int64_t mine = has_ts ? (orient == 0 ? tsq[rr] : tsk[rr]) : 0;
  const int64_t* other = orient == 0 ? tsk : tsq;
  uint8_t* dst = tile + orient * 16384;
  // valid entries of this row (index e along the other dimension): lo <= e < hi
  int lo = 0, hi = 128;
  if (MASKED) {
    const int i0 = iq * 128, j0 = jk * 128;
    if (orient == 0) { lo = 0; hi = (i0 + rr < (int) n) ? (i0 + rr - j0 + 1) : 0; }   // row = query, e = key
    else { lo = j0 + rr - i0; hi = (int) n - i0; }                                     // row = key, e = query
  }
  for (int ch = 0; ch < 8; ++ch) {
    uint32_t w[4];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      uint32_t word = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = ch * 16 + q4 * 4 + e;
        uint32_t bk = 0;
        if (has_ts) {
          int64_t d = mine - other[idx];
          d = d < 0 ? -d : d;
          bk = (uint32_t) bucket_wide(oct, thr, nb, slow, d);
        }
        if (MASKED && !(idx >= lo && idx < hi)) bk = 255u;
        word |= bk << (8 * e);
      }
      w[q4] = word;
    }
    *reinterpret_cast<uint4*>(dst + ((size_t) ch * 128 + rr) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

}  // namespace grb

using namespace grb;

extern "C" {

int64_t grb_hstu_bucket_cache_bytes(int64_t B, int64_t max_len) {
  const int64_t NT = (max_len + 127) / 128;
  return B * (NT * (NT + 1) / 2) * 32768;
}

static int bucket_tiles_impl(const void* offsets, int index_bits, const int64_t* timestamps, int64_t B,
                             int64_t N, int64_t max_len, const int64_t* thresholds, int32_t num_buckets,
                             const uint32_t* octaves, void* cache, bool masked, grb_stream_t stream) {
  GRB_REQUIRE(index_bits == 32 || index_bits == 64, GRB_ERR_INVALID_ARG,
              "bucket_tiles: index_bits must be 32 or 64");
  GRB_REQUIRE(offsets && cache && B >= 0 && N > 0 && max_len >= 0 && max_len <= N, GRB_ERR_INVALID_ARG,
              "bucket_tiles: bad arguments");
  GRB_REQUIRE(masked || timestamps, GRB_ERR_INVALID_ARG, "bucket_tiles: timestamps are null");
  if (timestamps)
    GRB_REQUIRE(thresholds && num_buckets > 0 && num_buckets <= (masked ? 254 : 255), GRB_ERR_INVALID_ARG,
                "bucket_tiles: thresholds / num_buckets");
  GRB_REQUIRE(B <= 65535, GRB_ERR_UNSUPPORTED, "bucket_tiles: B <= 65535");
  const int NT = (int) ((max_len + 127) / 128);
  if (B == 0 || NT == 0) return GRB_OK;
  dim3 grid((unsigned) (NT * (NT + 1) / 2), (unsigned) B);
  auto st = reinterpret_cast<cudaStream_t>(stream);
  if (masked)
    hstu_bucket_tiles_kernel<true><<<grid, 256, 0, st>>>(offsets, index_bits, timestamps, N, thresholds,
                                                         num_buckets, octaves, NT, reinterpret_cast<uint8_t*>(cache));
  else
    hstu_bucket_tiles_kernel<false><<<grid, 256, 0, st>>>(offsets, index_bits, timestamps, N, thresholds,
                                                          num_buckets, octaves, NT, reinterpret_cast<uint8_t*>(cache));
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_hstu_bucket_tiles(const void* offsets, int index_bits, const int64_t* timestamps, int64_t B,
                          int64_t N, int64_t max_len, const int64_t* thresholds, int32_t num_buckets,
                          const uint32_t* octaves, void* cache, grb_stream_t stream) {
  return bucket_tiles_impl(offsets, index_bits, timestamps, B, N, max_len, thresholds, num_buckets, octaves,
                           cache, false, stream);
}

int grb_hstu_bucket_tiles_masked(const void* offsets, int index_bits, const int64_t* timestamps, int64_t B,
                                 int64_t N, int64_t max_len, const int64_t* thresholds, int32_t num_buckets,
                                 const uint32_t* octaves, void* cache, grb_stream_t stream) {
  return bucket_tiles_impl(offsets, index_bits, timestamps, B, N, max_len, thresholds, num_buckets, octaves,
                           cache, true, stream);
}

}
