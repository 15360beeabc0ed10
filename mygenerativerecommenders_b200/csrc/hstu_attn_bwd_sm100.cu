// hstu_attn_bwd_sm100.cu — HSTU jagged attention backward on Blackwell tensor cores
// (bf16 in, fp32 accumulate in TMEM, dqk = dv = 64).
//
// Math (backward of hstu.py:186-204 with the bias of :96-128), per sequence / head, j <= i < n:
//   S = Q K^T + bias ; P = SiLU(S)/N ; dP = dO V^T ; dS = dP * SiLU'(S)/N
//   dV = P^T dO ; dK = dS^T Q ; dQ = dS K ; d pos_w[N-1+j-i] += dS ; d ts_w[bucket] += dS
//
// One CTA owns (sequence b, head h, key tile j) and walks the query tiles i = j .. last.  Scores
// are produced TRANSPOSED (thread = key row) so that P^T and dS^T land in shared memory as
// K-major A operands; the same dS^T tile read MN-major is the A operand of dQ.
//
//   TMA warp : K_j, V_j once; Q_i, dO_i through a 2-stage ring (128-byte swizzle, jagged rows).
//   MMA warp : S^T  = K_j Q_i^T   M128 N128 K64   (K-major x K-major)        -> TMEM [0,128)
//              dP^T = V_j dO_i^T  M128 N128 K64                              -> TMEM [128,256)
//              dV  += P^T  dO_i   M128 N64  K128  (smem K-major x MN-major)  -> TMEM [256,320)
//              dK  += dS^T Q_i    M128 N64  K128                             -> TMEM [320,384)
//              dQ_i = dS   K_j    M128 N64  K128  (MN-major x MN-major)      -> TMEM [384,448)
//   epilogue : 2 warpgroups, thread = key row, each warpgroup half of the 128 query columns:
//              bias (integer bucketing), tanh-based SiLU / SiLU', masks, bf16 P^T / dS^T into
//              swizzled shared memory, bias-gradient partial sums; then dQ_i is read back and
//              added to the fp32 dq accumulator with vector red.global.
#include "hstu_attn_sm100.cuh"

namespace grb {

using namespace ptx;

constexpr int AB_NWG = 4;                       // epilogue warpgroups: 32 query columns each
constexpr int AB_EPI = AB_NWG * 128;
constexpr int AB_THREADS = 128 + AB_EPI;

struct AttnBwdParams {
  int64_t N, T;
  int H, nb, index_bits, n_kt;
  const void* offsets;
  const int64_t* ts;
  const float* ts_w;
  const float* pos_w;
  const int64_t* thr;
  const uint32_t* octaves;   // optional precomputed octave table
  const uint8_t* bcache;     // optional bucket-index tiles (hstu_bucket_cache.cu)
  int cache_nt;
  __nv_bfloat16* dk; int64_t lddk;
  __nv_bfloat16* dv; int64_t lddv;
  float* dq_accum;          // (T, H*64) fp32, zero-filled by the caller
  float* d_ts_w; float* d_pos_w;
  int d_bias_copies;
};

struct AbSmem {
  static constexpr int k = 0;
  static constexpr int v = k + AT_TILE_BYTES;
  static constexpr int ring = v + AT_TILE_BYTES;                 // 2 x (Q, dO)
  static constexpr int pT = ring + 4 * AT_TILE_BYTES;            // 2 blocks [128 k][64 q]
  static constexpr int dsT = pT + 2 * AT_TILE_BYTES;
  // query-side tables are double buffered by tile parity: one named barrier per tile orders
  // staging against use
  static constexpr int tsq = dsT + 2 * AT_TILE_BYTES;            // 2 x 128 x int64
  static constexpr int tsq32 = tsq + 2 * 128 * 8;                // 2 x 128 x uint32
  static constexpr int red = tsq32 + 2 * 128 * 4;                // 2 x 16 x int64
  static constexpr int pos = red + 32 * 8;                       // 2 x 256 x float (pre-halved)
  static constexpr int tsw = pos + 2 * 256 * 4;                  // 136 x float (pre-halved)
  static constexpr int oct = tsw + 136 * 4;                      // 32 x OctRec
  static constexpr int h_ts = oct + 32 * 16;                     // 16 warps x 136 x float
  // plain (unswizzled) bf16 copy of dS^T [128 key rows][128 query cols], row stride 272 B: the
  // d pos_w diagonal sums read it after the tile barrier
  static constexpr int ds_plain = h_ts + 16 * 136 * 4;
  static constexpr int DS_STRIDE = 272;
  static constexpr int bkt = ds_plain + 128 * DS_STRIDE;       // 128 x 128 uint8 bucket tile
  static constexpr int bars = bkt + 128 * 128;
  static constexpr int total = bars + 256;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

template <bool HAS_BIAS>
__global__ void __launch_bounds__(AB_THREADS, 1) hstu_attn_bwd_sm100_kernel(
    const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
    AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  using L = AbSmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kt = (int) blockIdx.x;                 // key tile: early tiles have the most work
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int64_t off0 = load_index(p.offsets, b, p.index_bits);
  int64_t n64 = load_index(p.offsets, b + 1, p.index_bits) - off0;
  if (n64 > p.N) n64 = p.N;
  const int n = (int) n64;
  const int j0 = kt * AT_BN;
  if (j0 >= n) return;
  const int n_qt = (n + AT_BM - 1) / AT_BM;
  const int n_it = n_qt - kt;                      // query tiles kt .. n_qt-1

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  const uint32_t bar_kv = smem_u32(bars + 0);
  const uint32_t bar_ring_full = smem_u32(bars + 1);    // [2]
  const uint32_t bar_ring_empty = smem_u32(bars + 3);   // [2]
  const uint32_t bar_s_full = smem_u32(bars + 5);
  const uint32_t bar_s_free = smem_u32(bars + 6);
  const uint32_t bar_pds_full = smem_u32(bars + 7);
  const uint32_t bar_pds_free = smem_u32(bars + 8);
  const uint32_t bar_dq_full = smem_u32(bars + 9);
  const uint32_t bar_dq_free = smem_u32(bars + 10);
  const uint32_t bar_dkv = smem_u32(bars + 11);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  int* flags = reinterpret_cast<int*>(bars + 13);
  const uint32_t bar_bkt = smem_u32(bars + 14);

  if (tid == 0) {
    mbar_init(bar_kv, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(bar_ring_full + 8 * s, 1); mbar_init(bar_ring_empty + 8 * s, 1); }
    mbar_init(bar_s_full, 1);
    mbar_init(bar_s_free, AB_EPI / 32);
    mbar_init(bar_pds_full, AB_EPI / 32);
    mbar_init(bar_pds_free, 1);
    mbar_init(bar_dq_full, 1);
    mbar_init(bar_dq_free, AB_EPI / 32);
    mbar_init(bar_dkv, 1);
    mbar_init(bar_bkt, 1);
    fence_barrier_init();
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV);
    prefetch_tensormap(&tmdO);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  if (HAS_BIAS && warp == 2) {
    if (p.octaves) load_octave_table(reinterpret_cast<OctRec*>(smem + L::oct), flags, p.octaves, lane);
    else build_octave_table(reinterpret_cast<OctRec*>(smem + L::oct), flags, p.thr, p.nb, lane);
  }
  if (warp == 3) {
    float* tsw = reinterpret_cast<float*>(smem + L::tsw);
    for (int i = lane; i < 136; i += 32) tsw[i] = (HAS_BIAS && i <= p.nb) ? 0.5f * p.ts_w[i] : 0.f;
    float* hp = reinterpret_cast<float*>(smem + L::h_ts);
    for (int i = lane; i < 16 * 136; i += 32) hp[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_kv, 2 * AT_TILE_BYTES);
      tma_load_2d(smem_u32(smem + L::k), &tmK, h * AT_D, (int) (off0 + j0), bar_kv);
      tma_load_2d(smem_u32(smem + L::v), &tmV, h * AT_D, (int) (off0 + j0), bar_kv);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        mbar_wait_parked(bar_ring_empty + 8 * st, ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(bar_ring_full + 8 * st, 2 * AT_TILE_BYTES);
        const uint32_t dst = smem_u32(smem + L::ring + st * 2 * AT_TILE_BYTES);
        const int row = (int) (off0 + (kt + it) * AT_BM);
        tma_load_2d(dst, &tmQ, h * AT_D, row, bar_ring_full + 8 * st);
        tma_load_2d(dst + AT_TILE_BYTES, &tmdO, h * AT_D, row, bar_ring_full + 8 * st);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t id_kk = make_idesc_bf16(128, 128, false, false);   // S^T, dP^T
      const uint32_t id_kmn = make_idesc_bf16(128, AT_D, false, true);  // dV, dK
      const uint32_t id_mnmn = make_idesc_bf16(128, AT_D, true, true);  // dQ
      const uint32_t ka = smem_u32(smem + L::k), va = smem_u32(smem + L::v);
      const uint32_t pa = smem_u32(smem + L::pT), da = smem_u32(smem + L::dsT);
      auto issue_scores = [&](int it) {
        const int st = it & 1;
        mbar_wait_parked(bar_ring_full + 8 * st, (it >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(smem + L::ring + st * 2 * AT_TILE_BYTES);
        const uint32_t oa = qa + AT_TILE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss(tmem, make_smem_desc_sw128(ka + ks * 32, 0, 1024),
                  make_smem_desc_sw128(qa + ks * 32, 0, 1024), id_kk, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss(tmem + 128, make_smem_desc_sw128(va + ks * 32, 0, 1024),
                  make_smem_desc_sw128(oa + ks * 32, 0, 1024), id_kk, ks > 0);
        umma_commit(bar_s_full);
      };
      mbar_wait_parked(bar_kv, 0);
      issue_scores(0);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        mbar_wait_parked(bar_s_free, it & 1);             // epilogue has read S^T / dP^T(it)
        if (it + 1 < n_it) issue_scores(it + 1);
        mbar_wait_parked(bar_pds_full, it & 1);           // P^T / dS^T(it) are in shared memory
        if (it > 0) mbar_wait_parked(bar_dq_free, (it - 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(smem + L::ring + st * 2 * AT_TILE_BYTES);
        const uint32_t oa = qa + AT_TILE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // dV += P^T dO_i : A K-major (2 blocks), B = dO MN-major
          umma_ss(tmem + 256, make_smem_desc_sw128(pa + (ks >> 2) * AT_TILE_BYTES + (ks & 3) * 32, 0, 1024),
                  make_smem_desc_sw128(oa + ks * 2048, 0, 1024), id_kmn, (it > 0) || (ks > 0));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // dK += dS^T Q_i
          umma_ss(tmem + 320, make_smem_desc_sw128(da + (ks >> 2) * AT_TILE_BYTES + (ks & 3) * 32, 0, 1024),
                  make_smem_desc_sw128(qa + ks * 2048, 0, 1024), id_kmn, (it > 0) || (ks > 0));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // dQ_i = dS K_j : A = dS^T read MN-major, B = K MN-major
          umma_ss(tmem + 384, make_smem_desc_sw128(da + ks * 2048, AT_TILE_BYTES, 1024),
                  make_smem_desc_sw128(ka + ks * 2048, 0, 1024), id_mnmn, ks > 0);
        umma_commit(bar_pds_free);
        umma_commit(bar_dq_full);
        umma_commit(bar_ring_empty + 8 * st);
      }
      umma_commit(bar_dkv);
    }
  } else if (warp >= 4) {
    // ================= epilogue warpgroups =================
    const int g = (warp - 4) >> 2;                 // warpgroup: query columns [32g, 32g+32)
    const int r = ((warp & 3) << 5) | lane;        // key row inside the tile = TMEM lane
    const int wq = warp - 4;
    const int et = tid - 128;                      // 0 .. AB_EPI-1
    const uint32_t lane_base = (uint32_t) ((warp & 3) * 32) << 16;
    const int jk = j0 + r;                         // key position in the sequence
    int64_t* tsq_all = reinterpret_cast<int64_t*>(smem + L::tsq);
    uint32_t* tsq32_all = reinterpret_cast<uint32_t*>(smem + L::tsq32);
    float* pos_all = reinterpret_cast<float*>(smem + L::pos);
    const float* tsw_s = reinterpret_cast<const float*>(smem + L::tsw);
    const OctRec* oct = reinterpret_cast<const OctRec*>(smem + L::oct);
    float* h_ts = reinterpret_cast<float*>(smem + L::h_ts) + wq * 136;   // this warp's d ts_w
    // this thread's 64-byte slice (4 x 16-byte chunks) of its 128-byte row in block g / 2
    uint8_t* pT = smem + L::pT + (g >> 1) * AT_TILE_BYTES + r * 128;
    uint8_t* dsT = smem + L::dsT + (g >> 1) * AT_TILE_BYTES + r * 128;
    const int chunk0 = (g & 1) * 4;
    uint8_t* ds_plain = smem + L::ds_plain;
    const float inv_n = 1.0f / (float) p.N;
    const float half_inv_n = 0.5f * inv_n;
    // this CTA's private copies of d pos_w / d ts_w (the caller sums the copies)
    const int64_t copy = (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) %
                         (unsigned) p.d_bias_copies;
    float* d_pos_mine = p.d_pos_w + copy * (2 * p.N - 1);
    float* d_ts_mine = p.d_ts_w + copy * (p.nb + 1);
    int64_t ts_k = 0;
    uint32_t tk32 = 0;
    bool slow = false, narrow = false;
    int64_t tmin = 0;
    if (HAS_BIAS) {
      ts_k = ext_ts_at(p.ts, b, p.N, (int64_t) jk);
      slow = flags[0] != 0;
      const int cnt = (int) (n64 + 1 < p.N ? n64 + 1 : p.N);
      const TsRange tr = scan_ts_range<AB_EPI>(p.ts + (int64_t) b * p.N, cnt, et,
                                               reinterpret_cast<int64_t*>(smem + L::red), 3);
      narrow = tr.narrow && !slow;
      tmin = tr.tmin;
      tk32 = (uint32_t) (ts_k - tmin);
    }
    auto read_back_dq = [&](int it) {   // dQ of iteration `it`: lane = query row, 16 columns per warpgroup
      mbar_wait(bar_dq_full, it & 1);
      tc_fence_after();
      uint32_t qv[16];
      tmem_ld16(tmem + lane_base + 384 + 16 * g, qv);
      tmem_ld_wait();
      const int qi = (kt + it) * AT_BM + r;
      if (qi < n) {
        float* dst = p.dq_accum + (off0 + qi) * (int64_t) (p.H * AT_D) + h * AT_D + 16 * g;
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4)
          red_add_v4(dst + 4 * v4, __uint_as_float(qv[4 * v4]) * half_inv_n,
                     __uint_as_float(qv[4 * v4 + 1]) * half_inv_n,
                     __uint_as_float(qv[4 * v4 + 2]) * half_inv_n,
                     __uint_as_float(qv[4 * v4 + 3]) * half_inv_n);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dq_free);
    };
    // query-side tables of tile `it` (ext_ts[i0 + c + 1] and the pos_w window) into buffer it & 1
    auto stage_tables = [&](int it) {
      const int i0 = (kt + it) * AT_BM;
      const int pb = it & 1;
      if (g == 0) {
        const int64_t tq = ext_ts_at(p.ts, b, p.N, (int64_t) i0 + r + 1);
        tsq_all[pb * 128 + r] = tq;
        tsq32_all[pb * 128 + r] = (uint32_t) (tq - tmin);
      } else if (g == 1) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int x = r + 128 * t;               // pos[x] = 0.5 * pos_w[N-1 + j0 - i0 - 127 + x]
          const int64_t idx = p.N - 1 + j0 - i0 - 127 + x;
          pos_all[pb * 256 + x] = (idx >= 0 && idx < 2 * p.N - 1) ? 0.5f * p.pos_w[idx] : 0.f;
        }
      }
    };
    // d ts_w: each thread run-length accumulates along its row (buckets change rarely along a
    // row) and adds a finished run to its warp's private histogram.  Shared memory has no native
    // fp32 add (atomicAdd is a CAS loop), but a run ends only a few times per row and only lanes
    // of one warp can collide here, so the loop almost never retries.
    auto flush_run = [&](int bk, float val) {
      if (bk >= 0 && val != 0.f) atomicAdd(&h_ts[bk], val * half_inv_n);
    };

    // cached bucket tiles: slot (kt + it, kt) of this sequence, "K orientation" half
    const bool cached = HAS_BIAS && p.bcache != nullptr;
    const uint8_t* bkt_s = smem + L::bkt;
    auto load_bkt = [&](int it) {
      const int64_t tps = (int64_t) p.cache_nt * (p.cache_nt + 1) / 2;
      const int64_t iq = kt + it;
      const uint8_t* src = p.bcache + ((int64_t) b * tps + iq * (iq + 1) / 2 + kt) * 32768 + 16384;
      mbar_arrive_expect_tx(bar_bkt, 16384);
      bulk_load_1d(smem_u32(bkt_s), src, 16384, bar_bkt);
    };
    if (HAS_BIAS) {
      if (cached && et == 0) load_bkt(0);
      stage_tables(0);
      named_bar_sync(2, AB_EPI);
    }
    for (int it = 0; it < n_it; ++it) {
      const int i0 = (kt + it) * AT_BM;
      const int pb = it & 1;
      if (HAS_BIAS && it + 1 < n_it) stage_tables(it + 1);   // ordered by this tile's barrier 4
      const int64_t* tsq_s = tsq_all + pb * 128;
      const uint32_t* tsq32_s = tsq32_all + pb * 128;
      const float* pos_s = pos_all + pb * 256;
      const bool edge = (it == 0) || (i0 + AT_BM > n);   // diagonal tile or ragged last tile
      mbar_wait(bar_s_full, it & 1);
      tc_fence_after();
      if (HAS_BIAS && it > 0) named_bar_sync(5, AB_EPI);   // diagonal sums of tile it-1 are done
      if (cached) mbar_wait(bar_bkt, it & 1);
      int run_bk = -1;            // d ts_w: run-length accumulate along the row
      float run_acc = 0.f;
#pragma unroll 1
      for (int c16 = 0; c16 < 2; ++c16) {
        const int cb = 32 * g + 16 * c16;
        uint32_t sv[16], dv_[16];
        tmem_ld16(tmem + lane_base + cb, sv);
        tmem_ld16(tmem + lane_base + 128 + cb, dv_);
        tmem_ld_wait();
        if (c16 == 0 && it > 0) mbar_wait(bar_pds_free, (it - 1) & 1);   // smem tiles reusable
#pragma unroll
        for (int c8 = 0; c8 < 2; ++c8) {
          uint32_t ppk[4], dpk[4];
          int bk[8];
          float hb[8];
          const int c0 = cb + 8 * c8;              // first query column of this group
          if (HAS_BIAS) {
            if (cached) {
              // 8 bucket bytes of this key row: query chunk (c0 / 16), bytes (c0 % 16) .. +7
              const uint2 raw = *reinterpret_cast<const uint2*>(
                  bkt_s + ((size_t) (c0 >> 4) * 128 + r) * 16 + (c0 & 8));
              const uint32_t w2[2] = {raw.x, raw.y};
#pragma unroll
              for (int e = 0; e < 8; ++e) bk[e] = (int) ((w2[e >> 2] >> (8 * (e & 3))) & 0xffu);
            } else if (narrow) {
              const uint4 ta = *reinterpret_cast<const uint4*>(tsq32_s + c0);
              const uint4 tb = *reinterpret_cast<const uint4*>(tsq32_s + c0 + 4);
              const uint32_t tq[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) bk[e] = bucket_narrow(oct, __usad(tk32, tq[e], 0u));
            } else {
#pragma unroll 1
              for (int e = 0; e < 8; ++e) {
                int64_t d = tsq_s[c0 + e] - ts_k;
                d = d < 0 ? -d : d;
                bk[e] = bucket_wide(oct, p.thr, p.nb, slow, d);
              }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) hb[e] = pos_s[r - (c0 + e) + 127] + tsw_s[bk[e]];
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) { bk[e] = 0; hb[e] = 0.f; }
          }
          float dsv[8], pvv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int cc = 8 * c8 + e;             // column inside this 16-chunk
            const float hx = fmaf(__uint_as_float(sv[cc]), 0.5f, hb[e]);
            const float th = tanh_approx(hx);
            // unscaled: P' = SiLU(x) = N P ; dS' = dP * 2 SiLU'(x) = 2N dS.  The 1/N and 1/(2N)
            // factors are linear and applied once to dV, dK, dQ and the bias gradients.
            pvv[e] = fmaf(hx, th, hx);
            const float u1 = fmaf(-th, th, 1.0f);                        // 1 - tanh^2
            const float w2 = fmaf(hx, u1, 1.0f + th);                    // 2 * SiLU'(x)
            dsv[e] = __uint_as_float(dv_[cc]) * w2;
          }
          if (edge) {   // diagonal / ragged tiles only (warp-uniform): causal and length masks
            const int lo = jk - i0 - c0, hi = n - i0 - c0;   // valid columns: lo <= e < hi
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const bool ok = (e >= lo) && (e < hi);
              pvv[e] = ok ? pvv[e] : 0.f;
              dsv[e] = ok ? dsv[e] : 0.f;
            }
          }
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            ppk[e >> 1] = pack_bf16x2(pvv[e], pvv[e + 1]);
            dpk[e >> 1] = pack_bf16x2(dsv[e], dsv[e + 1]);
          }
          if (HAS_BIAS) {
            // d pos_w needs the sums of dS along the diagonals of the tile: stash a plain copy
            // (per-element global red costs ~41 LSU cycles per warp instruction; a shared
            // read-modify-write per element races between neighbouring lanes)
            *reinterpret_cast<uint4*>(ds_plain + r * L::DS_STRIDE + c0 * 2) =
                make_uint4(dpk[0], dpk[1], dpk[2], dpk[3]);
            // d ts_w: run-length accumulate along the row
            const bool same = (bk[0] == run_bk) & (bk[1] == run_bk) & (bk[2] == run_bk) &
                              (bk[3] == run_bk) & (bk[4] == run_bk) & (bk[5] == run_bk) &
                              (bk[6] == run_bk) & (bk[7] == run_bk);
            if (same) {
              run_acc += ((dsv[0] + dsv[1]) + (dsv[2] + dsv[3])) + ((dsv[4] + dsv[5]) + (dsv[6] + dsv[7]));
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                if (bk[e] != run_bk) {
                  flush_run(run_bk, run_acc);
                  run_bk = bk[e];
                  run_acc = 0.f;
                }
                run_acc += dsv[e];
              }
            }
          }
          // 16-byte chunk of this thread's row slice, 128-byte swizzle
          const int chunk = ((chunk0 + 2 * c16 + c8) ^ (r & 7)) * 16;
          *reinterpret_cast<uint4*>(pT + chunk) = make_uint4(ppk[0], ppk[1], ppk[2], ppk[3]);
          *reinterpret_cast<uint4*>(dsT + chunk) = make_uint4(dpk[0], dpk[1], dpk[2], dpk[3]);
        }
      }
      if (HAS_BIAS) flush_run(run_bk, run_acc);
      tc_fence_before();
      fence_proxy_async_smem();                    // st.shared -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar_s_free); mbar_arrive(bar_pds_full); }
      if (HAS_BIAS) {
        named_bar_sync(4, AB_EPI);   // dS^T copy complete; tables of tile it+1 staged
        if (cached && et == 0 && it + 1 < n_it) load_bkt(it + 1);   // everyone is done with bkt_s
        // d pos_w[N-1+j-i]: thread (x, half) sums diagonal x = r - c + 127 over 64 key rows
        const int x = et & 255, hf = et >> 8;
        if (x < 255) {
          int r0 = hf * 64, r1 = r0 + 64;
          const int lo = x - 127 > 0 ? x - 127 : 0;          // c = r - x + 127 >= 0
          const int hi = x + 1 < 128 ? x + 1 : 128;          // c <= 127
          r0 = r0 > lo ? r0 : lo;
          r1 = r1 < hi ? r1 : hi;
          float sum = 0.f;
          const uint8_t* ptr = ds_plain + r0 * L::DS_STRIDE + (r0 - x + 127) * 2;
          for (int rr = r0; rr < r1; ++rr, ptr += L::DS_STRIDE + 2)
            sum += __uint_as_float((uint32_t) (*reinterpret_cast<const uint16_t*>(ptr)) << 16);
          const int64_t idx = p.N - 1 + j0 - i0 - 127 + x;
          if (sum != 0.f && idx >= 0 && idx < 2 * p.N - 1)
            atomicAdd(d_pos_mine + idx, sum * half_inv_n);
        }
      }
      if (it > 0) read_back_dq(it - 1);
    }
    read_back_dq(n_it - 1);

    // ---- dV / dK: thread = key row; warpgroups 0,1 store dV halves, 2,3 store dK halves ----
    mbar_wait(bar_dkv, 0);
    tc_fence_after();
    {
      __nv_bfloat16* dst = (g < 2 ? p.dv + (off0 + jk) * p.lddv : p.dk + (off0 + jk) * p.lddk) +
                           h * AT_D + 32 * (g & 1);
      uint32_t ov[32];
      tmem_ld32(tmem + lane_base + 256 + 32 * g, ov);
      tmem_ld_wait();
      const float sc = g < 2 ? inv_n : half_inv_n;   // dV = P'^T dO / N ; dK = dS'^T Q / (2N)
      if (jk < n) {
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(ov[v4 * 8 + 0]) * sc, __uint_as_float(ov[v4 * 8 + 1]) * sc);
          o.y = pack_bf16x2(__uint_as_float(ov[v4 * 8 + 2]) * sc, __uint_as_float(ov[v4 * 8 + 3]) * sc);
          o.z = pack_bf16x2(__uint_as_float(ov[v4 * 8 + 4]) * sc, __uint_as_float(ov[v4 * 8 + 5]) * sc);
          o.w = pack_bf16x2(__uint_as_float(ov[v4 * 8 + 6]) * sc, __uint_as_float(ov[v4 * 8 + 7]) * sc);
          *reinterpret_cast<uint4*>(dst + v4 * 8) = o;
        }
      }
    }
    if (HAS_BIAS) {
      named_bar_sync(4, AB_EPI);                    // every warp's histogram is final
      if (et <= p.nb) {
        const float* hall = reinterpret_cast<const float*>(smem + L::h_ts);
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < AB_EPI / 32; ++w) v += hall[w * 136 + et];
        if (v != 0.f) atomicAdd(d_ts_mine + et, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// dq_accum (T, H*64) fp32 -> dq (T, lddq) bf16
__global__ void dq_to_bf16_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq,
                                  int64_t rows, int W, int64_t lddq) {
  const int64_t idx = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (idx >= rows * W) return;
  const int64_t r = idx / W;
  const int c = (int) (idx - r * W);
  const float4 v = *reinterpret_cast<const float4*>(acc + idx);
  uint2 o;
  o.x = pack_bf16x2(v.x, v.y);
  o.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(dq + r * lddq + c) = o;
}

bool hstu_attn_bwd_sm100_supported(const grb_hstu_attn_args* a) {
  if (a->dtype != GRB_BF16 || a->dqk != AT_D || a->dv != AT_D) return false;
  if (a->timestamps && a->num_buckets > 128) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(a->q) || !al16(a->k) || !al16(a->v) || !al16(a->dout) || !al16(a->dq) ||
      !al16(a->dk) || !al16(a->dv_grad) || !al16(a->dq_accum))
    return false;
  if ((a->ldq * 2) % 16 || (a->ldk * 2) % 16 || (a->ldv * 2) % 16 || (a->lddo * 2) % 16 ||
      (a->lddq * 2) % 16 || (a->lddk * 2) % 16 || (a->lddv * 2) % 16)
    return false;
  if (a->T >= (1ll << 31) || a->T == 0) return false;
  return true;
}

int hstu_attn_bwd_sm100(const grb_hstu_attn_args* a, cudaStream_t st) {
  if (a->B == 0 || a->max_len == 0) return GRB_OK;
  CUtensorMap tmQ, tmK, tmV, tmdO;
  int rc;
  const uint64_t W = (uint64_t) a->H * AT_D;
  if ((rc = make_tmap_bf16_2d(&tmQ, a->q, a->T, W, a->ldq, AT_BM)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, a->k, a->T, W, a->ldk, AT_BN)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, a->v, a->T, W, a->ldv, AT_BN)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmdO, a->dout, a->T, W, a->lddo, AT_BM)) != GRB_OK) return rc;
  AttnBwdParams p{};
  p.N = a->N; p.T = a->T; p.H = a->H; p.nb = a->num_buckets; p.index_bits = a->index_bits;
  p.n_kt = (int) ceil_div(a->max_len, AT_BN);
  p.offsets = a->offsets; p.ts = a->timestamps; p.ts_w = a->ts_w; p.pos_w = a->pos_w;
  p.thr = a->bucket_thresholds; p.octaves = a->bucket_octaves;
  p.bcache = a->timestamps ? a->bucket_cache : nullptr;
  p.cache_nt = (int) ceil_div(a->bucket_cache_max_len, AT_BM);
  p.dk = reinterpret_cast<__nv_bfloat16*>(a->dk); p.lddk = a->lddk;
  p.dv = reinterpret_cast<__nv_bfloat16*>(a->dv_grad); p.lddv = a->lddv;
  p.dq_accum = a->dq_accum; p.d_ts_w = a->d_ts_w; p.d_pos_w = a->d_pos_w;
  p.d_bias_copies = a->d_bias_copies > 0 ? a->d_bias_copies : 1;
  const size_t smem = AbSmem::total + 1024;
  dim3 grid((unsigned) p.n_kt, (unsigned) a->H, (unsigned) a->B);
  if (a->timestamps) {
    auto kern = hstu_attn_bwd_sm100_kernel<true>;
    GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    kern<<<grid, AB_THREADS, smem, st>>>(tmQ, tmK, tmV, tmdO, p);
  } else {
    auto kern = hstu_attn_bwd_sm100_kernel<false>;
    GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    kern<<<grid, AB_THREADS, smem, st>>>(tmQ, tmK, tmV, tmdO, p);
  }
  GRB_LAUNCH_OK();
  const int Wi = a->H * AT_D;
  const int64_t total4 = a->T * Wi / 4;
  dq_to_bf16_kernel<<<(unsigned) ceil_div(total4, 256), 256, 0, st>>>(
      a->dq_accum, reinterpret_cast<__nv_bfloat16*>(a->dq), a->T, Wi, a->lddq);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}  // namespace grb
