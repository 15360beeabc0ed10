// hstu_attn_bwd_sm100.cu — HSTU jagged attention backward on Blackwell tensor cores
// (bf16 in, fp32 accumulate in TMEM, dqk = dv = 64).
//
// Math (backward of hstu.py:186-204 with the bias of :96-128), per sequence / head, j <= i < n:
//   S = Q K^T + bias ; P = SiLU(S)/N ; dP = dO V^T ; dS = dP * SiLU'(S)/N
//   dV = P^T dO ; dK = dS^T Q ; dQ = dS K ; d pos_w[N-1+j-i] += dS ; d ts_w[bucket] += dS
//
// One CTA owns (sequence b, head h, key tile j) and walks the query tiles i = j .. last.  Scores
// are produced TRANSPOSED (thread = key row): P^T goes back to TMEM as the A operand of dV, dS^T
// lands in shared memory as the K-major A operand of dK and, read MN-major, the A operand of dQ.
//
//   warp 0 (TMA)  : K_j, V_j once; Q_i, dO_i through a 3-stage ring (128-byte swizzle, jagged
//                   rows); the cached 128x128 bucket-index tile of (i, j), double buffered.
//   warp 1 (MMA)  : S^T  = K_j Q_i^T   M128 2xN64 K64  (K-major x K-major)        -> TMEM [0,128)
//                   dP^T = V_j dO_i^T  M128 2xN64 K64                             -> TMEM [128,256)
//                   dV  += P^T  dO_i   M128 N64  K128  (TMEM A x MN-major)        -> TMEM [256,320)
//                   dK  += dS^T Q_i    M128 N64  K128  (smem K-major x MN-major)  -> TMEM [320,384)
//                   dQ_i = dS   K_j    M128 N64  K128  (MN-major x MN-major)      -> TMEM [384,448)
//                   (P^T as packed bf16 lives in TMEM [448,512))
//   warp 2        : stages the pos_w window of each query tile (double buffered, mbarriers)
//   epilogue      : 4 warpgroups, thread = key row: bias, tanh-based SiLU / SiLU', masks, bf16
//                   P^T / dS^T, bias-gradient partial sums; then dQ_i is read back and added to
//                   the fp32 dq accumulator with vector red.global.  No CTA-wide barrier inside
//                   the tile loop when the bucket cache is used: every hand-off is an mbarrier.
//
// Half-tile software pipeline: a query tile is processed as two halves of 64 query columns
// (warpgroup g owns columns 64*hf + 16g .. +16 of half hf).  Every buffer is naturally split the
// same way (S^T / dP^T / P^T column ranges in TMEM, the two 64-column blocks of dS^T in shared
// memory), so while the epilogue works on half B the tensor core already runs dV/dK of half A and
// the scores of the next tile's half A: neither side waits for a whole tile of the other.
//
// Bias gradients without shared-memory atomics:
//   d ts_w : each thread run-length accumulates along its row (the bucket changes rarely along a
//            row); a finished run is parked and reduced across the warp once per tile.
//   d pos_w: the 32x32 block of dS^T a warp owns is summed along its diagonals with lane
//            rotations: lane L collects diagonal r - c = L (and L - 32), 16 shuffles of packed
//            bf16 pairs per tile, then two red.global per thread.
#include "hstu_attn_sm100.cuh"
#include <cstdlib>
#include <cstdio>

namespace grb {

using namespace ptx;

// Build with -DGRB_BWD_TIMELINE and run with GRB_BWD_DEBUG=8 to get a clock64 timeline of CTA
// (0,0,0) on stderr (benchmarks/probes/bwd_timeline.py).
#ifdef GRB_BWD_TIMELINE
#define TL_STAMP(cond, slot) do { if (TLOG && (cond)) p.tl[slot] = clock64(); } while (0)
#else
#define TL_STAMP(cond, slot) do { } while (0)
#endif

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
  int dbg;
  long long* tl;            // timeline buffer (GRB_BWD_DEBUG=8)
};

template <bool HAS_BIAS>
struct AbSmem {
  // (Q, dO) stages: the kernel without bias has a short epilogue, so the loads must run further
  // ahead; with bias the bucket tiles take the space instead
  static constexpr int RING = HAS_BIAS ? 2 : 3;
  static constexpr int k = 0;
  static constexpr int v = k + AT_TILE_BYTES;
  static constexpr int ring = v + AT_TILE_BYTES;                 // RING x (Q, dO)
  // dS^T as [128 k][64 q] blocks: block A (query columns 0..63) double buffered by tile parity,
  // then block B.  dQ of a tile is issued late and reads both, so the next tile's half A must
  // not land in the buffer dQ is still reading.
  static constexpr int dsT = ring + RING * 2 * AT_TILE_BYTES;
  // dQ staging for the bulk reduce-add: 16 warps x [32 rows][16 floats], 64-byte swizzle
  static constexpr int dqs = dsT + 3 * AT_TILE_BYTES;
  static constexpr int bkt = dqs + 16 * 2048;                    // 2 x 128 x 128 uint8 bucket tiles
  // query-side tables are double buffered by tile parity.  The timestamp tables are only used
  // when there is no bucket cache, so they live inside the (then unused) bucket-tile space.
  static constexpr int tsq = bkt;                                // 2 x 128 x int64
  static constexpr int tsq32 = tsq + 2 * 128 * 8;                // 2 x 128 x uint32
  static constexpr int red = bkt + (HAS_BIAS ? 2 * 128 * 128 : 0);   // 2 x 16 x int64
  // pos_w window (pre-halved): 4 copies shifted by 0..3 floats so that every thread can fetch the
  // 4 consecutive values it needs with one aligned 16-byte load; copy stride = 8 banks
  static constexpr int POS_COPY = 264;                           // floats per copy
  static constexpr int pos = red + 32 * 8;                       // 2 x 4 x POS_COPY x float
  static constexpr int tsw = pos + 2 * 4 * POS_COPY * 4;         // 136 x float (pre-halved)
  static constexpr int oct = tsw + 136 * 4;                      // 32 x OctRec
  static constexpr int bars = oct + 32 * 16;
  static constexpr int total = bars + 224;   // 27 barrier / scratch words
};
static_assert(AbSmem<true>::total + 1024 <= 232448 && AbSmem<false>::total + 1024 <= 232448,
              "shared memory budget");

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}
// acc_lo += low bf16 half of `pair`, acc_hi += high half (fp32 accumulate, one FHADD each)
__device__ __forceinline__ void add_bf16_pair(float& acc_lo, float& acc_hi, uint32_t pair) {
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.bf16 %0, lo, %0;\n\t"
      "add.rn.f32.bf16 %1, hi, %1;\n\t}"
      : "+f"(acc_lo), "+f"(acc_hi)
      : "r"(pair));
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}

template <bool HAS_BIAS>
__global__ void __launch_bounds__(AB_THREADS, 1) hstu_attn_bwd_sm100_kernel(
    const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
    const __grid_constant__ CUtensorMap tmDQ, AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  using L = AbSmem<HAS_BIAS>;
  constexpr int AB_RING = L::RING;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kt = (int) blockIdx.z;                 // key tile, slowest grid index: the tiles with the most work start first
#ifdef GRB_BWD_TIMELINE
  const bool TLOG = (p.dbg & 8) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#endif
  const int h = blockIdx.x;
  const int b = blockIdx.y;
  const int64_t off0 = load_index(p.offsets, b, p.index_bits);
  int64_t n64 = load_index(p.offsets, b + 1, p.index_bits) - off0;
  if (n64 > p.N) n64 = p.N;
  const int n = (int) n64;
  const int j0 = kt * AT_BN;
  if (j0 >= n) return;
  const int n_qt = (n + AT_BM - 1) / AT_BM;
  const int n_it = n_qt - kt;                      // query tiles kt .. n_qt-1
  const bool cached = HAS_BIAS && p.bcache != nullptr;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  const uint32_t bar_kv = smem_u32(bars + 0);
  const uint32_t bar_ring_full = smem_u32(bars + 1);    // [3]
  const uint32_t bar_ring_empty = smem_u32(bars + 4);   // [3]
  const uint32_t bar_s_full = smem_u32(bars + 7);       // [2] scores of half hf are in TMEM
  const uint32_t bar_half_done = smem_u32(bars + 24);   // [2] half hf: S^T/dP^T read, P^T/dS^T written
  const uint32_t bar_pds_free = smem_u32(bars + 10);    // MMAs reading P^T / dS^T of a tile are done
  const uint32_t bar_a_free = smem_u32(bars + 26);      // dV / dK of half A (reads P^T half A) are done
  const uint32_t bar_dq_full = smem_u32(bars + 11);
  const uint32_t bar_dq_free = smem_u32(bars + 12);
  const uint32_t bar_dkv = smem_u32(bars + 13);
  const uint32_t bar_bkt_full = smem_u32(bars + 14);    // [2]
  const uint32_t bar_bkt_free = smem_u32(bars + 16);    // [2]
  const uint32_t bar_tab_full = smem_u32(bars + 18);    // [2]
  const uint32_t bar_tab_free = smem_u32(bars + 20);    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  int* flags = reinterpret_cast<int*>(bars + 23);

  if (tid == 0) {
    mbar_init(bar_kv, 1);
    for (int s = 0; s < AB_RING; ++s) { mbar_init(bar_ring_full + 8 * s, 1); mbar_init(bar_ring_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_s_full + 8 * s, 1); mbar_init(bar_half_done + 8 * s, AB_EPI / 32); }
    mbar_init(bar_pds_free, 1);
    mbar_init(bar_a_free, 1);
    mbar_init(bar_dq_full, 1);
    mbar_init(bar_dq_free, AB_EPI / 32);
    mbar_init(bar_dkv, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_bkt_full + 8 * s, 1);
      mbar_init(bar_bkt_free + 8 * s, AB_EPI / 32);
      mbar_init(bar_tab_full + 8 * s, 1);
      mbar_init(bar_tab_free + 8 * s, AB_EPI / 32);
    }
    fence_barrier_init();
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV);
    prefetch_tensormap(&tmdO); prefetch_tensormap(&tmDQ);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  if (HAS_BIAS && warp == 2) {
    if (p.octaves) load_octave_table(reinterpret_cast<OctRec*>(smem + L::oct), flags, p.octaves, lane);
    else build_octave_table(reinterpret_cast<OctRec*>(smem + L::oct), flags, p.thr, p.nb, lane);
  }
  if (warp == 3) {
    float* tsw = reinterpret_cast<float*>(smem + L::tsw);
    for (int i = lane; i < 136; i += 32) tsw[i] = (HAS_BIAS && i <= p.nb) ? 0.5f * p.ts_w[i] : 0.f;
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
      const int64_t tps = (int64_t) p.cache_nt * (p.cache_nt + 1) / 2;
      for (int it = 0; it < n_it; ++it) {
        if (cached) {   // bucket tile of (query tile kt + it, key tile kt), "K orientation" half
          const int sb = it & 1;
          mbar_wait_parked(bar_bkt_free + 8 * sb, ((it >> 1) & 1) ^ 1);
          const int64_t iq = kt + it;
          const uint8_t* src = p.bcache + ((int64_t) b * tps + iq * (iq + 1) / 2 + kt) * 32768 + 16384;
          mbar_arrive_expect_tx(bar_bkt_full + 8 * sb, 16384);
          bulk_load_1d(smem_u32(smem + L::bkt + sb * 16384), src, 16384, bar_bkt_full + 8 * sb);
        }
        const int st = it % AB_RING;
        mbar_wait_parked(bar_ring_empty + 8 * st, ((it / AB_RING) & 1) ^ 1);
        mbar_arrive_expect_tx(bar_ring_full + 8 * st, 2 * AT_TILE_BYTES);
        const uint32_t dst = smem_u32(smem + L::ring + st * 2 * AT_TILE_BYTES);
        const int row = (int) (off0 + (kt + it) * AT_BM);
        tma_load_2d(dst, &tmQ, h * AT_D, row, bar_ring_full + 8 * st);
        tma_load_2d(dst + AT_TILE_BYTES, &tmdO, h * AT_D, row, bar_ring_full + 8 * st);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    {   // whole warp, uniform control flow; one elected lane issues (umma_*_warp)
      const uint32_t id_kk = make_idesc_bf16(128, 64, false, false);    // S^T, dP^T halves
      const uint32_t id_kmn = make_idesc_bf16(128, AT_D, false, true);  // dV, dK
      const uint32_t id_mnmn = make_idesc_bf16(128, AT_D, true, true);  // dQ
      // descriptors are built once; stepping an operand is an add on the address field
      const uint64_t k_desc = make_smem_desc_sw128(smem_u32(smem + L::k), 0, 1024);
      const uint64_t v_desc = make_smem_desc_sw128(smem_u32(smem + L::v), 0, 1024);
      const uint64_t ds_k_desc = make_smem_desc_sw128(smem_u32(smem + L::dsT), 0, 1024);
      // dQ reads block A [parity] and block B as one MN-major operand: LBO = distance between them
      const uint64_t ds_mn_desc0 = make_smem_desc_sw128(smem_u32(smem + L::dsT), 2 * AT_TILE_BYTES, 1024);
      const uint64_t ds_mn_desc1 =
          make_smem_desc_sw128(smem_u32(smem + L::dsT) + AT_TILE_BYTES, AT_TILE_BYTES, 1024);
      const uint64_t ring_desc = make_smem_desc_sw128(smem_u32(smem + L::ring), 0, 1024);
      auto adv = [](uint64_t d, uint32_t bytes) { return d + (uint64_t) (bytes >> 4); };
      // scores of query columns 64*hf .. +64 of tile it (ring stage already resident)
      auto issue_scores = [&](int it, int hf) {
        const uint64_t q_desc = adv(ring_desc, (it % AB_RING) * 2 * AT_TILE_BYTES + hf * 8192);
        const uint64_t o_desc = adv(q_desc, AT_TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss_warp(tmem + 64 * hf, adv(k_desc, ks * 32), adv(q_desc, ks * 32), id_kk, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss_warp(tmem + 128 + 64 * hf, adv(v_desc, ks * 32), adv(o_desc, ks * 32), id_kk, ks > 0);
        umma_commit_warp(bar_s_full + 8 * hf);
      };
      // dV += P_hf^T dO_hf ; dK += dS_hf^T Q_hf  (K = the 64 query rows of half hf)
      auto issue_dvdk = [&](int it, int hf) {
        const uint64_t q_desc = adv(ring_desc, (it % AB_RING) * 2 * AT_TILE_BYTES);
        const uint64_t o_desc = adv(q_desc, AT_TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)   // A = packed bf16 P^T in TMEM, B = dO MN-major
          umma_ts_warp(tmem + 256, tmem + 448 + 32 * hf + ks * 8, adv(o_desc, (4 * hf + ks) * 2048), id_kmn,
                  (it > 0) || (hf > 0) || (ks > 0));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)   // A = dS^T block hf K-major, B = Q MN-major
          umma_ss_warp(tmem + 320, adv(ds_k_desc, (hf ? 2 : (it & 1)) * AT_TILE_BYTES + ks * 32),
                  adv(q_desc, (4 * hf + ks) * 2048), id_kmn, (it > 0) || (hf > 0) || (ks > 0));
      };
      mbar_wait_parked(bar_kv, 0);
      mbar_wait_parked(bar_ring_full, 0);
      tc_fence_after();
      issue_scores(0, 0);
      issue_scores(0, 1);
      for (int it = 0; it < n_it; ++it) {
        const int st = it % AB_RING;
        const bool more = it + 1 < n_it;
        mbar_wait_parked(bar_half_done, it & 1);          // half A of tile it
        TL_STAMP(lane == 0 && it < 16, it * 16 + 0);
        tc_fence_after();
        issue_dvdk(it, 0);
        umma_commit_warp(bar_a_free);
        if (more) {
          mbar_wait_parked(bar_ring_full + 8 * ((it + 1) % AB_RING), ((it + 1) / AB_RING) & 1);
          tc_fence_after();
          issue_scores(it + 1, 0);
        }
        TL_STAMP(lane == 0 && it < 16, it * 16 + 1);
        mbar_wait_parked(bar_half_done + 8, it & 1);      // half B of tile it
        TL_STAMP(lane == 0 && it < 16, it * 16 + 2);
        if (it > 0) mbar_wait_parked(bar_dq_free, (it - 1) & 1);
        TL_STAMP(lane == 0 && it < 16, it * 16 + 3);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // dQ_i = dS K_j : A = dS^T read MN-major, B = K MN-major
          umma_ss_warp(tmem + 384, adv((it & 1) ? ds_mn_desc1 : ds_mn_desc0, ks * 2048), adv(k_desc, ks * 2048),
                  id_mnmn, ks > 0);
        umma_commit_warp(bar_dq_full);             // first: the epilogue reads dQ back mid-tile
        issue_dvdk(it, 1);
        umma_commit_warp(bar_pds_free);
        umma_commit_warp(bar_ring_empty + 8 * st);
        if (more) issue_scores(it + 1, 1);
        TL_STAMP(lane == 0 && it < 16, it * 16 + 4);
      }
      umma_commit_warp(bar_dkv);
    }
  } else if (warp == 2) {
    // ================= pos_w window stager =================
    if (HAS_BIAS) {
      float* pos_all = reinterpret_cast<float*>(smem + L::pos);
      for (int it = 0; it < n_it; ++it) {
        const int pb = it & 1;
        const int i0 = (kt + it) * AT_BM;
        float vals[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) {              // pos[x] = 0.5 * pos_w[N-1 + j0 - i0 - 127 + x], x < 259
          const int x = lane + 32 * t;
          const int64_t idx = p.N - 1 + j0 - i0 - 127 + x;
          vals[t] = (x < 259 && idx >= 0 && idx < 2 * p.N - 1) ? 0.5f * p.pos_w[idx] : 0.f;
        }
        mbar_wait_parked(bar_tab_free + 8 * pb, ((it >> 1) & 1) ^ 1);
        float* dst = pos_all + pb * 4 * L::POS_COPY;
#pragma unroll
        for (int t = 0; t < 9; ++t) {              // copy s holds pos[y + s] at index y
          const int x = lane + 32 * t;
#pragma unroll
          for (int sft = 0; sft < 4; ++sft)
            if (x - sft >= 0 && x - sft < 256) dst[sft * L::POS_COPY + x - sft] = vals[t];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tab_full + 8 * pb);
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue warpgroups =================
    const int g = (warp - 4) >> 2;                 // warpgroup: query columns 64*hf + [16g, 16g+16)
    const int r = ((warp & 3) << 5) | lane;        // key row inside the tile = TMEM lane
    const int wq = warp - 4;
    const int et = tid - 128;                      // 0 .. AB_EPI-1
    const uint32_t lane_base = (uint32_t) ((warp & 3) * 32) << 16;
    const int jk = j0 + r;                         // key position in the sequence
    int64_t* tsq_all = reinterpret_cast<int64_t*>(smem + L::tsq);
    uint32_t* tsq32_all = reinterpret_cast<uint32_t*>(smem + L::tsq32);
    const float* pos_all = reinterpret_cast<const float*>(smem + L::pos);
    const float* tsw_s = reinterpret_cast<const float*>(smem + L::tsw);
    const OctRec* oct = reinterpret_cast<const OctRec*>(smem + L::oct);
    // this thread's 32-byte slice (2 x 16-byte chunks) of its 128-byte row, in block hf
    uint8_t* dsT = smem + L::dsT + r * 128;
    const int chunk0 = 2 * g;
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
    if (HAS_BIAS && !cached) {
      ts_k = ext_ts_at(p.ts, b, p.N, (int64_t) jk);
      slow = flags[0] != 0;
      const int cnt = (int) (n64 + 1 < p.N ? n64 + 1 : p.N);
      const TsRange tr = scan_ts_range<AB_EPI>(p.ts + (int64_t) b * p.N, cnt, et,
                                               reinterpret_cast<int64_t*>(smem + L::red), 3);
      narrow = tr.narrow && !slow;
      tmin = tr.tmin;
      tk32 = (uint32_t) (ts_k - tmin);
    }
    // dQ of iteration `it` (lane = query row, 16 columns per warpgroup) is added to the fp32
    // accumulator by a bulk reduce-add: the warp stages its [32 rows][16 floats] block in shared
    // memory (64-byte swizzle: 16-byte chunk k of row rr sits at chunk k ^ ((rr >> 1) & 3)) and
    // one lane hands it to the TMA unit, so no LSU time goes into 16-byte scattered atomics.
    // Rows past the end of the sequence are exactly zero (dS is masked), rows past T are clipped.
    uint8_t* dq_stage = smem + L::dqs + wq * 2048;
    auto read_back_dq = [&](int it) {
      mbar_wait(bar_dq_full, it & 1);
      tc_fence_after();
      uint32_t qv[16];
      tmem_ld16(tmem + lane_base + 384 + 16 * g, qv);
      tmem_ld_wait();
      if (lane == 0) bulk_wait_group_read0();      // the previous reduce has read the staging block
      __syncwarp();
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4)
        *reinterpret_cast<float4*>(dq_stage + lane * 64 + ((v4 ^ ((lane >> 1) & 3)) << 4)) =
            make_float4(__uint_as_float(qv[4 * v4]) * half_inv_n, __uint_as_float(qv[4 * v4 + 1]) * half_inv_n,
                        __uint_as_float(qv[4 * v4 + 2]) * half_inv_n, __uint_as_float(qv[4 * v4 + 3]) * half_inv_n);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_2d(&tmDQ, h * AT_D + 16 * g, (int) (off0 + (kt + it) * AT_BM + 32 * (warp & 3)),
                          smem_u32(dq_stage));
        bulk_commit_group();
        mbar_arrive(bar_dq_free);
      }
    };
    // Without the bucket cache the query-side timestamps of tile `it` (ext_ts[i0 + c + 1]) are
    // staged by warpgroup 0 into buffer it & 1: global loads a tile ahead (stage_fetch), stores at
    // the end of the tile (stage_commit), one named barrier per tile orders them against use.
    int64_t st_a = 0;
    auto stage_fetch = [&](int it) {
      if (g == 0) st_a = ext_ts_at(p.ts, b, p.N, (int64_t) (kt + it) * AT_BM + r + 1);
    };
    auto stage_commit = [&](int it) {
      if (g == 0) {
        tsq_all[(it & 1) * 128 + r] = st_a;
        tsq32_all[(it & 1) * 128 + r] = (uint32_t) (st_a - tmin);
      }
    };
    // d ts_w: a finished run is added to this CTA's private copy in global memory with a
    // fire-and-forget red (shared memory has no native fp32 add: atomicAdd there is a CAS loop)
    auto flush_run = [&](int bk, float val) {
      if (bk >= 0 && val != 0.f) atomicAdd(d_ts_mine + bk, val * half_inv_n);
    };

    if (HAS_BIAS && !cached) {
      stage_fetch(0);
      stage_commit(0);
      named_bar_sync(2, AB_EPI);
    }
    // d ts_w run state: carried across query tiles (the row continues to the right, where the
    // bucket is usually still the same), flushed once after the last tile
    // (run_bk4 = the run's bucket in all four bytes, 0xffffffff before the first element;
    //  run_tsw = 0.5 * ts_w[run bucket])
    uint32_t run_bk4 = 0xffffffffu;
    float run_acc = 0.f, run_tsw = 0.f;
    for (int it = 0; it < n_it; ++it) {
      const int i0 = (kt + it) * AT_BM;
      const int pb = it & 1;
      if (HAS_BIAS && !cached && it + 1 < n_it) stage_fetch(it + 1);
      const int64_t* tsq_s = tsq_all + pb * 128;
      const uint32_t* tsq32_s = tsq32_all + pb * 128;
      // this thread's copy (r & 3) of the pos_w window, positioned so that group c0 reads the two
      // aligned float4 at pos4 - c0 (elements e = 7..4) and pos4 - c0 + 4 (e = 3..0)
      const float* pos4 = pos_all + (pb * 4 + (r & 3)) * L::POS_COPY + (r + 120 - (r & 3));
      const uint8_t* bkt_s = smem + L::bkt + pb * 16384;
      const bool edge = (it == 0) || (i0 + AT_BM > n);   // diagonal tile or ragged last tile
      TL_STAMP(et == 0 && it < 16, it * 16 + 8);
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        const int cb = 64 * hf + 16 * g;           // first query column of this thread's chunk
        mbar_wait(bar_s_full + 8 * hf, it & 1);
        tc_fence_after();
        if (hf == 0) {
          if (HAS_BIAS) mbar_wait(bar_tab_full + 8 * pb, (it >> 1) & 1);
          if (cached) mbar_wait(bar_bkt_full + 8 * pb, (it >> 1) & 1);
        }
        TL_STAMP(et == 0 && it < 16, it * 16 + 9 + 3 * hf);
        uint32_t sv[16], dv_[16];
        tmem_ld16(tmem + lane_base + cb, sv);
        tmem_ld16(tmem + lane_base + 128 + cb, dv_);
        tmem_ld_wait();
        if (it > 0) {   // the MMAs that read the buffers this half is about to overwrite
          if (hf == 0) mbar_wait(bar_a_free, (it - 1) & 1);     // P^T half A (dV of half A, tile it-1)
          else mbar_wait(bar_pds_free, (it - 1) & 1);           // P^T half B, dS^T block B; and block A
        }                                                       // of this parity two tiles ago
        TL_STAMP(et == 0 && it < 16, it * 16 + 10 + 3 * hf);
        TL_STAMP(lane == 0 && it == 10, 256 + wq * 8 + 4 * hf);
        // d pos_w partial sums of this warp's 32x16 block: diagonal r - c = lane (am, bm -> lane - 1)
        // and lane - 32 (aw, bw -> lane - 33); a* take the even column of a pair, b* the odd one
        float am = 0.f, aw = 0.f, bm = 0.f, bw = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < 2; ++c8) {
          uint32_t ppk[4], dpk[4];
          int bk[8];
          float hb[8];
          const int c0 = cb + 8 * c8;              // first query column of this group
          if (edge) {
            // diagonal / ragged tile: a group whose 8 columns are masked for all 32 rows of the
            // warp (above the diagonal, or past the end of the sequence) only needs its zeros
            const int lo = jk - i0 - c0, hi = n - i0 - c0;   // valid columns: lo <= e < hi
            if (__all_sync(0xffffffffu, lo >= 8 || hi <= 0 || hi <= lo)) {
              const uint32_t z4[4] = {0u, 0u, 0u, 0u};
              tmem_st4(tmem + lane_base + 448 + (c0 >> 1), z4);
              *reinterpret_cast<uint4*>(dsT + (hf ? 2 : pb) * AT_TILE_BYTES + ((chunk0 + c8) ^ (r & 7)) * 16) =
                  make_uint4(0u, 0u, 0u, 0u);
              continue;
            }
          }
          bool uni = false;   // (warp-uniform) every lane's 8 buckets continue its current run
          if (HAS_BIAS) {
            uint2 raw = make_uint2(0u, 0u);
            if (cached) {
              // 8 bucket bytes of this key row: query chunk (c0 / 16), bytes (c0 % 16) .. +7
              raw = *reinterpret_cast<const uint2*>(bkt_s + ((size_t) (c0 >> 4) * 128 + r) * 16 + 8 * c8);
              uni = __all_sync(0xffffffffu, (raw.x == run_bk4) & (raw.y == run_bk4));
            }
            // x = r - (c0 + e) + 127 ; the thread's copy makes x - 3 (e = 0) and x - 7 (e = 4) aligned
            const float4 pa = *reinterpret_cast<const float4*>(pos4 - c0 + 4);   // e = 3, 2, 1, 0
            const float4 pc = *reinterpret_cast<const float4*>(pos4 - c0);       // e = 7, 6, 5, 4
            const float pz[8] = {pa.w, pa.z, pa.y, pa.x, pc.w, pc.z, pc.y, pc.x};
            if (uni) {
#pragma unroll
              for (int e = 0; e < 8; ++e) { bk[e] = 0; hb[e] = pz[e] + run_tsw; }
            } else {
              if (cached) {
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
#pragma unroll
                for (int e = 0; e < 8; ++e) {   // (unrolled: a runtime index would push bk[] to local memory)
                  int64_t d = tsq_s[c0 + e] - ts_k;
                  d = d < 0 ? -d : d;
                  bk[e] = bucket_wide(oct, p.thr, p.nb, slow, d);
                }
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) hb[e] = pz[e] + tsw_s[bk[e]];
            }
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
          // P^T: packed bf16 pairs, query columns c0 .. c0+7 -> TMEM columns 448 + c0/2 .. +3
          tmem_st4(tmem + lane_base + 448 + (c0 >> 1), ppk);
          // dS^T: 16-byte chunk of this thread's row slice, 128-byte swizzle, block hf
          *reinterpret_cast<uint4*>(dsT + (hf ? 2 : pb) * AT_TILE_BYTES + ((chunk0 + c8) ^ (r & 7)) * 16) =
              make_uint4(dpk[0], dpk[1], dpk[2], dpk[3]);
          if (HAS_BIAS) {
            // d pos_w: rotate each packed pair to the lane that owns its diagonal; the bf16 halves
            // are added to fp32 accumulators directly (add.f32.bf16)
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
              const int src = lane + 8 * c8 + 2 * k2;   // source lane (mod 32) of column pair k2
              const uint32_t got = __shfl_sync(0xffffffffu, dpk[k2], src);
              if (src < 32) add_bf16_pair(am, bm, got); else add_bf16_pair(aw, bw, got);
            }
            // d ts_w: run-length accumulate along the row
            const float sum8 = ((dsv[0] + dsv[1]) + (dsv[2] + dsv[3])) + ((dsv[4] + dsv[5]) + (dsv[6] + dsv[7]));
            if (uni) {
              run_acc += sum8;
            } else {
              int cur = (run_bk4 == 0xffffffffu) ? -1 : (int) (run_bk4 & 0xffu);
              const bool same = (bk[0] == cur) & (bk[1] == cur) & (bk[2] == cur) & (bk[3] == cur) &
                                (bk[4] == cur) & (bk[5] == cur) & (bk[6] == cur) & (bk[7] == cur);
              if (same) {
                run_acc += sum8;
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const bool chg = bk[e] != cur;
                  if (chg) flush_run(cur, run_acc);        // one predicated red, no loop
                  run_acc = (chg ? 0.f : run_acc) + dsv[e];
                  cur = bk[e];
                }
                run_bk4 = (uint32_t) cur * 0x01010101u;
                run_tsw = tsw_s[cur];
              }
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        fence_proxy_async_smem();                  // st.shared -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_half_done + 8 * hf);
          if (hf == 1) {
            if (HAS_BIAS) mbar_arrive(bar_tab_free + 8 * pb);
            if (cached) mbar_arrive(bar_bkt_free + 8 * pb);
          }
        }
        TL_STAMP(et == 0 && it < 16, it * 16 + 11 + 3 * hf);
        TL_STAMP(lane == 0 && it == 10, 256 + wq * 8 + 4 * hf + 1);
        if (HAS_BIAS) {
          // diagonal totals: b* of lane L+1 belong to the diagonals of lane L (lane 0's bm is
          // diagonal -1, the wrap diagonal of lane 31)
          const float tbm = __shfl_sync(0xffffffffu, bm, lane + 1);
          const float tbw = __shfl_sync(0xffffffffu, bw, lane + 1);
          const float tot_m = am + (lane < 31 ? tbm : 0.f);
          const float tot_w = aw + tbw + (lane == 31 ? tbm : 0.f);
          // diagonal r - c = rel  ->  pos_w index N-1 + (j0 + r) - (i0 + c)
          const int64_t idx_m = p.N - 1 + j0 - i0 + 32 * (warp & 3) - cb + lane;
          if (tot_m != 0.f && idx_m >= 0 && idx_m < 2 * p.N - 1)
            atomicAdd(d_pos_mine + idx_m, tot_m * half_inv_n);
          const int64_t idx_w = idx_m - 32;
          if (tot_w != 0.f && idx_w >= 0 && idx_w < 2 * p.N - 1)
            atomicAdd(d_pos_mine + idx_w, tot_w * half_inv_n);
        }
        TL_STAMP(lane == 0 && it == 10, 256 + wq * 8 + 4 * hf + 2);
        if (hf == 0 && it > 0) read_back_dq(it - 1);
        TL_STAMP(lane == 0 && it == 10, 256 + wq * 8 + 4 * hf + 3);
      }
      if (HAS_BIAS && !cached) {
        if (it + 1 < n_it) stage_commit(it + 1);   // other buffer: nobody reads it during tile it
        named_bar_sync(4, AB_EPI);
      }
      TL_STAMP(et == 0 && it < 16, it * 16 + 15);
    }
    read_back_dq(n_it - 1);
    if (lane == 0) bulk_wait_group0();           // shared memory must outlive the reduce
    if (HAS_BIAS) {
      // final runs of the 32 rows of this warp: mostly one bucket, so reduce per distinct bucket
      // across the warp and let one lane issue the red
      const int bk = run_bk4 == 0xffffffffu ? -1 : (int) (run_bk4 & 0xffu);
      unsigned todo = __ballot_sync(0xffffffffu, bk >= 0);
      while (todo) {
        const int bsel = __shfl_sync(0xffffffffu, bk, __ffs(todo) - 1);
        const bool mine = bk == bsel;
        const float v = warp_sum(mine ? run_acc : 0.f);
        if (lane == 0) flush_run(bsel, v);
        todo &= ~__ballot_sync(0xffffffffu, mine);
      }
    }

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
  if (a->B > 65535) return false;   // grid.y
  return true;
}

int hstu_attn_bwd_sm100(const grb_hstu_attn_args* a, cudaStream_t st) {
  if (a->B == 0 || a->max_len == 0) return GRB_OK;
  CUtensorMap tmQ, tmK, tmV, tmdO, tmDQ;
  int rc;
  const uint64_t W = (uint64_t) a->H * AT_D;
  if ((rc = make_tmap_bf16_2d(&tmQ, a->q, a->T, W, a->ldq, AT_BM)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, a->k, a->T, W, a->ldk, AT_BN)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, a->v, a->T, W, a->ldv, AT_BN)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmdO, a->dout, a->T, W, a->lddo, AT_BM)) != GRB_OK) return rc;
  if ((rc = make_tmap_f32_2d_sw64(&tmDQ, a->dq_accum, a->T, W, W, 32)) != GRB_OK) return rc;
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
#ifdef GRB_BWD_TIMELINE
  { const char* e = std::getenv("GRB_BWD_DEBUG"); p.dbg = e ? atoi(e) : 0; }
  if (p.dbg & 8) { static long long* tlbuf = nullptr; if (!tlbuf) cudaMalloc(&tlbuf, 512 * 8); p.tl = tlbuf; }
#endif
  dim3 grid((unsigned) a->H, (unsigned) a->B, (unsigned) p.n_kt);
  if (a->timestamps) {
    auto kern = hstu_attn_bwd_sm100_kernel<true>;
    const size_t smem = AbSmem<true>::total + 1024;
    GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    kern<<<grid, AB_THREADS, smem, st>>>(tmQ, tmK, tmV, tmdO, tmDQ, p);
  } else {
    auto kern = hstu_attn_bwd_sm100_kernel<false>;
    const size_t smem = AbSmem<false>::total + 1024;
    GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    kern<<<grid, AB_THREADS, smem, st>>>(tmQ, tmK, tmV, tmdO, tmDQ, p);
  }
  GRB_LAUNCH_OK();
#ifdef GRB_BWD_TIMELINE
  if (p.dbg & 8) {
    cudaStreamSynchronize(st);
    long long h[512];
    cudaMemcpy(h, p.tl, sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = h[8];
    for (int w = 0; w < 16; ++w)
      fprintf(stderr, "warp %2d (wg %d, smsp %d) tile 10: A go %6lld arrive %6lld flushed %6lld dq %6lld | B go %6lld arrive %6lld flushed %6lld\n", w, w / 4, w % 4,
              h[256+w*8]-t0, h[256+w*8+1]-t0, h[256+w*8+2]-t0, h[256+w*8+3]-t0, h[256+w*8+4]-t0, h[256+w*8+5]-t0, h[256+w*8+6]-t0);
    for (int it = 0; it < 12; ++it) {
      fprintf(stderr, "it %2d MMA: A_done %6lld issuedA %6lld B_done %6lld dq_free %6lld issuedB %6lld | EPI: top %6lld A: s_full %6lld go %6lld arrive %6lld  B: s_full %6lld go %6lld arrive %6lld  end %6lld\n",
              it, h[it*16+0]-t0, h[it*16+1]-t0, h[it*16+2]-t0, h[it*16+3]-t0, h[it*16+4]-t0,
              h[it*16+8]-t0, h[it*16+9]-t0, h[it*16+10]-t0, h[it*16+11]-t0, h[it*16+12]-t0, h[it*16+13]-t0, h[it*16+14]-t0, h[it*16+15]-t0);
    }
  }
#endif
  const int Wi = a->H * AT_D;
  const int64_t total4 = a->T * Wi / 4;
  dq_to_bf16_kernel<<<(unsigned) ceil_div(total4, 256), 256, 0, st>>>(
      a->dq_accum, reinterpret_cast<__nv_bfloat16*>(a->dq), a->T, Wi, a->lddq);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}  // namespace grb
