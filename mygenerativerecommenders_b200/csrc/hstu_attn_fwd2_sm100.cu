// hstu_attn_fwd2_sm100.cu — HSTU jagged attention forward, second generation: the forward
// counterpart of hstu_attn_bwd_sm100.cu's hand-off scheme.  Used whenever the bias comes from the
// per-batch bucket cache (or there is no bias); hstu_attn_sm100.cu keeps the path that derives
// the buckets from raw timestamps inside the kernel.
//
// Reference math: /root/reference/src/generative_recommenders_pl/models/sequential_encoders/
// hstu.py:96-128 (bias) + :134-205 (attention).  One CTA owns (sequence b, 128 query rows, a PAIR
// of heads) and walks units u = (key tile j, head hh), j = 0..iq (causal):
//
//   warp 0 (TMA)  : Q_h tiles once; per unit a 128x64 K tile and a 128x64 V tile into a 4-stage
//                   ring (128-byte swizzle, jagged row coordinate off[b] + j0); per key tile the
//                   cached 128x128 bucket-index tile (16 KiB bulk copy, double buffered).
//   warp 1 (MMA)  : S_u   = Q_h K_h^T   M128 N128 K64   (K-major x K-major)   -> TMEM S[u & 1]
//                   O_h  += P_u V_h     M128 N64  K128  (A = P in TMEM, B = V MN-major)
//                   whole warp, uniform control flow, one elected lane issues.
//   warp 2        : stages the pos_w window of each key tile (4 shifted copies so that every
//                   thread fetches its 8 values with two aligned 16-byte loads), double buffered.
//   epilogue      : 4 warpgroups, thread = query row, warpgroup g owns key columns [32g, 32g+32)
//                   of every unit: bias (bucket bytes + pos window + ts_w; the ts_w lookup is
//                   skipped when the warp's 8x32 block has one bucket per row), h = S/2 + bias/2,
//                   P = h + h*tanh(h) = SiLU(S + bias) on f16x2 (one MUFU per element), causal
//                   mask on the diagonal tile, bf16 P into its OWN TMEM columns (no aliasing with
//                   S), so S[u & 1] is free for unit u + 2 as soon as it has been read.
//   Every hand-off is an mbarrier; there is no CTA-wide barrier in the unit loop.
//
// TMEM: S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O_0 [384,448) O_1 [448,512).
#include "hstu_attn_sm100.cuh"
#include <cstdlib>

namespace grb {

using namespace ptx;

namespace {

constexpr int F2_HG = 2;
constexpr int F2_STAGES = 4;
constexpr int F2_EPI = 512;
constexpr int F2_THREADS = 128 + F2_EPI;

struct Fwd2Params {
  int64_t N, T;
  int H, nb, index_bits, n_qt;
  const void* offsets;
  const float* ts_w;
  const float* pos_w;
  const uint8_t* bcache;
  int cache_nt;
  __nv_bfloat16* out;
  int64_t ldo;
};

struct F2Smem {
  static constexpr int POS_COPY = 264;                              // floats per shifted copy
  static constexpr int q = 0;                                       // HG x 16 KiB
  static constexpr int kv = q + F2_HG * AT_TILE_BYTES;              // STAGES x (K, V)
  static constexpr int bkt = kv + F2_STAGES * 2 * AT_TILE_BYTES;    // 2 x 16 KiB bucket tiles
  static constexpr int pos = bkt + 2 * 16384;                       // 2 x 4 x POS_COPY floats
  static constexpr int tsw = pos + 2 * 4 * POS_COPY * 4;            // 136 floats (pre-halved)
  static constexpr int bars = tsw + 136 * 4;
  static constexpr int total = bars + 256;
};
static_assert(F2Smem::total + 1024 <= 232448, "shared memory budget");

// BF16M: the SiLU chain runs on packed bf16 pairs (the format P has to take anyway) instead of f16
// pairs converted to bf16 at the end: 3 instructions per element instead of 4.5.
template <bool HAS_BIAS, bool BF16M>
__global__ void __launch_bounds__(F2_THREADS, 1) hstu_attn_fwd2_kernel(
    const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
    const __grid_constant__ CUtensorMap tmV, Fwd2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  using L = F2Smem;
  constexpr int HG = F2_HG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_hg = p.H / HG;
  const int qt = p.n_qt - 1 - (int) (blockIdx.x / n_hg);   // heavy tiles first
  const int h0 = (int) (blockIdx.x % n_hg) * HG;
  const int b = blockIdx.y;
  const int64_t off0 = load_index(p.offsets, b, p.index_bits);
  int64_t n64 = load_index(p.offsets, b + 1, p.index_bits) - off0;
  if (n64 > p.N) n64 = p.N;
  const int n = (int) n64;
  const int i0 = qt * AT_BM;
  if (i0 >= n) return;
  const int n_kt = qt + 1;              // causal: key tiles 0..qt
  const int U = n_kt * HG;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  const uint32_t bar_q = smem_u32(bars + 0);
  const uint32_t bar_o = smem_u32(bars + 1);
  const uint32_t bar_kv_full = smem_u32(bars + 2);                  // [4]
  const uint32_t bar_kv_empty = smem_u32(bars + 6);                 // [4]
  const uint32_t bar_s_full = smem_u32(bars + 10);                  // [2]
  const uint32_t bar_p_full = smem_u32(bars + 12);                  // [2]
  const uint32_t bar_bkt_full = smem_u32(bars + 14);                // [2]
  const uint32_t bar_bkt_free = smem_u32(bars + 16);                // [2]
  const uint32_t bar_tab_full = smem_u32(bars + 18);                // [2]
  const uint32_t bar_tab_free = smem_u32(bars + 20);                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  if (tid == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_o, 1);
    for (int s = 0; s < F2_STAGES; ++s) { mbar_init(bar_kv_full + 8 * s, 1); mbar_init(bar_kv_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_s_full + 8 * s, 1);
      mbar_init(bar_p_full + 8 * s, F2_EPI / 32);
      mbar_init(bar_bkt_full + 8 * s, 1);
      mbar_init(bar_bkt_free + 8 * s, F2_EPI / 32);
      mbar_init(bar_tab_full + 8 * s, 1);
      mbar_init(bar_tab_free + 8 * s, F2_EPI / 32);
    }
    fence_barrier_init();
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  if (HAS_BIAS && warp == 3) {
    float* tsw = reinterpret_cast<float*>(smem + L::tsw);
    for (int i = lane; i < 136; i += 32) tsw[i] = i <= p.nb ? 0.5f * p.ts_w[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, HG * AT_TILE_BYTES);
      for (int hh = 0; hh < HG; ++hh)
        tma_load_2d(smem_u32(smem + L::q + hh * AT_TILE_BYTES), &tmQ, (h0 + hh) * AT_D,
                    (int) (off0 + i0), bar_q);
      const uint8_t* cache_seq = nullptr;
      if (HAS_BIAS) {   // slots (qt, 0..qt) of this sequence, "Q orientation" half of each
        const int64_t tps = (int64_t) p.cache_nt * (p.cache_nt + 1) / 2;
        cache_seq = p.bcache + ((int64_t) b * tps + (int64_t) qt * (qt + 1) / 2) * 32768;
      }
      for (int u = 0; u < U; ++u) {
        const int st = u % F2_STAGES, j = u / HG, hh = u % HG;
        if (HAS_BIAS && hh == 0) {
          const int pb = j & 1;
          mbar_wait_parked(bar_bkt_free + 8 * pb, ((j >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(bar_bkt_full + 8 * pb, 16384);
          bulk_load_1d(smem_u32(smem + L::bkt + pb * 16384), cache_seq + (int64_t) j * 32768, 16384,
                       bar_bkt_full + 8 * pb);
        }
        mbar_wait_parked(bar_kv_empty + 8 * st, ((u / F2_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(bar_kv_full + 8 * st, 2 * AT_TILE_BYTES);
        const uint32_t dst = smem_u32(smem + L::kv + st * 2 * AT_TILE_BYTES);
        tma_load_2d(dst, &tmK, (h0 + hh) * AT_D, (int) (off0 + j * AT_BN), bar_kv_full + 8 * st);
        tma_load_2d(dst + AT_TILE_BYTES, &tmV, (h0 + hh) * AT_D, (int) (off0 + j * AT_BN),
                    bar_kv_full + 8 * st);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp, one elected lane issues) =================
    const uint32_t idesc_qk = make_idesc_bf16(128, AT_BN, false, false);
    const uint32_t idesc_pv = make_idesc_bf16(128, AT_D, false, true);
    const uint64_t q_desc = make_smem_desc_sw128(smem_u32(smem + L::q), 0, 1024);
    const uint64_t kv_desc = make_smem_desc_sw128(smem_u32(smem + L::kv), 0, 1024);
    auto adv = [](uint64_t d, uint32_t bytes) { return d + (uint64_t) (bytes >> 4); };
    auto issue_qk = [&](int u) {
      const int st = u % F2_STAGES, hh = u % HG, sb = u & 1;
      mbar_wait_parked(bar_kv_full + 8 * st, (u / F2_STAGES) & 1);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < AT_D / 16; ++ks)
        umma_ss_warp(tmem + sb * 128, adv(q_desc, hh * AT_TILE_BYTES + ks * 32),
                     adv(kv_desc, st * 2 * AT_TILE_BYTES + ks * 32), idesc_qk, ks > 0);
      umma_commit_warp(bar_s_full + 8 * sb);
    };
    mbar_wait_parked(bar_q, 0);
    issue_qk(0);
    if (U > 1) issue_qk(1);
    for (int u = 0; u < U; ++u) {
      const int st = u % F2_STAGES, j = u / HG, hh = u % HG, sb = u & 1;
      mbar_wait_parked(bar_p_full + 8 * sb, (u >> 1) & 1);   // P_u written, S[sb] read
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < AT_BN / 16; ++ks)
        umma_ts_warp(tmem + 384 + hh * AT_D, tmem + 256 + sb * 64 + ks * 8,
                     adv(kv_desc, st * 2 * AT_TILE_BYTES + AT_TILE_BYTES + ks * 2048), idesc_pv,
                     (j > 0) || (ks > 0));
      umma_commit_warp(bar_kv_empty + 8 * st);
      // S[sb] is free (read by the epilogue before it arrived on p_full); the commit behind the
      // next scores also covers P V of unit u, which is what lets the epilogue overwrite P[sb]
      if (u + 2 < U) issue_qk(u + 2);
    }
    umma_commit_warp(bar_o);
  } else if (warp == 2) {
    // ================= pos_w window stager =================
    if (HAS_BIAS) {
      float* pos_all = reinterpret_cast<float*>(smem + L::pos);
      for (int j = 0; j < n_kt; ++j) {
        const int pb = j & 1;
        const int j0 = j * AT_BN;
        float vals[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) {              // pos[x] = 0.5 * pos_w[N-1 + j0 - i0 - 127 + x], x < 259
          const int x = lane + 32 * t;
          const int64_t idx = p.N - 1 + j0 - i0 - 127 + x;
          vals[t] = (x < 259 && idx >= 0 && idx < 2 * p.N - 1) ? 0.5f * p.pos_w[idx] : 0.f;
        }
        mbar_wait_parked(bar_tab_free + 8 * pb, ((j >> 1) & 1) ^ 1);
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
    const int g = (warp - 4) >> 2;                 // key columns [32g, 32g + 32) of every unit
    const int r = ((warp & 3) << 5) | lane;        // query row inside the tile = TMEM lane
    const uint32_t lane_base = (uint32_t) ((warp & 3) * 32) << 16;
    const int i = i0 + r;
    const float* tsw_s = reinterpret_cast<const float*>(smem + L::tsw);
    const float* pos_all = reinterpret_cast<const float*>(smem + L::pos);
    const uint32_t half_half = BF16M ? 0x3f003f00u : 0x38003800u;   // (0.5, 0.5) as bf16x2 / f16x2
    const int sft = (3 - r) & 3;                   // x(e) = c0 + e - r + 127 ; x(0) - sft is 4-aligned
    uint32_t hb2[16];                              // bias / 2 of the current key tile, packed pairs
#pragma unroll
    for (int w = 0; w < 16; ++w) hb2[w] = 0u;
    for (int u = 0; u < U; ++u) {
      const int j = u / HG, hh = u % HG, sb = u & 1, pb = j & 1;
      const bool diag = (j == qt);
      const uint8_t* bkt_s = smem + L::bkt + pb * 16384;
      const float* pos4 = pos_all + (pb * 4 + sft) * L::POS_COPY + (127 - r - sft);
      if (HAS_BIAS && hh == 0) {
        mbar_wait(bar_tab_full + 8 * pb, (j >> 1) & 1);
        mbar_wait(bar_bkt_full + 8 * pb, (j >> 1) & 1);
      }
      mbar_wait(bar_s_full + 8 * sb, (u >> 1) & 1);
      tc_fence_after();
      uint32_t pk[16];
      if (diag && (warp & 3) < g) {                // every column of the block is above every row
#pragma unroll
        for (int w = 0; w < 16; ++w) pk[w] = 0u;
      } else {
        uint32_t sv[32];
        tmem_ld32(tmem + lane_base + sb * 128 + 32 * g, sv);
        if (HAS_BIAS && hh == 0) {
          // bias / 2 of this thread's 32 (row, column) pairs: head independent, so it is built for the
          // first head of the pair (while the TMEM load is in flight) and kept in registers for the
          // second one
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            const int c0 = 32 * g + 8 * c8;
            // 8 bucket bytes of this query row: key chunk (c0 / 16), bytes (c0 % 16) .. +7
            const uint2 raw = *reinterpret_cast<const uint2*>(
                bkt_s + ((size_t) (c0 >> 4) * 128 + r) * 16 + (c0 & 8));
            const float4 pa = *reinterpret_cast<const float4*>(pos4 + c0);       // e = 0..3
            const float4 pc = *reinterpret_cast<const float4*>(pos4 + c0 + 4);   // e = 4..7
            const float pz[8] = {pa.x, pa.y, pa.z, pa.w, pc.x, pc.y, pc.z, pc.w};
            const uint32_t b0 = raw.x & 0xffu, b4 = b0 * 0x01010101u;
            float tv[8];
            if (__all_sync(0xffffffffu, (raw.x == b4) & (raw.y == b4))) {   // one bucket per row
              const float t = tsw_s[b0];
#pragma unroll
              for (int e = 0; e < 8; ++e) tv[e] = pz[e] + t;
            } else {
              const uint32_t w2[2] = {raw.x, raw.y};
#pragma unroll
              for (int e = 0; e < 8; ++e) tv[e] = pz[e] + tsw_s[(w2[e >> 2] >> (8 * (e & 3))) & 0xffu];
            }
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2)
              hb2[4 * c8 + e2] = BF16M ? pack_bf16x2(tv[2 * e2], tv[2 * e2 + 1]) : pack_f16x2(tv[2 * e2], tv[2 * e2 + 1]);
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int w = 0; w < 16; ++w) {
          // h = S/2 + bias/2 ; SiLU(S + bias) = h + h * tanh(h)
          if (BF16M) {
            const uint32_t s2 = pack_bf16x2(__uint_as_float(sv[2 * w]), __uint_as_float(sv[2 * w + 1]));
            uint32_t h2, t2;
            asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(h2) : "r"(s2), "r"(half_half), "r"(hb2[w]));
            asm("tanh.approx.bf16x2 %0, %1;" : "=r"(t2) : "r"(h2));
            asm("fma.rn.bf16x2 %0, %1, %2, %1;" : "=r"(pk[w]) : "r"(h2), "r"(t2));
          } else {
            const uint32_t s2 = pack_f16x2(__uint_as_float(sv[2 * w]), __uint_as_float(sv[2 * w + 1]));
            uint32_t h2, p2;
            asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(h2) : "r"(s2), "r"(half_half), "r"(hb2[w]));
            const uint32_t t2 = tanh_approx_f16x2(h2);
            asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(p2) : "r"(h2), "r"(t2));
            const float2 pf = __half22float2(*reinterpret_cast<const __half2*>(&p2));
            pk[w] = pack_bf16x2(pf.x, pf.y);
          }
        }
        if (diag) {   // causal mask, diagonal tile only (warp-uniform branch): keep columns <= r
          const int rel = r - 32 * g;
#pragma unroll
          for (int w = 0; w < 16; ++w)
            pk[w] &= (2 * w <= rel ? 0x0000ffffu : 0u) | (2 * w + 1 <= rel ? 0xffff0000u : 0u);
        }
      }
      // P_u: bf16 pairs of key columns 32g .. 32g+31 -> TMEM columns 256 + 64 sb + 16 g .. +15
      tmem_st16(tmem + lane_base + 256 + sb * 64 + 16 * g, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_p_full + 8 * sb);
        if (HAS_BIAS && hh == HG - 1) {
          mbar_arrive(bar_tab_free + 8 * pb);
          mbar_arrive(bar_bkt_free + 8 * pb);
        }
      }
    }
    // ---- O epilogue: scale by 1/N, bf16, store this thread's row: warpgroup g -> head g / 2,
    //      columns 32 (g & 1) .. + 31 ----
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv_n = 1.0f / (float) p.N;
    {
      const int hh = g >> 1, cpart = g & 1;
      uint32_t ov[32];
      __nv_bfloat16* dst = p.out + (off0 + i) * p.ldo + (h0 + hh) * AT_D + cpart * 32;
      tmem_ld32(tmem + lane_base + 384 + hh * AT_D + cpart * 32, ov);
      tmem_ld_wait();
      if (i < n) {
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(ov[v4 * 8 + 0]) * inv_n, __uint_as_float(ov[v4 * 8 + 1]) * inv_n);
          o.y = pack_bf16x2(__uint_as_float(ov[v4 * 8 + 2]) * inv_n, __uint_as_float(ov[v4 * 8 + 3]) * inv_n);
          o.z = pack_bf16x2(__uint_as_float(ov[v4 * 8 + 4]) * inv_n, __uint_as_float(ov[v4 * 8 + 5]) * inv_n);
          o.w = pack_bf16x2(__uint_as_float(ov[v4 * 8 + 6]) * inv_n, __uint_as_float(ov[v4 * 8 + 7]) * inv_n);
          *reinterpret_cast<uint4*>(dst + v4 * 8) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

// bias from the bucket cache (or no bias at all), even head count
bool hstu_attn_fwd2_sm100_usable(const grb_hstu_attn_args* a) {
  if (a->H % F2_HG != 0) return false;
  if (a->timestamps && a->bucket_cache == nullptr) return false;
  return true;
}

int hstu_attn_fwd2_sm100(const grb_hstu_attn_args* a, cudaStream_t st) {
  if (a->B == 0 || a->max_len == 0) return GRB_OK;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  const uint64_t W = (uint64_t) a->H * AT_D;
  if ((rc = make_tmap_bf16_2d(&tmQ, a->q, a->T, W, a->ldq, AT_BM)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, a->k, a->T, W, a->ldk, AT_BN)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, a->v, a->T, W, a->ldv, AT_BN)) != GRB_OK) return rc;
  Fwd2Params p{};
  p.N = a->N; p.T = a->T; p.H = a->H; p.nb = a->num_buckets; p.index_bits = a->index_bits;
  p.n_qt = (int) ceil_div(a->max_len, AT_BM);
  p.offsets = a->offsets; p.ts_w = a->ts_w; p.pos_w = a->pos_w;
  p.bcache = a->timestamps ? a->bucket_cache : nullptr;
  p.cache_nt = (int) ceil_div(a->bucket_cache_max_len, AT_BM);
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out); p.ldo = a->ldo;
  const size_t smem = F2Smem::total + 1024;
  dim3 grid((unsigned) (p.n_qt * (a->H / F2_HG)), (unsigned) a->B);
  // GRB_FWD2_F16=1 (A/B switch): the f16 SiLU chain with a conversion to bf16 at the end
  static const bool f16_chain = [] { const char* e = getenv("GRB_FWD2_F16"); return e && e[0] == '1'; }();
  auto launch = [&](auto kern) -> int {
    GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    kern<<<grid, F2_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
    return GRB_OK;
  };
  if (a->timestamps) rc = f16_chain ? launch(hstu_attn_fwd2_kernel<true, false>) : launch(hstu_attn_fwd2_kernel<true, true>);
  else rc = f16_chain ? launch(hstu_attn_fwd2_kernel<false, false>) : launch(hstu_attn_fwd2_kernel<false, true>);
  if (rc != GRB_OK) return rc;
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}  // namespace grb
