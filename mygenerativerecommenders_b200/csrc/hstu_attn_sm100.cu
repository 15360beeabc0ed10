// hstu_attn_sm100.cu — HSTU jagged attention forward on Blackwell tensor cores
// (bf16 in, fp32 accumulate, dqk = dv = 64).
//
// Reference math: /root/reference/src/generative_recommenders_pl/models/sequential_encoders/
// hstu.py:96-128 (bias) + :134-205 (attention).  One CTA owns (sequence b, 128 query rows,
// a group of HG heads) and walks the causal key tiles j = 0..iq:
//
//   TMA warp   : Q_h tiles once; then per unit (key tile j, head h) a 128x64 K tile and a
//                128x64 V tile (128-byte swizzle) into a 3-stage ring, rows addressed through
//                the jagged offsets (row coordinate = off[b] + j0).
//   MMA warp   : S = Q_h K_h^T   tcgen05.mma  M128 N128 K64  (both operands K-major smem)
//                O_h += P V_h    tcgen05.mma  M128 N64  K128 (A = P from TMEM, B = V MN-major)
//                S is double buffered in TMEM; P (bf16) aliases the first 64 columns of its S.
//   epilogue   : 2 warpgroups (thread = query row).  Per key tile the head-independent bias
//                tile  pos_w[N-1+j-i] + ts_w[bucket(|ts[i+1]-ts[j]|)]  is computed ONCE into
//                shared memory (fp16, pre-halved) with integer-only bucketing (clz octave table
//                built from the reference's tabulated thresholds), then every head of the group
//                does  h = S/2 + bias/2 ; P = h + h*tanh(h) = SiLU(S+bias)  (one MUFU per
//                element), causal mask on the diagonal tile, bf16 pack, tcgen05.st.
//                The 1/N scale is applied once to O.
//
// TMEM: S0 [0,128) S1 [128,256) O_h [256 + 64h, +64)  -> 512 columns at HG = 4.
#include "hstu_attn_sm100.cuh"

namespace grb {

using namespace ptx;

constexpr int AT_STAGES = 3;
constexpr int AT_EPI = 512;         // 4 epilogue warpgroups: (S buffer 0/1) x (column half 0/1)
constexpr int AT_THREADS = 128 + AT_EPI;   // warp 0 TMA, warp 1 MMA, warps 2-3 setup

struct AttnFwdParams {
  int64_t N, T;
  int H, nb, index_bits, n_qt;
  const void* offsets;
  const int64_t* ts;
  const float* ts_w;
  const float* pos_w;
  const int64_t* thr;
  const uint32_t* octaves;   // optional precomputed octave table
  const uint8_t* bcache;     // optional bucket-index tiles (hstu_bucket_cache.cu)
  int cache_nt;              // query tiles per sequence the cache was laid out for
  __nv_bfloat16* out;
  int64_t ldo;
};

template <int HG>
struct AttnSmem {
  // offsets into dynamic smem (1024-byte aligned base)
  static constexpr int q = 0;                                   // HG x 16 KiB
  static constexpr int kv = q + HG * AT_TILE_BYTES;             // STAGES x (K 16 KiB + V 16 KiB)
  static constexpr int bias = kv + AT_STAGES * 2 * AT_TILE_BYTES;  // 128 x 128 fp16 = 32 KiB
  static constexpr int bkt = bias + AT_BM * AT_BN * 2;          // 128 x 128 uint8 bucket tile
  static constexpr int tsk = bkt + AT_BM * AT_BN;               // 128 x int64
  static constexpr int tsk32 = tsk + 128 * 8;                   // 128 x uint32 (ts - tmin)
  static constexpr int red = tsk32 + 128 * 4;                   // 32 x int64 scratch
  static constexpr int pos = red + 32 * 8;                      // 256 x float
  static constexpr int tsw = pos + 256 * 4;                     // up to 4097 floats -> cap 132
  static constexpr int oct = tsw + 136 * 4;                     // 32 x OctRec
  static constexpr int bars = oct + 32 * 16;                    // barriers
  static constexpr int total = bars + 256;
};

template <int HG, bool HAS_BIAS>
__global__ void __launch_bounds__(AT_THREADS, 1) hstu_attn_fwd_sm100_kernel(
    const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
    const __grid_constant__ CUtensorMap tmV, AttnFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the 128-byte swizzle atoms need 1024-byte aligned tiles: align by hand (1 KiB of slack)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  using L = AttnSmem<HG>;
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
  const int U = n_kt * HG;              // units

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  const uint32_t bar_q = smem_u32(bars + 0);
  const uint32_t bar_o = smem_u32(bars + 1);
  const uint32_t bar_kv_full = smem_u32(bars + 2);                 // [STAGES]
  const uint32_t bar_kv_empty = smem_u32(bars + 2 + AT_STAGES);    // [STAGES]
  const uint32_t bar_s_full = smem_u32(bars + 2 + 2 * AT_STAGES);  // [2]
  const uint32_t bar_p_full = smem_u32(bars + 4 + 2 * AT_STAGES);  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 2 * AT_STAGES);
  int* flags = reinterpret_cast<int*>(bars + 7 + 2 * AT_STAGES);   // [0] slow, [1] b_zero
  const uint32_t bar_bkt = smem_u32(bars + 8 + 2 * AT_STAGES);

  if (tid == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_o, 1);
    for (int s = 0; s < AT_STAGES; ++s) { mbar_init(bar_kv_full + 8 * s, 1); mbar_init(bar_kv_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_s_full + 8 * s, 1); mbar_init(bar_p_full + 8 * s, 8); }
    mbar_init(bar_bkt, 1);
    fence_barrier_init();
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  if (HAS_BIAS && warp == 2) {
    if (p.octaves) load_octave_table(reinterpret_cast<OctRec*>(smem + L::oct), flags, p.octaves, lane);
    else build_octave_table(reinterpret_cast<OctRec*>(smem + L::oct), flags, p.thr, p.nb, lane);
  }
  if (HAS_BIAS && warp == 3) {
    float* tsw = reinterpret_cast<float*>(smem + L::tsw);
    for (int i = lane; i <= p.nb && i < 136; i += 32) tsw[i] = 0.5f * p.ts_w[i];
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
      for (int u = 0; u < U; ++u) {
        const int st = u % AT_STAGES, j = u / HG, hh = u % HG;
        mbar_wait_parked(bar_kv_empty + 8 * st, ((u / AT_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(bar_kv_full + 8 * st, 2 * AT_TILE_BYTES);
        const uint32_t dst = smem_u32(smem + L::kv + st * 2 * AT_TILE_BYTES);
        tma_load_2d(dst, &tmK, (h0 + hh) * AT_D, (int) (off0 + j * AT_BN), bar_kv_full + 8 * st);
        tma_load_2d(dst + AT_TILE_BYTES, &tmV, (h0 + hh) * AT_D, (int) (off0 + j * AT_BN),
                    bar_kv_full + 8 * st);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc_qk = make_idesc_bf16(128, AT_BN, false, false);
      const uint32_t idesc_pv = make_idesc_bf16(128, AT_D, false, true);
      mbar_wait_parked(bar_q, 0);
      auto issue_qk = [&](int u) {
        const int st = u % AT_STAGES, hh = u % HG, sb = u & 1;
        mbar_wait_parked(bar_kv_full + 8 * st, (u / AT_STAGES) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(smem + L::q + hh * AT_TILE_BYTES);
        const uint32_t ka = smem_u32(smem + L::kv + st * 2 * AT_TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < AT_D / 16; ++ks)
          umma_ss(tmem + sb * 128, make_smem_desc_sw128(qa + ks * 32, 0, 1024),
                  make_smem_desc_sw128(ka + ks * 32, 0, 1024), idesc_qk, ks > 0);
        umma_commit(bar_s_full + 8 * sb);
      };
      issue_qk(0);
      if (U > 1) issue_qk(1);
      for (int u = 0; u < U; ++u) {
        const int st = u % AT_STAGES, j = u / HG, hh = u % HG, sb = u & 1;
        mbar_wait_parked(bar_p_full + 8 * sb, (u >> 1) & 1);
        tc_fence_after();
        const uint32_t va = smem_u32(smem + L::kv + st * 2 * AT_TILE_BYTES + AT_TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < AT_BN / 16; ++ks)
          umma_ts(tmem + 256 + hh * AT_D, tmem + sb * 128 + ks * 8,
                  make_smem_desc_sw128(va + ks * 2048, 0, 1024), idesc_pv, (j > 0) || (ks > 0));
        umma_commit(bar_kv_empty + 8 * st);
        if (u + 2 < U) issue_qk(u + 2);
      }
      umma_commit(bar_o);
    }
  } else if (warp >= 4) {
    // ================= epilogue warpgroups =================
    const int wg = (warp - 4) >> 2;                // 0..3
    const int g = wg & 1;                          // S buffer: this warpgroup works on units u = g (mod 2)
    const int ch = wg >> 1;                        // which 64 of the 128 key columns of a unit
    const int r = ((warp & 3) << 5) | lane;        // query row inside the tile = TMEM lane
    const uint32_t lane_base = (uint32_t) ((warp & 3) * 32) << 16;
    const int i = i0 + r;
    __half* bias_s = reinterpret_cast<__half*>(smem + L::bias);
    int64_t* tsk_s = reinterpret_cast<int64_t*>(smem + L::tsk);
    float* pos_s = reinterpret_cast<float*>(smem + L::pos);
    const float* tsw_s = reinterpret_cast<const float*>(smem + L::tsw);
    const OctRec* oct = reinterpret_cast<const OctRec*>(smem + L::oct);
    uint32_t* tsk32_s = reinterpret_cast<uint32_t*>(smem + L::tsk32);
    int64_t ts_q = 0;
    uint32_t tq32 = 0;
    bool slow = false, narrow = false;
    int64_t tmin = 0;
    if (HAS_BIAS) {
      ts_q = ext_ts_at(p.ts, b, p.N, (int64_t) i + 1);
      slow = flags[0] != 0;
      const int cnt = (int) (n64 + 1 < p.N ? n64 + 1 : p.N);     // indices 0..min(n, N-1)
      const TsRange tr = scan_ts_range<AT_EPI>(p.ts + (int64_t) b * p.N, cnt, tid - 128,
                                       reinterpret_cast<int64_t*>(smem + L::red), 3);
      narrow = tr.narrow && !slow;
      tmin = tr.tmin;
      tq32 = (uint32_t) (ts_q - tmin);   // garbage for rows >= n: their output is never stored
    }
    auto stage_tables = [&](int j) {   // key timestamps and the pos_w window of key tile j
      const int j0 = j * AT_BN;
      if (wg == 0) {
        const int64_t tk = ext_ts_at(p.ts, b, p.N, (int64_t) j0 + r);
        tsk_s[r] = tk;
        tsk32_s[r] = (uint32_t) (tk - tmin);
      } else if (wg == 1) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int x = r + 128 * t;               // pos_s[x] = 0.5 * pos_w[N-1 + j0 - i0 - 127 + x]
          const int64_t idx = p.N - 1 + j0 - i0 - 127 + x;
          pos_s[x] = (idx >= 0 && idx < 2 * p.N - 1) ? 0.5f * p.pos_w[idx] : 0.f;
        }
      }
    };
    if (HAS_BIAS) stage_tables(0);
    const uint32_t half_half = 0x38003800u;        // (0.5h, 0.5h)
    // cached bucket tiles: slot (qt, j) of this sequence, "Q orientation" half
    const bool cached = HAS_BIAS && p.bcache != nullptr;
    const uint8_t* bkt_s = smem + L::bkt;
    const uint8_t* cache_seq = nullptr;
    if (cached) {
      const int64_t tps = (int64_t) p.cache_nt * (p.cache_nt + 1) / 2;
      cache_seq = p.bcache + ((int64_t) b * tps + (int64_t) qt * (qt + 1) / 2) * 32768;
      if (tid == 128) {
        mbar_arrive_expect_tx(bar_bkt, 16384);
        bulk_load_1d(smem_u32(bkt_s), cache_seq, 16384, bar_bkt);
      }
    }
    for (int j = 0; j < n_kt; ++j) {
      if (HAS_BIAS) {
        named_bar_sync(2, AT_EPI);                 // previous tile's bias fully consumed; tables visible
        // this warpgroup's quarter of the bias tile: columns [32wg, 32wg+32), pre-halved fp16
#pragma unroll 2
        for (int c8 = 0; c8 < 4; ++c8) {
          const int cb = 32 * wg + 8 * c8;
          float v[8];
          if (cached) {
            if (c8 == 0) mbar_wait(bar_bkt, j & 1);
            // 8 bucket bytes of this row: chunk (cb / 16), bytes (cb % 16) .. +7
            const uint2 raw = *reinterpret_cast<const uint2*>(
                bkt_s + ((size_t) (cb >> 4) * 128 + r) * 16 + (cb & 8));
            const uint32_t w2[2] = {raw.x, raw.y};
#pragma unroll
            for (int e = 0; e < 8; ++e)
              v[e] = pos_s[cb + e - r + 127] + tsw_s[(w2[e >> 2] >> (8 * (e & 3))) & 0xffu];
          } else if (narrow) {
            const uint4 ta = *reinterpret_cast<const uint4*>(tsk32_s + cb);
            const uint4 tb = *reinterpret_cast<const uint4*>(tsk32_s + cb + 4);
            const uint32_t tk[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
            for (int e = 0; e < 8; ++e)
              v[e] = pos_s[cb + e - r + 127] + tsw_s[bucket_narrow(oct, __usad(tq32, tk[e], 0u))];
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              int64_t d = ts_q - tsk_s[cb + e];
              d = d < 0 ? -d : d;
              v[e] = pos_s[cb + e - r + 127] + tsw_s[bucket_wide(oct, p.thr, p.nb, slow, d)];
            }
          }
          uint4 pk;
          pk.x = pack_f16x2(v[0], v[1]); pk.y = pack_f16x2(v[2], v[3]);
          pk.z = pack_f16x2(v[4], v[5]); pk.w = pack_f16x2(v[6], v[7]);
          *reinterpret_cast<uint4*>(bias_s + ((size_t) (cb >> 3) * 128 + r) * 8) = pk;
        }
        named_bar_sync(1, AT_EPI);                 // bias tile complete
        if (j + 1 < n_kt) stage_tables(j + 1);     // tables are free again after barrier 1
        if (cached && tid == 128 && j + 1 < n_kt) {   // so is the bucket tile buffer
          mbar_arrive_expect_tx(bar_bkt, 16384);
          bulk_load_1d(smem_u32(bkt_s), cache_seq + (int64_t) (j + 1) * 32768, 16384, bar_bkt);
        }
      }
      const bool diag = (j == qt);
      for (int hh = g; hh < HG; hh += 2) {
        const int u = j * HG + hh;                 // u & 1 == g because HG is even
        const uint32_t s_addr = tmem + lane_base + g * 128;
        mbar_wait(bar_s_full + 8 * g, (u >> 1) & 1);
        tc_fence_after();
        // P (bf16, 64 TMEM columns) overlays S columns [0, 64).  P chunk c (16 columns) lands on
        // S chunk c / 2, so: the lower-half warpgroup (S chunks 0, 1) reads chunk 1 FIRST, tells
        // the upper-half warpgroup, and only then reads chunk 0 and writes P0, P1 (both over S
        // chunk 0, which only it reads); the upper-half warpgroup (S chunks 2, 3, never
        // overwritten) waits for that signal before its first P write (P2 lands on S chunk 1).
        auto compute_chunk = [&](int c32, uint32_t (&pk)[16]) {
          if (diag && (warp & 3) < c32) {   // diagonal tile: all 32 key columns lie above every
#pragma unroll                             // row of this warp -> P is zero, nothing to compute
            for (int w = 0; w < 16; ++w) pk[w] = 0u;
            return;
          }
          uint32_t sv[32];
          tmem_ld32(s_addr + c32 * 32, sv);
          tmem_ld_wait();
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            uint4 raw = make_uint4(0u, 0u, 0u, 0u);
            if (HAS_BIAS)
              raw = *reinterpret_cast<const uint4*>(bias_s + ((size_t) (c32 * 4 + c8) * 128 + r) * 8);
            const uint32_t hb[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              const int cc = c8 * 8 + 2 * e2;
              // h = S/2 + bias/2 ; SiLU(S + bias) = h + h * tanh(h)
              const uint32_t s2 = pack_f16x2(__uint_as_float(sv[cc]), __uint_as_float(sv[cc + 1]));
              uint32_t h2, p2;
              asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(h2) : "r"(s2), "r"(half_half), "r"(hb[e2]));
              const uint32_t t2 = tanh_approx_f16x2(h2);
              asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(p2) : "r"(h2), "r"(t2));
              const float2 pf = __half22float2(*reinterpret_cast<const __half2*>(&p2));
              pk[cc >> 1] = pack_bf16x2(pf.x, pf.y);
            }
          }
          if (diag) {   // causal mask, diagonal tile only (warp-uniform branch): keep columns <= r
            const int rel = r - c32 * 32;
#pragma unroll
            for (int w = 0; w < 16; ++w)
              pk[w] &= (2 * w <= rel ? 0x0000ffffu : 0u) | (2 * w + 1 <= rel ? 0xffff0000u : 0u);
          }
        };
        uint32_t pk0[16], pk1[16];
        if (ch == 0) {
          compute_chunk(1, pk1);
          tc_fence_before();
          asm volatile("bar.arrive %0, 256;" ::"r"(5 + g) : "memory");   // S chunk 1 is consumed
          compute_chunk(0, pk0);
          tmem_st16(s_addr + 0, pk0);
          tmem_st16(s_addr + 16, pk1);
        } else {
          compute_chunk(2, pk0);
          named_bar_sync(5 + g, 256);
          tc_fence_after();
          tmem_st16(s_addr + 32, pk0);
          compute_chunk(3, pk1);
          tmem_st16(s_addr + 48, pk1);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p_full + 8 * g);
      }
    }
    // ---- O epilogue: scale by 1/N, bf16, store this thread's row ----
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv_n = 1.0f / (float) p.N;
    for (int hh = g; hh < HG; hh += 2) {
      uint32_t ov[32];
      __nv_bfloat16* dst = p.out + (off0 + i) * p.ldo + (h0 + hh) * AT_D + ch * 32;
      tmem_ld32(tmem + lane_base + 256 + hh * AT_D + ch * 32, ov);
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

bool hstu_attn_fwd_sm100_supported(const grb_hstu_attn_args* a) {
  if (a->dtype != GRB_BF16 || a->dqk != AT_D || a->dv != AT_D) return false;
  if (a->H % 2 != 0) return false;
  if (a->timestamps && a->num_buckets > 128) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(a->q) || !al16(a->k) || !al16(a->v) || !al16(a->out)) return false;
  if ((a->ldq * 2) % 16 || (a->ldk * 2) % 16 || (a->ldv * 2) % 16 || (a->ldo * 2) % 16) return false;
  if (a->T >= (1ll << 31) || a->T == 0) return false;
  return true;
}

template <int HG>
static int launch_fwd_sm100(const grb_hstu_attn_args* a, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  const uint64_t W = (uint64_t) a->H * AT_D;
  if ((rc = make_tmap_bf16_2d(&tmQ, a->q, a->T, W, a->ldq, AT_BM)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, a->k, a->T, W, a->ldk, AT_BN)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, a->v, a->T, W, a->ldv, AT_BN)) != GRB_OK) return rc;
  AttnFwdParams p{};
  p.N = a->N; p.T = a->T; p.H = a->H; p.nb = a->num_buckets; p.index_bits = a->index_bits;
  p.n_qt = (int) ceil_div(a->max_len, AT_BM);
  p.offsets = a->offsets; p.ts = a->timestamps; p.ts_w = a->ts_w; p.pos_w = a->pos_w;
  p.thr = a->bucket_thresholds; p.octaves = a->bucket_octaves;
  p.bcache = a->timestamps ? a->bucket_cache : nullptr;
  p.cache_nt = (int) ceil_div(a->bucket_cache_max_len, AT_BM);
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out); p.ldo = a->ldo;
  const size_t smem = AttnSmem<HG>::total + 1024;
  dim3 grid((unsigned) (p.n_qt * (a->H / HG)), (unsigned) a->B);
  if (a->timestamps) {
    auto kern = hstu_attn_fwd_sm100_kernel<HG, true>;
    GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    kern<<<grid, AT_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
  } else {
    auto kern = hstu_attn_fwd_sm100_kernel<HG, false>;
    GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    kern<<<grid, AT_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
  }
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int hstu_attn_fwd_sm100(const grb_hstu_attn_args* a, cudaStream_t st) {
  if (a->B == 0 || a->max_len == 0) return GRB_OK;
  if (a->H % 4 == 0) return launch_fwd_sm100<4>(a, st);
  return launch_fwd_sm100<2>(a, st);
}

}  // namespace grb
