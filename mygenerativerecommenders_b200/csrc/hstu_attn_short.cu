// hstu_attn_short.cu — HSTU jagged attention for SHORT sequences (n <= 256 tokens: the ml-1m /
// ml-20m / amazon-books shapes), forward and backward, bf16, head dim 64.
//
// Reference math: /root/reference/src/generative_recommenders_pl/models/sequential_encoders/
// hstu.py:96-128 (bias) + :134-205 (attention) and its backward.
//
// Why a second pair of kernels: at these lengths every 128x128 tile is a diagonal or ragged tile,
// a (sequence, head) is 1-3 tiles of work, and the long-sequence kernels (one CTA per SM: 512 TMEM
// columns, 224 KiB of shared memory, a pipeline that needs many tiles to fill) spend their time in
// prologue / load / drain latency: 1.9 % of the tensor peak, 29 % issue-slot utilisation (ncu,
// profiles/r1_hstu_attn_bwd_sm100_c2shape_ncu.txt).  Here a CTA is small enough that TWO are
// resident per SM (256 TMEM columns, < 100 KiB of shared memory, 320 threads), so one CTA's
// epilogue overlaps the other's loads and MMAs, and the per-element work is cut:
//
//   * the causal and length masks live IN the per-batch bucket tiles (grb_hstu_bucket_tiles_masked):
//     a masked pair holds bucket 255, which the kernels' ts_w table maps to a bias of -15000;
//     tanh.approx saturates to exactly -1 there (benchmarks/probes/tanh_sat_probe.cu: exact for
//     every x <= -8, f32 and f16x2), so P = h + h tanh(h) and dS = dP (1+t + h (1-t^2)) are exactly
//     0 without a single compare or select in the epilogue.  The bias itself is two shared-memory
//     lookups per element (ts_w[bucket], pos_w by distance; tables staged once per persistent CTA).
//     (A first version tabulated fp16 bias tiles per layer with an extra kernel: 26 us per layer
//     for the kernel and 25 MB of tiles, more than it saved in the epilogues.)
//   * backward: one CTA owns a whole (sequence, head): dQ of a query tile is complete inside the
//     CTA, so there is no fp32 dQ accumulator to zero, no bulk reduce-add and no conversion pass
//     (for 128 < n <= 256 the partial dQ of query tile 1 makes one plain fp32 round trip through
//     the caller's dq_accum scratch, written and read back by the same thread).
//
// Work decomposition: CTA = (sequence b, head h); units = causal 128x128 tiles: one for n <= 128,
// three for n <= 256.  The grid is two phases of B*H CTAs: the first only takes sequences with
// three units, the second the others, so the heavy items start first.
#include "hstu_attn_sm100.cuh"
#include <cstdlib>
#include <cstdio>

namespace grb {

using namespace ptx;

namespace {

constexpr int SH_THREADS = 320;                 // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int SH_EPI_WARPS = 8;
constexpr float SH_MASK = -15000.f;             // pre-halved bias of a masked pair

// developer probe: globaltimer stamps (ns) of the first SH_TL_CTAS CTAs that do work, 32 slots each
constexpr int SH_TL_CTAS = 64, SH_TL_SLOTS = 32;
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define SH_STAMP(cond, slot) do { if (p.tl && (cond) && tl_id < SH_TL_CTAS) p.tl[tl_id * SH_TL_SLOTS + (slot)] = gtime(); } while (0)

// bounded mbarrier wait: a protocol bug traps (the launch fails with an error) instead of hanging
// the GPU.  Plain try_wait: the hardware suspends the thread until the phase completes or an
// internal time limit expires, so the loop body runs rarely (the form with a suspend-time hint
// compiles to a TRYWAIT + NANOSLEEP loop that re-issued every ~80 cycles: 47 % of all executed
// instructions of the first version of these kernels, ncu).
__device__ __forceinline__ void mbar_wait_g(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1u << 24)) {
      printf("grb200 hstu_attn_short: barrier timeout (block %d thread %d bar %u parity %u)\n",
             (int) blockIdx.x, (int) threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// shared -> global element-wise bf16 add of `bytes` (multiple of 16), bulk async-group
__device__ __forceinline__ void bulk_reduce_add_bf16(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.noftz.bf16 [%0], [%1], %2;" ::"l"(dst),
               "r"(src_smem), "r"(bytes)
               : "memory");
}
// predicated fire-and-forget fp32 add (no branch around it)
__device__ __forceinline__ void red_add_f32_pred(float* addr, float v, bool p) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q red.global.add.f32 [%0], %1;\n\t}" ::"l"(addr),
               "f"(v), "r"((int) p)
               : "memory");
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
// acc_lo += low bf16 half of `pair`, acc_hi += high half (fp32 accumulate)
__device__ __forceinline__ void add_bf16_pair(float& acc_lo, float& acc_hi, uint32_t pair) {
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.bf16 %0, lo, %0;\n\t"
      "add.rn.f32.bf16 %1, hi, %1;\n\t}"
      : "+f"(acc_lo), "+f"(acc_hi)
      : "r"(pair));
}

// ------------------------------------------------------------------------------------------------
// work items
// ------------------------------------------------------------------------------------------------
// Both kernels are PERSISTENT: the grid is two CTAs per SM and every CTA walks a list of
// (sequence, head) items.  (A first version launched one CTA per item: launching 1 024 CTAs of
// 100 KiB / 320 threads took 8 us of the forward's 23 us before the last ones even started,
// profiles/r2_short_attn_timeline_v1.txt.)  The item order comes from a per-batch schedule written
// behind the bias tiles (hstu_short_schedule_kernel): sequences of three units first, then those of
// one; CTA c takes items c, 2G-1-c, 2G+c, ... (snake order), so the heavy items are spread evenly.
// bias tables of one layer, staged once per (persistent) CTA.  tsw[b] = ts_w[b] / 2, tsw[255] = the
// mask; pos[256 + d] = pos_w[N-1-d] / 2 for the causal distances d = i - j in [0, N), 0 elsewhere.
__device__ __forceinline__ void stage_bias_tables(float* tsw, float* pos, const float* __restrict__ ts_w,
                                                  const float* __restrict__ pos_w, int nb, int64_t N,
                                                  int tid, int nthreads) {
  for (int i = tid; i < 256; i += nthreads)
    tsw[i] = i == 255 ? SH_MASK : ((ts_w && i <= nb) ? 0.5f * ts_w[i] : 0.f);
  for (int x = tid; x < 512; x += nthreads) {
    const int d = x - 256;
    pos[x] = (pos_w && d >= 0 && d < N) ? 0.5f * pos_w[N - 1 - d] : 0.f;
  }
}

struct ItemIter {
  int r, G, c, n_items;
  __device__ ItemIter(int G_, int c_, int n_items_) : r(-1), G(G_), c(c_), n_items(n_items_) {}
  __device__ int next() {
    for (;;) {
      ++r;
      if ((int64_t) r * G >= n_items) return -1;
      const int idx = r * G + ((r & 1) ? G - 1 - c : c);
      if (idx < n_items) return idx;
    }
  }
};

struct Item { int seq, h, n, nu; int64_t off0; };

__device__ __forceinline__ Item decode_item(int idx, int H, const int* __restrict__ order,
                                            const void* __restrict__ offsets, int index_bits, int64_t N) {
  Item it;
  it.seq = order[idx / H];
  it.h = idx % H;
  it.off0 = load_index(offsets, it.seq, index_bits);
  int64_t n64 = load_index(offsets, it.seq + 1, index_bits) - it.off0;
  if (n64 > N) n64 = N;
  it.n = (int) n64;
  it.nu = it.n > 128 ? 3 : 1;
  return it;
}

// sched[0] = sequences with n > 0, sched[1] = of those with n > 128, sched[2 ..] = their indices,
// the n > 128 ones first.  One CTA.
__global__ void __launch_bounds__(1024) hstu_short_schedule_kernel(const void* __restrict__ offsets,
                                                                   int index_bits, int B, int64_t N,
                                                                   int* __restrict__ sched) {
  __shared__ int n_heavy, n_light, c_heavy, c_light;
  if (threadIdx.x == 0) { n_heavy = 0; n_light = 0; c_heavy = 0; c_light = 0; }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int64_t n = load_index(offsets, b + 1, index_bits) - load_index(offsets, b, index_bits);
    if (n > N) n = N;
    if (n > 128) atomicAdd(&n_heavy, 1);
    else if (n > 0) atomicAdd(&n_light, 1);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int64_t n = load_index(offsets, b + 1, index_bits) - load_index(offsets, b, index_bits);
    if (n > N) n = N;
    if (n > 128) sched[2 + atomicAdd(&c_heavy, 1)] = b;
    else if (n > 0) sched[2 + n_heavy + atomicAdd(&c_light, 1)] = b;
  }
  if (threadIdx.x == 0) { sched[0] = n_heavy + n_light; sched[1] = n_heavy; }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
//   warp 0 (TMA) : per unit Q(qt), K(kt), V(kt) 128x64 tiles into a 2-slot ring (128-byte swizzle);
//                  runs ahead across items, so the next item's tiles arrive during this one's math
//   warp 1 (MMA) : S = Q K^T  M128 N128 K64 -> TMEM [0,128) ; O (+)= P V  M128 N64 K128, A = P in
//                  TMEM [128,192), B = V MN-major -> TMEM [192,256)
//   warps 2..9   : thread = query row, warpgroup g owns key columns [64g, 64g+64): h = S/2 + bias/2
//                  (fp16 pairs), P = h + h tanh(h), bf16 P -> TMEM; 16-column blocks that are masked
//                  for all 32 rows of the warp only write zeros.  O of a finished query tile is read
//                  back at the start of the next unit (its scores' commit covers the last P V),
//                  scaled by 1/N and stored.
struct ShortFwdParams {
  int64_t N;
  int B, H, index_bits, tps;
  long long* tl;            // developer probe (GRB_SHORT_TIMELINE=1)
  const void* offsets;
  const uint8_t* bcache;    // masked bucket tiles (grb_hstu_bucket_tiles_masked), 32 KiB per slot
  const int* sched;
  const float* ts_w;        // (nb + 1) or nullptr (no relative bias)
  const float* pos_w;       // (2N - 1) or nullptr
  int nb;
  __nv_bfloat16* out;
  int64_t ldo;
  int64_t T;                // rows of out; rows >= offsets[B] are zeroed here when zero_tail is set
  int zero_tail;
};

// Rows [offsets[B], T) of a (T, H * 64) bf16 matrix := 0, spread over the whole grid (fixed row buckets:
// at most one bucket's worth of rows, a few hundred KB).
__device__ __forceinline__ void zero_tail_rows_of(__nv_bfloat16* m, int64_t ld, int64_t t0, int64_t T, int H) {
  const int vpr = H * 8;                                  // 16-byte stores per row
  const int64_t nvec = (T - t0) * vpr;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t) gridDim.x * blockDim.x) {
    const int64_t r = t0 + i / vpr;
    const int c = (int) (i % vpr);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(m + r * ld) + 16 * c) = make_uint4(0u, 0u, 0u, 0u);
  }
}

struct SfSmem {
  static constexpr int ring = 0;                                   // 2 x (Q, K, V)
  static constexpr int tsw = ring + 2 * 3 * AT_TILE_BYTES;         // 256 floats: ts_w / 2, [255] = mask
  static constexpr int pos = tsw + 256 * 4;                        // 512 floats: pos[256 + (i - j)] = pos_w[N-1-(i-j)] / 2
  static constexpr int bars = pos + 512 * 4;
  static constexpr int total = bars + 128;
};

__global__ void __launch_bounds__(SH_THREADS, 2) hstu_attn_short_fwd_kernel(
    const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
    const __grid_constant__ CUtensorMap tmV, ShortFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  using L = SfSmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (p.zero_tail) zero_tail_rows_of(p.out, p.ldo, load_index(p.offsets, p.B, p.index_bits), p.T, p.H);
  const int n_items = p.sched[0] * p.H;
  const int G = (int) gridDim.x, cta = (int) blockIdx.x;
  if (cta >= n_items) return;
  const int* order = p.sched + 2;
  const int tl_id = cta < 32 ? cta : (cta >= G - 32 ? 32 + cta - (G - 32) : 1 << 20);   // first and last 32 CTAs
  SH_STAMP(tid == 0, 0);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  const uint32_t bar_kv_full = smem_u32(bars + 0);     // [2]
  const uint32_t bar_slot_free = smem_u32(bars + 2);   // [2]
  const uint32_t bar_s_full = smem_u32(bars + 4);
  const uint32_t bar_p_full = smem_u32(bars + 5);
  const uint32_t bar_o_full = smem_u32(bars + 6);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(bar_kv_full + 8 * s, 1); mbar_init(bar_slot_free + 8 * s, 1); }
    mbar_init(bar_s_full, 1);
    mbar_init(bar_p_full, SH_EPI_WARPS);
    mbar_init(bar_o_full, 1);
    fence_barrier_init();
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 256);
  if (warp >= 2)
    stage_bias_tables(reinterpret_cast<float*>(smem + L::tsw), reinterpret_cast<float*>(smem + L::pos), p.ts_w,
                      p.pos_w, p.nb, p.N, tid - 64, SH_THREADS - 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      ItemIter iter(G, cta, n_items);
      uint32_t gu = 0;
      for (int idx = iter.next(); idx >= 0; idx = iter.next()) {
        const Item it = decode_item(idx, p.H, order, p.offsets, p.index_bits, p.N);
        for (int u = 0; u < it.nu; ++u, ++gu) {
          const int sl = gu & 1, qt = u >= 1, kt = u == 2;
          mbar_wait_g(bar_slot_free + 8 * sl, ((gu >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(bar_kv_full + 8 * sl, 3 * AT_TILE_BYTES);
          const uint32_t dst = smem_u32(smem + L::ring + sl * 3 * AT_TILE_BYTES);
          tma_load_2d(dst, &tmQ, it.h * AT_D, (int) (it.off0 + qt * AT_BM), bar_kv_full + 8 * sl);
          tma_load_2d(dst + AT_TILE_BYTES, &tmK, it.h * AT_D, (int) (it.off0 + kt * AT_BN), bar_kv_full + 8 * sl);
          tma_load_2d(dst + 2 * AT_TILE_BYTES, &tmV, it.h * AT_D, (int) (it.off0 + kt * AT_BN), bar_kv_full + 8 * sl);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp, one elected lane issues) =================
    const uint32_t idesc_qk = make_idesc_bf16(128, AT_BN, false, false);
    const uint32_t idesc_pv = make_idesc_bf16(128, AT_D, false, true);
    const uint64_t ring_desc = make_smem_desc_sw128(smem_u32(smem + L::ring), 0, 1024);
    auto adv = [](uint64_t d, uint32_t bytes) { return d + (uint64_t) (bytes >> 4); };
    auto issue_qk = [&](uint32_t gu) {
      const int sl = gu & 1;
      mbar_wait_g(bar_kv_full + 8 * sl, (gu >> 1) & 1);
      tc_fence_after();
      const uint64_t q_desc = adv(ring_desc, sl * 3 * AT_TILE_BYTES);
#pragma unroll
      for (int ks = 0; ks < AT_D / 16; ++ks)
        umma_ss_warp(tmem, adv(q_desc, ks * 32), adv(q_desc, AT_TILE_BYTES + ks * 32), idesc_qk, ks > 0);
      umma_commit_warp(bar_s_full);
    };
    ItemIter iter(G, cta, n_items);
    int idx = iter.next();
    int nu = decode_item(idx, p.H, order, p.offsets, p.index_bits, p.N).nu;
    int u = 0;
    uint32_t gu = 0;
    issue_qk(0);
    for (;;) {
      const int sl = gu & 1;
      mbar_wait_g(bar_p_full, gu & 1);                // P of this unit written, S read
      tc_fence_after();
      const uint64_t v_desc = adv(ring_desc, sl * 3 * AT_TILE_BYTES + 2 * AT_TILE_BYTES);
#pragma unroll
      for (int ks = 0; ks < AT_BN / 16; ++ks)
        umma_ts_warp(tmem + 192, tmem + 128 + ks * 8, adv(v_desc, ks * 2048), idesc_pv, (u == 2) || (ks > 0));
      umma_commit_warp(bar_slot_free + 8 * sl);
      if (++u == nu) {
        idx = iter.next();
        if (idx < 0) break;
        nu = decode_item(idx, p.H, order, p.offsets, p.index_bits, p.N).nu;
        u = 0;
      }
      ++gu;
      issue_qk(gu);                                   // its commit also covers the P V just issued
    }
    umma_commit_warp(bar_o_full);
  } else {
    // ================= epilogue =================
    const int wq = warp & 3;                          // TMEM lane quarter = rows 32 wq .. +31
    const int g = (warp - 2) >> 2;                    // key columns [64 g, 64 g + 64)
    const int r = (wq << 5) | lane;
    const uint32_t lane_base = (uint32_t) (wq * 32) << 16;
    const uint32_t half_half = 0x38003800u;           // (0.5h, 0.5h)
    const float inv_n = 1.0f / (float) p.N;
    const float* tsw_s = reinterpret_cast<const float*>(smem + L::tsw);
    const float* pos_s = reinterpret_cast<const float*>(smem + L::pos);
    // O of a finished query tile, stored at the start of the next unit
    bool pend = false;
    int pend_valid = 0;
    __nv_bfloat16* pend_dst = nullptr;
    auto store_o = [&]() {
      uint32_t ov[32];
      tmem_ld32(tmem + lane_base + 192 + 32 * g, ov);
      tmem_ld_wait();
      if (r < pend_valid) {
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(ov[v4 * 8 + 0]) * inv_n, __uint_as_float(ov[v4 * 8 + 1]) * inv_n);
          o.y = pack_bf16x2(__uint_as_float(ov[v4 * 8 + 2]) * inv_n, __uint_as_float(ov[v4 * 8 + 3]) * inv_n);
          o.z = pack_bf16x2(__uint_as_float(ov[v4 * 8 + 4]) * inv_n, __uint_as_float(ov[v4 * 8 + 5]) * inv_n);
          o.w = pack_bf16x2(__uint_as_float(ov[v4 * 8 + 6]) * inv_n, __uint_as_float(ov[v4 * 8 + 7]) * inv_n);
          *reinterpret_cast<uint4*>(pend_dst + v4 * 8) = o;
        }
      }
      pend = false;
    };
    ItemIter iter(G, cta, n_items);
    uint32_t gu = 0;
    int k_item = 0;
    for (int idx = iter.next(); idx >= 0; idx = iter.next(), ++k_item) {
      const Item it = decode_item(idx, p.H, order, p.offsets, p.index_bits, p.N);
      SH_STAMP(tid == 64 && k_item < 7, 1 + 2 * k_item);
      for (int u = 0; u < it.nu; ++u, ++gu) {
        const int qt = u >= 1, kt = u == 2;
        const int rows_valid = it.n - qt * AT_BM;        // > 0
        const int cols_valid = it.n - kt * AT_BN;        // > 0 (may exceed 128)
        const uint8_t* bkt = p.bcache + ((int64_t) it.seq * p.tps + u) * 32768;   // slot index == u, Q orientation
        // distance i - j of (this row, column c) = dist0 - c
        const float* pos_r = pos_s + 256 + (qt - kt) * AT_BM + r;
        // which of this thread's four 16-column blocks have any unmasked pair in the warp's 32 rows
        bool live[4];
        uint4 bq[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int c0 = 64 * g + 16 * c;
          live[c] = (32 * wq < rows_valid) && (c0 < cols_valid) && (qt != kt || c0 <= 32 * wq + 31);
          if (live[c])   // bucket bytes of columns c0 .. c0+15 of this row (255 = masked pair)
            bq[c] = ldg_nc_v4(bkt + ((size_t) (c0 >> 4) * 128 + r) * 16);
        }
        mbar_wait_g(bar_s_full, gu & 1);
        tc_fence_after();
        if (pend) store_o();                            // this commit also covers the previous P V
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int c0 = 64 * g + 16 * c;
          uint32_t pk[8];
          if (!live[c]) {
#pragma unroll
            for (int w = 0; w < 8; ++w) pk[w] = 0u;
          } else {
            uint32_t sv[16];
            tmem_ld16(tmem + lane_base + c0, sv);
            tmem_ld_wait();
            const uint32_t bw[4] = {bq[c].x, bq[c].y, bq[c].z, bq[c].w};
#pragma unroll
            for (int e2 = 0; e2 < 8; ++e2) {
              // bias / 2 of the pair: ts_w[bucket] / 2 + pos_w[N-1+j-i] / 2
              const float b0 = tsw_s[(bw[e2 >> 1] >> (16 * (e2 & 1))) & 0xffu] + pos_r[-(c0 + 2 * e2)];
              const float b1 = tsw_s[(bw[e2 >> 1] >> (16 * (e2 & 1) + 8)) & 0xffu] + pos_r[-(c0 + 2 * e2 + 1)];
              const uint32_t hb2 = pack_f16x2(b0, b1);
              const uint32_t s2 = pack_f16x2(__uint_as_float(sv[2 * e2]), __uint_as_float(sv[2 * e2 + 1]));
              uint32_t h2, p2;
              asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(h2) : "r"(s2), "r"(half_half), "r"(hb2));
              const uint32_t t2 = tanh_approx_f16x2(h2);
              asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(p2) : "r"(h2), "r"(t2));
              const float2 pf = __half22float2(*reinterpret_cast<const __half2*>(&p2));
              pk[e2] = pack_bf16x2(pf.x, pf.y);
            }
          }
          tmem_st8(tmem + lane_base + 128 + (c0 >> 1), pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p_full);
        if (u == 0 || u == 2) {                         // last unit of its query tile
          pend = true;
          pend_valid = rows_valid;
          pend_dst = p.out + (it.off0 + qt * AT_BM + r) * p.ldo + it.h * AT_D + 32 * g;
        }
      }
      SH_STAMP(tid == 64 && k_item < 7, 2 + 2 * k_item);
    }
    mbar_wait_g(bar_o_full, 0);
    tc_fence_after();
    if (pend) store_o();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// Scores are produced TRANSPOSED (thread = key row), a query tile as two halves of 64 columns:
//   S^T_hf  = K Q_hf^T   M128 N64 K64 -> TMEM [0,64)      dP^T_hf = V dO_hf^T -> TMEM [64,128)
//   dV     += P^T_hf dO_hf   (A = bf16 P^T in TMEM, written over the S^T columns its thread owns)
//   dK     += dS^T_hf Q_hf   (A = dS^T block hf in shared memory, K-major)
//   dQ      = dS K           (A = both dS^T blocks read MN-major)            -> TMEM [64,128)
//   dV -> TMEM [128,192), dK -> TMEM [192,256).
// Units of an item: (kt, qt) = (0,0) [, (0,1), (1,1)].  K/V are loaded once per key tile, Q/dO once
// per query tile (unit 2 reuses the Q/dO of unit 1); the loads of the next item start as soon as the
// last MMA of this one has completed, i.e. while the epilogue still drains dQ / dV / dK.
struct ShortBwdParams {
  int64_t N;
  int B, H, index_bits, tps;
  long long* tl;
  const void* offsets;
  const uint8_t* bcache;
  const int* sched;
  const float* ts_w;
  const float* pos_w;
  int nb;
  __nv_bfloat16* dq; int64_t lddq;
  __nv_bfloat16* dk; int64_t lddk;
  __nv_bfloat16* dv; int64_t lddv;
  float* dq_accum;          // (T, H*64) fp32 scratch (no zero fill needed)
  float* d_ts_w; float* d_pos_w;   // (copies, nb + 1) / (copies, 2N - 1) fp32, accumulated (+=)
  int d_bias_copies;
  int64_t T;                // rows of dq / dk / dv; rows >= offsets[B] are zeroed here when zero_tail is set
  int zero_tail;
};

struct SbSmem {
  static constexpr int k = 0;
  static constexpr int v = k + AT_TILE_BYTES;
  static constexpr int q = v + AT_TILE_BYTES;
  static constexpr int dout = q + AT_TILE_BYTES;
  static constexpr int dsT = dout + AT_TILE_BYTES;                 // blocks A, B: [128 k][64 q] bf16
  static constexpr int tsw = dsT + 2 * AT_TILE_BYTES;              // as in the forward
  static constexpr int pos = tsw + 256 * 4;
  static constexpr int bars = pos + 512 * 4;
  static constexpr int total = bars + 128;
};

template <bool HAS_BIAS>
__global__ void __launch_bounds__(SH_THREADS, 2) hstu_attn_short_bwd_kernel(
    const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
    ShortBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  using L = SbSmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (p.zero_tail) {
    const int64_t t0 = load_index(p.offsets, p.B, p.index_bits);
    zero_tail_rows_of(p.dq, p.lddq, t0, p.T, p.H);
    zero_tail_rows_of(p.dk, p.lddk, t0, p.T, p.H);
    zero_tail_rows_of(p.dv, p.lddv, t0, p.T, p.H);
  }
  const int n_items = p.sched[0] * p.H;
  const int G = (int) gridDim.x, cta = (int) blockIdx.x;
  if (cta >= n_items) return;
  const int* order = p.sched + 2;
  const int tl_id = cta < 32 ? cta : (cta >= G - 32 ? 32 + cta - (G - 32) : 1 << 20);   // first and last 32 CTAs
  SH_STAMP(tid == 0, 0);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  const uint32_t bar_kv_full = smem_u32(bars + 0);
  const uint32_t bar_qdo_full = smem_u32(bars + 1);
  const uint32_t bar_s_full = smem_u32(bars + 2);      // [2]
  const uint32_t bar_half_done = smem_u32(bars + 4);   // [2]
  const uint32_t bar_pa_free = smem_u32(bars + 6);
  const uint32_t bar_dq_full = smem_u32(bars + 7);
  const uint32_t bar_dq_read = smem_u32(bars + 8);
  const uint32_t bar_dkv_full = smem_u32(bars + 9);
  const uint32_t bar_dkv_read = smem_u32(bars + 10);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  if (tid == 0) {
    mbar_init(bar_kv_full, 1);
    mbar_init(bar_qdo_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(bar_s_full + 8 * s, 1); mbar_init(bar_half_done + 8 * s, SH_EPI_WARPS); }
    mbar_init(bar_pa_free, 1);
    mbar_init(bar_dq_full, 1);
    mbar_init(bar_dq_read, SH_EPI_WARPS);
    mbar_init(bar_dkv_full, 1);
    mbar_init(bar_dkv_read, SH_EPI_WARPS);
    fence_barrier_init();
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); prefetch_tensormap(&tmdO);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 256);
  if (warp >= 2)
    stage_bias_tables(reinterpret_cast<float*>(smem + L::tsw), reinterpret_cast<float*>(smem + L::pos), p.ts_w,
                      p.pos_w, p.nb, p.N, tid - 64, SH_THREADS - 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      ItemIter iter(G, cta, n_items);
      uint32_t gu = 0;                                 // units before this item
      for (int idx = iter.next(); idx >= 0; idx = iter.next()) {
        const Item it = decode_item(idx, p.H, order, p.offsets, p.index_bits, p.N);
        // every MMA of the previous unit has completed: the operand tiles may be overwritten
        if (gu > 0) mbar_wait_g(bar_dq_full, (gu - 1) & 1);
        mbar_arrive_expect_tx(bar_kv_full, 2 * AT_TILE_BYTES);
        tma_load_2d(smem_u32(smem + L::k), &tmK, it.h * AT_D, (int) it.off0, bar_kv_full);
        tma_load_2d(smem_u32(smem + L::v), &tmV, it.h * AT_D, (int) it.off0, bar_kv_full);
        mbar_arrive_expect_tx(bar_qdo_full, 2 * AT_TILE_BYTES);
        tma_load_2d(smem_u32(smem + L::q), &tmQ, it.h * AT_D, (int) it.off0, bar_qdo_full);
        tma_load_2d(smem_u32(smem + L::dout), &tmdO, it.h * AT_D, (int) it.off0, bar_qdo_full);
        if (it.nu == 3) {
          mbar_wait_g(bar_dq_full, gu & 1);            // unit 0 done
          mbar_arrive_expect_tx(bar_qdo_full, 2 * AT_TILE_BYTES);
          tma_load_2d(smem_u32(smem + L::q), &tmQ, it.h * AT_D, (int) (it.off0 + AT_BM), bar_qdo_full);
          tma_load_2d(smem_u32(smem + L::dout), &tmdO, it.h * AT_D, (int) (it.off0 + AT_BM), bar_qdo_full);
          mbar_wait_g(bar_dq_full, (gu + 1) & 1);      // unit 1 done
          mbar_arrive_expect_tx(bar_kv_full, 2 * AT_TILE_BYTES);
          tma_load_2d(smem_u32(smem + L::k), &tmK, it.h * AT_D, (int) (it.off0 + AT_BN), bar_kv_full);
          tma_load_2d(smem_u32(smem + L::v), &tmV, it.h * AT_D, (int) (it.off0 + AT_BN), bar_kv_full);
        }
        gu += it.nu;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp, one elected lane issues) =================
    const uint32_t id_kk = make_idesc_bf16(128, 64, false, false);    // S^T, dP^T halves
    const uint32_t id_kmn = make_idesc_bf16(128, AT_D, false, true);  // dV, dK
    const uint32_t id_mnmn = make_idesc_bf16(128, AT_D, true, true);  // dQ
    const uint64_t k_desc = make_smem_desc_sw128(smem_u32(smem + L::k), 0, 1024);
    const uint64_t v_desc = make_smem_desc_sw128(smem_u32(smem + L::v), 0, 1024);
    const uint64_t q_desc = make_smem_desc_sw128(smem_u32(smem + L::q), 0, 1024);
    const uint64_t o_desc = make_smem_desc_sw128(smem_u32(smem + L::dout), 0, 1024);
    const uint64_t ds_k_desc = make_smem_desc_sw128(smem_u32(smem + L::dsT), 0, 1024);
    const uint64_t ds_mn_desc = make_smem_desc_sw128(smem_u32(smem + L::dsT), AT_TILE_BYTES, 1024);
    auto adv = [](uint64_t d, uint32_t bytes) { return d + (uint64_t) (bytes >> 4); };
    ItemIter iter(G, cta, n_items);
    uint32_t gu = 0, gk = 0, c_kv = 0, c_qdo = 0;
    for (int idx = iter.next(); idx >= 0; idx = iter.next()) {
      const Item it = decode_item(idx, p.H, order, p.offsets, p.index_bits, p.N);
      for (int u = 0; u < it.nu; ++u, ++gu) {
        const bool first_of_kt = (u != 1);
        const bool last_of_kt = (it.nu == 1) || (u >= 1);
        if (u != 1) { mbar_wait_g(bar_kv_full, c_kv & 1); ++c_kv; }
        if (u != 2) { mbar_wait_g(bar_qdo_full, c_qdo & 1); ++c_qdo; }
        if (gu > 0) mbar_wait_g(bar_dq_read, (gu - 1) & 1);    // dQ of the previous unit has left TMEM [64,128)
        tc_fence_after();
        // ---- half A scores ----
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss_warp(tmem, adv(k_desc, ks * 32), adv(q_desc, ks * 32), id_kk, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss_warp(tmem + 64, adv(v_desc, ks * 32), adv(o_desc, ks * 32), id_kk, ks > 0);
        umma_commit_warp(bar_s_full);
        mbar_wait_g(bar_half_done, gu & 1);
        if (first_of_kt && gk > 0) mbar_wait_g(bar_dkv_read, (gk - 1) & 1);   // dV / dK of the previous key tile stored
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)     // dV += P^T_A dO_A : P^T of queries 16 ks .. at TMEM col 32 (ks/2) + 8 (ks%2)
          umma_ts_warp(tmem + 128, tmem + 32 * (ks >> 1) + 8 * (ks & 1), adv(o_desc, ks * 2048), id_kmn,
                       !first_of_kt || ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)     // dK += dS^T_A Q_A
          umma_ss_warp(tmem + 192, adv(ds_k_desc, ks * 32), adv(q_desc, ks * 2048), id_kmn,
                       !first_of_kt || ks > 0);
        umma_commit_warp(bar_pa_free);
        // ---- half B scores: dP^T first (its TMEM columns are free), S^T once P^T_A has been consumed ----
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss_warp(tmem + 64, adv(v_desc, ks * 32), adv(o_desc, 8192 + ks * 32), id_kk, ks > 0);
        mbar_wait_g(bar_pa_free, gu & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss_warp(tmem, adv(k_desc, ks * 32), adv(q_desc, 8192 + ks * 32), id_kk, ks > 0);
        umma_commit_warp(bar_s_full + 8);
        mbar_wait_g(bar_half_done + 8, gu & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ts_warp(tmem + 128, tmem + 32 * (ks >> 1) + 8 * (ks & 1), adv(o_desc, (4 + ks) * 2048), id_kmn, true);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss_warp(tmem + 192, adv(ds_k_desc, AT_TILE_BYTES + ks * 32), adv(q_desc, (4 + ks) * 2048), id_kmn, true);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)     // dQ = dS K : A = dS^T (both blocks) MN-major, B = K MN-major
          umma_ss_warp(tmem + 64, adv(ds_mn_desc, ks * 2048), adv(k_desc, ks * 2048), id_mnmn, ks > 0);
        umma_commit_warp(bar_dq_full);
        if (last_of_kt) { umma_commit_warp(bar_dkv_full); ++gk; }
      }
    }
  } else {
    // ================= epilogue: thread = key row =================
    const int wq = warp & 3;
    const int g = (warp - 2) >> 2;                    // query columns [32 g, 32 g + 32) of each half
    const int r = (wq << 5) | lane;
    const uint32_t lane_base = (uint32_t) (wq * 32) << 16;
    const float inv_n = 1.0f / (float) p.N;
    const float half_inv_n = 0.5f * inv_n;
    const float* tsw_s = reinterpret_cast<const float*>(smem + L::tsw);
    const float* pos_s = reinterpret_cast<const float*>(smem + L::pos);
    uint8_t* dsT_row = smem + L::dsT + r * 128;
    // bias gradients (fp32, per head, from the unrounded dS'): this CTA's private copies
    const int64_t copy = cta % p.d_bias_copies;
    float* d_pos_mine = HAS_BIAS ? p.d_pos_w + copy * (2 * p.N - 1) : nullptr;
    float* d_ts_mine = HAS_BIAS ? p.d_ts_w + copy * (p.nb + 1) : nullptr;
    ItemIter iter(G, cta, n_items);
    uint32_t gu = 0, gk = 0;
    int k_item = 0;
    for (int idx = iter.next(); idx >= 0; idx = iter.next(), ++k_item) {
      const Item it = decode_item(idx, p.H, order, p.offsets, p.index_bits, p.N);
      const int n = it.n, h = it.h;
      const int64_t off0 = it.off0;
      SH_STAMP(tid == 64 && k_item < 7, 1 + 2 * k_item);
      for (int u = 0; u < it.nu; ++u, ++gu) {
        const int qt = u >= 1, kt = u == 2;
        const int i0 = qt * AT_BM, j0 = kt * AT_BN;
        const int jk = j0 + r;
        const bool last_of_kt = (it.nu == 1) || (u >= 1);
        const uint8_t* bkt = p.bcache + ((int64_t) it.seq * p.tps + u) * 32768 + 16384;   // K orientation
        // distance i - j of (query column c, this key row) = c + dist0
        const float* pos_r = pos_s + 256 + (qt - kt) * AT_BM - r;
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          const int cb = 64 * hf + 32 * g;               // first query column (in the tile) of this thread
          // prefetch the bucket bytes (32) of this half before the scores arrive
          uint4 bqv[2];
#pragma unroll
          for (int sc = 0; sc < 2; ++sc)
            bqv[sc] = ldg_nc_v4(bkt + ((size_t) ((cb >> 4) + sc) * 128 + r) * 16);
          mbar_wait_g(bar_s_full + 8 * hf, gu & 1);
          tc_fence_after();
          // d ts_w: along a key row the bucket is a step function of the query position: the dS' of a
          // run are summed in a register and leave with one red when the bucket changes (255 = masked
          // pair, value exactly 0).  The run is closed at the end of the thread's 32 columns.
          uint32_t run_bk = 255u;
          float run_acc = 0.f;
#pragma unroll
          for (int sc = 0; sc < 2; ++sc) {               // 16-column sub-chunks
            const int c16 = cb + 16 * sc;
            const int tcol = 32 * g + 16 * sc;           // column inside the half's TMEM region
            // valid columns of this row in the sub-chunk: lo <= e < hi; nothing to do for the warp?
            const int lo = jk - i0 - c16, hi = n - i0 - c16;
            const bool dead = __all_sync(0xffffffffu, lo >= 16 || hi <= 0 || hi <= lo);
            float am = 0.f, aw = 0.f, bm = 0.f, bw = 0.f;    // d pos_w partial sums of this 32 x 16 block
            if (dead) {
              const uint32_t z4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
              for (int c8 = 0; c8 < 2; ++c8) {
                tmem_st4(tmem + lane_base + 32 * g + 8 * sc + 4 * c8, z4);
                *reinterpret_cast<uint4*>(dsT_row + hf * AT_TILE_BYTES + ((4 * g + 2 * sc + c8) ^ (r & 7)) * 16) =
                    make_uint4(0u, 0u, 0u, 0u);
              }
              continue;
            }
            uint32_t sv[16], dv_[16];
            tmem_ld16(tmem + lane_base + tcol, sv);
            tmem_ld16(tmem + lane_base + 64 + tcol, dv_);
            tmem_ld_wait();
#pragma unroll
            for (int c8 = 0; c8 < 2; ++c8) {
              const uint32_t bw2[2] = {c8 ? bqv[sc].z : bqv[sc].x, c8 ? bqv[sc].w : bqv[sc].y};
              uint32_t ppk[4], dpk[4];
              float dsv[8];
#pragma unroll
              for (int e2 = 0; e2 < 4; ++e2) {
                float pv2[2];
                float* ds2 = dsv + 2 * e2;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                  const int cc = 8 * c8 + 2 * e2 + t;
                  // bias / 2 = ts_w[bucket] / 2 + pos_w[N-1+j-i] / 2 ; bucket 255 (masked) -> -15000
                  const float hb = tsw_s[(bw2[e2 >> 1] >> (16 * (e2 & 1) + 8 * t)) & 0xffu] + pos_r[c16 + cc];
                  const float hx = fmaf(__uint_as_float(sv[cc]), 0.5f, hb);
                  const float th = tanh_approx(hx);
                  // unscaled: P' = SiLU(x) = N P ; dS' = dP * 2 SiLU'(x) = 2N dS (factors applied at the end)
                  pv2[t] = fmaf(hx, th, hx);
                  const float u1 = fmaf(-th, th, 1.0f);
                  const float w2 = fmaf(hx, u1, 1.0f + th);
                  ds2[t] = __uint_as_float(dv_[cc]) * w2;
                }
                ppk[e2] = pack_bf16x2(pv2[0], pv2[1]);
                dpk[e2] = pack_bf16x2(ds2[0], ds2[1]);
              }
              // P^T: queries c16 + 8 c8 .. +7 of this key row -> TMEM columns 32 g + 8 sc + 4 c8 .. +3
              tmem_st4(tmem + lane_base + 32 * g + 8 * sc + 4 * c8, ppk);
              // dS^T: 16-byte chunk (4 g + 2 sc + c8) of this key row, 128-byte swizzle, block hf
              *reinterpret_cast<uint4*>(dsT_row + hf * AT_TILE_BYTES + ((4 * g + 2 * sc + c8) ^ (r & 7)) * 16) =
                  make_uint4(dpk[0], dpk[1], dpk[2], dpk[3]);
              if (HAS_BIAS) {
                // d pos_w: the warp's 32 x 16 block is summed along its diagonals: each packed pair is
                // rotated to the lane that owns its diagonal (r - c = lane: am / bm, lane - 32: aw / bw)
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                  const int src = lane + 8 * c8 + 2 * k2;
                  const uint32_t got = __shfl_sync(0xffffffffu, dpk[k2], src);
                  if (src < 32) add_bf16_pair(am, bm, got); else add_bf16_pair(aw, bw, got);
                }
                // d ts_w: branch-free run-length walk
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const uint32_t bk = (bw2[e >> 2] >> (8 * (e & 3))) & 0xffu;
                  const bool chg = bk != run_bk;
                  red_add_f32_pred(d_ts_mine + run_bk, run_acc * half_inv_n, chg && run_bk != 255u && run_acc != 0.f);
                  run_acc = (chg ? 0.f : run_acc) + dsv[e];
                  run_bk = bk;
                }
              }
            }
            if (HAS_BIAS) {
              // diagonal totals: b* of lane L+1 belong to the diagonals of lane L
              const float tbm = __shfl_sync(0xffffffffu, bm, lane + 1);
              const float tbw = __shfl_sync(0xffffffffu, bw, lane + 1);
              const float tot_m = am + (lane < 31 ? tbm : 0.f);
              const float tot_w = aw + tbw + (lane == 31 ? tbm : 0.f);
              // diagonal r - c = rel  ->  pos_w index N-1 + (j0 + r) - (i0 + c)
              const int64_t idx_m = p.N - 1 + j0 - i0 + 32 * wq - c16 + lane;
              red_add_f32_pred(d_pos_mine + idx_m, tot_m * half_inv_n, tot_m != 0.f && idx_m >= 0 && idx_m < 2 * p.N - 1);
              const int64_t idx_w = idx_m - 32;
              red_add_f32_pred(d_pos_mine + idx_w, tot_w * half_inv_n, tot_w != 0.f && idx_w >= 0 && idx_w < 2 * p.N - 1);
            }
          }
          if (HAS_BIAS)
            red_add_f32_pred(d_ts_mine + run_bk, run_acc * half_inv_n, run_bk != 255u && run_acc != 0.f);
          tmem_st_wait();
          tc_fence_before();
          fence_proxy_async_smem();                  // st.shared -> visible to the MMA / bulk reduce (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_half_done + 8 * hf);
        }
        // ---- dQ of this unit: thread = query row r of tile qt, columns 32 g .. +31 ----
        mbar_wait_g(bar_dq_full, gu & 1);
        tc_fence_after();
        {
          uint32_t qv[32];
          tmem_ld32(tmem + lane_base + 64 + 32 * g, qv);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_dq_read);
          const int i = i0 + r;
          if (i < n) {
            float* part = p.dq_accum + (off0 + i) * ((int64_t) p.H * AT_D) + h * AT_D + 32 * g;
            if (u == 1) {                              // partial of query tile 1 (key tile 0)
#pragma unroll
              for (int v4 = 0; v4 < 8; ++v4)
                *reinterpret_cast<float4*>(part + 4 * v4) =
                    make_float4(__uint_as_float(qv[4 * v4]), __uint_as_float(qv[4 * v4 + 1]),
                                __uint_as_float(qv[4 * v4 + 2]), __uint_as_float(qv[4 * v4 + 3]));
            } else {
              if (u == 2) {
#pragma unroll
                for (int v4 = 0; v4 < 8; ++v4) {
                  const float4 a = *reinterpret_cast<const float4*>(part + 4 * v4);
                  qv[4 * v4] = __float_as_uint(__uint_as_float(qv[4 * v4]) + a.x);
                  qv[4 * v4 + 1] = __float_as_uint(__uint_as_float(qv[4 * v4 + 1]) + a.y);
                  qv[4 * v4 + 2] = __float_as_uint(__uint_as_float(qv[4 * v4 + 2]) + a.z);
                  qv[4 * v4 + 3] = __float_as_uint(__uint_as_float(qv[4 * v4 + 3]) + a.w);
                }
              }
              __nv_bfloat16* dst = p.dq + (off0 + i) * p.lddq + h * AT_D + 32 * g;
#pragma unroll
              for (int v4 = 0; v4 < 4; ++v4) {
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(qv[v4 * 8 + 0]) * half_inv_n, __uint_as_float(qv[v4 * 8 + 1]) * half_inv_n);
                o.y = pack_bf16x2(__uint_as_float(qv[v4 * 8 + 2]) * half_inv_n, __uint_as_float(qv[v4 * 8 + 3]) * half_inv_n);
                o.z = pack_bf16x2(__uint_as_float(qv[v4 * 8 + 4]) * half_inv_n, __uint_as_float(qv[v4 * 8 + 5]) * half_inv_n);
                o.w = pack_bf16x2(__uint_as_float(qv[v4 * 8 + 6]) * half_inv_n, __uint_as_float(qv[v4 * 8 + 7]) * half_inv_n);
                *reinterpret_cast<uint4*>(dst + v4 * 8) = o;
              }
            }
          }
        }
        // ---- dV / dK of a finished key tile: thread = key row, columns 32 g .. +31 of each ----
        if (last_of_kt) {
          mbar_wait_g(bar_dkv_full, gk & 1);
          ++gk;
          tc_fence_after();
#pragma unroll
          for (int which = 0; which < 2; ++which) {
            uint32_t ov[32];
            tmem_ld32(tmem + lane_base + 128 + 64 * which + 32 * g, ov);
            tmem_ld_wait();
            const float sc = which == 0 ? inv_n : half_inv_n;   // dV = P'^T dO / N ; dK = dS'^T Q / (2N)
            if (jk < n) {
              __nv_bfloat16* dst = (which == 0 ? p.dv + (off0 + jk) * p.lddv : p.dk + (off0 + jk) * p.lddk) +
                                   h * AT_D + 32 * g;
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
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_dkv_read);
        }
      }
      SH_STAMP(tid == 64 && k_item < 7, 2 + 2 * k_item);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------
// Self-contained: any head count (the long-sequence forward wants an even one).
bool hstu_attn_short_usable(const grb_hstu_attn_args* a, bool bwd) {
  if (a->short_schedule == nullptr || a->max_len > 256) return false;
  if (a->dtype != GRB_BF16 || a->dqk != AT_D || a->dv != AT_D) return false;
  if (a->bucket_cache == nullptr || !a->bucket_cache_masked || a->bucket_cache_max_len != a->max_len) return false;
  if (a->timestamps && a->num_buckets > 254) return false;
  if (a->T >= (1ll << 31) || a->T == 0) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(a->q) || !al16(a->k) || !al16(a->v) || !al16(a->bucket_cache)) return false;
  if ((a->ldq * 2) % 16 || (a->ldk * 2) % 16 || (a->ldv * 2) % 16) return false;
  if (!bwd) {
    if (!al16(a->out) || (a->ldo * 2) % 16) return false;
  } else {
    if (!al16(a->dout) || !al16(a->dq) || !al16(a->dk) || !al16(a->dv_grad) || !al16(a->dq_accum)) return false;
    if ((a->lddo * 2) % 16 || (a->lddq * 2) % 16 || (a->lddk * 2) % 16 || (a->lddv * 2) % 16) return false;
    if (a->max_len > 128 && a->dq_accum == nullptr) return false;
  }
  return true;
}

// GRB_SHORT_TIMELINE=1: dump the stamps of the probed CTAs to stderr after the launch (synchronises)
static long long* timeline_buffer() {
  static long long* buf = nullptr;
  static const bool on = [] { const char* e = std::getenv("GRB_SHORT_TIMELINE"); return e && e[0] == '1'; }();
  if (!on) return nullptr;      // (read once per process, not on every launch)
  if (!buf) cudaMalloc(&buf, SH_TL_CTAS * SH_TL_SLOTS * sizeof(long long));
  cudaMemset(buf, 0, SH_TL_CTAS * SH_TL_SLOTS * sizeof(long long));
  return buf;
}
static void timeline_dump(const char* what, long long* buf, cudaStream_t st) {
  if (!buf) return;
  cudaStreamSynchronize(st);
  static long long h[SH_TL_CTAS * SH_TL_SLOTS];
  cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost);
  long long t0 = 0;
  for (int c = 0; c < SH_TL_CTAS; ++c) if (h[c * SH_TL_SLOTS] && (!t0 || h[c * SH_TL_SLOTS] < t0)) t0 = h[c * SH_TL_SLOTS];
  for (int c = 0; c < SH_TL_CTAS; ++c) {
    const long long* r = h + c * SH_TL_SLOTS;
    if (!r[0]) continue;
    fprintf(stderr, "%s cta %2d nu %lld:", what, c, r[15]);
    for (int s2 = 0; s2 < 15; ++s2) fprintf(stderr, " %d:%lld", s2, r[s2] ? r[s2] - t0 : -1);
    fprintf(stderr, "\n");
  }
}

// persistent grid: two CTAs per SM (the kernels' resource limits), never more than there are items
static unsigned short_grid(const grb_hstu_attn_args* a) {
  const int64_t items = a->B * a->H;
  const int64_t slots = 2 * (int64_t) num_sms();
  return (unsigned) (items < slots ? items : slots);
}

int hstu_attn_short_fwd(const grb_hstu_attn_args* a, cudaStream_t st) {
  if (a->B == 0 || a->max_len == 0) return GRB_OK;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  const uint64_t W = (uint64_t) a->H * AT_D;
  if ((rc = make_tmap_bf16_2d(&tmQ, a->q, a->T, W, a->ldq, AT_BM)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, a->k, a->T, W, a->ldk, AT_BN)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, a->v, a->T, W, a->ldv, AT_BN)) != GRB_OK) return rc;
  ShortFwdParams p{};
  p.N = a->N; p.B = (int) a->B; p.H = a->H; p.index_bits = a->index_bits;
  const int NT = (int) ceil_div(a->max_len, AT_BM);
  p.tps = NT * (NT + 1) / 2;
  p.offsets = a->offsets;
  p.bcache = a->bucket_cache;
  p.sched = a->short_schedule;
  p.ts_w = a->timestamps ? a->ts_w : nullptr; p.pos_w = a->timestamps ? a->pos_w : nullptr;
  p.nb = a->timestamps ? a->num_buckets : 0;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out); p.ldo = a->ldo;
  p.T = a->T; p.zero_tail = a->zero_tail_rows;
  GRB_REQUIRE(!p.zero_tail || ((reinterpret_cast<uintptr_t>(a->out) | (uintptr_t) (a->ldo * 2)) & 15) == 0,
              GRB_ERR_INVALID_ARG, "hstu_attn_short_fwd: zero_tail_rows needs 16-byte aligned rows");
  const size_t smem = SfSmem::total + 1024;
  auto kern = hstu_attn_short_fwd_kernel;
  GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   (int) cudaSharedmemCarveoutMaxShared));   // two CTAs per SM
  p.tl = timeline_buffer();
  kern<<<short_grid(a), SH_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
  timeline_dump("fwd", p.tl, st);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int hstu_attn_short_bwd(const grb_hstu_attn_args* a, cudaStream_t st) {
  if (a->B == 0 || a->max_len == 0) return GRB_OK;
  CUtensorMap tmQ, tmK, tmV, tmdO;
  int rc;
  const uint64_t W = (uint64_t) a->H * AT_D;
  if ((rc = make_tmap_bf16_2d(&tmQ, a->q, a->T, W, a->ldq, AT_BM)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, a->k, a->T, W, a->ldk, AT_BN)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, a->v, a->T, W, a->ldv, AT_BN)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmdO, a->dout, a->T, W, a->lddo, AT_BM)) != GRB_OK) return rc;
  ShortBwdParams p{};
  p.N = a->N; p.B = (int) a->B; p.H = a->H; p.index_bits = a->index_bits;
  const int NT = (int) ceil_div(a->max_len, AT_BM);
  p.tps = NT * (NT + 1) / 2;
  p.offsets = a->offsets;
  p.bcache = a->bucket_cache;
  p.sched = a->short_schedule;
  p.ts_w = a->timestamps ? a->ts_w : nullptr; p.pos_w = a->timestamps ? a->pos_w : nullptr;
  p.nb = a->timestamps ? a->num_buckets : 0;
  p.dq = reinterpret_cast<__nv_bfloat16*>(a->dq); p.lddq = a->lddq;
  p.dk = reinterpret_cast<__nv_bfloat16*>(a->dk); p.lddk = a->lddk;
  p.dv = reinterpret_cast<__nv_bfloat16*>(a->dv_grad); p.lddv = a->lddv;
  p.dq_accum = a->dq_accum;
  p.d_ts_w = a->d_ts_w; p.d_pos_w = a->d_pos_w;
  p.d_bias_copies = a->d_bias_copies > 0 ? a->d_bias_copies : 1;
  p.T = a->T; p.zero_tail = a->zero_tail_rows;
  GRB_REQUIRE(!p.zero_tail || ((reinterpret_cast<uintptr_t>(a->dq) | reinterpret_cast<uintptr_t>(a->dk) |
                                reinterpret_cast<uintptr_t>(a->dv_grad) | (uintptr_t) (a->lddq * 2) |
                                (uintptr_t) (a->lddk * 2) | (uintptr_t) (a->lddv * 2)) & 15) == 0,
              GRB_ERR_INVALID_ARG, "hstu_attn_short_bwd: zero_tail_rows needs 16-byte aligned rows");
  const size_t smem = SbSmem::total + 1024;
  p.tl = timeline_buffer();
  const unsigned items = short_grid(a);
  if (a->timestamps) {
    auto kern = hstu_attn_short_bwd_kernel<true>;
    GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int) cudaSharedmemCarveoutMaxShared));
    kern<<<items, SH_THREADS, smem, st>>>(tmQ, tmK, tmV, tmdO, p);
    timeline_dump("bwd", p.tl, st);
  } else {
    auto kern = hstu_attn_short_bwd_kernel<false>;
    GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int) cudaSharedmemCarveoutMaxShared));
    kern<<<items, SH_THREADS, smem, st>>>(tmQ, tmK, tmV, tmdO, p);
  }
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}  // namespace grb

using namespace grb;

extern "C" {

int grb_hstu_short_schedule(const void* offsets, int index_bits, int64_t B, int64_t N, int32_t* schedule,
                            grb_stream_t stream) {
  GRB_REQUIRE(index_bits == 32 || index_bits == 64, GRB_ERR_INVALID_ARG,
              "short_schedule: index_bits must be 32 or 64");
  GRB_REQUIRE(offsets && schedule && B >= 0 && N > 0, GRB_ERR_INVALID_ARG, "short_schedule: bad arguments");
  GRB_REQUIRE(B <= 65535, GRB_ERR_UNSUPPORTED, "short_schedule: B <= 65535");
  hstu_short_schedule_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(offsets, index_bits, (int) B, N,
                                                                                     schedule);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}
