// mips_sm100.cu — retrieval score GEMM on Blackwell tensor cores (bf16 queries x bf16 items,
// fp32 accumulate in TMEM), fused with the sample-store / threshold-filter epilogue of
// mips_topk.cu, so the (B, X) logits of the reference (indexing/top_k.py:62) never reach HBM.
//
// Persistent kernel, one CTA per SM.  Work item = (chunk of consecutive item tiles, query block
// of 128 rows), ordered chunk-major so the CTAs that run concurrently read the SAME item tiles
// for different query blocks: the item table streams from HBM once and is re-used out of L2.
//
//   TMA warp  : NQB query blocks [128 x D] each (K-major, 64-column 128-byte-swizzled slabs) when
//               the query-block group changes; item slabs [128 items x 64] through a TMA ring.
//   MMA warp  : per slab, tcgen05.mma M128 N128 K16 x 4 for EACH of the NQB query blocks, so one
//               L2 -> smem item slab feeds NQB x 128 queries (L2 traffic, not the tensor pipe, is
//               the first limit at B = 4096: 32 query blocks x 5.12 GB with one block per CTA).
//               2 TMEM accumulator sets of NQB x 128 columns.
//   epilogue  : 2 x NQB warpgroups (two 64-column halves per query block), thread = query row:
//               tcgen05.ld the scores of the tile and either store them (sample pass) or append
//               (score, index) >= tau[row] to the row's candidate list (one atomic per tile half).
//               One query block for the whole launch (B <= 128, instantiation <1, SMALL>, the two
//               passes of mips_small.cu): 16 epilogue warps (32-column quarters), group maxima of
//               the sample tiles (GMAX) or hits appended to the thread's private sub-list (PRIVATE).
#include "common.cuh"
#include "mips_epilogue.cuh"
#include "sm100_ptx.cuh"
#include "hstu_attn_sm100.cuh"

namespace grb {

using namespace ptx;

constexpr int MS_SLAB = 128 * 64 * 2;   // 16 KiB: 128 rows x 64 bf16
template <int NQB> struct MsCfg {
  // TMA warp, MMA warp, and per query block TWO epilogue warpgroups (64 of the 128 score columns
  // each): with one, a tile's epilogue (TMEM loads, reject tree, the atomic's round trip) takes
  // longer than its 1024 MMA cycles as soon as candidates are dense (small shards, early phases);
  // thread counts: ms_threads<NQB, SMALL>() below
  static constexpr int stages = NQB == 1 ? 8 : 6;
  static constexpr int acc = 4 / NQB;                   // accumulator sets (NQB x 128 columns each)
  static constexpr int q = 0;                           // NQB x 4 slabs
  static constexpr int ring = NQB * 4 * MS_SLAB;
  static constexpr int bars = ring + stages * MS_SLAB;
  static constexpr int total = bars + 512;
};

struct MipsSmParams {
  int64_t B, X;
  int D;                 // multiple of 64, <= 256
  int64_t n_launch_tiles;
  int64_t chunk;         // launch tiles per work item
  int64_t n_chunks;
  int64_t n_qb;
  ScoreEpi epi;
};

// SMALL = the small-batch plan's two passes (GMAX, PRIVATE: one query block, NQB = 1); the other
// instantiations carry STORE and FILTER only, so neither pays for the other's code
// epilogue warps: 8 per query block (two column halves x four TMEM lane quadrants); the small-batch
// passes run 16 (four 32-column quarters): their epilogue is a chain of dependent steps per tile
// (TMEM load, max tree, compare walk) that two warps per scheduler cannot overlap
template <int NQB, bool SMALL> constexpr int ms_epi_warps() { return SMALL ? 16 : 8 * NQB; }
template <int NQB, bool SMALL> constexpr int ms_threads() { return 64 + 32 * ms_epi_warps<NQB, SMALL>(); }

template <int NQB, bool SMALL>
__global__ void __launch_bounds__(ms_threads<NQB, SMALL>(), 1) mips_scores_sm100_kernel(
    const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmI,
    MipsSmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  using MsSmem = MsCfg<NQB>;
  constexpr int MS_STAGES = MsSmem::stages, MS_ACC = MsSmem::acc;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MsSmem::bars);
  const uint32_t bar_full = smem_u32(bars);                        // [STAGES]
  const uint32_t bar_empty = smem_u32(bars + MS_STAGES);           // [STAGES]
  const uint32_t bar_acc_full = smem_u32(bars + 2 * MS_STAGES);    // [ACC]
  const uint32_t bar_acc_empty = smem_u32(bars + 2 * MS_STAGES + MS_ACC);  // [ACC]
  const uint32_t bar_q_full = smem_u32(bars + 2 * MS_STAGES + 2 * MS_ACC);
  const uint32_t bar_q_empty = bar_q_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MS_STAGES + 2 * MS_ACC + 2);

  if (tid == 0) {
    for (int s = 0; s < MS_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < MS_ACC; ++s) { mbar_init(bar_acc_full + 8 * s, 1); mbar_init(bar_acc_empty + 8 * s, ms_epi_warps<NQB, SMALL>()); }
    mbar_init(bar_q_full, 1);
    mbar_init(bar_q_empty, 1);
    fence_barrier_init();
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmI);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int kslabs = p.D / 64;
  const int64_t n_items = p.n_chunks * p.n_qb;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;          // slab counter (ring position)
      int64_t cur_qb = -1;
      uint32_t q_loads = 0;
      for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int64_t chunk = w / p.n_qb, qb = w % p.n_qb;
        if (qb != cur_qb) {
          mbar_wait_parked(bar_q_empty, (q_loads & 1) ^ 1);   // MMAs of the previous Q are done
          mbar_arrive_expect_tx(bar_q_full, NQB * kslabs * MS_SLAB);
          for (int qq = 0; qq < NQB; ++qq)
            for (int kc = 0; kc < kslabs; ++kc)
              tma_load_2d(smem_u32(smem + MsSmem::q + (qq * 4 + kc) * MS_SLAB), &tmQ, kc * 64,
                          (int) ((qb * NQB + qq) * 128), bar_q_full);
          cur_qb = qb;
          ++q_loads;
        }
        const int64_t u0 = chunk * p.chunk;
        const int64_t u1 = (u0 + p.chunk < p.n_launch_tiles) ? u0 + p.chunk : p.n_launch_tiles;
        for (int64_t u = u0; u < u1; ++u) {
          const int64_t row0 = epi_item_tile(p.epi, u) * MIPS_TILE_N;
          for (int kc = 0; kc < kslabs; ++kc, ++it) {
            const uint32_t st = it % MS_STAGES;
            mbar_wait_parked(bar_empty + 8 * st, ((it / MS_STAGES) & 1) ^ 1);
            mbar_arrive_expect_tx(bar_full + 8 * st, MS_SLAB);
            tma_load_2d(smem_u32(smem + MsSmem::ring + st * MS_SLAB), &tmI, kc * 64, (int) row0,
                        bar_full + 8 * st);
          }
        }
      }
    }
  } else if (warp == 1) {
    {   // whole warp, uniform control flow; one elected lane issues (umma_*_warp): an
        // `if (lane == 0)` region costs ~85 cycles of issue per MMA, more than its 64 cycles of work
      const uint32_t idesc = make_idesc_bf16(128, MIPS_TILE_N, false, false);
      const uint64_t q_desc = make_smem_desc_sw128(smem_u32(smem + MsSmem::q), 0, 1024);
      const uint64_t ring_desc = make_smem_desc_sw128(smem_u32(smem + MsSmem::ring), 0, 1024);
      auto adv = [](uint64_t d, uint32_t bytes) { return d + (uint64_t) (bytes >> 4); };
      uint32_t it = 0, tile = 0, q_loads = 0;
      int64_t cur_qb = -1;
      for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int64_t chunk = w / p.n_qb, qb = w % p.n_qb;
        if (qb != cur_qb) {
          if (cur_qb >= 0) umma_commit_warp(bar_q_empty);     // all MMAs reading the old Q are issued
          mbar_wait_parked(bar_q_full, q_loads & 1);
          tc_fence_after();
          cur_qb = qb;
          ++q_loads;
        }
        const int64_t u0 = chunk * p.chunk;
        const int64_t u1 = (u0 + p.chunk < p.n_launch_tiles) ? u0 + p.chunk : p.n_launch_tiles;
        for (int64_t u = u0; u < u1; ++u, ++tile) {
          const uint32_t ab = tile % MS_ACC;
          mbar_wait_parked(bar_acc_empty + 8 * ab, ((tile / MS_ACC) & 1) ^ 1);
          tc_fence_after();
          for (int kc = 0; kc < kslabs; ++kc, ++it) {
            const uint32_t st = it % MS_STAGES;
            mbar_wait_parked(bar_full + 8 * st, (it / MS_STAGES) & 1);
            tc_fence_after();
            const uint64_t i_desc = adv(ring_desc, st * MS_SLAB);
#pragma unroll
            for (int qq = 0; qq < NQB; ++qq) {
              const uint64_t qd = adv(q_desc, (qq * 4 + kc) * MS_SLAB);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_ss_warp(tmem + (ab * NQB + qq) * MIPS_TILE_N, adv(qd, ks * 32), adv(i_desc, ks * 32),
                             idesc, (kc > 0) || (ks > 0));
            }
            umma_commit_warp(bar_empty + 8 * st);
          }
          umma_commit_warp(bar_acc_full + 8 * ab);
        }
      }
    }
  } else {
    // epilogue: warps 2 .. 2+8*NQB, TMEM lane quadrant = warp % 4; four consecutive warps form a
    // group (query block qq, column half)
    const int r = ((warp & 3) << 5) | lane;
    const int qq = (warp - 2) >> 3;
    const int half = ((warp - 2) >> 2) & 1;
    const uint32_t lane_base = (uint32_t) ((warp & 3) * 32) << 16;
    uint32_t tile = 0;
    if constexpr (SMALL) {
      // One query block for the whole launch: this thread keeps query row r and the 32-column
      // quarter `cq` of every tile its CTA scores.  GMAX writes group maxima of the sample tiles;
      // PRIVATE appends hits to the thread's own sub-list — the slot counter is a register, there
      // is no atomic and no second pass over the accumulator.
      const int cq = (warp - 2) >> 2;
      const int64_t row = r;
      const bool row_ok = row < p.B;
      const bool priv = p.epi.mode == MIPS_EPI_PRIVATE;
      const float tau = (priv && row_ok) ? p.epi.tau[row] : INFINITY;
      const int sub = MIPS_SUB_PER_CTA * (int) blockIdx.x + cq;
      const int cap = p.epi.sub_cap;
      // a sub-list has cap + MIPS_SUB_SPARE entries: the write position is clamped to cap once per
      // tile, not per hit, so a full list takes at most 32 stray writes into its spare tail
      const uint64_t list = reinterpret_cast<uint64_t>(
          p.epi.sub_cand + (row * p.epi.n_sub + sub) * (int64_t) (cap + MIPS_SUB_SPARE));
      const int G = p.epi.group;
      int slot = 0;
      for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int64_t u0 = w * p.chunk;       // n_qb == 1: work item = chunk
        const int64_t u1 = (u0 + p.chunk < p.n_launch_tiles) ? u0 + p.chunk : p.n_launch_tiles;
        for (int64_t u = u0; u < u1; ++u, ++tile) {
          const uint32_t ab = tile % MS_ACC;
          mbar_wait(bar_acc_full + 8 * ab, (tile / MS_ACC) & 1);
          tc_fence_after();
          const int64_t item0 = epi_item_tile(p.epi, u) * MIPS_TILE_N + cq * 32;
          uint32_t sv[32];
          tmem_ld32(tmem + lane_base + ab * MIPS_TILE_N + cq * 32, sv);
          tmem_ld_wait();
          // the accumulator is in registers: the MMA warp may refill it while this warp walks
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * ab);
          if (item0 + 32 > p.X) {   // the last item tile only: columns past X never win
            const int lim = (int) (p.X - item0);
#pragma unroll
            for (int c = 0; c < 32; ++c) sv[c] = c < lim ? sv[c] : 0xff800000u;
          }
          float m4[8], m8[4];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            m4[k] = fmaxf(fmaxf(__uint_as_float(sv[4 * k]), __uint_as_float(sv[4 * k + 1])),
                          fmaxf(__uint_as_float(sv[4 * k + 2]), __uint_as_float(sv[4 * k + 3])));
#pragma unroll
          for (int k = 0; k < 4; ++k) m8[k] = fmaxf(m4[2 * k], m4[2 * k + 1]);
          const float m16a = fmaxf(m8[0], m8[1]), m16b = fmaxf(m8[2], m8[3]);
          const float m32 = fmaxf(m16a, m16b);
          if (!priv) {
            if (row_ok) {
              float* const go = p.epi.gmax + row * p.epi.n_groups + u * (MIPS_TILE_N / G) + cq * (32 / G);
              if (G == 8) *reinterpret_cast<float4*>(go) = make_float4(m8[0], m8[1], m8[2], m8[3]);
              else if (G == 16) *reinterpret_cast<float2*>(go) = make_float2(m16a, m16b);
              else go[0] = m32;
            }
          } else if (__any_sync(0xffffffffu, m32 >= tau)) {   // tau = +inf on rows past B
            // the two reject levels (32 columns, 4 columns) are warp votes: real branches.  Left to
            // itself the compiler predicates the whole walk and issues every column's store sequence
            // on every tile (ncu: 790 instructions per warp and tile); a walked column costs 8
            // instructions, and with ~0.4 hits per (row, tile) two thirds of the groups are skipped
            const int32_t ibase = (int32_t) item0;
            // entry `slot` of this tile's view of the list: the real one while slot <= cap
            uint64_t wbase = list + (int64_t) (min(slot, cap) - slot) * (int64_t) sizeof(MipsCand);
            asm volatile("" : "+l"(wbase));   // kept as one base: a hit is multiply-add, store, increment
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (__any_sync(0xffffffffu, m4[k] >= tau)) {
#pragma unroll
                for (int c = 4 * k; c < 4 * k + 4; ++c) {
                  if (__uint_as_float(sv[c]) >= tau) {
                    asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(wbase + (uint64_t) (uint32_t) slot * sizeof(MipsCand)),
                                 "r"(sv[c]), "r"(ibase + c) : "memory");
                    ++slot;
                  }
                }
              }
            }
          }
        }
      }
      if (priv && row_ok) p.epi.sub_counts[row * p.epi.n_sub + sub] = slot;
    } else
    for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int64_t chunk = w / p.n_qb, qb = w % p.n_qb;
      const int64_t row = (qb * NQB + qq) * 128 + r;
      const bool row_ok = row < p.B;
      const float tau = (p.epi.mode == MIPS_EPI_FILTER && row_ok) ? p.epi.tau[row] : INFINITY;
      const int64_t u0 = chunk * p.chunk;
      const int64_t u1 = (u0 + p.chunk < p.n_launch_tiles) ? u0 + p.chunk : p.n_launch_tiles;
      for (int64_t u = u0; u < u1; ++u, ++tile) {
        const uint32_t ab = tile % MS_ACC;
        mbar_wait(bar_acc_full + 8 * ab, (tile / MS_ACC) & 1);
        tc_fence_after();
        const int64_t item0 = epi_item_tile(p.epi, u) * MIPS_TILE_N;
        // only the last item tile can run past X: everywhere else the per-score bound checks
        // (64-bit compares) are skipped
        const bool full = item0 + MIPS_TILE_N <= p.X;
        const uint32_t acc_addr = tmem + lane_base + (ab * NQB + qq) * MIPS_TILE_N;
        if (p.epi.mode == MIPS_EPI_STORE) {
#pragma unroll 1
          for (int c32 = 2 * half; c32 < 2 * half + 2; ++c32) {
            uint32_t sv[32];
            tmem_ld32(acc_addr + c32 * 32, sv);
            tmem_ld_wait();
            if (!row_ok) continue;
            float* o = p.epi.out + row * p.epi.Xs + u * MIPS_TILE_N + c32 * 32;
            if (full) {
#pragma unroll
              for (int v4 = 0; v4 < 8; ++v4)
                *reinterpret_cast<uint4*>(o + v4 * 4) =
                    make_uint4(sv[v4 * 4 + 0], sv[v4 * 4 + 1], sv[v4 * 4 + 2], sv[v4 * 4 + 3]);
            } else {
#pragma unroll
              for (int v4 = 0; v4 < 8; ++v4) {
                float4 f;
                const int64_t ib = item0 + c32 * 32 + v4 * 4;
                f.x = (ib + 0 < p.X) ? __uint_as_float(sv[v4 * 4 + 0]) : -INFINITY;
                f.y = (ib + 1 < p.X) ? __uint_as_float(sv[v4 * 4 + 1]) : -INFINITY;
                f.z = (ib + 2 < p.X) ? __uint_as_float(sv[v4 * 4 + 2]) : -INFINITY;
                f.w = (ib + 3 < p.X) ? __uint_as_float(sv[v4 * 4 + 3]) : -INFINITY;
                *reinterpret_cast<float4*>(o + v4 * 4) = f;
              }
            }
          }
        } else {
          // FILTER, two passes over the 128 scores of this row so that a tile costs ONE atomic (its
          // ~700-cycle round trip, paid per 8-score sub-group, was what held the dense early phases
          // back).  Pass 1: cheap reject at two levels (max of 32, max of 8) and a hit bitmask per
          // 32-score chunk; the compare walk of a sub-group is executed by the whole warp as soon
          // as ONE of its 32 rows needs it, so it is kept to 8 scores.
          // (the chunk loops stay rolled — unrolled, the code no longer fits the instruction cache
          //  and the kernel runs 2x slower — so the hit masks live in scalars, not in an array)
          uint32_t h0 = 0u, h1 = 0u;
#pragma unroll 1
          for (int c32 = 2 * half; c32 < 2 * half + 2; ++c32) {
            uint32_t sv[32];
            tmem_ld32(acc_addr + c32 * 32, sv);
            tmem_ld_wait();
            float m8[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float m = __uint_as_float(sv[8 * k]);
#pragma unroll
              for (int c = 1; c < 8; ++c) m = fmaxf(m, __uint_as_float(sv[8 * k + c]));
              m8[k] = m;
            }
            uint32_t h = 0u;
            if (row_ok && fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])) >= tau) {
              const int lim = full ? 32 : (int) (p.X - (item0 + c32 * 32));   // valid scores in this chunk
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (m8[k] >= tau) {
#pragma unroll
                  for (int c = 8 * k; c < 8 * k + 8; ++c)
                    h |= (__uint_as_float(sv[c]) >= tau && c < lim) ? (1u << c) : 0u;
                }
              }
            }
            h0 = (c32 & 1) == 0 ? h : h0;
            h1 = (c32 & 1) == 1 ? h : h1;
          }
          const int cnt = __popc(h0) + __popc(h1);
          if (__any_sync(0xffffffffu, cnt > 0)) {
            int slot = cnt ? atomicAdd(p.epi.counts + row, cnt) : 0;   // one atomic per (row, tile half)
            // pass 2: only chunks / sub-groups with hits are read again and walked
#pragma unroll 1
            for (int c32 = 2 * half; c32 < 2 * half + 2; ++c32) {
              const uint32_t h = (c32 & 1) == 0 ? h0 : h1;
              if (!__any_sync(0xffffffffu, h != 0u)) continue;
              uint32_t sv[32];
              tmem_ld32(acc_addr + c32 * 32, sv);
              tmem_ld_wait();
              const int32_t ibase = (int32_t) (item0 + c32 * 32);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if ((h >> (8 * k)) & 0xffu) {
#pragma unroll
                  for (int c = 8 * k; c < 8 * k + 8; ++c) {
                    if ((h >> c) & 1u) {
                      if (slot < p.epi.cap) {
                        p.epi.cscores[row * p.epi.cap + slot] = __uint_as_float(sv[c]);
                        p.epi.cidx[row * p.epi.cap + slot] = ibase + c;
                      }
                      ++slot;
                    }
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * ab);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

bool mips_sm100_supported(const grb_mips_topk_args* a) {
  if (a->dtype != GRB_BF16) return false;
  if (a->D % 64 != 0 || a->D > 256 || a->D <= 0) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(a->queries) || !al16(a->items)) return false;
  if ((a->ldq * 2) % 16 || (a->ldi * 2) % 16) return false;
  return true;
}

// how a launch over n_launch_tiles is cut into work items: ~8 per CTA for balance, at most 128 tiles
// (8 MiB of items) each; the small-batch plan sizes its private sub-lists from the same grid
void mips_sm100_work_split(int64_t n_launch_tiles, int64_t n_qb, int64_t* chunk_out, int64_t* n_chunks_out,
                           unsigned* grid_out) {
  const int sms = num_sms();
  int64_t chunk = ceil_div(n_launch_tiles * n_qb, (int64_t) sms * 8);
  if (chunk < 1) chunk = 1;
  if (chunk > 128) chunk = 128;
  const int64_t n_chunks = ceil_div(n_launch_tiles, chunk);
  const int64_t n_items = n_chunks * n_qb;
  *chunk_out = chunk;
  *n_chunks_out = n_chunks;
  *grid_out = (unsigned) (n_items < sms ? n_items : sms);
}

template <int NQB, bool SMALL>
static int launch_ms(const CUtensorMap& tmQ, const CUtensorMap& tmI, const MipsSmParams& p, unsigned grid,
                     cudaStream_t st) {
  const size_t smem = MsCfg<NQB>::total + 1024;
  GRB_CUDA_OK(cudaFuncSetAttribute(mips_scores_sm100_kernel<NQB, SMALL>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  mips_scores_sm100_kernel<NQB, SMALL><<<grid, ms_threads<NQB, SMALL>(), smem, st>>>(tmQ, tmI, p);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int mips_scores_sm100(const grb_mips_topk_args* a, const ScoreEpi& epi, int64_t n_launch_tiles,
                      cudaStream_t st) {
  CUtensorMap tmQ, tmI;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmQ, a->queries, a->B, a->D, a->ldq, 128)) != GRB_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tmI, a->items, a->X, a->D, a->ldi, 128)) != GRB_OK) return rc;
  MipsSmParams p{};
  p.B = a->B; p.X = a->X; p.D = (int) a->D;
  p.n_launch_tiles = n_launch_tiles;
  const int nqb = a->B > 128 ? 2 : 1;     // query blocks per CTA
  p.n_qb = ceil_div(a->B, 128 * nqb);     // query-block groups
  unsigned grid;
  mips_sm100_work_split(n_launch_tiles, p.n_qb, &p.chunk, &p.n_chunks, &grid);
  p.epi = epi;
  if (epi.mode == MIPS_EPI_GMAX || epi.mode == MIPS_EPI_PRIVATE) {
    GRB_REQUIRE(nqb == 1 && p.n_qb == 1, GRB_ERR_INVALID_ARG, "mips_topk: small-batch pass with B=%lld",
                (long long) a->B);
    GRB_REQUIRE(epi.mode == MIPS_EPI_GMAX || epi.n_sub == MIPS_SUB_PER_CTA * (int) grid, GRB_ERR_INVALID_ARG,
                "mips_topk: %d sub-lists for a grid of %u", epi.n_sub, grid);
    return launch_ms<1, true>(tmQ, tmI, p, grid, st);
  }
  return nqb == 2 ? launch_ms<2, false>(tmQ, tmI, p, grid, st) : launch_ms<1, false>(tmQ, tmI, p, grid, st);
}

}  // namespace grb
