// hstu_attn_decode.cu — the incremental ("delta_x_offsets" + cache) form of the HSTU attention:
// one query row per sequence against that sequence's cached keys and values.
//
// Reference: /root/reference/src/generative_recommenders_pl/models/sequential_encoders/hstu.py
//   :151-177  the new q / k rows are index_copy_'d into the cached padded (B, N, H*dqk) tensors,
//   :179-204  the WHOLE (B, H, N, N) attention is recomputed from the caches,
//   :397-401  and only the rows delta_x_offsets[0] of the result are kept.
// Row p of the result depends on keys / values 0..p of the same sequence only (causal mask :667), so
// this kernel computes exactly those rows:
//     out[b, h, :] = sum_{j <= p_b} SiLU(q[b,h].k[b,j,h] + pos_w[N-1+j-p_b]
//                                        + ts_w[bucket(|ts[b,p_b+1] - ts[b,j]|)]) / N * v[off_b + j, h, :]
// It is a batched GEMV pair: HBM-bound on the cached K and V rows, (p_b+1) * H * (dqk+dv) elements
// per sequence, each read once.  CTA = (sequence, head); keys are spread over the threads for the
// score pass (one 16-byte-vectorised dot product per thread and key), then over the warps for the
// value pass with the value columns across the lanes (coalesced rows).
#include "common.cuh"

namespace grb {
namespace {

constexpr int DEC_THREADS = 128;

struct DecodeParams {
  int64_t N;
  int H, dqk, dv, nb, index_bits, pos_bits;
  const void* q; int64_t ldq;          // (B, H*dqk) the new query rows
  const void* kc; int64_t ldk;         // padded key cache (B, N, H*dqk), row stride ldk
  const void* v; int64_t ldv;          // jagged values (T, H*dv)
  const void* offsets; const void* pos;
  const int64_t* ts; const float* ts_w; const float* pos_w; const int64_t* thr;
  void* out; int64_t ldo;              // (B, H*dv)
};

template <typename T> __device__ __forceinline__ float ld_f32(const T* p);
template <> __device__ __forceinline__ float ld_f32<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f32<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_f32(T* p, float v);
template <> __device__ __forceinline__ void st_f32<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f32<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16(v);
}

// dot product of a shared-memory fp32 query with one global key row; VEC: 16-byte loads
template <typename T, bool VEC>
__device__ __forceinline__ float dot_row(const float* __restrict__ qs, const T* __restrict__ kr, int d) {
  float acc = 0.f;
  if (VEC) {
    constexpr int PER = 16 / (int) sizeof(T);
    for (int c = 0; c < d; c += PER) {
      const uint4 raw = *reinterpret_cast<const uint4*>(kr + c);
      if (sizeof(T) == 4) {
        const float* f = reinterpret_cast<const float*>(&raw);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc = fmaf(qs[c + e], f[e], acc);
      } else {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h[e]);
          acc = fmaf(qs[c + 2 * e], f.x, acc);
          acc = fmaf(qs[c + 2 * e + 1], f.y, acc);
        }
      }
    }
  } else {
    for (int c = 0; c < d; ++c) acc = fmaf(qs[c], ld_f32(kr + c), acc);
  }
  return acc;
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(DEC_THREADS) hstu_attn_decode_kernel(DecodeParams p) {
  extern __shared__ __align__(16) float dsm[];
  const int b = blockIdx.x, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t off0 = load_index(p.offsets, b, p.index_bits);
  const int64_t n = load_index(p.offsets, b + 1, p.index_bits) - off0;
  const int64_t pos = load_index(p.pos, b, p.pos_bits);
  T* out = reinterpret_cast<T*>(p.out) + (int64_t) b * p.ldo + (int64_t) h * p.dv;
  if (pos < 0 || pos >= p.N || pos >= n) {   // not a row of this sequence: the reference's row would be
    for (int c = tid; c < p.dv; c += DEC_THREADS) st_f32(out + c, 0.f);   // a padded (zero) one
    return;
  }
  const int nk = (int) pos + 1;                       // keys 0 .. pos
  float* qs = dsm;                                    // dqk (padded to 4)
  float* ps = qs + ((p.dqk + 3) & ~3);                // nk probabilities
  float* red = ps + ((nk + 3) & ~3);                  // 4 warps x dv partial sums
  const T* q = reinterpret_cast<const T*>(p.q) + (int64_t) b * p.ldq + (int64_t) h * p.dqk;
  for (int c = tid; c < p.dqk; c += DEC_THREADS) qs[c] = ld_f32(q + c);
  __syncthreads();

  const bool has_bias = p.ts != nullptr;
  int64_t tq = 0;
  if (has_bias) {
    int64_t qi = pos + 1;                             // ext_ts[b, pos + 1]; index N reads ts[b, N-1]
    if (qi >= p.N) qi = p.N - 1;
    tq = p.ts[(int64_t) b * p.N + qi];
  }
  const float inv_n = 1.0f / (float) p.N;
  const T* kbase = reinterpret_cast<const T*>(p.kc) + ((int64_t) b * p.N) * p.ldk + (int64_t) h * p.dqk;
  for (int j = tid; j < nk; j += DEC_THREADS) {
    float x = dot_row<T, VEC>(qs, kbase + (int64_t) j * p.ldk, p.dqk);
    if (has_bias) {
      int64_t d = tq - p.ts[(int64_t) b * p.N + j];
      d = d < 0 ? -d : d;
      x += p.pos_w[p.N - 1 + j - pos] + p.ts_w[bucket_of(p.thr, p.nb, d)];
    }
    ps[j] = silu_f32(x) * inv_n;
  }
  __syncthreads();

  // out[c] = sum_j ps[j] * v[off0 + j, h, c]: warp w takes keys j = w (mod 4), lanes take columns
  const T* vbase = reinterpret_cast<const T*>(p.v) + off0 * p.ldv + (int64_t) h * p.dv;
  for (int c0 = 0; c0 < p.dv; c0 += 32 * 4) {         // <= 4 columns per lane and pass
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = warp; j < nk; j += DEC_THREADS / 32) {
      const float pj = ps[j];
      const T* vr = vbase + (int64_t) j * p.ldv;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = c0 + lane + 32 * e;
        if (c < p.dv) acc[e] = fmaf(pj, ld_f32(vr + c), acc[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) red[warp * 128 + lane + 32 * e] = acc[e];
    __syncthreads();
    if (tid < 128) {
      const int c = c0 + tid;
      if (c < p.dv) st_f32(out + c, (red[tid] + red[128 + tid]) + (red[256 + tid] + red[384 + tid]));
    }
    __syncthreads();
  }
}

// ---- vector path: rows of 16 * G bytes (G a power of two <= 32), G lanes per row ---------------
// A warp reads 32 / G whole rows per load instruction (coalesced 16-byte chunks); the query slice of a
// lane stays in registers; partial dot products meet through shuffles.  The value pass keeps the same
// lane <-> chunk mapping, every lane accumulating its 16 bytes of columns over the keys of its group.
constexpr int DECV_THREADS = 256;
// independent 16-byte loads per lane and loop trip.  Measured on the 8 k-token shape (B200, bf16, d = 64):
// 1 -> 0.74 ms (64 registers, 4 CTAs / SM), 2 -> 0.67 ms (79 registers, 3 CTAs / SM), 4 -> slower
// (128 registers: the lost occupancy outweighs the loads in flight).
constexpr int DEC_UNROLL = 2;

template <typename T> struct Chunk;           // 16 bytes of T as floats
template <> struct Chunk<float> {
  static constexpr int E = 4;
  static __device__ __forceinline__ void load(const float* p, float (&f)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
};
template <> struct Chunk<__nv_bfloat16> {
  static constexpr int E = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 v = __bfloat1622float2(h[e]);
      f[2 * e] = v.x; f[2 * e + 1] = v.y;
    }
  }
};

template <typename T>
__global__ void __launch_bounds__(DECV_THREADS) hstu_attn_decode_vec_kernel(DecodeParams p, int GQ, int GV,
                                                                            int heads_per_cta) {
  extern __shared__ __align__(16) float dsm[];
  constexpr int E = Chunk<T>::E;
  constexpr int NW = DECV_THREADS / 32;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t off0 = load_index(p.offsets, b, p.index_bits);
  const int64_t n = load_index(p.offsets, b + 1, p.index_bits) - off0;
  const int64_t pos = load_index(p.pos, b, p.pos_bits);
  const int h_begin = blockIdx.y * heads_per_cta;
  const int h_end = min(p.H, h_begin + heads_per_cta);
  if (pos < 0 || pos >= p.N || pos >= n) {
    T* out = reinterpret_cast<T*>(p.out) + (int64_t) b * p.ldo;
    for (int c = h_begin * p.dv + tid; c < h_end * p.dv; c += DECV_THREADS) st_f32(out + c, 0.f);
    return;
  }
  const int nk = (int) pos + 1;
  const int nk4 = (nk + 3) & ~3;
  float* ps = dsm;                                    // nk probabilities of the current head
  float* bias_s = ps + nk4;                           // nk bias values: the same for every head
  float* red = bias_s + nk4;                          // NW x dv partial sums
  const bool has_bias = p.ts != nullptr;
  if (has_bias) {
    // pos_w[N-1+j-pos] + ts_w[bucket(|ts[b,pos+1] - ts[b,j]|)], one key per thread, once per CTA
    int64_t* thr_s = reinterpret_cast<int64_t*>(red + NW * p.dv + ((NW * p.dv) & 1));
    const bool thr_in_smem = p.nb <= 256;
    if (thr_in_smem) {
      for (int i = tid; i < p.nb; i += DECV_THREADS) thr_s[i] = p.thr[i];
      __syncthreads();
    }
    const int64_t* thr = thr_in_smem ? thr_s : p.thr;
    int64_t qi = pos + 1;                             // ext_ts[b, pos + 1]; index N reads ts[b, N-1]
    if (qi >= p.N) qi = p.N - 1;
    const int64_t tq = p.ts[(int64_t) b * p.N + qi];
    for (int j = tid; j < nk; j += DECV_THREADS) {
      int64_t d = tq - p.ts[(int64_t) b * p.N + j];
      d = d < 0 ? -d : d;
      bias_s[j] = p.pos_w[p.N - 1 + j - pos] + p.ts_w[bucket_of(thr, p.nb, d)];
    }
  } else {
    for (int j = tid; j < nk; j += DECV_THREADS) bias_s[j] = 0.f;
  }
  __syncthreads();
  const float inv_n = 1.0f / (float) p.N;
  for (int h = h_begin; h < h_end; ++h) {
    {   // scores: GQ lanes per key, 32 / GQ keys per warp and step
      const int sub = lane & (GQ - 1), grp = lane / GQ, per = 32 / GQ;
      float qf[E];
      Chunk<T>::load(reinterpret_cast<const T*>(p.q) + (int64_t) b * p.ldq + (int64_t) h * p.dqk + sub * E, qf);
      const T* kbase = reinterpret_cast<const T*>(p.kc) + ((int64_t) b * p.N) * p.ldk + (int64_t) h * p.dqk + sub * E;
      for (int j0 = warp * per; j0 < nk; j0 += DEC_UNROLL * NW * per) {
        float kf[DEC_UNROLL][E];
#pragma unroll
        for (int u = 0; u < DEC_UNROLL; ++u) {
          const int j = j0 + u * NW * per + grp;
          if (j < nk) Chunk<T>::load(kbase + (int64_t) j * p.ldk, kf[u]);
        }
#pragma unroll
        for (int u = 0; u < DEC_UNROLL; ++u) {
          const int j = j0 + u * NW * per + grp;
          float acc = 0.f;
          if (j < nk) {
#pragma unroll
            for (int e = 0; e < E; ++e) acc = fmaf(qf[e], kf[u][e], acc);
          }
          for (int o = GQ >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          if (sub == 0 && j < nk) ps[j] = silu_f32(acc + bias_s[j]) * inv_n;
        }
      }
    }
    __syncthreads();
    {   // values: GV lanes per row; a lane keeps columns [sub * E, sub * E + E) of the keys of its group
      const int sub = lane & (GV - 1), grp = lane / GV, per = 32 / GV;
      const T* vbase = reinterpret_cast<const T*>(p.v) + off0 * p.ldv + (int64_t) h * p.dv + sub * E;
      float acc[E];
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = 0.f;
      for (int j0 = warp * per + grp; j0 < nk; j0 += DEC_UNROLL * NW * per) {
        float vf[DEC_UNROLL][E];
#pragma unroll
        for (int u = 0; u < DEC_UNROLL; ++u) {
          const int j = j0 + u * NW * per;
          if (j < nk) Chunk<T>::load(vbase + (int64_t) j * p.ldv, vf[u]);
        }
#pragma unroll
        for (int u = 0; u < DEC_UNROLL; ++u) {
          const int j = j0 + u * NW * per;
          if (j < nk) {
            const float pj = ps[j];
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = fmaf(pj, vf[u][e], acc[e]);
          }
        }
      }
      for (int o = GV; o < 32; o <<= 1)               // fold the key groups of the warp
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
      if (grp == 0)
#pragma unroll
        for (int e = 0; e < E; ++e) red[warp * p.dv + sub * E + e] = acc[e];
    }
    __syncthreads();
    T* out = reinterpret_cast<T*>(p.out) + (int64_t) b * p.ldo + (int64_t) h * p.dv;
    for (int c = tid; c < p.dv; c += DECV_THREADS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) t += red[w * p.dv + c];
      st_f32(out + c, t);
    }
    __syncthreads();                                  // ps / red are reused by the next head
  }
}

// ---- short sequences: CTA = sequence, one warp per head, no block barrier after the bias ----------
// With a few hundred keys the block-wide kernel above is a chain of short loops separated by
// __syncthreads (13 per sequence at H = 4) and runs at a third of the HBM rate.  Here the head-independent
// bias is computed once by the whole CTA; after that every warp owns one (sequence, head): scores into
// its private slice of shared memory, __syncwarp, value pass, and the lanes of key group 0 store the
// output row as 16-byte chunks.  Latency is hidden by the number of independent warps per SM.
template <typename T> __device__ __forceinline__ void store_chunk(T* dst, const float* f);
template <> __device__ __forceinline__ void store_chunk<float>(float* dst, const float* f) {
  *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
}
template <> __device__ __forceinline__ void store_chunk<__nv_bfloat16>(__nv_bfloat16* dst, const float* f) {
  uint4 o;
  __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
  o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
  o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(dst) = o;
}

template <typename T>
__global__ void __launch_bounds__(512) hstu_attn_decode_warp_kernel(DecodeParams p, int GQ, int GV) {
  extern __shared__ __align__(16) float dsm[];
  constexpr int E = Chunk<T>::E;
  constexpr int U = 2;                                // independent loads per lane and trip
  const int b = blockIdx.x;
  const int tid = threadIdx.x, h = tid >> 5, lane = tid & 31;   // blockDim.x = 32 * H
  const int64_t off0 = load_index(p.offsets, b, p.index_bits);
  const int64_t n = load_index(p.offsets, b + 1, p.index_bits) - off0;
  const int64_t pos = load_index(p.pos, b, p.pos_bits);
  T* out = reinterpret_cast<T*>(p.out) + (int64_t) b * p.ldo + (int64_t) h * p.dv;
  if (pos < 0 || pos >= p.N || pos >= n) {
    for (int c = lane; c < p.dv; c += 32) st_f32(out + c, 0.f);
    return;
  }
  const int nk = (int) pos + 1;
  const int nk4 = (nk + 3) & ~3;
  float* bias_s = dsm;
  float* ps = dsm + nk4 * (1 + h);                    // this warp's probabilities
  if (p.ts != nullptr) {
    int64_t qi = pos + 1;                             // ext_ts[b, pos + 1]; index N reads ts[b, N-1]
    if (qi >= p.N) qi = p.N - 1;
    const int64_t tq = p.ts[(int64_t) b * p.N + qi];
    for (int j = tid; j < nk; j += blockDim.x) {
      int64_t d = tq - p.ts[(int64_t) b * p.N + j];
      d = d < 0 ? -d : d;
      bias_s[j] = p.pos_w[p.N - 1 + j - pos] + p.ts_w[bucket_of(p.thr, p.nb, d)];
    }
  } else {
    for (int j = tid; j < nk; j += blockDim.x) bias_s[j] = 0.f;
  }
  __syncthreads();
  const float inv_n = 1.0f / (float) p.N;
  {
    const int sub = lane & (GQ - 1), grp = lane / GQ, per = 32 / GQ;
    float qf[E];
    Chunk<T>::load(reinterpret_cast<const T*>(p.q) + (int64_t) b * p.ldq + (int64_t) h * p.dqk + sub * E, qf);
    const T* kbase = reinterpret_cast<const T*>(p.kc) + ((int64_t) b * p.N) * p.ldk + (int64_t) h * p.dqk + sub * E;
    for (int j0 = 0; j0 < nk; j0 += U * per) {
      float kf[U][E];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * per + grp;
        if (j < nk) Chunk<T>::load(kbase + (int64_t) j * p.ldk, kf[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * per + grp;
        float acc = 0.f;
        if (j < nk) {
#pragma unroll
          for (int e = 0; e < E; ++e) acc = fmaf(qf[e], kf[u][e], acc);
        }
        for (int o = GQ >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (sub == 0 && j < nk) ps[j] = silu_f32(acc + bias_s[j]) * inv_n;
      }
    }
  }
  __syncwarp();
  {
    const int sub = lane & (GV - 1), grp = lane / GV, per = 32 / GV;
    const T* vbase = reinterpret_cast<const T*>(p.v) + off0 * p.ldv + (int64_t) h * p.dv + sub * E;
    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    for (int j0 = grp; j0 < nk; j0 += U * per) {
      float vf[U][E];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (j0 + u * per < nk) Chunk<T>::load(vbase + (int64_t) (j0 + u * per) * p.ldv, vf[u]);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (j0 + u * per < nk) {
          const float pj = ps[j0 + u * per];
#pragma unroll
          for (int e = 0; e < E; ++e) acc[e] = fmaf(pj, vf[u][e], acc[e]);
        }
    }
    for (int o = GV; o < 32; o <<= 1)
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
    if (grp == 0) store_chunk<T>(out + sub * E, acc);
  }
}

}  // namespace
}  // namespace grb

extern "C" int grb_hstu_attn_decode(const grb_hstu_attn_decode_args* a, grb_stream_t stream) {
  using namespace grb;
  GRB_REQUIRE(a != nullptr, GRB_ERR_INVALID_ARG, "hstu_attn_decode: null args");
  GRB_REQUIRE(a->B >= 0 && a->N > 0 && a->H > 0 && a->dqk > 0 && a->dv > 0, GRB_ERR_INVALID_ARG,
              "hstu_attn_decode: bad sizes");
  GRB_REQUIRE(a->B <= 0x7fffffff && a->H <= 65535, GRB_ERR_UNSUPPORTED, "hstu_attn_decode: grid too large");
  GRB_REQUIRE(a->N <= 24000, GRB_ERR_UNSUPPORTED,
              "hstu_attn_decode: N = %lld exceeds the shared-memory score buffers (24000)", (long long) a->N);
  GRB_REQUIRE(a->dtype == GRB_F32 || a->dtype == GRB_BF16, GRB_ERR_INVALID_ARG,
              "hstu_attn_decode: dtype must be GRB_F32 or GRB_BF16");
  GRB_REQUIRE((a->index_bits == 32 || a->index_bits == 64) && (a->pos_bits == 32 || a->pos_bits == 64),
              GRB_ERR_INVALID_ARG, "hstu_attn_decode: index_bits / pos_bits must be 32 or 64");
  if (a->B == 0) return GRB_OK;
  GRB_REQUIRE(a->q && a->k_cache && a->v && a->offsets && a->positions && a->out, GRB_ERR_INVALID_ARG,
              "hstu_attn_decode: null tensor");
  if (a->timestamps)
    GRB_REQUIRE(a->ts_w && a->pos_w && a->bucket_thresholds && a->num_buckets > 0, GRB_ERR_INVALID_ARG,
                "hstu_attn_decode: bias tables missing");
  DecodeParams p{};
  p.N = a->N; p.H = a->H; p.dqk = a->dqk; p.dv = a->dv; p.nb = a->timestamps ? a->num_buckets : 0;
  p.index_bits = a->index_bits; p.pos_bits = a->pos_bits;
  p.q = a->q; p.ldq = a->ldq; p.kc = a->k_cache; p.ldk = a->ldk; p.v = a->v; p.ldv = a->ldv;
  p.offsets = a->offsets; p.pos = a->positions;
  p.ts = a->timestamps; p.ts_w = a->ts_w; p.pos_w = a->pos_w; p.thr = a->bucket_thresholds;
  p.out = a->out; p.ldo = a->ldo;
  const size_t es = a->dtype == GRB_F32 ? 4 : 2;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((unsigned) a->B, (unsigned) a->H);
  {   // vector path: rows of 16 * 2^k bytes, 16-byte aligned everywhere
    auto pow2_lanes = [&](int d) {
      const size_t bytes = (size_t) d * es;
      if (bytes % 16) return 0;
      const size_t g = bytes / 16;
      return (g <= 32 && (g & (g - 1)) == 0) ? (int) g : 0;
    };
    auto al16 = [&](const void* ptr, int64_t ld) {
      return reinterpret_cast<uintptr_t>(ptr) % 16 == 0 && (ld * (int64_t) es) % 16 == 0;
    };
    const int gq = pow2_lanes(a->dqk), gv = pow2_lanes(a->dv);
    const bool aligned = gq && gv && al16(a->q, a->ldq) && al16(a->k_cache, a->ldk) && al16(a->v, a->ldv);
    const size_t smemw = sizeof(float) * (size_t) (a->H + 1) * ((a->N + 3) & ~3);
    if (aligned && a->H <= 16 && smemw <= 48 * 1024 && al16(a->out, a->ldo) &&
        ((size_t) a->dv * es) % 16 == 0) {
      // short sequences: one warp per (sequence, head)
      dim3 gridw((unsigned) a->B);
      if (a->dtype == GRB_F32)
        hstu_attn_decode_warp_kernel<float><<<gridw, 32 * a->H, smemw, st>>>(p, gq, gv);
      else
        hstu_attn_decode_warp_kernel<__nv_bfloat16><<<gridw, 32 * a->H, smemw, st>>>(p, gq, gv);
      GRB_LAUNCH_OK();
      return GRB_OK;
    }
    if (aligned) {
      const size_t nwdv = (size_t) (DECV_THREADS / 32) * a->dv;
      const size_t smemv = sizeof(float) * (2 * ((a->N + 3) & ~3) + nwdv + (nwdv & 1)) + 8 * 256;
      // the bias is the same for every head: a CTA takes all heads of a sequence once there are
      // enough sequences to fill the machine, else one head each
      const int hpc = a->B >= 4 * (int64_t) num_sms() ? a->H : 1;
      dim3 gridv((unsigned) a->B, (unsigned) ceil_div(a->H, hpc));
      if (a->dtype == GRB_F32) {
        auto kern = hstu_attn_decode_vec_kernel<float>;
        GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smemv));
        kern<<<gridv, DECV_THREADS, smemv, st>>>(p, gq, gv, hpc);
      } else {
        auto kern = hstu_attn_decode_vec_kernel<__nv_bfloat16>;
        GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smemv));
        kern<<<gridv, DECV_THREADS, smemv, st>>>(p, gq, gv, hpc);
      }
      GRB_LAUNCH_OK();
      return GRB_OK;
    }
  }
  const bool vec = (reinterpret_cast<uintptr_t>(a->k_cache) % 16 == 0) && ((a->ldk * es) % 16 == 0) &&
                   ((a->dqk * es) % 16 == 0);
  const size_t smem = sizeof(float) * (((a->dqk + 3) & ~3) + ((a->N + 3) & ~3) + 4 * 128);
#define GRB_DEC_LAUNCH(T, V)                                                                         \
  do {                                                                                               \
    auto kern = hstu_attn_decode_kernel<T, V>;                                                       \
    GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
    kern<<<grid, DEC_THREADS, smem, st>>>(p);                                                        \
  } while (0)
  if (a->dtype == GRB_F32) { if (vec) GRB_DEC_LAUNCH(float, true); else GRB_DEC_LAUNCH(float, false); }
  else { if (vec) GRB_DEC_LAUNCH(__nv_bfloat16, true); else GRB_DEC_LAUNCH(__nv_bfloat16, false); }
#undef GRB_DEC_LAUNCH
  GRB_LAUNCH_OK();
  return GRB_OK;
}
