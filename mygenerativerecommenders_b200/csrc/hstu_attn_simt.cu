// hstu_attn_simt.cu — HSTU jagged pointwise-SiLU attention, CUDA-core path (fp32 math).
//
// Follows /root/reference/src/generative_recommenders_pl/models/sequential_encoders/hstu.py:
//   :96-128  RelativeBucketedTimeAndPositionBasedBias.forward  (pos_w[N-1+j-i] + ts_w[bucket])
//   :134-205 _hstu_attention_maybe_from_cache                   (QK^T, +bias, SiLU/N, mask, PV)
// but jagged end-to-end: no padded q/k/v, no (B,H,N,N) score tensor, no (B,N,N) bucket tensor.
// This path serves the fp32 configurations (ml-1m d=50: head dims that are not tensor-core
// shaped, and fp32 parity at 1e-5) and any bf16 shape the tcgen05 path does not cover.
//
// Tiling: 64 query rows x 64 key rows per step, 256 threads, each thread a 4x4 patch of S
// (rows 4*ty+a, cols tx+16*c) and a 4 x NC patch of the output (cols tx+16*c, c < NC).
#include "common.cuh"

namespace grb {

constexpr int BM = 64;
constexpr int BN = 64;
constexpr int ATT_THREADS = 256;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// rows [row0, row0+64) of a jagged (T, ld) matrix slice [col0, col0+d) -> smem (64, lds) fp32;
// rows >= nvalid are zero.
template <typename T>
__device__ __forceinline__ void load_tile(float* __restrict__ dst, int lds,
                                          const T* __restrict__ src, int64_t ld, int64_t row0,
                                          int nvalid, int d) {
  for (int idx = threadIdx.x; idx < 64 * d; idx += ATT_THREADS) {
    int r = idx / d, c = idx - r * d;
    float v = 0.f;
    if (r < nvalid) v = to_f32<T>(src[(row0 + r) * ld + c]);
    dst[r * lds + c] = v;
  }
}

struct SimtParams {
  int64_t N, T;
  int H, dqk, dv, nb;
  int index_bits;
  const void* q; const void* k; const void* v;
  int64_t ldq, ldk, ldv;
  const void* offsets;
  const int64_t* ts;
  const float* ts_w; const float* pos_w; const int64_t* thr;
  void* out; int64_t ldo;
  const void* dout; int64_t lddo;
  void* dq; void* dk; void* dvg; int64_t lddq, lddk, lddv;
  float* dq_accum; float* d_ts_w; float* d_pos_w;
};

// ext_ts[b, idx] with the reference's (B, N+1) extension: idx == N reads ts[b, N-1].
__device__ __forceinline__ int64_t ext_ts(const int64_t* ts, int64_t b, int64_t N, int64_t idx) {
  if (idx >= N) idx = N - 1;
  return ts[b * N + idx];
}

template <typename T, int NC>
__global__ void __launch_bounds__(ATT_THREADS) hstu_attn_fwd_simt(SimtParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qt = gridDim.x - 1 - blockIdx.x;  // heavy (late) tiles first
  const int64_t off0 = load_index(p.offsets, b, p.index_bits);
  int64_t n64 = load_index(p.offsets, b + 1, p.index_bits) - off0;
  if (n64 > p.N) n64 = p.N;
  const int n = (int) n64;
  const int i0 = qt * BM;
  if (i0 >= n) return;

  const int ldsq = p.dqk | 1, ldsv = p.dv | 1;
  float* Qs = reinterpret_cast<float*>(smem_raw);
  float* Ks = Qs + 64 * ldsq;
  float* Vs = Ks + 64 * ldsq;
  float* Ps = Vs + 64 * ldsv;                      // 64 x 65
  float* tsw = Ps + 64 * 65;                       // nb+1
  int64_t* thr = reinterpret_cast<int64_t*>(tsw + ((p.nb + 1 + 1) & ~1));  // 8B aligned
  int64_t* tsq = thr + p.nb;                       // 64
  int64_t* tsk = tsq + 64;                         // 64

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const bool has_bias = p.ts != nullptr;
  const T* qg = reinterpret_cast<const T*>(p.q) + h * p.dqk;
  const T* kg = reinterpret_cast<const T*>(p.k) + h * p.dqk;
  const T* vg = reinterpret_cast<const T*>(p.v) + h * p.dv;

  load_tile<T>(Qs, ldsq, qg, p.ldq, off0 + i0, min(64, n - i0), p.dqk);
  if (has_bias) {
    for (int i = tid; i <= p.nb; i += ATT_THREADS) tsw[i] = p.ts_w[i];
    for (int i = tid; i < p.nb; i += ATT_THREADS) thr[i] = p.thr[i];
    if (tid < 64) tsq[tid] = ext_ts(p.ts, b, p.N, i0 + tid + 1);
  }

  float acc[4][NC];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[a][c] = 0.f;

  const float fN = (float) p.N;
  const int last_kt = min(i0 + BM - 1, n - 1) / BN;
  for (int kt = 0; kt <= last_kt; ++kt) {
    const int j0 = kt * BN;
    __syncthreads();  // previous iteration finished with Ks/Vs/Ps
    load_tile<T>(Ks, ldsq, kg, p.ldk, off0 + j0, min(64, n - j0), p.dqk);
    load_tile<T>(Vs, ldsv, vg, p.ldv, off0 + j0, min(64, n - j0), p.dv);
    if (has_bias && tid < 64) tsk[tid] = ext_ts(p.ts, b, p.N, j0 + tid);
    __syncthreads();

    float s[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[a][c] = 0.f;
    for (int kk = 0; kk < p.dqk; ++kk) {
      float qv[4], kv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) qv[a] = Qs[(4 * ty + a) * ldsq + kk];
#pragma unroll
      for (int c = 0; c < 4; ++c) kv[c] = Ks[(tx + 16 * c) * ldsq + kk];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) s[a][c] = fmaf(qv[a], kv[c], s[a][c]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int il = 4 * ty + a, i = i0 + il;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int jl = tx + 16 * c, j = j0 + jl;
        float pv = 0.f;
        if (j <= i && i < n) {
          float x = s[a][c];
          if (has_bias) {
            int64_t d = tsq[il] - tsk[jl];
            d = d < 0 ? -d : d;
            x += p.pos_w[p.N - 1 + j - i] + tsw[bucket_of(thr, p.nb, d)];
          }
          pv = silu_f32(x) / fN;
        }
        Ps[il * 65 + jl] = pv;
      }
    }
    __syncthreads();
    for (int j = 0; j < BN; ++j) {
      float pr[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) pr[a] = Ps[(4 * ty + a) * 65 + j];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col = tx + 16 * c;
        const float vv = (col < p.dv) ? Vs[j * ldsv + col] : 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a][c] = fmaf(pr[a], vv, acc[a][c]);
      }
    }
  }
  T* og = reinterpret_cast<T*>(p.out) + h * p.dv;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + 4 * ty + a;
    if (i >= n) continue;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = tx + 16 * c;
      if (col < p.dv) og[(off0 + i) * p.ldo + col] = from_f32<T>(acc[a][c]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward.  One CTA per (sequence, head, key tile j); loops over query tiles i >= j.
//   dP = dO V^T ; dS = dP * SiLU'(S+bias)/N (j<=i) ; dV_j += P^T dO ; dK_j += dS^T Q ;
//   dQ_i += dS K (fp32 atomics) ; d pos_w / d ts_w: histograms of dS (hstu.py:125-128 backward).
// ---------------------------------------------------------------------------------------------
template <typename T, int NC>
__global__ void __launch_bounds__(ATT_THREADS) hstu_attn_bwd_simt(SimtParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.z, h = blockIdx.y, kt = blockIdx.x;
  const int64_t off0 = load_index(p.offsets, b, p.index_bits);
  int64_t n64 = load_index(p.offsets, b + 1, p.index_bits) - off0;
  if (n64 > p.N) n64 = p.N;
  const int n = (int) n64;
  const int j0 = kt * BN;
  if (j0 >= n) return;

  const int ldsq = p.dqk | 1, ldsv = p.dv | 1;
  float* Ks = reinterpret_cast<float*>(smem_raw);
  float* Vs = Ks + 64 * ldsq;
  float* Qs = Vs + 64 * ldsv;
  float* dOs = Qs + 64 * ldsq;
  float* Ps = dOs + 64 * ldsv;     // 64 x 65, [q][k]
  float* dSs = Ps + 64 * 65;       // 64 x 65, [q][k]
  float* tsw = dSs + 64 * 65;      // nb+1
  float* h_ts = tsw + (p.nb + 1);  // nb+1 histogram (whole CTA lifetime)
  float* h_pos = h_ts + (p.nb + 1);  // 128 (per q tile): index (j-i) - (j0-i0) + 63
  int64_t* thr = reinterpret_cast<int64_t*>(h_pos + 128 + (((p.nb + 1) * 2) & 1));
  int64_t* tsq = thr + p.nb;
  int64_t* tsk = tsq + 64;

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const bool has_bias = p.ts != nullptr;
  const T* qg = reinterpret_cast<const T*>(p.q) + h * p.dqk;
  const T* kg = reinterpret_cast<const T*>(p.k) + h * p.dqk;
  const T* vg = reinterpret_cast<const T*>(p.v) + h * p.dv;
  const T* dog = reinterpret_cast<const T*>(p.dout) + h * p.dv;

  load_tile<T>(Ks, ldsq, kg, p.ldk, off0 + j0, min(64, n - j0), p.dqk);
  load_tile<T>(Vs, ldsv, vg, p.ldv, off0 + j0, min(64, n - j0), p.dv);
  if (has_bias) {
    for (int i = tid; i <= p.nb; i += ATT_THREADS) { tsw[i] = p.ts_w[i]; h_ts[i] = 0.f; }
    for (int i = tid; i < p.nb; i += ATT_THREADS) thr[i] = p.thr[i];
    if (tid < 64) tsk[tid] = ext_ts(p.ts, b, p.N, j0 + tid);
  }

  float dKa[4][NC], dVa[4][NC];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < NC; ++c) { dKa[a][c] = 0.f; dVa[a][c] = 0.f; }

  const float fN = (float) p.N;
  const int n_qt = (n + BM - 1) / BM;
  for (int qt = kt; qt < n_qt; ++qt) {
    const int i0 = qt * BM;
    __syncthreads();
    load_tile<T>(Qs, ldsq, qg, p.ldq, off0 + i0, min(64, n - i0), p.dqk);
    load_tile<T>(dOs, ldsv, dog, p.lddo, off0 + i0, min(64, n - i0), p.dv);
    if (has_bias) {
      if (tid < 64) tsq[tid] = ext_ts(p.ts, b, p.N, i0 + tid + 1);
      if (tid < 128) h_pos[tid] = 0.f;
    }
    __syncthreads();

    float s[4][4], dp[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) { s[a][c] = 0.f; dp[a][c] = 0.f; }
    for (int kk = 0; kk < p.dqk; ++kk) {
      float qv[4], kv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) qv[a] = Qs[(4 * ty + a) * ldsq + kk];
#pragma unroll
      for (int c = 0; c < 4; ++c) kv[c] = Ks[(tx + 16 * c) * ldsq + kk];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) s[a][c] = fmaf(qv[a], kv[c], s[a][c]);
    }
    for (int kk = 0; kk < p.dv; ++kk) {
      float ov[4], vv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) ov[a] = dOs[(4 * ty + a) * ldsv + kk];
#pragma unroll
      for (int c = 0; c < 4; ++c) vv[c] = Vs[(tx + 16 * c) * ldsv + kk];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) dp[a][c] = fmaf(ov[a], vv[c], dp[a][c]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int il = 4 * ty + a, i = i0 + il;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int jl = tx + 16 * c, j = j0 + jl;
        float pv = 0.f, ds = 0.f;
        if (j <= i && i < n) {
          float x = s[a][c];
          int bk = 0;
          if (has_bias) {
            int64_t d = tsq[il] - tsk[jl];
            d = d < 0 ? -d : d;
            bk = bucket_of(thr, p.nb, d);
            x += p.pos_w[p.N - 1 + j - i] + tsw[bk];
          }
          const float sg = 1.0f / (1.0f + expf(-x));
          pv = x * sg / fN;
          ds = dp[a][c] * (sg * (1.0f + x * (1.0f - sg))) / fN;
          if (has_bias) {
            atomicAdd(&h_ts[bk], ds);
            atomicAdd(&h_pos[(jl - il) + 63], ds);
          }
        }
        Ps[il * 65 + jl] = pv;
        dSs[il * 65 + jl] = ds;
      }
    }
    __syncthreads();
    // dV[k][c] += sum_q P[q][k] dO[q][c] ; dK[k][c] += sum_q dS[q][k] Q[q][c]
    for (int qq = 0; qq < BM; ++qq) {
      float pr[4], dr[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        pr[a] = Ps[qq * 65 + 4 * ty + a];
        dr[a] = dSs[qq * 65 + 4 * ty + a];
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col = tx + 16 * c;
        const float ov = (col < p.dv) ? dOs[qq * ldsv + col] : 0.f;
        const float qv = (col < p.dqk) ? Qs[qq * ldsq + col] : 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          dVa[a][c] = fmaf(pr[a], ov, dVa[a][c]);
          dKa[a][c] = fmaf(dr[a], qv, dKa[a][c]);
        }
      }
    }
    // dQ[q][c] = sum_k dS[q][k] K[k][c]  -> fp32 atomics
    {
      float dqa[4][NC];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < NC; ++c) dqa[a][c] = 0.f;
      for (int kk = 0; kk < BN; ++kk) {
        float dr[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) dr[a] = dSs[(4 * ty + a) * 65 + kk];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int col = tx + 16 * c;
          const float kv = (col < p.dqk) ? Ks[kk * ldsq + col] : 0.f;
#pragma unroll
          for (int a = 0; a < 4; ++a) dqa[a][c] = fmaf(dr[a], kv, dqa[a][c]);
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = i0 + 4 * ty + a;
        if (i >= n) continue;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int col = tx + 16 * c;
          if (col < p.dqk)
            atomicAdd(&p.dq_accum[(off0 + i) * (int64_t) (p.H * p.dqk) + h * p.dqk + col],
                      dqa[a][c]);
        }
      }
    }
    if (has_bias) {
      __syncthreads();  // h_pos complete (it was filled before the previous barrier; be explicit)
      if (tid < 127) {
        const float v = h_pos[tid];
        // tid = (jl - il) + 63  ->  j - i = tid - 63 + (j0 - i0)
        const int64_t rel = (int64_t) tid - 63 + (j0 - i0) + (p.N - 1);
        if (v != 0.f && rel >= 0 && rel < 2 * p.N - 1) atomicAdd(&p.d_pos_w[rel], v);
      }
    }
  }
  __syncthreads();
  if (has_bias) {
    for (int i = tid; i <= p.nb; i += ATT_THREADS) {
      const float v = h_ts[i];
      if (v != 0.f) atomicAdd(&p.d_ts_w[i], v);
    }
  }
  T* dkg = reinterpret_cast<T*>(p.dk) + h * p.dqk;
  T* dvg = reinterpret_cast<T*>(p.dvg) + h * p.dv;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int j = j0 + 4 * ty + a;
    if (j >= n) continue;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = tx + 16 * c;
      if (col < p.dqk) dkg[(off0 + j) * p.lddk + col] = from_f32<T>(dKa[a][c]);
      if (col < p.dv) dvg[(off0 + j) * p.lddv + col] = from_f32<T>(dVa[a][c]);
    }
  }
}

// dq_accum (T, H*dqk) fp32 -> dq (T, lddq) of type T
template <typename T>
__global__ void dq_convert_kernel(const float* __restrict__ acc, T* __restrict__ dq, int64_t rows,
                                  int W, int64_t lddq) {
  const int64_t idx = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * W) return;
  const int64_t r = idx / W;
  const int c = (int) (idx - r * W);
  dq[r * lddq + c] = from_f32<T>(acc[idx]);
}

static SimtParams make_params(const grb_hstu_attn_args* a) {
  SimtParams p{};
  p.N = a->N; p.T = a->T; p.H = a->H; p.dqk = a->dqk; p.dv = a->dv; p.nb = a->num_buckets;
  p.index_bits = a->index_bits;
  p.q = a->q; p.k = a->k; p.v = a->v; p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv;
  p.offsets = a->offsets; p.ts = a->timestamps; p.ts_w = a->ts_w; p.pos_w = a->pos_w;
  p.thr = a->bucket_thresholds;
  p.out = a->out; p.ldo = a->ldo; p.dout = a->dout; p.lddo = a->lddo;
  p.dq = a->dq; p.dk = a->dk; p.dvg = a->dv_grad; p.lddq = a->lddq; p.lddk = a->lddk;
  p.lddv = a->lddv; p.dq_accum = a->dq_accum; p.d_ts_w = a->d_ts_w; p.d_pos_w = a->d_pos_w;
  if (p.ts == nullptr) p.nb = 0;
  return p;
}

template <typename T, int NC>
static int launch_fwd(const grb_hstu_attn_args* a, cudaStream_t st) {
  SimtParams p = make_params(a);
  const int ldsq = p.dqk | 1, ldsv = p.dv | 1;
  size_t smem = sizeof(float) * (2 * 64 * ldsq + 64 * ldsv + 64 * 65 + ((p.nb + 2) & ~1)) +
                sizeof(int64_t) * (p.nb + 128);
  auto kern = hstu_attn_fwd_simt<T, NC>;
  GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  dim3 grid((unsigned) ceil_div(a->max_len, BM), (unsigned) a->H, (unsigned) a->B);
  kern<<<grid, ATT_THREADS, smem, st>>>(p);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

template <typename T, int NC>
static int launch_bwd(const grb_hstu_attn_args* a, cudaStream_t st) {
  SimtParams p = make_params(a);
  const int ldsq = p.dqk | 1, ldsv = p.dv | 1;
  size_t smem = sizeof(float) * (2 * 64 * ldsq + 2 * 64 * ldsv + 2 * 64 * 65 + 2 * (p.nb + 1) +
                                 128 + 2) +
                sizeof(int64_t) * (p.nb + 128);
  auto kern = hstu_attn_bwd_simt<T, NC>;
  GRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  dim3 grid((unsigned) ceil_div(a->max_len, BN), (unsigned) a->H, (unsigned) a->B);
  kern<<<grid, ATT_THREADS, smem, st>>>(p);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int check_attn_args(const grb_hstu_attn_args* a, bool bwd) {
  GRB_REQUIRE(a != nullptr, GRB_ERR_INVALID_ARG, "hstu_attn: null args");
  GRB_REQUIRE(a->index_bits == 32 || a->index_bits == 64, GRB_ERR_INVALID_ARG,
              "hstu_attn: index_bits must be 32 or 64");
  GRB_REQUIRE(a->B >= 0 && a->N > 0 && a->T >= 0 && a->H > 0 && a->dqk > 0 && a->dv > 0,
              GRB_ERR_INVALID_ARG, "hstu_attn: bad sizes");
  GRB_REQUIRE(a->max_len >= 0 && a->max_len <= a->N, GRB_ERR_INVALID_ARG,
              "hstu_attn: max_len %lld must be in [0, N=%lld]", (long long) a->max_len,
              (long long) a->N);
  GRB_REQUIRE(a->dtype == GRB_F32 || a->dtype == GRB_BF16, GRB_ERR_INVALID_ARG,
              "hstu_attn: dtype must be GRB_F32 or GRB_BF16");
  GRB_REQUIRE(a->q && a->k && a->v && a->offsets, GRB_ERR_INVALID_ARG,
              "hstu_attn: q/k/v/offsets must be non-null");
  if (a->timestamps) {
    GRB_REQUIRE(a->ts_w && a->pos_w && a->bucket_thresholds && a->num_buckets > 0 &&
                    a->num_buckets <= 4096,
                GRB_ERR_INVALID_ARG, "hstu_attn: bias tables missing");
  }
  if (!bwd) {
    GRB_REQUIRE(a->out, GRB_ERR_INVALID_ARG, "hstu_attn_fwd: out is null");
  } else {
    GRB_REQUIRE(a->dout && a->dq && a->dk && a->dv_grad, GRB_ERR_INVALID_ARG,
                "hstu_attn_bwd: dout/dq/dk/dv_grad must be non-null");
    GRB_REQUIRE(a->dq_accum || (a->short_schedule && a->max_len <= 128), GRB_ERR_INVALID_ARG,
                "hstu_attn_bwd: dq_accum workspace is null");
    if (a->timestamps)
      GRB_REQUIRE(a->d_ts_w && a->d_pos_w, GRB_ERR_INVALID_ARG,
                  "hstu_attn_bwd: d_ts_w/d_pos_w must be non-null with timestamps");
  }
  return GRB_OK;
}

int hstu_attn_fwd_simt_dispatch(const grb_hstu_attn_args* a, cudaStream_t st) {
  const int dmax = a->dv;
  GRB_REQUIRE(a->dqk <= 256 && a->dv <= 256, GRB_ERR_UNSUPPORTED,
              "hstu_attn_fwd (CUDA-core path): head dims must be <= 256 (dqk=%d dv=%d)", a->dqk,
              a->dv);
  if (a->B == 0 || a->T == 0 || a->max_len == 0) return GRB_OK;
#define GRB_FWD(TT)                                        \
  (dmax <= 64 ? launch_fwd<TT, 4>(a, st)                   \
              : dmax <= 128 ? launch_fwd<TT, 8>(a, st) : launch_fwd<TT, 16>(a, st))
  return a->dtype == GRB_F32 ? GRB_FWD(float) : GRB_FWD(__nv_bfloat16);
#undef GRB_FWD
}

int hstu_attn_bwd_simt_dispatch(const grb_hstu_attn_args* a, cudaStream_t st) {
  const int dmax = a->dv > a->dqk ? a->dv : a->dqk;
  GRB_REQUIRE(dmax <= 128, GRB_ERR_UNSUPPORTED,
              "hstu_attn_bwd (CUDA-core path): head dims must be <= 128 (dqk=%d dv=%d)", a->dqk,
              a->dv);
  if (a->B == 0 || a->T == 0 || a->max_len == 0) return GRB_OK;
  int rc;
  if (a->dtype == GRB_F32)
    rc = dmax <= 64 ? launch_bwd<float, 4>(a, st) : launch_bwd<float, 8>(a, st);
  else
    rc = dmax <= 64 ? launch_bwd<__nv_bfloat16, 4>(a, st) : launch_bwd<__nv_bfloat16, 8>(a, st);
  if (rc != GRB_OK) return rc;
  const int W = a->H * a->dqk;
  const int64_t total = a->T * W;
  const unsigned blocks = (unsigned) ceil_div(total, 256);
  if (a->dtype == GRB_F32)
    dq_convert_kernel<float><<<blocks, 256, 0, st>>>(a->dq_accum, reinterpret_cast<float*>(a->dq),
                                                      a->T, W, a->lddq);
  else
    dq_convert_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
        a->dq_accum, reinterpret_cast<__nv_bfloat16*>(a->dq), a->T, W, a->lddq);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}  // namespace grb
