// ssl_bwd_csr.cu — backward of the fused sampled softmax for the in-batch negatives path (one table that
// is already L2-normalised: negatives_samples/negative_sampler.py:192-212, losses/
// autoregressive_losses.py:279-306), without atomics on the table gradient and with 16-bit gathers.
//
// The row-major backward (neg_gather.cu: ssl_bwd_vec_kernel) gathers every sampled cache row once more
// (N' R D fp32 = 1.84 GB at C2, served by L2) and scatter-adds dl * q[n] into d cache[idx] with vector
// reds (another 1.84 GB of atomic traffic on 11 k rows): 0.36 ms, the largest kernel of the C2 step.
// Here the two halves of that work are separated and both become plain reads of bf16 rows:
//
//   prep    : bf16 copies of q (N', D) and of the cache (X0, D); zero the per-row counters.
//   rows    : warp per supervised position n.  coef[n, r] = g[n] p[n, r] / T (0 for a masked collision),
//             written out; dq[n] = dzpos p[n] + sum_r coef e_r with e_r gathered as 16-byte bf16 chunks;
//             dp[n] = dzpos q[n]; counts[idx[n, r]] += 1 for every non-zero coefficient.
//   (exclusive scan of the counts: grb_complete_cumsum)
//   scatter : pair (n, r) -> its slot in the list of cache row idx[n, r] (counting sort; the slot inside the
//             list is the old value of the counter, remembered by `rows`).
//   cols    : warp per cache row c: d cache[c] = sum over its list of coef * q[n] (bf16 q rows, fp32
//             accumulation in registers, one plain store per row — rows nobody sampled get zeros, so the
//             gradient needs no zero fill).
//
// Gradients therefore see q, the cache and the coefficients rounded to bf16 inside the two sums (the
// products are mixed-precision FMAs, bf16 x bf16 -> fp32 accumulate; relative 2^-9 per factor, independent
// signs): well inside the 2e-2 gradient tolerance of the bf16 path; the loss and the
// probabilities are the forward kernel's, computed from the fp32 values.
#include "common.cuh"

namespace grb {
namespace {

constexpr int CSR_WARPS = 8;

struct CsrP {
  int64_t n_rows, X0;
  int R, D;
  float temp;
  const float* q; int64_t ldq;
  const float* p; int64_t ldp;
  const float* t0; int64_t ldt0;
  const int64_t* idx; const int64_t* pos_ids; const int64_t* neg_ids;
  const float* probs; const float* g;
  float* dq; float* dp; float* dt0;
  __nv_bfloat16* q16; __nv_bfloat16* t16;
  float* coef; int32_t* counts; int32_t* ptr; int32_t* slot; int32_t* pairs;
};

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// acc[0..7] += c * (the 8 bf16 values of v), with c given as a bf16 pair (c, c): eight mixed-precision
// FMAs (bf16 x bf16 -> fp32 accumulate, FHFMA) on the register halves, no unpacking
__device__ __forceinline__ void fma8_bf16(float (&acc)[8], uint32_t c2, const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm("{\n\t.reg .b16 cl, ch, vl, vh;\n\tmov.b32 {cl, ch}, %2;\n\tmov.b32 {vl, vh}, %3;\n\t"
        "fma.rn.f32.bf16 %0, cl, vl, %0;\n\tfma.rn.f32.bf16 %1, ch, vh, %1;\n\t}"
        : "+f"(acc[2 * i]), "+f"(acc[2 * i + 1])
        : "r"(c2), "r"(w[i]));
}
__device__ __forceinline__ uint32_t bf16_pair(float c) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(c));
  return r;
}

__global__ void __launch_bounds__(256) csr_prep_kernel(CsrP P) {
  const int64_t nq = P.n_rows * (P.D / 4), nt = P.X0 * (P.D / 4);
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  const int d4 = P.D / 4;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < nq + nt; i += stride) {
    const bool is_q = i < nq;
    const int64_t j = is_q ? i : i - nq;
    const int64_t row = j / d4;
    const int c = (int) (j - row * d4) * 4;
    const float4 v = *reinterpret_cast<const float4*>((is_q ? P.q + row * P.ldq : P.t0 + row * P.ldt0) + c);
    uint2 o;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.x) : "f"(v.y), "f"(v.x));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.y) : "f"(v.w), "f"(v.z));
    *reinterpret_cast<uint2*>((is_q ? P.q16 : P.t16) + row * P.D + c) = o;
  }
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < P.X0; i += stride) P.counts[i] = 0;
}

__global__ void __launch_bounds__(CSR_WARPS * 32) csr_rows_kernel(CsrP P) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t) blockIdx.x * CSR_WARPS + (threadIdx.x >> 5);
  if (n >= P.n_rows) return;
  const bool act = 8 * lane < P.D;             // lane owns columns [8 lane, 8 lane + 8)
  const float g = P.g[n];
  const float* pr = P.probs + n * (int64_t) (P.R + 1);
  const float dzpos = g * (pr[0] - 1.0f) / P.temp;
  float dqa[8];
  if (act) {
    const float* qr = P.q + n * P.ldq + 8 * lane;
    const float* prow = P.p + n * P.ldp + 8 * lane;
    const float4 q0 = *reinterpret_cast<const float4*>(qr), q1 = *reinterpret_cast<const float4*>(qr + 4);
    const float4 p0 = *reinterpret_cast<const float4*>(prow), p1 = *reinterpret_cast<const float4*>(prow + 4);
    float* dpr = P.dp + n * (int64_t) P.D + 8 * lane;
    *reinterpret_cast<float4*>(dpr) = make_float4(dzpos * q0.x, dzpos * q0.y, dzpos * q0.z, dzpos * q0.w);
    *reinterpret_cast<float4*>(dpr + 4) = make_float4(dzpos * q1.x, dzpos * q1.y, dzpos * q1.z, dzpos * q1.w);
    dqa[0] = dzpos * p0.x; dqa[1] = dzpos * p0.y; dqa[2] = dzpos * p0.z; dqa[3] = dzpos * p0.w;
    dqa[4] = dzpos * p1.x; dqa[5] = dzpos * p1.y; dqa[6] = dzpos * p1.z; dqa[7] = dzpos * p1.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) dqa[k] = 0.f;
  }
  const int64_t pid = P.pos_ids[n];
  const int64_t* idx = P.idx + n * (int64_t) P.R;
  const int64_t* nid = P.neg_ids + n * (int64_t) P.R;
  float* cf = P.coef + n * (int64_t) P.R;
  for (int r0 = 0; r0 < P.R; r0 += 32) {
    const int r = r0 + lane;
    float my_c = 0.f;
    int64_t my_i = 0;
    if (r < P.R) {
      my_i = idx[r];
      if (g != 0.f && nid[r] != pid) my_c = g * pr[r + 1] / P.temp;
      cf[r] = my_c;
      // the counter's old value is this pair's position inside the list of its cache row: the counting
      // sort below needs no second round of atomics
      if (my_c != 0.f) P.slot[n * (int64_t) P.R + r] = atomicAdd(P.counts + my_i, 1);
    }
    if (!__any_sync(0xffffffffu, my_c != 0.f)) continue;
#pragma unroll 2
    for (int j4 = 0; j4 < 32; j4 += 4) {
      float c4[4];
      int64_t i4[4];
      uint4 e4[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        c4[t] = __shfl_sync(0xffffffffu, my_c, j4 + t);
        i4[t] = __shfl_sync(0xffffffffu, my_i, j4 + t);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)      // all four 16-byte gathers in flight before the first use
        e4[t] = act ? __ldg(reinterpret_cast<const uint4*>(P.t16 + i4[t] * P.D + 8 * lane)) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (c4[t] == 0.f) continue;    // warp-uniform
        fma8_bf16(dqa, bf16_pair(c4[t]), e4[t]);
      }
    }
  }
  if (act) {
    float* dqr = P.dq + n * (int64_t) P.D + 8 * lane;
    *reinterpret_cast<float4*>(dqr) = make_float4(dqa[0], dqa[1], dqa[2], dqa[3]);
    *reinterpret_cast<float4*>(dqr + 4) = make_float4(dqa[4], dqa[5], dqa[6], dqa[7]);
  }
}

__global__ void __launch_bounds__(256) csr_scatter_kernel(CsrP P) {
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_rows * P.R) return;
  if (P.coef[i] == 0.f) return;
  P.pairs[P.ptr[P.idx[i]] + P.slot[i]] = (int32_t) i;
}

__global__ void __launch_bounds__(CSR_WARPS * 32) csr_cols_kernel(CsrP P) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t) blockIdx.x * CSR_WARPS + (threadIdx.x >> 5);
  if (c >= P.X0) return;
  const bool act = 8 * lane < P.D;
  const int beg = P.ptr[c], end = P.ptr[c + 1];
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j0 = beg; j0 < end; j0 += 32) {
    // this lane's pair of the batch: its coefficient and supervised position
    const int j = j0 + lane;
    float my_c = 0.f;
    int my_n = 0;
    if (j < end) {
      const int pr = P.pairs[j];
      my_c = P.coef[pr];
      my_n = pr / P.R;
    }
    const int cnt = end - j0 < 32 ? end - j0 : 32;
    for (int t0 = 0; t0 < cnt; t0 += 4) {
      float c4[4];
      uint4 q4[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        c4[t] = __shfl_sync(0xffffffffu, my_c, t0 + t);         // lanes past cnt hold 0
        const int nn = __shfl_sync(0xffffffffu, my_n, t0 + t);
        q4[t] = (act && t0 + t < cnt) ? __ldg(reinterpret_cast<const uint4*>(P.q16 + (int64_t) nn * P.D + 8 * lane))
                                      : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) fma8_bf16(acc, bf16_pair(c4[t]), q4[t]);
    }
  }
  if (act) {
    float* o = P.dt0 + c * (int64_t) P.D + 8 * lane;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

int64_t al256(int64_t x) { return (x + 255) & ~255ll; }

struct CsrPlan { int64_t q16, t16, coef, counts, ptr, slot, pairs, total; };
CsrPlan plan(int64_t n, int R, int D, int64_t X0) {
  CsrPlan L{};
  int64_t o = 0;
  L.q16 = o;    o = al256(o + n * D * 2);
  L.t16 = o;    o = al256(o + X0 * D * 2);
  L.coef = o;   o = al256(o + n * R * 4);
  L.counts = o; o = al256(o + X0 * 4);
  L.ptr = o;    o = al256(o + (X0 + 1) * 4);
  L.slot = o;   o = al256(o + n * R * 4);
  L.pairs = o;  o = al256(o + n * R * 4);
  L.total = o;
  return L;
}

}  // namespace
}  // namespace grb

using namespace grb;

extern "C" {

int grb_complete_cumsum(const void* lengths, void* offsets, int64_t B, int index_bits, grb_stream_t stream);

int64_t grb_sampled_softmax_bwd_csr_workspace_bytes(int64_t n_rows, int32_t R, int32_t D, int64_t table_rows) {
  if (n_rows < 0 || R <= 0 || D <= 0 || table_rows <= 0) return GRB_ERR_INVALID_ARG;
  return plan(n_rows, R, D, table_rows).total;
}

int grb_sampled_softmax_bwd_csr(const grb_ssl_args* a, int64_t table_rows, void* workspace,
                                int64_t workspace_bytes, grb_stream_t stream) {
  GRB_REQUIRE(a != nullptr && workspace != nullptr, GRB_ERR_INVALID_ARG, "sampled_softmax_bwd_csr: null args");
  GRB_REQUIRE(a->dtype == GRB_F32 && a->l2_norm == 0 && a->d1 == 0 && a->table1 == nullptr,
              GRB_ERR_UNSUPPORTED, "sampled_softmax_bwd_csr: one fp32 table of normalised rows only");
  GRB_REQUIRE(a->D == a->d0 && a->D % 8 == 0 && a->D <= 256 && a->R > 0 && a->n_rows >= 0 && table_rows > 0,
              GRB_ERR_UNSUPPORTED, "sampled_softmax_bwd_csr: D must be a multiple of 8, at most 256 (D=%d)", a->D);
  GRB_REQUIRE(a->n_rows * (int64_t) a->R < (1ll << 31) && table_rows < (1ll << 31), GRB_ERR_UNSUPPORTED,
              "sampled_softmax_bwd_csr: more than 2^31 (row, negative) pairs");
  GRB_REQUIRE(a->q && a->p && a->table0 && a->idx0 && a->pos_ids && a->neg_ids && a->probs && a->g && a->dq &&
                  a->dp && a->dtable0,
              GRB_ERR_INVALID_ARG, "sampled_softmax_bwd_csr: null tensor");
  GRB_REQUIRE(a->ldq_ % 4 == 0 && a->ldp % 4 == 0 && a->ldt0 % 4 == 0 &&
                  ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->p) |
                    reinterpret_cast<uintptr_t>(a->table0) | reinterpret_cast<uintptr_t>(a->dq) |
                    reinterpret_cast<uintptr_t>(a->dp) | reinterpret_cast<uintptr_t>(a->dtable0) |
                    reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
              GRB_ERR_INVALID_ARG, "sampled_softmax_bwd_csr: 16-byte aligned rows required");
  GRB_REQUIRE(a->temperature > 0.f, GRB_ERR_INVALID_ARG, "sampled_softmax_bwd_csr: temperature <= 0");
  const CsrPlan L = plan(a->n_rows, a->R, a->D, table_rows);
  GRB_REQUIRE(workspace_bytes >= L.total, GRB_ERR_WORKSPACE, "sampled_softmax_bwd_csr: workspace %lld < %lld",
              (long long) workspace_bytes, (long long) L.total);
  auto st = reinterpret_cast<cudaStream_t>(stream);
  auto ws = reinterpret_cast<unsigned char*>(workspace);
  CsrP P{};
  P.n_rows = a->n_rows; P.X0 = table_rows; P.R = a->R; P.D = a->D; P.temp = a->temperature;
  P.q = (const float*) a->q; P.ldq = a->ldq_; P.p = (const float*) a->p; P.ldp = a->ldp;
  P.t0 = (const float*) a->table0; P.ldt0 = a->ldt0;
  P.idx = a->idx0; P.pos_ids = a->pos_ids; P.neg_ids = a->neg_ids; P.probs = a->probs; P.g = a->g;
  P.dq = a->dq; P.dp = a->dp; P.dt0 = a->dtable0;
  P.q16 = reinterpret_cast<__nv_bfloat16*>(ws + L.q16); P.t16 = reinterpret_cast<__nv_bfloat16*>(ws + L.t16);
  P.coef = reinterpret_cast<float*>(ws + L.coef); P.counts = reinterpret_cast<int32_t*>(ws + L.counts);
  P.ptr = reinterpret_cast<int32_t*>(ws + L.ptr); P.slot = reinterpret_cast<int32_t*>(ws + L.slot);
  P.pairs = reinterpret_cast<int32_t*>(ws + L.pairs);
  csr_prep_kernel<<<1184, 256, 0, st>>>(P);
  GRB_LAUNCH_OK();
  if (a->n_rows > 0) {
    csr_rows_kernel<<<(unsigned) ceil_div(a->n_rows, (int64_t) CSR_WARPS), CSR_WARPS * 32, 0, st>>>(P);
    GRB_LAUNCH_OK();
  }
  int rc = grb_complete_cumsum(P.counts, P.ptr, table_rows, 32, stream);
  if (rc != GRB_OK) return rc;
  if (a->n_rows > 0) {
    csr_scatter_kernel<<<(unsigned) ceil_div(a->n_rows * a->R, (int64_t) 256), 256, 0, st>>>(P);
    GRB_LAUNCH_OK();
  }
  csr_cols_kernel<<<(unsigned) ceil_div(table_rows, (int64_t) CSR_WARPS), CSR_WARPS * 32, 0, st>>>(P);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}
