// neg_gather.cu — fused sampled-softmax: negative gather . dot (+L2 norm) + collision mask +
// log-softmax, forward and backward.  The (N', R, D) negatives tensor is never materialised.
//
// Reference (under /root/reference/src/generative_recommenders_pl/models/):
//   negatives_samples/negative_sampler.py:31-37   _maybe_l2_norm: x / clamp(||x||, eps)
//   negatives_samples/negative_sampler.py:123-131 gather of sampled ids (Local sampler; this
//        fork's embedding = concat(item_emb[id], year_emb[year_lookup[id]]), embeddings.py:94-97)
//   negatives_samples/negative_sampler.py:208-211 gather from the in-batch cache
//   similarity/dot_product.py:61-64               bmm((N',R,D),(N',D,1))
//   losses/autoregressive_losses.py:279-306       /T, where(==, -5e4), -log_softmax[:,0]
//
// Mapping: one warp per supervised position n; lanes stride over the D embedding columns, so
// every gathered row is read with coalesced 128-byte requests; dot and squared norm are reduced
// with warp shuffles.  HBM/L2-bound: bytes = N'.R.(D.4 + 16) + 2.N'.D.4 + N'.(R+1).4.
#include "common.cuh"

namespace grb {

constexpr int SSL_WARPS = 4;

struct SslP {
  int64_t n_rows;
  int R, D, d0, d1, l2;
  float eps, temp;
  const float* q; int64_t ldq;
  const float* p; int64_t ldp;
  const float* t0; int64_t ldt0;
  const float* t1; int64_t ldt1;
  const int64_t* idx0; const int64_t* idx1;
  const int64_t* pos_ids; const int64_t* neg_ids;
  float* loss_rows; float* probs;
  const float* g; float* dq; float* dp; float* dt0; float* dt1;
};

template <int NPL>
__device__ __forceinline__ void gather_row(const SslP& P, int64_t i0, int64_t i1, int lane,
                                           float (&e)[NPL]) {
#pragma unroll
  for (int u = 0; u < NPL; ++u) {
    const int c = lane + 32 * u;
    float v = 0.f;
    if (c < P.d0) v = __ldg(P.t0 + i0 * P.ldt0 + c);
    else if (c < P.D) v = __ldg(P.t1 + i1 * P.ldt1 + (c - P.d0));
    e[u] = v;
  }
}

template <int NPL>
__global__ void __launch_bounds__(SSL_WARPS * 32) ssl_fwd_kernel(SslP P) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t) blockIdx.x * SSL_WARPS + (threadIdx.x >> 5);
  if (n >= P.n_rows) return;
  float qv[NPL], e[NPL];
  float pd = 0.f;
#pragma unroll
  for (int u = 0; u < NPL; ++u) {
    const int c = lane + 32 * u;
    qv[u] = c < P.D ? P.q[n * P.ldq + c] : 0.f;
    const float pv = c < P.D ? P.p[n * P.ldp + c] : 0.f;
    pd = fmaf(qv[u], pv, pd);
  }
  const float zpos = warp_sum(pd) / P.temp;
  const int64_t pid = P.pos_ids[n];
  float* pr = P.probs + n * (int64_t) (P.R + 1);
  float m = zpos;
  for (int r = 0; r < P.R; ++r) {
    const int64_t i0 = P.idx0[n * P.R + r];
    const int64_t i1 = P.idx1 ? P.idx1[n * P.R + r] : 0;
    gather_row<NPL>(P, i0, i1, lane, e);
    float dot = 0.f, nn = 0.f;
#pragma unroll
    for (int u = 0; u < NPL; ++u) { dot = fmaf(qv[u], e[u], dot); nn = fmaf(e[u], e[u], nn); }
    dot = warp_sum(dot);
    float z;
    if (P.l2) {
      nn = warp_sum(nn);
      z = dot / fmaxf(sqrtf(nn), P.eps);
    } else {
      z = dot;
    }
    z = (P.neg_ids[n * P.R + r] == pid) ? -5e4f : z / P.temp;
    if (lane == 0) pr[r + 1] = z;  // stash logits, normalised below
    m = fmaxf(m, z);
  }
  __syncwarp();
  float se = 0.f;
  for (int r = lane; r < P.R; r += 32) se += expf(pr[r + 1] - m);
  se = warp_sum(se) + expf(zpos - m);
  const float lse = logf(se);
  for (int r = lane; r < P.R; r += 32) pr[r + 1] = expf(pr[r + 1] - m - lse);
  if (lane == 0) {
    pr[0] = expf(zpos - m - lse);
    P.loss_rows[n] = -(zpos - m - lse);
  }
}

template <int NPL>
__global__ void __launch_bounds__(SSL_WARPS * 32) ssl_bwd_kernel(SslP P) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t) blockIdx.x * SSL_WARPS + (threadIdx.x >> 5);
  if (n >= P.n_rows) return;
  float qv[NPL], e[NPL], dqa[NPL];
  const float g = P.g[n];
  const float* pr = P.probs + n * (int64_t) (P.R + 1);
  const float dzpos = g * (pr[0] - 1.0f) / P.temp;
#pragma unroll
  for (int u = 0; u < NPL; ++u) {
    const int c = lane + 32 * u;
    qv[u] = c < P.D ? P.q[n * P.ldq + c] : 0.f;
    const float pv = c < P.D ? P.p[n * P.ldp + c] : 0.f;
    dqa[u] = dzpos * pv;
    if (c < P.D) P.dp[n * (int64_t) P.D + c] = dzpos * qv[u];
  }
  const int64_t pid = P.pos_ids[n];
  if (g != 0.f) {
    for (int r = 0; r < P.R; ++r) {
      if (P.neg_ids[n * P.R + r] == pid) continue;  // constant -5e4: no gradient (warp-uniform)
      const float dl = g * pr[r + 1] / P.temp;
      if (dl == 0.f) continue;                       // warp-uniform
      const int64_t i0 = P.idx0[n * P.R + r];
      const int64_t i1 = P.idx1 ? P.idx1[n * P.R + r] : 0;
      gather_row<NPL>(P, i0, i1, lane, e);
      float a = 1.0f, bcoef = 0.f;  // d e = dl * (a * q - bcoef * e) ; d q += dl * a * e
      if (P.l2) {
        float dot = 0.f, nn = 0.f;
#pragma unroll
        for (int u = 0; u < NPL; ++u) { dot = fmaf(qv[u], e[u], dot); nn = fmaf(e[u], e[u], nn); }
        dot = warp_sum(dot);
        nn = warp_sum(nn);
        const float nrm = sqrtf(nn);
        if (nrm > P.eps) { a = 1.0f / nrm; bcoef = dot / (nrm * nn); }
        else { a = 1.0f / P.eps; bcoef = 0.f; }      // clamp branch: denominator is constant
      }
#pragma unroll
      for (int u = 0; u < NPL; ++u) {
        const int c = lane + 32 * u;
        dqa[u] = fmaf(dl * a, e[u], dqa[u]);
        const float de = dl * (a * qv[u] - bcoef * e[u]);
        if (c < P.d0) atomicAdd(P.dt0 + i0 * (int64_t) P.d0 + c, de);
        else if (c < P.D) atomicAdd(P.dt1 + i1 * (int64_t) P.d1 + (c - P.d0), de);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < NPL; ++u) {
    const int c = lane + 32 * u;
    if (c < P.D) P.dq[n * (int64_t) P.D + c] = dqa[u];
  }
}

// ---- vector path: one table, D = 128 * NV, 16-byte aligned rows --------------------------------
// Lane l owns columns 128 v + 4 l .. + 3 (float4 gathers: two 512-byte requests per 1 KiB row
// instead of eight 128-byte ones).  The forward reduces the dot products of 32 negatives together
// with a transposing butterfly (31 shuffles for 32 sums instead of 160): afterwards lane j holds
// the logit of negative r0 + j, which is also the layout the softmax and the stores want.
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc); return fmaf(a.w, b.w, acc);
}
// in: v[j] = this lane's partial sum for item j.  out: v[0] on lane j = sum over lanes of item j.
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = up ? v[i + s] : v[i];
      const float send = up ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

template <int NV, bool L2>
__global__ void __launch_bounds__(SSL_WARPS * 32) ssl_fwd_vec_kernel(SslP P) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t) blockIdx.x * SSL_WARPS + (threadIdx.x >> 5);
  if (n >= P.n_rows) return;
  float4 q[NV];
  float pd = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    q[v] = *reinterpret_cast<const float4*>(P.q + n * P.ldq + 128 * v + 4 * lane);
    pd = dot4(q[v], *reinterpret_cast<const float4*>(P.p + n * P.ldp + 128 * v + 4 * lane), pd);
  }
  const float zpos = warp_sum(pd) / P.temp;
  const int64_t pid = P.pos_ids[n];
  float* pr = P.probs + n * (int64_t) (P.R + 1);
  const int64_t* idx = P.idx0 + n * (int64_t) P.R;
  const int64_t* nid = P.neg_ids + n * (int64_t) P.R;
  float m = zpos;
  for (int r0 = 0; r0 < P.R; r0 += 32) {
    float dots[32], nrm[L2 ? 32 : 1];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float d = 0.f, nn = 0.f;
      if (r0 + j < P.R) {                            // warp-uniform
        const float* row = P.t0 + idx[r0 + j] * P.ldt0 + 4 * lane;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const float4 e = __ldg(reinterpret_cast<const float4*>(row + 128 * v));
          d = dot4(q[v], e, d);
          if (L2) nn = dot4(e, e, nn);
        }
      }
      dots[j] = d;
      if (L2) nrm[j] = nn;
    }
    float z = transpose_reduce32(dots, lane);
    if (L2) {
      float (&nr)[32] = reinterpret_cast<float (&)[32]>(nrm);
      z = z / fmaxf(sqrtf(transpose_reduce32(nr, lane)), P.eps);
    }
    const int r = r0 + lane;
    if (r < P.R) {
      z = (nid[r] == pid) ? -5e4f : z / P.temp;
      pr[r + 1] = z;                                 // stash logits, normalised below
      m = fmaxf(m, z);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __syncwarp();
  float se = 0.f;
  for (int r = lane; r < P.R; r += 32) se += expf(pr[r + 1] - m);
  se = warp_sum(se) + expf(zpos - m);
  const float lse = logf(se);
  for (int r = lane; r < P.R; r += 32) pr[r + 1] = expf(pr[r + 1] - m - lse);
  if (lane == 0) {
    pr[0] = expf(zpos - m - lse);
    P.loss_rows[n] = -(zpos - m - lse);
  }
}

__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

template <int NV, bool L2>
__global__ void __launch_bounds__(SSL_WARPS * 32) ssl_bwd_vec_kernel(SslP P) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t) blockIdx.x * SSL_WARPS + (threadIdx.x >> 5);
  if (n >= P.n_rows) return;
  const float g = P.g[n];
  const float* pr = P.probs + n * (int64_t) (P.R + 1);
  const float dzpos = g * (pr[0] - 1.0f) / P.temp;
  float4 q[NV], dqa[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = 128 * v + 4 * lane;
    q[v] = *reinterpret_cast<const float4*>(P.q + n * P.ldq + c);
    const float4 pv = *reinterpret_cast<const float4*>(P.p + n * P.ldp + c);
    dqa[v] = make_float4(dzpos * pv.x, dzpos * pv.y, dzpos * pv.z, dzpos * pv.w);
    *reinterpret_cast<float4*>(P.dp + n * (int64_t) P.D + c) =
        make_float4(dzpos * q[v].x, dzpos * q[v].y, dzpos * q[v].z, dzpos * q[v].w);
  }
  const int64_t pid = P.pos_ids[n];
  const int64_t* idx = P.idx0 + n * (int64_t) P.R;
  const int64_t* nid = P.neg_ids + n * (int64_t) P.R;
  if (g != 0.f) {
    for (int r0 = 0; r0 < P.R; r0 += 32) {
      // this lane's negative of the batch: coefficient (0: masked collision / out of range) and row
      const int r = r0 + lane;
      float my_dl = 0.f;
      int64_t my_i = 0;
      if (r < P.R) {
        my_i = idx[r];
        if (nid[r] != pid) my_dl = g * pr[r + 1] / P.temp;
      }
      // four negatives at a time: all eight 16-byte gathers are issued before any is consumed
#pragma unroll 2
      for (int j4 = 0; j4 < 32; j4 += 4) {
        float dl[4];
        int64_t i0[4];
        float4 e[4][NV];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          dl[t] = __shfl_sync(0xffffffffu, my_dl, j4 + t);
          i0[t] = __shfl_sync(0xffffffffu, my_i, j4 + t);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float* row = P.t0 + i0[t] * P.ldt0 + 4 * lane;    // (masked slots point at row 0)
#pragma unroll
          for (int v = 0; v < NV; ++v) e[t][v] = __ldg(reinterpret_cast<const float4*>(row + 128 * v));
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (dl[t] == 0.f) continue;                  // warp-uniform
          float* drow = P.dt0 + i0[t] * (int64_t) P.d0 + 4 * lane;
          float a = 1.0f, bcoef = 0.f;  // d e = dl * (a * q - bcoef * e) ; d q += dl * a * e
          if (L2) {
            float dot = 0.f, nn = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) { dot = dot4(q[v], e[t][v], dot); nn = dot4(e[t][v], e[t][v], nn); }
            dot = warp_sum(dot);
            nn = warp_sum(nn);
            const float nrm = sqrtf(nn);
            if (nrm > P.eps) { a = 1.0f / nrm; bcoef = dot / (nrm * nn); }
            else { a = 1.0f / P.eps; bcoef = 0.f; }    // clamp branch: denominator is constant
          }
          const float da = dl[t] * a, db = dl[t] * bcoef;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const float4 ev = e[t][v];
            dqa[v].x = fmaf(da, ev.x, dqa[v].x); dqa[v].y = fmaf(da, ev.y, dqa[v].y);
            dqa[v].z = fmaf(da, ev.z, dqa[v].z); dqa[v].w = fmaf(da, ev.w, dqa[v].w);
            red_add4(drow + 128 * v, da * q[v].x - db * ev.x, da * q[v].y - db * ev.y,
                     da * q[v].z - db * ev.z, da * q[v].w - db * ev.w);
          }
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v)
    *reinterpret_cast<float4*>(P.dq + n * (int64_t) P.D + 128 * v + 4 * lane) = dqa[v];
}

// single table, D = 128, 256 or 512, everything 16-byte aligned
static bool ssl_vec_ok(const SslP& P, bool bwd) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (P.d1 != 0 || (P.D != 128 && P.D != 256 && P.D != 512)) return false;
  if (!al(P.q) || !al(P.p) || !al(P.t0) || P.ldq % 4 || P.ldp % 4 || P.ldt0 % 4) return false;
  if (bwd && (!al(P.dq) || !al(P.dp) || !al(P.dt0))) return false;
  return true;
}

// table_grad[ids[i], :] += grad[i, :]: one warp per row, 16-byte vector reductions
template <bool VEC>
__global__ void __launch_bounds__(256) rows_scatter_add_kernel(const float* __restrict__ grad, int64_t ld,
                                                               const int64_t* __restrict__ ids,
                                                               float* __restrict__ tg, int64_t n, int D,
                                                               int64_t num_rows, int64_t skip) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t) blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const int64_t id = ids[i];
  if (id == skip || id < 0 || id >= num_rows) return;
  const float* src = grad + i * ld;
  float* dst = tg + id * (int64_t) D;
  if (VEC) {
    for (int c = 4 * lane; c < D; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(src + c);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v.x), "f"(v.y),
                   "f"(v.z), "f"(v.w)
                   : "memory");
    }
  } else {
    for (int c = lane; c < D; c += 32) atomicAdd(dst + c, src[c]);
  }
}

__global__ void __launch_bounds__(256) rows_scale_kernel(float* __restrict__ table, const int64_t* __restrict__ ids,
                                                         int64_t n, int D, int64_t num_rows, int64_t skip,
                                                         float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t) blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const int64_t id = ids[i];
  if (id == skip || id < 0 || id >= num_rows) return;
  float* row = table + id * (int64_t) D;
  for (int c = 4 * lane; c < D; c += 128) {
    float4 v = *reinterpret_cast<float4*>(row + c);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    *reinterpret_cast<float4*>(row + c) = v;
  }
}

// sender side of the table-gradient exchange: one warp per (slot, destination), 16-byte stores
struct RowsDst { float* rows[16]; int64_t* ids[16]; };
__global__ void __launch_bounds__(256) p2p_put_table_rows_kernel(const float* __restrict__ src,
                                                                 const int64_t* __restrict__ ids, int64_t n,
                                                                 int D, int64_t num_rows, int64_t skip,
                                                                 float scale, RowsDst dst, int64_t slot0) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t) blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const int64_t id = ids[i];
  const bool real = id != skip && id >= 0 && id < num_rows;
  if (lane == 0) dst.ids[blockIdx.y][slot0 + i] = real ? id : skip;
  if (!real) return;
  const float* row = src + id * (int64_t) D;
  float* out = dst.rows[blockIdx.y] + (slot0 + i) * (int64_t) D;
  for (int c = 4 * lane; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    *reinterpret_cast<float4*>(out + c) = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
  }
}

static int make(const grb_ssl_args* a, SslP* P, bool bwd) {
  GRB_REQUIRE(a != nullptr, GRB_ERR_INVALID_ARG, "sampled_softmax: null args");
  GRB_REQUIRE(a->dtype == GRB_F32, GRB_ERR_UNSUPPORTED, "sampled_softmax: only fp32 tables");
  GRB_REQUIRE(a->n_rows >= 0 && a->R > 0 && a->D > 0 && (a->D <= 256 || (a->D == 512 && a->d1 == 0)) && a->d0 > 0 &&
                  a->d0 + a->d1 == a->D && a->d1 >= 0,
              GRB_ERR_INVALID_ARG, "sampled_softmax: bad sizes (D=%d d0=%d d1=%d R=%d)", a->D,
              a->d0, a->d1, a->R);
  GRB_REQUIRE(a->q && a->p && a->table0 && a->idx0 && a->pos_ids && a->neg_ids && a->probs,
              GRB_ERR_INVALID_ARG, "sampled_softmax: null tensor");
  GRB_REQUIRE((a->d1 == 0) == (a->table1 == nullptr), GRB_ERR_INVALID_ARG,
              "sampled_softmax: table1/d1 mismatch");
  GRB_REQUIRE(a->d1 == 0 || a->idx1, GRB_ERR_INVALID_ARG, "sampled_softmax: idx1 is null");
  GRB_REQUIRE(a->temperature > 0.f, GRB_ERR_INVALID_ARG, "sampled_softmax: temperature <= 0");
  if (!bwd) GRB_REQUIRE(a->loss_rows, GRB_ERR_INVALID_ARG, "sampled_softmax_fwd: loss_rows null");
  else
    GRB_REQUIRE(a->g && a->dq && a->dp && a->dtable0 && (a->d1 == 0 || a->dtable1),
                GRB_ERR_INVALID_ARG, "sampled_softmax_bwd: null gradient buffer");
  P->n_rows = a->n_rows; P->R = a->R; P->D = a->D; P->d0 = a->d0; P->d1 = a->d1;
  P->l2 = a->l2_norm; P->eps = a->l2_eps; P->temp = a->temperature;
  P->q = (const float*) a->q; P->ldq = a->ldq_; P->p = (const float*) a->p; P->ldp = a->ldp;
  P->t0 = (const float*) a->table0; P->ldt0 = a->ldt0;
  P->t1 = (const float*) a->table1; P->ldt1 = a->ldt1;
  P->idx0 = a->idx0; P->idx1 = a->idx1; P->pos_ids = a->pos_ids; P->neg_ids = a->neg_ids;
  P->loss_rows = a->loss_rows; P->probs = a->probs; P->g = a->g; P->dq = a->dq; P->dp = a->dp;
  P->dt0 = a->dtable0; P->dt1 = a->dtable1;
  return GRB_OK;
}

}  // namespace grb

using namespace grb;

extern "C" {

int grb_sampled_softmax_fwd(const grb_ssl_args* a, grb_stream_t stream) {
  SslP P{};
  int rc = make(a, &P, false);
  if (rc != GRB_OK) return rc;
  if (P.n_rows == 0) return GRB_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned) ceil_div(P.n_rows, SSL_WARPS);
  if (ssl_vec_ok(P, false)) {
    if (P.D == 128) {
      if (P.l2) ssl_fwd_vec_kernel<1, true><<<grid, SSL_WARPS * 32, 0, st>>>(P);
      else ssl_fwd_vec_kernel<1, false><<<grid, SSL_WARPS * 32, 0, st>>>(P);
    } else if (P.D == 512) {   // C5: D = 512
      if (P.l2) ssl_fwd_vec_kernel<4, true><<<grid, SSL_WARPS * 32, 0, st>>>(P);
      else ssl_fwd_vec_kernel<4, false><<<grid, SSL_WARPS * 32, 0, st>>>(P);
    } else {
      if (P.l2) ssl_fwd_vec_kernel<2, true><<<grid, SSL_WARPS * 32, 0, st>>>(P);
      else ssl_fwd_vec_kernel<2, false><<<grid, SSL_WARPS * 32, 0, st>>>(P);
    }
    GRB_LAUNCH_OK();
    return GRB_OK;
  }
  if (P.D <= 64) ssl_fwd_kernel<2><<<grid, SSL_WARPS * 32, 0, st>>>(P);
  else if (P.D <= 128) ssl_fwd_kernel<4><<<grid, SSL_WARPS * 32, 0, st>>>(P);
  else ssl_fwd_kernel<8><<<grid, SSL_WARPS * 32, 0, st>>>(P);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_sampled_softmax_bwd(const grb_ssl_args* a, grb_stream_t stream) {
  SslP P{};
  int rc = make(a, &P, true);
  if (rc != GRB_OK) return rc;
  if (P.n_rows == 0) return GRB_OK;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned) ceil_div(P.n_rows, SSL_WARPS);
  if (ssl_vec_ok(P, true)) {
    if (P.D == 128) {
      if (P.l2) ssl_bwd_vec_kernel<1, true><<<grid, SSL_WARPS * 32, 0, st>>>(P);
      else ssl_bwd_vec_kernel<1, false><<<grid, SSL_WARPS * 32, 0, st>>>(P);
    } else if (P.D == 512) {   // C5: D = 512
      if (P.l2) ssl_bwd_vec_kernel<4, true><<<grid, SSL_WARPS * 32, 0, st>>>(P);
      else ssl_bwd_vec_kernel<4, false><<<grid, SSL_WARPS * 32, 0, st>>>(P);
    } else {
      if (P.l2) ssl_bwd_vec_kernel<2, true><<<grid, SSL_WARPS * 32, 0, st>>>(P);
      else ssl_bwd_vec_kernel<2, false><<<grid, SSL_WARPS * 32, 0, st>>>(P);
    }
    GRB_LAUNCH_OK();
    return GRB_OK;
  }
  if (P.D <= 64) ssl_bwd_kernel<2><<<grid, SSL_WARPS * 32, 0, st>>>(P);
  else if (P.D <= 128) ssl_bwd_kernel<4><<<grid, SSL_WARPS * 32, 0, st>>>(P);
  else ssl_bwd_kernel<8><<<grid, SSL_WARPS * 32, 0, st>>>(P);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_rows_scatter_add(const float* grad, int64_t ld_grad, const int64_t* ids, float* table_grad,
                         int64_t n, int32_t D, int64_t num_rows, int64_t skip_id,
                         grb_stream_t stream) {
  using namespace grb;
  GRB_REQUIRE(n >= 0 && D > 0 && num_rows > 0 && ld_grad >= D, GRB_ERR_INVALID_ARG,
              "rows_scatter_add: bad sizes");
  if (n == 0) return GRB_OK;
  GRB_REQUIRE(grad && ids && table_grad, GRB_ERR_INVALID_ARG, "rows_scatter_add: null tensor");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned) ceil_div(n, 8);
  const bool vec = (D % 4 == 0) && (ld_grad % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(table_grad)) & 15) == 0;
  if (vec) rows_scatter_add_kernel<true><<<grid, 256, 0, st>>>(grad, ld_grad, ids, table_grad, n, D, num_rows, skip_id);
  else rows_scatter_add_kernel<false><<<grid, 256, 0, st>>>(grad, ld_grad, ids, table_grad, n, D, num_rows, skip_id);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_rows_scale(float* table, const int64_t* ids, int64_t n, int32_t D, int64_t num_rows,
                   int64_t skip_id, float scale, grb_stream_t stream) {
  using namespace grb;
  GRB_REQUIRE(table && ids && n >= 0 && D > 0 && D % 4 == 0 && num_rows > 0 &&
                  (reinterpret_cast<uintptr_t>(table) & 15) == 0,
              GRB_ERR_INVALID_ARG, "rows_scale: bad arguments (D must be a multiple of 4)");
  if (n == 0) return GRB_OK;
  rows_scale_kernel<<<(unsigned) ceil_div(n, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      table, ids, n, D, num_rows, skip_id, scale);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

int grb_p2p_put_table_rows(const float* grad_table, const int64_t* ids, int64_t n, int32_t D,
                           int64_t num_rows, int64_t skip_id, float scale, void* const* dst_rows,
                           void* const* dst_ids, int32_t n_dst, int64_t slot_offset,
                           grb_stream_t stream) {
  using namespace grb;
  GRB_REQUIRE(grad_table && ids && dst_rows && dst_ids && n >= 0 && D > 0 && D % 4 == 0 && num_rows > 0 &&
                  n_dst > 0 && n_dst <= 16 && slot_offset >= 0,
              GRB_ERR_INVALID_ARG, "p2p_put_table_rows: bad arguments (D must be a multiple of 4)");
  if (n == 0) return GRB_OK;
  RowsDst d{};
  for (int i = 0; i < n_dst; ++i) {
    GRB_REQUIRE(dst_rows[i] != nullptr && dst_ids[i] != nullptr &&
                    (reinterpret_cast<uintptr_t>(dst_rows[i]) & 15) == 0,
                GRB_ERR_INVALID_ARG, "p2p_put_table_rows: bad destination");
    d.rows[i] = reinterpret_cast<float*>(dst_rows[i]);
    d.ids[i] = reinterpret_cast<int64_t*>(dst_ids[i]);
  }
  p2p_put_table_rows_kernel<<<dim3((unsigned) ceil_div(n, 8), (unsigned) n_dst), 256, 0,
                              reinterpret_cast<cudaStream_t>(stream)>>>(grad_table, ids, n, D, num_rows,
                                                                        skip_id, scale, d, slot_offset);
  GRB_LAUNCH_OK();
  return GRB_OK;
}

}
