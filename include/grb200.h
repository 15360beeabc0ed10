/*
 * grb200.h — C ABI of libgrb200.so: B200 (sm_100a) kernels for the two hot paths of the
 * generative-recommender pipeline (HSTU jagged attention, candidate retrieval).
 *
 * Every entry point is what the reference's Python would bind for that call site
 * (reference paths are relative to /root/reference/src/generative_recommenders_pl/):
 * raw device pointers, explicit sizes / strides / dtype codes, a cudaStream_t.
 * No torch types, no C++ exceptions: functions return GRB_OK (0) or a negative error
 * code; grb_last_error_string() gives the thread-local message.  All functions are
 * stream-ordered and never synchronise the device or read device memory on the host.
 * The caller owns every buffer (outputs and workspaces included).
 */
#ifndef GRB200_H_
#define GRB200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRB_VERSION 100

typedef void* grb_stream_t; /* cudaStream_t */

enum grb_status {
  GRB_OK = 0,
  GRB_ERR_INVALID_ARG = -1,
  GRB_ERR_UNSUPPORTED = -2,
  GRB_ERR_CUDA = -3,
  GRB_ERR_WORKSPACE = -4
};

enum grb_dtype { GRB_F32 = 0, GRB_BF16 = 1 };

int grb_version(void);
const char* grb_last_error_string(void);
/* Number of kernels this library has launched in this process (for bench.py's gpu_launches). */
int64_t grb_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * a1  models/utils/ops.py:18-38  asynchronous_complete_cumsum  (replaces torch.ops.fbgemm.* :27)
 *     offsets[0] = 0, offsets[i+1] = sum(lengths[0..i]);  index_bits = 32 | 64.  Bit-exact.
 * ------------------------------------------------------------------------------------------- */
int grb_complete_cumsum(const void* lengths, void* offsets, int64_t B, int index_bits,
                        grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a2  models/utils/ops.py:41-64  dense_to_jagged  (replaces torch.ops.fbgemm.dense_to_jagged :51)
 *     dense (B, N, row_bytes) -> jagged (T, row_bytes), out[off[b]+i] = dense[b,i], i < n_b.
 *     Rows are opaque bytes (any dtype).  n_b = off[b+1]-off[b] is clipped to N.
 *     dense_batch_stride_bytes: distance between dense[b] and dense[b+1] (0 = N*row_bytes), so
 *     the slices x[:, :-1] / x[:, 1:] of generative_recommenders.py:409-424 need no copy.
 *     Also the backward of a3.
 * ------------------------------------------------------------------------------------------- */
int grb_dense_to_jagged(const void* dense, const void* offsets, void* jagged, int64_t B,
                        int64_t N, int64_t row_bytes, int64_t dense_batch_stride_bytes,
                        int index_bits, grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a3  models/utils/ops.py:67-114  jagged_to_padded_dense  (replaces torch.ops.fbgemm.* :87)
 *     jagged (T, row_bytes) -> dense (B, N, row_bytes); rows i >= n_b are filled with the
 *     elem_bytes-wide pattern `pad_pattern` (1, 2, 4 or 8 bytes; host memory, copied by value).
 *     Also the backward of a2 (pad = 0).
 * ------------------------------------------------------------------------------------------- */
int grb_jagged_to_padded_dense(const void* jagged, const void* offsets, void* dense, int64_t B,
                               int64_t N, int64_t row_bytes, int64_t dense_batch_stride_bytes,
                               const void* pad_pattern, int elem_bytes, int index_bits,
                               grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a10 models/utils/ops.py:171-187  get_current_embeddings: out[b] = dense[b, lengths[b]-1]
 *     (index wraps like the reference's flattened gather: lengths[b]==0 reads row b*N-1).
 *     scatter=1 runs the transpose (backward): dense[b, lengths[b]-1] += out[b] is NOT needed
 *     because each b hits a distinct row; it writes instead (dense must be pre-zeroed).
 * ------------------------------------------------------------------------------------------- */
int grb_gather_last_rows(const void* dense, const void* lengths, void* out, int64_t B, int64_t N,
                         int64_t row_bytes, int index_bits, int scatter, grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a4+a5  models/sequential_encoders/hstu.py:71-128 (bias) and :134-205 (attention), fused and
 *     jagged (no padded intermediates):  for each sequence b, head h, 0 <= j <= i < n_b
 *        S = q_h[i].k_h[j] + pos_w[N-1+j-i] + ts_w[bucket(|ts[b,i+1] - ts[b,j]|)]
 *        out_h[i] = sum_j SiLU(S)/N * v_h[j]
 *     bucket(d) = #{t : thresholds[t] <= d}  (thresholds: the reference's bucketization_fn
 *     tabulated by the host module; monotone int64 table of num_buckets entries),
 *     ts index i+1 == N reads ts[b, N-1] (hstu.py:113-115).
 *     q,k: (T, H*dqk) with row stride ldq/ldk elements; v: (T, H*dv) stride ldv; out (T, H*dv)
 *     stride ldo.  dtype GRB_F32 (any dqk,dv <= 256; CUDA-core path, fp32 parity) or GRB_BF16
 *     (dqk = dv = 64 ; tcgen05/TMEM/TMA path, fp32 accumulate).
 *     timestamps == NULL  <=>  all_timestamps is None in the reference (no bias at all).
 * ------------------------------------------------------------------------------------------- */
typedef struct grb_hstu_attn_args {
  int64_t B;            /* sequences */
  int64_t N;            /* padded max length: the 1/N scale and the pos_w origin (hstu.py:150,193) */
  int64_t T;            /* total jagged rows = offsets[B] */
  int64_t max_len;      /* upper bound of n_b known to the host (<= N); sizes the grid */
  int32_t H, dqk, dv;
  int32_t dtype;        /* grb_dtype of q,k,v,out,(dout,dq,dk,dv_grad) */
  int32_t index_bits;   /* offsets: 32 | 64 */
  int32_t num_buckets;  /* ts_w has num_buckets+1 entries; thresholds has num_buckets */
  const void* q; const void* k; const void* v;
  int64_t ldq, ldk, ldv;
  const void* offsets;            /* (B+1) */
  const int64_t* timestamps;      /* (B, N) or NULL */
  const float* ts_w;              /* (num_buckets+1) fp32 */
  const float* pos_w;             /* (2N-1) fp32 */
  const int64_t* bucket_thresholds; /* (num_buckets) ascending */
  void* out; int64_t ldo;         /* forward output (T, H*dv) */
  /* backward only */
  const void* dout; int64_t lddo; /* (T, H*dv) */
  void* dq; void* dk; void* dv_grad; int64_t lddq, lddk, lddv; /* same dtype as q/k/v */
  float* dq_accum;                /* fp32 (T, H*dqk) contiguous workspace, zero-filled by caller */
  float* d_ts_w;                  /* fp32 (d_bias_copies, num_buckets+1), accumulated (+=) */
  float* d_pos_w;                 /* fp32 (d_bias_copies, 2N-1), accumulated (+=) */
  int32_t d_bias_copies;          /* >= 1 (0 is read as 1): CTAs spread their bias-gradient atomics
                                     over this many private copies; the caller sums the copies */
  /* optional: grb_bucket_octaves(thresholds) computed once by the host (else built per CTA) */
  const uint32_t* bucket_octaves; /* device, GRB_OCTAVE_WORDS uint32 */
  /* optional: grb_hstu_bucket_tiles output for THIS (offsets, timestamps, thresholds): the bucket
   * indices are layer-, head- and direction-independent, so one tabulation serves every attention
   * launch of a step.  bucket_cache_max_len = the max_len it was built with. */
  const uint8_t* bucket_cache;
  int64_t bucket_cache_max_len;
  /* optional, short sequences (max_len <= 256): the item schedule of this batch
   * (grb_hstu_short_schedule).  When set, together with a MASKED bucket cache
   * (grb_hstu_bucket_tiles_masked; needed with or without timestamps), the launch takes the
   * short-sequence kernels (csrc/hstu_attn_short.cu): persistent, two CTAs per SM, one work item per
   * (sequence, head).  Backward with max_len > 128 also needs dq_accum (T, H*dqk) fp32 as plain
   * scratch (no zero fill). */
  const int32_t* short_schedule;
  int32_t bucket_cache_masked;    /* bucket_cache came from grb_hstu_bucket_tiles_masked */
  /* short-sequence kernels only: T counts rows past offsets[B] (a fixed-size row bucket) and the launch
   * itself writes zeros to those rows of out (forward) / dq, dk, dv_grad (backward) — the kernels never
   * touch them otherwise, and the row-wise consumers (GEMMs, column sums) read them.  Rows 16-byte
   * aligned.  0: the caller has zero-filled them. */
  int32_t zero_tail_rows;
} grb_hstu_attn_args;

/* Host helper: the integer bucketing table the tcgen05 kernels use, from the ascending threshold
 * table (HOST memory, num_buckets entries).  out: GRB_OCTAVE_WORDS uint32 of host memory:
 * 32 records {base, t1, t2, t3} (bucket(d) = base + (d>=t1)+(d>=t2)+(d>=t3) for d in
 * [2^e, 2^(e+1))), then word 128 = 1 if the table cannot express the thresholds, word 129 =
 * bucket(0). */
#define GRB_OCTAVE_WORDS 130
int grb_bucket_octaves(const int64_t* thresholds_host, int32_t num_buckets, uint32_t* out_host);

/* Bucket-index tiles, once per batch (hstu.py:113-123 hoisted out of the layers).  cache must hold
 * grb_hstu_bucket_cache_bytes(B, max_len) bytes; layout in csrc/hstu_bucket_cache.cu. */
int64_t grb_hstu_bucket_cache_bytes(int64_t B, int64_t max_len);
int grb_hstu_bucket_tiles(const void* offsets, int index_bits, const int64_t* timestamps, int64_t B,
                          int64_t N, int64_t max_len, const int64_t* thresholds, int32_t num_buckets,
                          const uint32_t* octaves, void* cache, grb_stream_t stream);

/* Short-sequence attention (max_len <= 256), per batch:
 *   grb_hstu_bucket_tiles_masked: as grb_hstu_bucket_tiles, but pairs outside the causal triangle or
 *     past the end of the sequence hold 255 (the kernels map it to a bias of -15000, where the
 *     approximate tanh saturates, so masked scores contribute exact zeros); timestamps == NULL gives
 *     mask-only tiles (valid pairs 0).  num_buckets <= 254.
 *   grb_hstu_short_schedule: schedule (B + 2) int32: [0] = sequences with n > 0, [1] = of those with
 *     n > 128, [2..] their indices, the n > 128 ones first: the item order of the persistent kernels. */
int grb_hstu_bucket_tiles_masked(const void* offsets, int index_bits, const int64_t* timestamps, int64_t B,
                                 int64_t N, int64_t max_len, const int64_t* thresholds, int32_t num_buckets,
                                 const uint32_t* octaves, void* cache, grb_stream_t stream);
int grb_hstu_short_schedule(const void* offsets, int index_bits, int64_t B, int64_t N, int32_t* schedule,
                            grb_stream_t stream);

int grb_hstu_attn_fwd(const grb_hstu_attn_args* a, grb_stream_t stream);
int grb_hstu_attn_bwd(const grb_hstu_attn_args* a, grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a5/a6 incremental form  hstu.py:151-177 (+ :293-298, :321-322, :397-401, :415-418): the
 *     delta_x_offsets + cache path.  The reference index_copy_'s the new q / k rows into the cached
 *     padded tensors, recomputes the whole (B, H, N, N) attention and keeps one row per sequence;
 *     this computes that row only (it depends on keys / values 0..p_b alone, causal mask :667):
 *        out[b, h] = sum_{j <= p_b} SiLU(q[b,h].k_cache[b,j,h] + pos_w[N-1+j-p_b]
 *                                         + ts_w[bucket(|ts[b,p_b+1] - ts[b,j]|)]) / N * v[off_b+j, h]
 *     q (B, H*dqk) row stride ldq: the new query rows; k_cache (B, N, H*dqk) padded, row stride ldk
 *     (already holding the new key at [b, p_b]); v (T, H*dv) jagged, stride ldv (new value at
 *     off_b + p_b); positions (B) = delta_x_offsets[1] (int32 | int64, pos_bits); out (B, H*dv)
 *     stride ldo.  A position outside [0, n_b) yields a zero row (the reference's padded row).
 *     Forward only (inference).  HBM-bound: (p_b+1) * H * (dqk+dv) elements per sequence.
 * ------------------------------------------------------------------------------------------- */
typedef struct grb_hstu_attn_decode_args {
  int64_t B, N;
  int32_t H, dqk, dv;
  int32_t dtype;        /* grb_dtype of q, k_cache, v, out */
  int32_t index_bits;   /* offsets: 32 | 64 */
  int32_t pos_bits;     /* positions: 32 | 64 */
  int32_t num_buckets;
  const void* q; int64_t ldq;
  const void* k_cache; int64_t ldk;
  const void* v; int64_t ldv;
  const void* offsets;            /* (B+1) */
  const void* positions;          /* (B) */
  const int64_t* timestamps;      /* (B, N) or NULL */
  const float* ts_w; const float* pos_w; const int64_t* bucket_thresholds;
  void* out; int64_t ldo;
} grb_hstu_attn_decode_args;
int grb_hstu_attn_decode(const grb_hstu_attn_decode_args* a, grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * f2  the steps either side of the HSTU stack, jagged (csrc/input_path.cu):
 *     embeddings/embeddings.py:94-97 + preprocessors/learnable_positional_embedding.py:42-58 +
 *     sequential_encoders/hstu.py:502:
 *        io[t, :] = dropout_p(table[ids[b, i], :] * scale + pos[i, :]) * (ids[b, i] != 0)
 *     for the jagged row t = offsets[b] + i, i < n_b; rows >= offsets[B] (row buckets) are zero.
 *     io (rows, D) in dtype (fp32 | bf16).  Dropout: counter-based (Philox4x32-10) keyed by *seed
 *     (device int64), counter = element index / 4; p_drop == 0: no dropout, seed may be NULL.
 *     Backward: io holds d(io); d_table (V, D) and d_pos (N, D) fp32 are accumulated (+=) with
 *     16-byte vector reds (either may be NULL); the same seed regenerates the mask.
 *     D % 4 == 0, rows 16-byte aligned.
 *     grb_l2norm_cast_fwd/bwd: postprocessors.py:47-55 on the encoder's jagged rows in the compute
 *     dtype: y fp32 = x / max(||x||, eps) with x fp32 | bf16; backward writes dx in x's dtype.
 * ------------------------------------------------------------------------------------------- */
typedef struct grb_jagged_input_args {
  int64_t B, N, V, rows;
  int32_t D, dtype, index_bits, reserved;
  const float* table; int64_t ldt;
  const int64_t* ids;              /* (B, N) */
  const void* offsets;             /* (B+1) */
  const float* pos; int64_t ldp;   /* (>= N, D) */
  float scale, p_drop;
  const int64_t* seed;
  void* io; int64_t ldio;
  float* d_table; float* d_pos;
} grb_jagged_input_args;
/* rows [offsets[B], min(rows, offsets[B] + max_tail_rows)) of n_mats (rows, row_bytes) matrices := 0
 * (row stride ld_bytes, matrices mat_stride_bytes apart; everything 16-byte aligned).  Fixed row buckets:
 * the padding rows past offsets[B] are the only ones the attention kernels leave unwritten. */
int grb_zero_tail_rows(void* base, int64_t ld_bytes, int64_t rows, int64_t row_bytes, int32_t n_mats,
                       int64_t mat_stride_bytes, const void* offsets, int32_t index_bits, int64_t B,
                       int64_t max_tail_rows, grb_stream_t stream);

/* b5  negatives_samples/negative_sampler.py:208-211  in-batch draw + id gather in one pass:
 *     offsets[i] = Philox4x32-10(seed, i / 2) (62 bits) mod max(count[0], 1);  ids[i] = cached_ids[offsets[i]].
 *     seed, count: device int64 scalars (no host read).  Same distribution as randint(0, count), not
 *     the same stream. */
int grb_draw_negatives(const int64_t* seed, const int64_t* count, const int64_t* cached_ids, int64_t n,
                       int64_t* offsets, int64_t* ids, grb_stream_t stream);

/* b5  negatives_samples/negative_sampler.py:187-196: the de-duplicated id list of the in-batch cache,
 *     torch.unique(ids[presences]) (ascending: the draw of :208-211 indexes it), for a small id space, in two
 *     launches and without a device sync.  ids (B, N) int64; the valid ids of row b are its first
 *     offsets[b+1] - offsets[b] + rows_extra entries (id 0 = padding; ids >= n_flags are skipped).
 *     flags: persistent int32[n_flags + ceil(n_flags / 1024)] (membership table + one counter per 1024
 *     ids), zero-initialised ONCE by the caller and never cleared by it (entries are stamped with epoch[0],
 *     a device int32 starting at 1 that this call increments; the counters are left at zero).  n_flags <= 2^22.
 *     uniq[0 .. count) = the distinct ids ascending, uniq[count .. n_out) = 0, count[0] = their number. */
int grb_inbatch_distinct_ids(const int64_t* ids, int64_t B, int64_t N, const void* offsets, int32_t index_bits,
                             int32_t rows_extra, int32_t* flags, int64_t n_flags, int32_t* epoch, int64_t* uniq,
                             int64_t n_out, int64_t* count, grb_stream_t stream);

int grb_jagged_input_fwd(const grb_jagged_input_args* a, grb_stream_t stream);
int grb_jagged_input_bwd(const grb_jagged_input_args* a, grb_stream_t stream);
int grb_l2norm_cast_fwd(const void* x, int64_t ldx, int dtype, float* y, int64_t ldy, float* inv, int64_t rows,
                        int64_t W, float eps, grb_stream_t stream);
int grb_l2norm_cast_bwd(const float* y, int64_t ldy, const float* dy, int64_t lddy, const float* inv, void* dx,
                        int64_t lddx, int dtype, int64_t rows, int64_t W, grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * f1  hstu.py:302-320 (UVQK projection + SiLU) and :404-413 (output projection + bias + residual),
 *     forward and backward, as ONE tcgen05 GEMM with fused epilogues (csrc/proj_gemm.cu):
 *        C[M, N] = epilogue(A[M, K] * B[K, N]),  bf16 operands, fp32 accumulation.
 *     a_mn = 0: A stored (M, K) row-major, row stride lda;  a_mn = 1: A stored (K, M) row-major (a
 *     transposed operand read in place: the weight gradients contract over the tokens).
 *     b_mn = 0: B stored (N, K) row-major (nn.Linear weight);  b_mn = 1: B stored (K, N) row-major.
 *     N % 256 == 0; K % 64 == 0 unless F32_ADD; all pointers / row strides 16-byte aligned.
 *     Epilogues: PLAIN    out0 (M, N) bf16 = C
 *                SILU2    out0 = C (bf16), out1 = SiLU(out0) (bf16)
 *                BIAS_RES out0 = bf16(C + bias[n] + res[m, n])   bias fp32 (N) or NULL, res bf16 or NULL
 *                F32_ADD  out0 (M, N) fp32 += C   (split-K over the SMs, vector reds; caller zero-fills)
 *     grb_colsum_bf16: out[W] (fp32) += column sums of a bf16 (rows, W) matrix (bias gradient).
 * ------------------------------------------------------------------------------------------- */
enum { GRB_GEMM_EPI_PLAIN = 0, GRB_GEMM_EPI_SILU2 = 1, GRB_GEMM_EPI_BIAS_RES = 2, GRB_GEMM_EPI_F32_ADD = 3 };
typedef struct grb_proj_gemm_args {
  int64_t M, N, K;
  int32_t a_mn, b_mn, epi, reserved;
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* out0; int64_t ldo0;
  void* out1; int64_t ldo1;
  const float* bias;
  const void* res; int64_t ldres;
} grb_proj_gemm_args;
int grb_proj_gemm(const grb_proj_gemm_args* a, grb_stream_t stream);
int grb_colsum_bf16(const void* x, int64_t ldx, int64_t rows, int32_t W, float* out, grb_stream_t stream);
/* out_a[wa] / out_b[wb] = column sums of the fp32 matrices a (rows, wa) / b (rows, wb), one launch: sums the
 * privatised copies of d ts_w and d pos_w that grb_hstu_attn_bwd fills (d_bias_copies of them). */
int grb_colsum_f32_pair(const float* a, int32_t wa, float* out_a, const float* b, int32_t wb, float* out_b,
                        int32_t rows, grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a6  hstu.py:304-320  the activation between the UVQK projection and its consumers:
 *     F.silu(batched_mm_output) followed by torch.split into u, v, q, k.
 *     grb_silu_fwd: y = x / (1 + exp(-x)), (rows, W), fp32 math.
 *     grb_silu_split_bwd: the consumers return n_blocks (<= 4) separate gradients grads[i] of shape
 *     (rows, widths[i]) (row stride ld_grads[i]; NULL = zero), the column blocks of the activation in
 *     order; dx (rows, sum widths) = cat(grads) * silu'(x) in ONE pass (autograd: cat, then
 *     silu_backward).  grads / ld_grads / widths are HOST arrays.  Rows 16-byte aligned, widths
 *     multiples of 16 bytes.
 * ------------------------------------------------------------------------------------------- */
int grb_silu_fwd(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int32_t W,
                 int32_t dtype, grb_stream_t stream);
int grb_silu_split_bwd(const void* x, int64_t ldx, int32_t n_blocks, const void* const* grads,
                       const int64_t* ld_grads, const int32_t* widths, void* dx, int64_t lddx,
                       int64_t rows, int32_t dtype, grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a6  hstu.py:258-264,402  y = gate * LayerNorm_W(x; eps, no affine)   (gate == NULL: y = LN(x))
 *     x,gate,y: (rows, W) row strides ldx/ldg/ldy elements.  mean/rstd (rows) fp32 are saved
 *     for backward.  Backward: dx, dgate from dy (dgate == NULL when gate == NULL).
 * ------------------------------------------------------------------------------------------- */
int grb_ln_gate_fwd(const void* x, int64_t ldx, const void* gate, int64_t ldg, void* y,
                    int64_t ldy, float* mean, float* rstd, int64_t rows, int64_t W, float eps,
                    int dtype, grb_stream_t stream);
int grb_ln_gate_bwd(const void* x, int64_t ldx, const void* gate, int64_t ldg, const void* dy,
                    int64_t lddy, const float* mean, const float* rstd, void* dx, int64_t lddx,
                    void* dgate, int64_t lddg, int64_t rows, int64_t W, int dtype,
                    grb_stream_t stream);

/* The same pair with the two elementwise neighbours of the STU layer folded in (struct form):
 *   p_drop > 0 (forward and backward): dropout on the output, hstu.py:404-408 dropout(u * norm(a)) in front
 *     of the output projection.  The mask is drawn inside the kernels (Philox4x32-10 keyed by the device
 *     int64 `seed` the caller draws per step, counter = (row, 8-element chunk, `salt` of the call site)) and
 *     regenerated in the backward: no mask tensor, no dropout / masked-scale passes.  bf16 rows of 256 /
 *     512 / 1024 elements only (GRB_ERR_UNSUPPORTED otherwise: apply the dropout outside).
 *   res (backward): dx = LayerNorm backward + res, the gradient reaching x through the residual branch
 *     (hstu.py:413 new_outputs = o(...) + x) — the add autograd would launch on its own.
 * In the backward call `y` / `ldy` carry dy. */
typedef struct grb_ln_gate_args {
  const void* x; int64_t ldx;
  const void* gate; int64_t ldg;         /* NULL: plain LayerNorm */
  void* y; int64_t ldy;                  /* forward: output; backward: dy (read only) */
  float* mean; float* rstd;              /* (rows) saved by the forward, read by the backward */
  int64_t rows, W;
  float eps;
  int32_t dtype;
  float p_drop; int32_t _pad0;
  const int64_t* seed; int64_t salt;
  void* dx; int64_t lddx;                /* backward outputs */
  void* dgate; int64_t lddg;
  const void* res; int64_t ldres;        /* backward: added to dx; NULL = none */
} grb_ln_gate_args;
int grb_ln_gate_fwd_ex(const grb_ln_gate_args* a, grb_stream_t stream);
int grb_ln_gate_bwd_ex(const grb_ln_gate_args* a, grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * b4  negatives_samples/negative_sampler.py:31-37 (_maybe_l2_norm) and postprocessors.py:47-55:
 *     y = x / clamp(||x||_2, min=eps) per row, fp32.  inv (rows) is saved for backward:
 *     1 / max(||x||, eps), negated when the clamp was active.  Backward: dx from dy, y, inv.
 * ------------------------------------------------------------------------------------------- */
int grb_l2norm_fwd(const float* x, int64_t ldx, float* y, int64_t ldy, float* inv, int64_t rows,
                   int64_t W, float eps, grb_stream_t stream);
int grb_l2norm_bwd(const float* y, int64_t ldy, const float* dy, int64_t lddy, const float* inv,
                   float* dx, int64_t lddx, int64_t rows, int64_t W, grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * b2  models/indexing/top_k.py:44-70  MIPSBruteForceTopK  (mm + topk + id gather), fused:
 *     the (B, X) score matrix never reaches HBM.
 *
 *     grb_mips_topk chains, stream-ordered and without host sync (L = refinement levels):
 *       0. scores of every 4^L-th 128-item tile                -> ws (B, Xs) fp32 ;
 *          tau[b] = k-th largest sample score (exact)           (>= k items score >= tau[b]);
 *          sample scores >= tau[b] become the first candidates
 *       p = 1..L. scores of the tiles that are multiples of 4^(L-p) but not of 4^(L-p+1), keep
 *          (score, index) >= tau[b] in row b's candidate list, then tighten tau[b] to the k-th
 *          largest candidate so far
 *       last. exact top-k of each candidate list, sorted descending, ties -> lowest index,
 *          then ids[b, r] = item_ids[index]  (item_ids == NULL: ids = index)
 *     Every tile is scored exactly once.  status[0] (device int32) is set to the largest candidate
 *     count if any row overflowed `cand_cap`, else left 0; the host wrapper re-runs with a larger
 *     workspace in that case.  sample_stride: 0 = auto, else 4^L is the largest power of 4 <= it.
 *
 *     One query block (B <= 128, GRB_BF16, a corpus of >= ~500 tiles, cand_cap left automatic) takes
 *     the one-query-block plan instead (csrc/mips_small.cu), four launches, the corpus streamed 1 + 1/s
 *     times (s = 8 or 16):  group maxima of every s-th tile -> tau[b] = k'-th largest group maximum
 *     (each the score of a distinct item) -> all tiles, hits >= tau[b] appended to per-thread private
 *     sub-lists (no atomics) -> exact select.  Same outputs, same status word; a capacity passed
 *     explicitly (cand_cap != the automatic one: the wrapper's re-run after an overflow) selects the
 *     phased plan.  grb_mips_topk_workspace_bytes covers both layouts.
 *
 *     queries (B, D) row stride ldq; items (X, D) row stride ldi (the reference keeps the
 *     transposed *view* of this contiguous table, candidate_index.py:29).
 *     dtype GRB_BF16: tcgen05 path (D % 64 == 0, D <= 256);  GRB_F32: CUDA-core path (any D).
 * ------------------------------------------------------------------------------------------- */
typedef struct grb_mips_topk_args {
  int64_t B, X, D;
  int32_t k;
  int32_t dtype;
  const void* queries; int64_t ldq;
  const void* items; int64_t ldi;
  const int64_t* item_ids;      /* (X) or NULL */
  float* out_scores;            /* (B, k) fp32 */
  int64_t* out_ids;             /* (B, k) int64 */
  void* workspace; int64_t workspace_bytes;
  int64_t sample_stride;        /* every sample_stride-th item tile is sampled (>=1); 0 = auto */
  int64_t cand_cap;             /* candidate capacity per row; 0 = auto */
  int32_t* status;              /* device int32[2]: [0] overflow count, [1] reserved */
  /* f3 (candidate_index.py:125-158 + metrics/retrieval.py:40-68), fused into the final selection;
   * all optional (NULL / 0 = plain top-k).  invalid_ids (B, n_invalid) int64, row stride
   * ld_invalid: ids that must not appear in row b's result.  The thresholds are taken for
   * k' = min(k + n_invalid, X) exactly as the reference over-selects, the last kernel sorts the
   * best k' of a row in shared memory, drops the invalid ones and writes the first k (rows left
   * with fewer than k valid entries are padded with (-inf, -1); the reference raises there).
   * Requires k + n_invalid <= 2048 and n_invalid <= 1024.
   * target_ids (B) int64 + out_ranks (B) int32: 1 + position of target_ids[b] in row b's result,
   * k + 1 when absent: the `ranks` every metric of RetrievalMetrics.compute is a function of. */
  const int64_t* invalid_ids; int64_t ld_invalid; int32_t n_invalid;
  const int64_t* target_ids;
  int32_t* out_ranks;
} grb_mips_topk_args;

/* Bytes of workspace grb_mips_topk needs for these sizes (fills sample_stride/cand_cap if 0). */
int64_t grb_mips_topk_workspace_bytes(grb_mips_topk_args* a);
int grb_mips_topk(const grb_mips_topk_args* a, grb_stream_t stream);

/* Exact top-k of per-row candidate lists (also the N3 merge of per-shard top-k after the
 * NCCL all-gather).  cand_scores/cand_ids: (B, cap) ; counts: (B) int32 or NULL (= cap each).
 * Sorted descending by score, ties -> lowest id.  If id_map != NULL, out_ids = id_map[cand_id].
 * Rows with fewer than k candidates are padded with (-inf, -1). */
int grb_topk_select(const float* cand_scores, const int64_t* cand_ids, const int32_t* counts,
                    int64_t B, int64_t cap, int32_t k, const int64_t* id_map, float* out_scores,
                    int64_t* out_ids, grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * N3 exchange (new; the reference keeps the whole corpus on every rank): each rank stores its
 *     (rows, row_bytes) block of local top-k results into the same column block of EVERY peer's
 *     gather buffer with peer-to-peer stores over NVLink / NVSwitch, so that after one cross-rank
 *     barrier every rank holds the (rows, world * k) candidate matrix grb_topk_select merges.
 *     dst: host array of n_dst device pointers (the peers' buffers, from a symmetric-memory
 *     rendezvous; may include the local one).  row_bytes, strides and offsets multiples of 4.
 * ------------------------------------------------------------------------------------------- */
int grb_p2p_put_rows(const void* src, int64_t src_row_stride_bytes, void* const* dst, int32_t n_dst,
                     int64_t dst_row_stride_bytes, int64_t dst_col_offset_bytes, int64_t rows,
                     int64_t row_bytes, grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * b4/b5 + b1 + b6  negatives_samples/negative_sampler.py:105-131,192-212 ; similarity/
 *     dot_product.py:61-64 ; losses/autoregressive_losses.py:259-306 — fused sampled softmax:
 *       e_r   = concat(table0[idx0[n,r]], table1[idx1[n,r]])        (table1 may be NULL)
 *       neg_r = q[n].e_r / max(||e_r||, eps)   (l2_norm != 0)   |  q[n].e_r  (l2_norm == 0)
 *       pos   = q[n].p[n]                       (p already normalised by the caller)
 *       z     = [pos/T, where(neg_id[n,r]==pos_id[n], -5e4, neg_r/T)]   (T = temperature)
 *       loss_row[n] = -log_softmax(z)[0]
 *     The (N', R, D) negatives tensor is never materialised.  probs (N', R+1) fp32 = softmax(z)
 *     is saved for backward.  Backward takes g[n] = dL/dloss_row[n] and produces dq, dp and
 *     scatter-adds d table0 / d table1 (fp32 atomics; tables' grads must be pre-zeroed or
 *     accumulate).
 * ------------------------------------------------------------------------------------------- */
typedef struct grb_ssl_args {
  int64_t n_rows;       /* N' */
  int32_t R;            /* negatives per row */
  int32_t D;            /* = d0 + d1 */
  int32_t d0, d1;
  int32_t l2_norm;      /* normalise gathered negatives */
  int32_t dtype;        /* GRB_F32 (tables, q, p) */
  float l2_eps;
  float temperature;
  const void* q; int64_t ldq_;          /* (N', D) */
  const void* p; int64_t ldp;           /* (N', D) positives (normalised) */
  const void* table0; int64_t ldt0;     /* (X0, d0) */
  const void* table1; int64_t ldt1;     /* (X1, d1) or NULL */
  const int64_t* idx0;                  /* (N', R) rows of table0 */
  const int64_t* idx1;                  /* (N', R) rows of table1 or NULL */
  const int64_t* pos_ids;               /* (N') */
  const int64_t* neg_ids;               /* (N', R) ids compared with pos_ids */
  float* loss_rows;                     /* (N') out */
  float* probs;                         /* (N', R+1) out (fwd) / in (bwd) */
  /* backward */
  const float* g;                       /* (N') */
  float* dq; float* dp;                 /* (N', D) fp32, written */
  float* dtable0; float* dtable1;       /* fp32 (X0,d0)/(X1,d1) contiguous, accumulated */
} grb_ssl_args;

int grb_sampled_softmax_fwd(const grb_ssl_args* a, grb_stream_t stream);
int grb_sampled_softmax_bwd(const grb_ssl_args* a, grb_stream_t stream);

/* The same backward for ONE table of already-normalised rows (l2_norm == 0, d1 == 0: the in-batch cache,
 * negative_sampler.py:192-212), D % 8 == 0, D <= 256, without atomics on the table gradient: the row-wise
 * half (dq, dp, the coefficients) gathers bf16 copies of the table rows; the (row, negative) pairs are
 * counting-sorted by table row; one warp per table row then sums coef * q[n] over its list from a bf16
 * copy of q and stores the row (rows nobody sampled get zeros: dtable0 needs no zero fill and is
 * OVERWRITTEN, not accumulated).  table_rows = rows of table0 / dtable0.  Gradients see q and the table
 * rounded to bf16 inside the sums.  Five stream-ordered launches; workspace from the function below. */
int64_t grb_sampled_softmax_bwd_csr_workspace_bytes(int64_t n_rows, int32_t R, int32_t D, int64_t table_rows);
int grb_sampled_softmax_bwd_csr(const grb_ssl_args* a, int64_t table_rows, void* workspace,
                                int64_t workspace_bytes, grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Embedding-table gradient: table_grad[ids[i], :] += grad[i, :] for every i with ids[i] != skip_id.
 *     Replaces aten::embedding_dense_backward under the nn.Embedding tables of
 *     models/embeddings/embeddings.py:40-101 (padding_idx = 0 -> skip_id = 0; pass -1 to skip
 *     nothing).  grad (n, D) fp32 with row stride ld_grad; table_grad (V, D) fp32 contiguous,
 *     zero-filled or accumulating.  fp32 atomics: the summation order is not fixed.
 * ------------------------------------------------------------------------------------------- */
int grb_rows_scatter_add(const float* grad, int64_t ld_grad, const int64_t* ids, float* table_grad,
                         int64_t n, int32_t D, int64_t num_rows, int64_t skip_id,
                         grb_stream_t stream);

/* table[ids[i], :] *= scale for every i with a real id (!= skip_id, in range); ids must be distinct
 * (the list of rows a batch touched).  Pre-scales the local rows of a data-parallel table gradient
 * in place, so that the dense (V, D) tensor autograd produced becomes the reduced gradient once the
 * peers' rows are added (grb_p2p_put_table_rows + grb_rows_scatter_add): no second dense buffer. */
int grb_rows_scale(float* table, const int64_t* ids, int64_t n, int32_t D, int64_t num_rows,
                   int64_t skip_id, float scale, grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * AdamW step over a list of fp32 tensors, one launch per 64 tensors (the optimizer of the train
 *     step: torch.optim.AdamW as the reference configures it, configs/model/hstu.yaml +
 *     models/generative_recommenders.py:254-322 configure_optimizers).  Per element, in fp32:
 *         p *= 1 - lr * weight_decay
 *         m  = m + (g - m) * (1 - beta1)           v = beta2 * v + (1 - beta2) * g * g
 *         p -= (lr / bias_correction1) * m / (sqrt(v) / sqrt(bias_correction2) + eps)
 *     with bias_correction{1,2} = 1 - beta{1,2}^step computed by the caller.  p, g, m, v, numel are
 *     HOST arrays of n device pointers / element counts.  28 bytes of HBM traffic per element.
 * ------------------------------------------------------------------------------------------- */
int grb_adamw_step(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                   const int64_t* numel, double lr, double beta1, double beta2, double eps,
                   double weight_decay, double bias_correction1, double bias_correction2,
                   grb_stream_t stream);

/* autoregressive_losses.py:306 (and :172), the loss of the step from its per-position terms:
 *     out2[0] = sum(x * w) / sum(w),  out2[1] = sum(w)        (x, w fp32 (n); one CTA, fixed order)
 * and its backward gx[i] = gout[0] * w[i] / out2[1]  (w = supervision_weights carries no gradient). */
int grb_weighted_mean_fwd(const float* x, const float* w, int64_t n, float* out2, grb_stream_t stream);
int grb_weighted_mean_bwd(const float* w, const float* out2, const float* gout, int64_t n, float* gx,
                          grb_stream_t stream);

/* dst[i] (bf16, numel[i]) = src[i] (fp32) for a HOST list of n device tensors, one launch per 64 tensors:
 * the compute-dtype copies of the fp32 master weights of all STU layers of a forward pass (the per-layer
 * casts of hstu.py:300-305 / :404-413 under autocast, made once). */
int grb_cast_f32_bf16_many(int n, const float* const* src, void* const* dst, const int64_t* numel,
                           grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Cross-rank barrier on the stream, over peer memory: every rank adds 1 to word `slot` of each
 *     rank's signal array (system-scope release after its earlier stores / reds), then waits until
 *     its own word reaches n_ranks * epoch.  signals: host array of n_ranks device pointers to
 *     symmetric uint64 arrays (zero-initialised once); epoch counts this slot's barriers from 1.
 * ------------------------------------------------------------------------------------------- */
int grb_p2p_barrier(void* const* signals, int32_t n_ranks, int32_t rank, int32_t slot, int64_t epoch,
                    grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel all-reduce of the dense gradients over peer memory (new; the reference relies on
 *     Lightning DDP's NCCL all-reduce, configs/trainer/ddp.yaml:4).  Two-shot, in place: rank r reads
 *     float range [r, r + 1) * ceil(numel / 4 / n_ranks) * 4 of EVERY rank's buffer, sums in rank order,
 *     multiplies by scale and stores the result into the same range of EVERY rank's buffer.  bufs:
 *     host array of n_ranks device pointers (symmetric fp32 buffers of numel elements, numel % 4 == 0,
 *     16-byte aligned).  The caller orders it with grb_p2p_barrier before (every rank's buffer is
 *     filled) and after (every range has landed everywhere).
 * ------------------------------------------------------------------------------------------- */
int grb_p2p_allreduce(void* const* bufs, int32_t n_ranks, int32_t rank, int64_t numel, float scale,
                      grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel exchange of an embedding-table gradient over peer memory (new; the reference
 *     all-reduces the dense table, configs/trainer/ddp.yaml).  Sender side: for i < n, slot
 *     s = slot_offset + i of every destination r gets  dst_ids[r][s] = ids[i]  and, when ids[i] is a
 *     real row (!= skip_id, in range),  dst_rows[r][s, :] = scale * grad_table[ids[i], :]  — plain
 *     16-byte stores over NVLink / NVSwitch.  After a barrier the receiver scatter-adds its staging
 *     buffer into its dense gradient with grb_rows_scatter_add (local atomics only).
 *     grad_table (V, D) fp32 contiguous; dst_rows[r] (slots, D) fp32; dst_ids[r] (slots) int64.
 * ------------------------------------------------------------------------------------------- */
int grb_p2p_put_table_rows(const float* grad_table, const int64_t* ids, int64_t n, int32_t D,
                           int64_t num_rows, int64_t skip_id, float scale, void* const* dst_rows,
                           void* const* dst_ids, int32_t n_dst, int64_t slot_offset,
                           grb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Building-block self test (tcgen05 descriptors, TMA swizzle, TMEM layouts).  Runs tiny GEMMs
 * in every operand mode the attention / retrieval kernels use and writes max-abs errors to
 * host array errs[n_modes].  Returns the number of modes run, or a negative error code.
 * This one synchronises the stream (test helper).
 * ------------------------------------------------------------------------------------------- */
int grb_selftest_umma(float* errs, int max_modes, grb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GRB200_H_ */
