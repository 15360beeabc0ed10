// hstu_attn_api.cu — C-ABI entry points of the fused HSTU attention; picks the tcgen05 path
// (bf16, head dims 64) or the CUDA-core path (fp32 / other shapes).  No CPU fallback.
#include "common.cuh"
#include <cstdlib>

namespace grb {
int check_attn_args(const grb_hstu_attn_args* a, bool bwd);
int hstu_attn_fwd_simt_dispatch(const grb_hstu_attn_args* a, cudaStream_t st);
int hstu_attn_bwd_simt_dispatch(const grb_hstu_attn_args* a, cudaStream_t st);
bool hstu_attn_fwd_sm100_supported(const grb_hstu_attn_args* a);
bool hstu_attn_bwd_sm100_supported(const grb_hstu_attn_args* a);
int hstu_attn_fwd_sm100(const grb_hstu_attn_args* a, cudaStream_t st);
bool hstu_attn_fwd2_sm100_usable(const grb_hstu_attn_args* a);
int hstu_attn_fwd2_sm100(const grb_hstu_attn_args* a, cudaStream_t st);
int hstu_attn_bwd_sm100(const grb_hstu_attn_args* a, cudaStream_t st);
bool hstu_attn_short_usable(const grb_hstu_attn_args* a, bool bwd);
int hstu_attn_short_fwd(const grb_hstu_attn_args* a, cudaStream_t st);
int hstu_attn_short_bwd(const grb_hstu_attn_args* a, cudaStream_t st);

// developer switches, read once per process (not on every launch)
static bool env_flag(const char* name) {
  const char* e = std::getenv(name);
  return e && e[0] == '1';
}
static bool force_cuda_core() {
  static const bool v = env_flag("GRB_FORCE_CUDA_CORE");
  return v;
}
static bool force_fwd_v1() {
  static const bool v = env_flag("GRB_FWD_V1");
  return v;
}
}  // namespace grb

using namespace grb;

extern "C" {

int grb_hstu_attn_fwd(const grb_hstu_attn_args* a, grb_stream_t stream) {
  int rc = check_attn_args(a, false);
  if (rc != GRB_OK) return rc;
  GRB_REQUIRE(a->B <= 65535 && a->H <= 65535, GRB_ERR_UNSUPPORTED, "hstu_attn: B,H <= 65535");
  GRB_REQUIRE(a->short_schedule || !a->bucket_cache_masked, GRB_ERR_INVALID_ARG,
              "hstu_attn: a masked bucket cache only serves the short-sequence kernels (short_schedule is null)");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  if (a->short_schedule) {   // the caller asked for the short-sequence kernels: take them or say why not
    GRB_REQUIRE(hstu_attn_short_usable(a, false), GRB_ERR_UNSUPPORTED,
                "hstu_attn_fwd: short_schedule given but the short-sequence path does not apply (bf16, head "
                "dims 64, max_len <= 256, 16-byte aligned rows, masked bucket_cache built for max_len)");
    return hstu_attn_short_fwd(a, st);
  }
  GRB_REQUIRE(!a->zero_tail_rows, GRB_ERR_UNSUPPORTED,
              "hstu_attn_fwd: zero_tail_rows is a service of the short-sequence kernels (zero-fill the rows)");
  if (!force_cuda_core() && hstu_attn_fwd_sm100_supported(a)) {
    // second-generation kernel when the buckets come from the per-batch cache (or no bias);
    // GRB_FWD_V1=1 keeps the first one (developer switch)
    if (!force_fwd_v1() && hstu_attn_fwd2_sm100_usable(a)) return hstu_attn_fwd2_sm100(a, st);
    return hstu_attn_fwd_sm100(a, st);
  }
  return hstu_attn_fwd_simt_dispatch(a, st);
}

int grb_hstu_attn_bwd(const grb_hstu_attn_args* a, grb_stream_t stream) {
  int rc = check_attn_args(a, true);
  if (rc != GRB_OK) return rc;
  GRB_REQUIRE(a->B <= 65535 && a->H <= 65535, GRB_ERR_UNSUPPORTED, "hstu_attn: B,H <= 65535");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  if (a->short_schedule) {
    GRB_REQUIRE(hstu_attn_short_usable(a, true), GRB_ERR_UNSUPPORTED,
                "hstu_attn_bwd: short_schedule given but the short-sequence path does not apply (bf16, head "
                "dims 64, max_len <= 256, 16-byte aligned rows, masked bucket_cache built for max_len, "
                "dq_accum scratch when max_len > 128)");
    return hstu_attn_short_bwd(a, st);
  }
  GRB_REQUIRE(!a->zero_tail_rows, GRB_ERR_UNSUPPORTED,
              "hstu_attn_bwd: zero_tail_rows is a service of the short-sequence kernels (zero-fill the rows)");
  if (!force_cuda_core() && hstu_attn_bwd_sm100_supported(a)) return hstu_attn_bwd_sm100(a, st);
  return hstu_attn_bwd_simt_dispatch(a, st);
}

}
