// TEMPORARY stubs until the tcgen05 kernels land.
#include "common.cuh"
#include "mips_epilogue.cuh"
namespace grb {
bool hstu_attn_bwd_sm100_supported(const grb_hstu_attn_args*) { return false; }
int hstu_attn_bwd_sm100(const grb_hstu_attn_args*, cudaStream_t) { return GRB_ERR_UNSUPPORTED; }
}
