// tmem_ld_probe.cu — measures tcgen05.ld (32x32b) read throughput per SM on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_probe tmem_ld_probe.cu
// Measured on B200: one warp 46 B/clk (a 32x32b.x16 load + wait = 44 cycles), 4 warps (one per lane
// quadrant) 184 B/clk/SM, 16 warps 450-470 B/clk/SM: TMEM reads are not what bounds the attention
// epilogues (160 KiB per backward tile = ~350 cycles).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld(uint32_t addr, uint32_t* v);
template <>
__device__ __forceinline__ void ld<16>(uint32_t addr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr));
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t addr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(addr));
}

template <int X>
__global__ void probe(int reps, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  const uint32_t base = tmem + ((uint32_t) ((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t v[X];
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < reps; ++i) {
#pragma unroll
    for (int c = 0; c < 512 / X / 4; ++c) {    // each warp sweeps a quarter of the columns per rep
      ld<X>(base + (uint32_t) (((warp >> 2) * 128 + c * X) & 511), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int k = 0; k < X; ++k) acc ^= v[k];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int X>
void run(int warps) {
  long long* d_c; uint32_t* d_s;
  cudaMalloc(&d_c, 8 * 148); cudaMalloc(&d_s, 4 * 1024);
  const int reps = 2000;
  probe<X><<<148, warps * 32>>>(reps, d_c, d_s);
  probe<X><<<148, warps * 32>>>(reps, d_c, d_s);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
  const double bytes = (double) reps * (512 / X / 4) * warps * X * 32 * 4;
  printf("x%d warps=%2d: %lld cycles, %.1f B/clk/SM  (%s)\n", X, warps, c, bytes / c, cudaGetErrorString(e));
  cudaFree(d_c); cudaFree(d_s);
}

int main() {
  for (int w : {1, 4, 8, 16}) run<16>(w);
  for (int w : {1, 4, 8, 16}) run<32>(w);
  return 0;
}
