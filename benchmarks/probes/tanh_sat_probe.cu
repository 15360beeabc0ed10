// Does MUFU.TANH saturate to exactly -1 for large negative inputs?  (The short-sequence attention
// kernels fold the causal / length mask into the bias tile as a large negative value and rely on
// h + h * tanh(h) == 0 there.)  Prints the residual for f32 and f16x2.
#include <cstdio>
#include <cuda_fp16.h>
__global__ void probe(const float* in, float* out32, float* out16, float* res32, float* res16, int n) {
  int i = threadIdx.x;
  if (i >= n) return;
  float x = in[i], y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  out32[i] = y;
  res32[i] = fmaf(x, y, x);
  __half2 h = __floats2half2_rn(x, x * 0.5f);
  unsigned hu = *reinterpret_cast<unsigned*>(&h), tu, pu;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(tu) : "r"(hu));
  asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(pu) : "r"(hu), "r"(tu));
  out16[i] = __half2float(__low2half(*reinterpret_cast<__half2*>(&tu)));
  res16[i] = __half2float(__low2half(*reinterpret_cast<__half2*>(&pu)));
}
int main() {
  const int n = 12;
  float h[n] = {-5.f, -8.f, -9.f, -10.f, -12.f, -16.f, -20.f, -50.f, -100.f, -1000.f, -15000.f, -30000.f};
  float *d, *o32, *o16, *r32, *r16;
  cudaMalloc(&d, 4 * n); cudaMalloc(&o32, 4 * n); cudaMalloc(&o16, 4 * n); cudaMalloc(&r32, 4 * n); cudaMalloc(&r16, 4 * n);
  cudaMemcpy(d, h, 4 * n, cudaMemcpyHostToDevice);
  probe<<<1, 32>>>(d, o32, o16, r32, r16, n);
  float a[n], b[n], c[n], e[n];
  cudaMemcpy(a, o32, 4 * n, cudaMemcpyDeviceToHost); cudaMemcpy(b, o16, 4 * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(c, r32, 4 * n, cudaMemcpyDeviceToHost); cudaMemcpy(e, r16, 4 * n, cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; ++i)
    printf("x=%9.1f tanh32=%.9g (1+t=%.3g) x+x*t=%.6g | tanh16=%.9g x+x*t(f16)=%.6g\n", h[i], a[i], 1.0 + (double) a[i], c[i], b[i], e[i]);
  return cudaGetLastError() != cudaSuccess;
}
