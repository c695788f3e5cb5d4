// FP64 issue-rate microbenchmark: separately rounded DMUL + DADD (the replay's instruction mix)
// vs DFMA, at several occupancies.  Build: nvcc -arch=sm_100a -O3 -fmad=false -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int kChains, bool kFma>
__global__ void k(double* out, double a, double r, int iters) {
  double x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; c++) x[c] = threadIdx.x + c;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < kChains; c++) {
      // the multiplier is another chain's running value, so neither form can be hoisted out of the loop
      if (kFma) x[c] = __fma_rn(-x[(c + 1) % kChains], r, x[c]);
      else x[c] = __dsub_rn(x[c], __dmul_rn(x[(c + 1) % kChains], r));
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < kChains; c++) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int kChains, bool kFma>
void run(int threads, int ctas_per_sm, int sms, double* out) {
  int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<kChains, kFma><<<sms * ctas_per_sm, threads>>>(out, 1.000001, 0.5, 16);
  cudaEventRecord(e0);
  k<kChains, kFma><<<sms * ctas_per_sm, threads>>>(out, 1.000001, 0.5, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double inst = (double)sms * ctas_per_sm * threads * iters * kChains * (kFma ? 1 : 2);
  printf("chains=%d fma=%d threads=%d ctas/sm=%d  %.3f ms  %.2f T thread-inst/s  (%.1f lanes/clk/SM at 1.9 GHz)\n", kChains,
         (int)kFma, threads, ctas_per_sm, ms, inst / ms / 1e9, inst / (ms * 1e-3) / sms / 1.9e9);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 16 * 1024);
  run<32, false>(128, 1, sms, out);
  run<32, false>(128, 3, sms, out);
  run<32, false>(384, 1, sms, out);
  run<32, false>(256, 4, sms, out);
  run<32, true>(128, 3, sms, out);
  run<32, true>(256, 4, sms, out);
  run<8, false>(256, 4, sms, out);
  run<8, true>(256, 4, sms, out);
  return 0;
}
