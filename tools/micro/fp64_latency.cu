// Dependent-chain latencies that bound the panel step: DMUL->DADD, __ddiv_rn, __threadfence, an L2
// round trip (ld.global.cg), a release/acquire flag hop between two CTAs.
// Build: nvcc -arch=sm_100a -O3 -fmad=false -o fp64_latency.bin fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_chain(double* out, const double* in, int n, long long* cyc) {
  double x = in[0], a = in[1], r = in[2];
  long long t0 = clock64();
  for (int i = 0; i < n; i++) x = __dsub_rn(x, __dmul_rn(a, r + x));   // DADD -> DMUL -> DADD fully dependent
  long long t1 = clock64();
  out[0] = x; cyc[0] = t1 - t0;
}
__global__ void k_div(double* out, const double* in, int n, long long* cyc) {
  double x = in[0], p = in[1];
  long long t0 = clock64();
  for (int i = 0; i < n; i++) x = __ddiv_rn(x, p);
  long long t1 = clock64();
  out[0] = x; cyc[0] = t1 - t0;
}
__global__ void k_fence(double* out, int n, long long* cyc) {
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { out[threadIdx.x + 32 * (i & 7)] = i; __threadfence(); }
  long long t1 = clock64();
  cyc[0] = t1 - t0;
}
__global__ void k_l2(const int* chase, int n, long long* cyc, int* out) {
  int j = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { int v; asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(chase + j)); j = v; }
  long long t1 = clock64();
  cyc[0] = t1 - t0; out[0] = j;
}
// ping-pong between CTA 0 and CTA 1 through two flags (release/acquire at gpu scope)
__global__ void k_pingpong(unsigned int* flags, int n, long long* cyc) {
  if (threadIdx.x != 0) return;
  long long t0 = clock64();
  for (unsigned int i = 1; i <= (unsigned int)n; i++) {
    if (blockIdx.x == 0) {
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags), "r"(i) : "memory");
      unsigned int v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + 32) : "memory"); } while (v != i);
    } else {
      unsigned int v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags) : "memory"); } while (v != i);
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + 32), "r"(i) : "memory");
    }
  }
  long long t1 = clock64();
  if (blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double *in, *out; long long* cyc; int* chase; unsigned int* flags; int* iout;
  cudaMalloc(&in, 64); cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 64); cudaMalloc(&chase, 1 << 20); cudaMalloc(&flags, 512);
  cudaMalloc(&iout, 64);
  double h[3] = {1.0, 1.0000001, 0.5};
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  int hc[1 << 18];
  for (int i = 0; i < (1 << 18); i++) hc[i] = (i + 4099 * 8) & ((1 << 18) - 1);
  cudaMemcpy(chase, hc, sizeof(hc), cudaMemcpyHostToDevice);
  cudaMemset(flags, 0, 512);
  int n = 4096;
  for (int rep = 0; rep < 2; rep++) {
    k_chain<<<1, 1>>>(out, in, n, cyc); cudaDeviceSynchronize();
    if (rep) printf("DMUL+DADD dependent pair: %.1f cycles (%.1f per op)\n", (double)cyc[0] / n, (double)cyc[0] / n / 2);
    k_div<<<1, 1>>>(out, in, n, cyc); cudaDeviceSynchronize();
    if (rep) printf("__ddiv_rn dependent: %.1f cycles\n", (double)cyc[0] / n);
    k_fence<<<1, 32>>>(out, n, cyc); cudaDeviceSynchronize();
    if (rep) printf("store + __threadfence: %.1f cycles\n", (double)cyc[0] / n);
    k_l2<<<1, 1>>>(chase, n, cyc, iout); cudaDeviceSynchronize();
    if (rep) printf("ld.global.cg dependent (L2 hit): %.1f cycles\n", (double)cyc[0] / n);
    k_pingpong<<<2, 32>>>(flags, 2000, cyc); cudaDeviceSynchronize();
    if (rep) printf("release/acquire flag round trip between two CTAs: %.1f cycles\n", (double)cyc[0] / 2000);
    cudaMemset(flags, 0, 512);
  }
  return 0;
}
