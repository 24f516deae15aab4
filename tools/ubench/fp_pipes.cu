// fp_pipes.cu -- micro-benchmark of the CUDA-core pipes the FX kernels lean on (B200, sm_100a):
// scalar FFMA (3-register form), packed fma.rn.f32x2, DFMA, FMNMX mix, SHFL, MUFU lg2/ex2, log10f, and dependent-chain
// latencies.  One CTA of 1024 threads per SM (8 warps per scheduler), clock64 around an unrolled loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/fp_pipes tools/ubench/fp_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, float seed) {
  float a[8], b = seed, c = seed * 0.5f;
  double d[8];
  unsigned long long p[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; d[i] = a[i]; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
  unsigned long long bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
  unsigned long long cc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fmaf(a[i], b, c);
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc));
      if (MODE == 2) d[i] = fma(d[i], (double)b, (double)c);
      if (MODE == 3) { a[i] = fmaf(a[i], b, c); a[i] = fmaxf(a[i], c); }
      if (MODE == 4) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);
      if (MODE == 5) a[i] = __log2f(a[i]) + 2.f;
      if (MODE == 6) a[i] = log10f(a[i]) + 11.f;
      if (MODE == 7) a[i] = exp2f(a[i]) * 0.25f;
    }
    if (MODE == 8) a[0] = fmaf(a[0], b, c);                       // dependent FFMA chain (latency)
    if (MODE == 9) d[0] = fma(d[0], (double)b, (double)c);         // dependent DFMA chain
    if (MODE == 10) a[0] = fmaxf(fmaf(a[0], b, c), c);            // FFMA -> FMNMX chain
    if (MODE == 11) d[0] = __shfl_up_sync(0xffffffffu, d[0], 1) + 1.0;   // 64-bit shuffle + DADD chain
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i] + (float)d[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, int ops_per_iter, float* out, long long* cyc) {
  int sms = 148;
  k<MODE><<<sms, threads>>>(out, cyc, 1.0001f);
  cudaDeviceSynchronize();
  k<MODE><<<sms, threads>>>(out, cyc, 1.0001f);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += h[i];
  avg /= sms;
  double warp_instr = (double)(threads / 32) * ITERS * ops_per_iter;
  printf("%-28s threads/SM %4d  cycles %10.0f  warp-instr/clk/SM %.3f  cycles/iter %.2f  (%s)\n", name, threads, avg,
         warp_instr / avg, avg / ITERS, cudaGetErrorString(e));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>("ffma 3-reg", 1024, 8, out, cyc);
  run<0>("ffma 3-reg", 512, 8, out, cyc);
  run<0>("ffma 3-reg", 128, 8, out, cyc);
  run<1>("fma.f32x2", 1024, 8, out, cyc);
  run<1>("fma.f32x2", 128, 8, out, cyc);
  run<2>("dfma", 1024, 8, out, cyc);
  run<2>("dfma", 128, 8, out, cyc);
  run<3>("ffma+fmnmx", 1024, 16, out, cyc);
  run<4>("shfl.up 32b", 1024, 8, out, cyc);
  run<5>("lg2.approx+fadd", 1024, 16, out, cyc);
  run<6>("log10f+fadd", 1024, 8, out, cyc);
  run<7>("exp2f+fmul", 1024, 8, out, cyc);
  run<8>("lat: ffma chain", 32, 1, out, cyc);
  run<9>("lat: dfma chain", 32, 1, out, cyc);
  run<10>("lat: ffma->fmnmx chain", 32, 1, out, cyc);
  run<11>("lat: shfl64+dadd chain", 32, 1, out, cyc);
  return 0;
}
