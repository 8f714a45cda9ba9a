// Issue-rate microbenchmark: scalar FFMA (register and immediate forms) vs packed FFMA2 per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);                 // 3-register FFMA
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], x[i], 0.999f);         // immediate addend
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {                                     // packed, register operands
        unsigned long long v, aa, bb;
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
        asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
        asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
        asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
      }
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = x[i] * a;                         // FMUL
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));   // MUFU
    }
  }
  unsigned long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(t1 - t0) * 0.f;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
template <int MODE>
void run(const char* name, int warps, int per_iter) {
  float* d;
  cudaMalloc(&d, 148 * 1024 * 4);
  const int iters = 4096;
  k<MODE><<<148, warps * 32>>>(d, iters, 1.0001f, 1e-7f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, warps * 32>>>(d, iters, 1.0001f, 1e-7f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float cyc;
  cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
  // warp-instructions per SM sub-partition per cycle
  double wi = (double)iters * per_iter * warps / 4.0;
  printf("%-28s warps/SM %2d: %.3f warp-instr/clk/SMSP (%.1f elem-ops/clk/SM)\n", name, warps, wi / cyc, wi / cyc * 4 * 32 * (MODE == 2 ? 2 : 1));
  cudaFree(d);
}
int main() {
  for (int w : {4, 8, 16}) {
    run<0>("FFMA reg,reg,reg", w, 16);
    run<1>("FFMA reg,reg,imm", w, 16);
    run<2>("FFMA2 (8 per 16 elems)", w, 8);
    run<3>("FMUL reg,reg", w, 16);
    run<4>("MUFU.EX2", w, 16);
  }
  return 0;
}
