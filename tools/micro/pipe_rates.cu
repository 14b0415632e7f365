// Micro-benchmark: sustained per-SM throughput of DFMA, F2F.F64.F32, FRND+F2I, FFMA and an
// integer-ALU float->double widening, to decide which pipe bounds the reduce / finalize kernels.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double *out, float seed, int iters) {
  float f0 = seed + threadIdx.x, f1 = f0 * 1.5f, f2 = f0 * 2.5f, f3 = f0 * 3.5f;
  double a0 = f0, a1 = f1, a2 = f2, a3 = f3, m = 1.0000001, c = 1e-9;
  int acc = 0;
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) { a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c); }
    if (OP == 1) { a0 += (double)f0; a1 += (double)f1; a2 += (double)f2; a3 += (double)f3; f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f; }
    if (OP == 2) { acc += (int)ceilf(f0) + (int)ceilf(f1) + (int)ceilf(f2) + (int)ceilf(f3); f0 += 0.37f; f1 += 0.37f; f2 += 0.37f; f3 += 0.37f; }
    if (OP == 3) { f0 = fmaf(f0, 1.0000001f, 1e-9f); f1 = fmaf(f1, 1.0000001f, 1e-9f); f2 = fmaf(f2, 1.0000001f, 1e-9f); f3 = fmaf(f3, 1.0000001f, 1e-9f); }
    if (OP == 4) {
      float fs[4] = {f0, f1, f2, f3}; double *as[4] = {&a0, &a1, &a2, &a3};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        unsigned u = __float_as_uint(fs[q]);
        unsigned hi = (u & 0x80000000u) | (((u & 0x7FFFFFFFu) >> 3) + (896u << 20));
        *as[q] += __hiloint2double(((u >> 23) & 0xFF) ? hi : (u & 0x80000000u), ((u >> 23) & 0xFF) ? (u << 29) : 0u);
      }
      f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + f0 + f1 + f2 + f3 + acc;
}
template <int OP> void run(const char *name, double ops_per_iter) {
  int dev_sms; cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double *out; cudaMalloc(&out, sizeof(double) * dev_sms * 8 * 256);
  const int iters = 20000;
  k<OP><<<dev_sms * 8, 256>>>(out, 1.f, 100);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<OP><<<dev_sms * 8, 256>>>(out, 1.f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double thread_ops = (double)dev_sms * 8 * 256 * iters * ops_per_iter;
  printf("%-28s %8.3f ms  %8.1f Gop/s  %6.1f lane-ops/clk/SM (at %d MHz)\n", name, ms, thread_ops / ms / 1e6,
         thread_ops / (ms * 1e-3) / dev_sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}
int main() {
  run<0>("DFMA", 4); run<1>("F2F.F64.F32 + DADD", 4); run<2>("FRND.CEIL + F2I (+IADD)", 4); run<3>("FFMA", 4);
  run<4>("ALU widen f32->f64 + DADD", 4);
  return 0;
}
