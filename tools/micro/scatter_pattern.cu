// Micro-benchmark: how fast can B200 move 16-B records with the partition pass's memory pattern
// and NOTHING else?  Every CTA reads one contiguous tile (3072 records) and writes it as `bins`
// contiguous segments, one per output stream (stream d starts at d * N / bins; tile t appends
// at t * seg inside it) — the access pattern of a radix partition pass with a uniform digit
// distribution, without keys, ranking or look-back.  Compare with a plain copy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scatter_pattern scatter_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kTile = 3072, kThreads = 384;
__global__ void __launch_bounds__(kThreads, 3) scatter(const float4 *in, float4 *out, size_t n, int bins, int use_smem) {
  extern __shared__ float4 sm[];
  const size_t t = blockIdx.x, base = t * kTile;
  const int seg = kTile / bins;
  const size_t stream_len = n / bins;
  if (use_smem) {
    for (int j = threadIdx.x; j < kTile; j += kThreads) sm[j] = in[base + j];
    __syncthreads();
  }
  for (int j = threadIdx.x; j < kTile; j += kThreads) {
    const int d = j / seg, o = j - d * seg;
    const float4 v = use_smem ? sm[j] : in[base + j];
    out[(size_t)d * stream_len + t * seg + o] = v;
  }
}
__global__ void copyk(const float4 *in, float4 *out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
int main() {
  const size_t tiles = 3256, n = tiles * kTile;  // ~10 M records, 160 MB
  float4 *a, *b;
  cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMemset(a, 1, n * 16);
  cudaFuncSetAttribute(scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, kTile * 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](auto fn, const char *name) {
    fn(); fn();
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) fn();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("%-44s %7.1f us  %6.2f TB/s (read + write)\n", name, ms * 1e3, 2.0 * n * 16 / ms / 1e9);
  };
  time([&] { copyk<<<148 * 8, 512>>>(a, b, n); }, "plain copy");
  for (int bins : {1, 64, 128, 256}) {
    for (int sm : {0, 1}) {
      char name[96]; snprintf(name, sizeof name, "tile scatter, %3d streams, %s", bins, sm ? "staged in smem" : "direct");
      time([&] { scatter<<<(unsigned)tiles, kThreads, kTile * 16>>>(a, b, n, bins, sm); }, name);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
