// Shared-memory atomic throughput on random addresses (what K1's slot reservation, the window scatter and the window
// insert are made of): returning add, non-returning add, CAS, plain load+store, for several table sizes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template<int MODE>
__global__ void __launch_bounds__(1024, 1) k(uint32_t nb_mask, uint32_t iters, uint32_t* out) {
  extern __shared__ uint32_t sm[];
  for(uint32_t i = threadIdx.x; i <= nb_mask; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  uint32_t x = mix(blockIdx.x * 1024u + threadIdx.x + 1u), acc = 0;
  for(uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for(int u = 0; u < 4; ++u) {
      x = x * 1664525u + 1013904223u;
      const uint32_t a = (x >> 8) & nb_mask;
      if(MODE == 0) acc += atomicAdd(&sm[a], 1u);                 // returning
      else if(MODE == 1) atomicAdd(&sm[a], 1u);                   // result unused
      else if(MODE == 2) acc += atomicCAS(&sm[a], 0u, x | 1u);     // CAS
      else if(MODE == 3) { uint32_t v = sm[a]; sm[a] = v + 1; acc += v; }   // plain load + store (racy on purpose)
      else { acc += sm[a]; }                                      // plain load
    }
  }
  __syncthreads();
  if(acc == 0x12345678u || sm[threadIdx.x & nb_mask] == 0xFFFFFFFFu) out[0] = acc;
}
template<int MODE> void run(const char* name, uint32_t nb, uint32_t* d) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const uint32_t iters = 4096;
  k<MODE><<<148, 1024, nb * 4, 0>>>(nb - 1, 64, d);
  cudaEventRecord(a);
  k<MODE><<<148, 1024, nb * 4, 0>>>(nb - 1, iters, d);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double ops = 148.0 * 1024 * iters * 4;
  printf("%-22s bins %6u : %7.2f G ops/s  = %5.2f ops/clk/SM (at 1.95 GHz)  [%s]\n", name, nb, ops / ms / 1e6, ops / ms / 1e6 / 148 / 1.95, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  uint32_t* d; cudaMalloc(&d, 64);
  for(uint32_t nb : {512u, 2048u, 16384u}) {
    run<0>("atomicAdd returning", nb, d); run<1>("atomicAdd no result", nb, d); run<2>("atomicCAS", nb, d);
    run<3>("load+store", nb, d); run<4>("load", nb, d);
  }
  return 0;
}
