// Microbenchmark: the L2 operation mix of K2 (first-probe CAS, then for ~36% of the keys a
// look-ahead of 4 loads + one more CAS) over 32 MB regions, region after region.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }
template<int MODE, int ILP>
__global__ void k(uint32_t* tab, uint64_t mask, uint64_t base, uint64_t per_thread, uint32_t* sink, uint64_t seed) {
  uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for(uint64_t i = 0; i < per_thread; i += ILP) {
    uint32_t old[ILP]; uint64_t idx[ILP];
#pragma unroll
    for(int j = 0; j < ILP; ++j) { idx[j] = base + (mix(seed + gid * per_thread + i + j) & mask); old[j] = atomicCAS(&tab[idx[j]], 0u, (uint32_t)(idx[j] | 1)); }
#pragma unroll
    for(int j = 0; j < ILP; ++j) {
      if(old[j] != 0) {       // slot taken (another key): walk
        if(MODE == 1) {
          uint32_t s0 = __ldcg(&tab[idx[j] + 1]), s1 = __ldcg(&tab[idx[j] + 3]), s2 = __ldcg(&tab[idx[j] + 6]), s3 = __ldcg(&tab[idx[j] + 10]);
          uint64_t t = s0 == 0 ? idx[j] + 1 : s1 == 0 ? idx[j] + 3 : s2 == 0 ? idx[j] + 6 : idx[j] + 10;
          acc += atomicCAS(&tab[t], 0u, (uint32_t)(t | 1)) + s3;
        } else {
          uint64_t t = idx[j]; uint32_t o = old[j];
          for(int p = 1; p < 12 && o != 0; ++p) { t = idx[j] + p * (p + 1) / 2; o = atomicCAS(&tab[t], 0u, (uint32_t)(t | 1)); }
          acc += o;
        }
      }
    }
  }
  if(acc == 0x12345678) *sink = acc;
}
template<int MODE, int ILP>
float run(uint32_t* tab, uint64_t slots, uint64_t total_slots, int tps, double load, uint32_t* sink, int regions) {
  int block = 512, grid = 148 * tps / block;
  uint64_t per_thread = (uint64_t)(slots * load) / ((uint64_t)grid * block);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaMemset(tab, 0, total_slots * 4);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for(int r = 0; r < regions; ++r) k<MODE, ILP><<<grid, block>>>(tab, slots - 1, (uint64_t)r * slots, per_thread, sink, 1234 + r);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return (float)(per_thread * grid * block * (double)regions / ms / 1e6);   // G keys/s
}
int main() {
  uint64_t total = 1ull << 31;
  uint32_t* tab; cudaMalloc(&tab, total * 4 + (1 << 20)); uint32_t* sink; cudaMalloc(&sink, 4);
  for(int lg : {22, 23, 24}) {
    uint64_t slots = 1ull << lg; int regions = 64;
    for(double load : {0.46, 0.58}) for(int tps : {1024, 2048}) {
      printf("region %3.0f MB load %.2f thr/SM %d: lookahead ilp4 %.1f  ilp1 %.1f | sequential-cas ilp4 %.1f G keys/s\n", slots * 4.0 / 1048576, load, tps,
             run<1, 4>(tab, slots, total, tps, load, sink, regions), run<1, 1>(tab, slots, total, tps, load, sink, regions), run<0, 4>(tab, slots, total, tps, load, sink, regions));
      fflush(stdout);
    }
  }
  return 0;
}
