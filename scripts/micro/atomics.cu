// Microbenchmark: random 32-bit atomics / loads over a region of a given size (B200 ceilings).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }
template<int OP, int ILP>
__global__ void k(uint32_t* tab, uint64_t mask, uint64_t base, uint64_t per_thread, uint32_t* sink, uint64_t seed) {
  uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for(uint64_t i = 0; i < per_thread; i += ILP) {
    uint32_t r[ILP];
#pragma unroll
    for(int j = 0; j < ILP; ++j) {
      uint64_t idx = base + (mix(seed + gid * per_thread + i + j) & mask);
      if(OP == 0) r[j] = atomicCAS(&tab[idx], 0u, (uint32_t)(idx | 1));
      else if(OP == 1) { atomicAdd(&tab[idx], 1u); r[j] = 0; }
      else if(OP == 2) r[j] = __ldcg(&tab[idx]);
      else r[j] = atomicAdd(&tab[idx], 1u);
    }
#pragma unroll
    for(int j = 0; j < ILP; ++j) acc += r[j];
  }
  if(acc == 0x12345678) *sink = acc;
}
template<int OP, int ILP>
float run(uint32_t* tab, uint64_t slots, uint64_t total_slots, int threads_per_sm, uint64_t ops, uint32_t* sink, int regions) {
  int nsm = 148;
  int block = 512, grid = nsm * threads_per_sm / block;
  uint64_t per_thread = ops / ((uint64_t)grid * block) / regions;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaMemset(tab, 0, total_slots * 4);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for(int r = 0; r < regions; ++r) k<OP, ILP><<<grid, block>>>(tab, slots - 1, (uint64_t)r * slots, per_thread, sink, 1234 + r);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return (float)(per_thread * grid * block * (double)regions / ms / 1e6);   // G ops/s
}
int main() {
  uint64_t total = 1ull << 33;   // 32 GB of uint32
  uint32_t* tab; cudaMalloc(&tab, total * 4); uint32_t* sink; cudaMalloc(&sink, 4);
  printf("region_MB regions op ilp thr/SM  Gops/s\n");
  const char* names[] = {"cas", "red", "ld", "atomadd"};
  for(int lg = 21; lg <= 33; lg += (lg < 27 ? 2 : 3)) {   // slots per region: 2^21 (8MB) ... 2^33 (32GB)
    uint64_t slots = 1ull << lg;
    // touch each region with ~0.5 ops per slot (like a table at load 0.5), several regions in sequence
    int regions = (int)((total / slots) < 64 ? (total / slots) : 64);
    uint64_t ops = slots / 2 * regions;
    if(ops < (1ull << 28)) { regions = 64; ops = slots / 2 * regions; }
    for(int tps : {1024, 2048}) {
      float v0 = run<0, 1>(tab, slots, total, tps, ops, sink, regions);
      float v1 = run<0, 4>(tab, slots, total, tps, ops, sink, regions);
      float v2 = run<1, 4>(tab, slots, total, tps, ops, sink, regions);
      float v3 = run<2, 4>(tab, slots, total, tps, ops, sink, regions);
      float v4 = run<3, 4>(tab, slots, total, tps, ops, sink, regions);
      printf("%8.0f %4d thr/SM=%d  cas1=%.1f cas4=%.1f red4=%.1f ld4=%.1f atomadd4=%.1f\n", slots * 4.0 / 1048576, regions, tps, v0, v1, v2, v3, v4);
      fflush(stdout);
    }
  }
  // heavier load per region: 2 ops per slot (hits dominate)
  for(int lg : {23, 24, 25}) {
    uint64_t slots = 1ull << lg; int regions = 64; uint64_t ops = slots * 2 * regions;
    float v1 = run<0, 4>(tab, slots, total, 2048, ops, sink, regions);
    float v2 = run<1, 4>(tab, slots, total, 2048, ops, sink, regions);
    printf("hot %8.0f MB: cas4=%.1f red4=%.1f\n", slots * 4.0 / 1048576, v1, v2);
  }
  return 0;
}
