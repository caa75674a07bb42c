// Microbenchmark for the round-2 K2 candidate, first half ("window partition"): the 4-byte records
// of one L2 region (here: a contiguous run of R records) are split into NB window bins with a
// shared-memory staged tile sort, so that each window's records end up contiguous.
// pass 1: per-region histogram of the bins; pass 2: exclusive scan (tiny); pass 3: per tile, rank
// by shared-memory atomics, reserve output ranges with one global atomic per (tile, bin), local
// sort in shared memory, write runs. Reports G records/s for pass 1+3 over all regions.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o subpart subpart.cu && ./subpart
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }
__global__ void gen(uint32_t* rec, uint64_t n) {
  for(uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) rec[i] = (uint32_t)mix(i);
}
template<int NBLG, int SHIFT, int NTH, int PER>
__global__ void __launch_bounds__(NTH) hist(const uint32_t* __restrict__ rec, uint64_t R, uint32_t tiles_per_region, uint32_t* __restrict__ counts) {
  constexpr int NB = 1 << NBLG, TILE = NTH * PER;
  __shared__ uint32_t h[NB];
  uint32_t region = blockIdx.x / tiles_per_region, tile = blockIdx.x % tiles_per_region;
  for(int i = threadIdx.x; i < NB; i += NTH) h[i] = 0;
  __syncthreads();
  const uint32_t* src = rec + (uint64_t)region * R + (uint64_t)tile * TILE;
  uint64_t left = R - (uint64_t)tile * TILE; if(left > TILE) left = TILE;
  for(uint32_t i = threadIdx.x * 4; i < left; i += NTH * 4) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i));
    atomicAdd(&h[(v.x >> SHIFT) & (NB - 1)], 1); atomicAdd(&h[(v.y >> SHIFT) & (NB - 1)], 1);
    atomicAdd(&h[(v.z >> SHIFT) & (NB - 1)], 1); atomicAdd(&h[(v.w >> SHIFT) & (NB - 1)], 1);
  }
  __syncthreads();
  for(int i = threadIdx.x; i < NB; i += NTH) if(h[i]) atomicAdd(&counts[(uint64_t)region * NB + i], h[i]);
}
template<int NBLG>
__global__ void scan(const uint32_t* __restrict__ counts, uint32_t* __restrict__ cursor, uint64_t R) {   // one CTA of NB threads per region
  constexpr int NB = 1 << NBLG;
  __shared__ uint32_t s[NB];
  uint32_t c = counts[(uint64_t)blockIdx.x * NB + threadIdx.x];
  s[threadIdx.x] = c; __syncthreads();
  for(int d = 1; d < NB; d <<= 1) { uint32_t v = threadIdx.x >= d ? s[threadIdx.x - d] : 0; __syncthreads(); s[threadIdx.x] += v; __syncthreads(); }
  cursor[(uint64_t)blockIdx.x * NB + threadIdx.x] = s[threadIdx.x] - c;     // exclusive, relative to the region's output base
}
template<int NBLG, int SHIFT, int NTH, int PER>
__global__ void __launch_bounds__(NTH) scatter(const uint32_t* __restrict__ rec, uint32_t* __restrict__ out, uint64_t R, uint32_t tiles_per_region, uint32_t* __restrict__ cursor) {
  constexpr int NB = 1 << NBLG, TILE = NTH * PER;
  __shared__ uint32_t cnt[NB], lbase[NB], gbase[NB];
  extern __shared__ uint32_t stage[];                      // TILE records
  uint32_t region = blockIdx.x / tiles_per_region, tile = blockIdx.x % tiles_per_region;
  for(int i = threadIdx.x; i < NB; i += NTH) cnt[i] = 0;
  __syncthreads();
  const uint32_t* src = rec + (uint64_t)region * R + (uint64_t)tile * TILE;
  uint64_t left = R - (uint64_t)tile * TILE; if(left > TILE) left = TILE;
  uint32_t r[PER], rank[PER];
#pragma unroll
  for(int j = 0; j < PER / 4; ++j) {
    uint32_t i = (j * NTH + threadIdx.x) * 4;
    uint4 v = i < left ? __ldcs(reinterpret_cast<const uint4*>(src + i)) : make_uint4(0, 0, 0, 0);
    r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
#pragma unroll
    for(int q = 0; q < 4; ++q) rank[4 * j + q] = i + q < left ? atomicAdd(&cnt[(r[4 * j + q] >> SHIFT) & (NB - 1)], 1) : 0;
  }
  __syncthreads();
  // exclusive scan of cnt -> lbase (NB <= NTH assumed), reserve global ranges
  if(threadIdx.x < NB) lbase[threadIdx.x] = cnt[threadIdx.x];
  __syncthreads();
  for(int d = 1; d < NB; d <<= 1) { uint32_t v = 0; if(threadIdx.x < NB && threadIdx.x >= d) v = lbase[threadIdx.x - d]; __syncthreads(); if(threadIdx.x < NB) lbase[threadIdx.x] += v; __syncthreads(); }
  if(threadIdx.x < NB) {
    uint32_t c = cnt[threadIdx.x];
    lbase[threadIdx.x] -= c;
    gbase[threadIdx.x] = c ? atomicAdd(&cursor[(uint64_t)region * NB + threadIdx.x], c) : 0;
  }
  __syncthreads();
#pragma unroll
  for(int j = 0; j < PER / 4; ++j) {
    uint32_t i = (j * NTH + threadIdx.x) * 4;
#pragma unroll
    for(int q = 0; q < 4; ++q) if(i + q < left) stage[lbase[(r[4 * j + q] >> SHIFT) & (NB - 1)] + rank[4 * j + q]] = r[4 * j + q];
  }
  __syncthreads();
  uint32_t* dst = out + (uint64_t)region * R;
  for(uint32_t i = threadIdx.x; i < left; i += NTH) {
    uint32_t v = stage[i], b = (v >> SHIFT) & (NB - 1);
    dst[gbase[b] + (i - lbase[b])] = v;
  }
}
template<int NBLG, int SHIFT, int NTH, int PER>
void run(uint32_t* rec, uint32_t* out, uint32_t* counts, uint32_t* cursor, uint64_t n, uint64_t R) {
  constexpr int NB = 1 << NBLG, TILE = NTH * PER;
  uint32_t regions = (uint32_t)(n / R), tpr = (uint32_t)((R + TILE - 1) / TILE);
  size_t smem = (size_t)TILE * 4;
  cudaFuncSetAttribute(scatter<NBLG, SHIFT, NTH, PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemset(counts, 0, (uint64_t)regions * NB * 4);
  cudaEvent_t a, b, c; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventCreate(&c);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  hist<NBLG, SHIFT, NTH, PER><<<regions * tpr, NTH>>>(rec, R, tpr, counts);
  scan<NBLG><<<regions, NB>>>(counts, cursor, R);
  cudaEventRecord(b);
  scatter<NBLG, SHIFT, NTH, PER><<<regions * tpr, NTH, smem>>>(rec, out, R, tpr, cursor);
  cudaEventRecord(c); cudaEventSynchronize(c);
  float m1, m2; cudaEventElapsedTime(&m1, a, b); cudaEventElapsedTime(&m2, b, c);
  cudaError_t e = cudaGetLastError();
  // check: every output record of region 0 lies in a non-decreasing bin sequence
  uint32_t* h = (uint32_t*)malloc(R * 4); cudaMemcpy(h, out, R * 4, cudaMemcpyDeviceToHost);
  uint64_t bad = 0; for(uint64_t i = 1; i < R; ++i) if(((h[i] >> SHIFT) & (NB - 1)) < ((h[i - 1] >> SHIFT) & (NB - 1))) ++bad;
  free(h);
  printf("%4d bins, tile %5d (%d thr x %d): hist+scan %6.2f ms, scatter %6.2f ms -> %6.1f G rec/s (%4.0f GB/s of 12 B/rec)  order violations %llu %s\n", NB, TILE, NTH, PER, m1, m2,
         n / (m1 + m2) / 1e6, 12.0 * n / (m1 + m2) / 1e6, (unsigned long long)bad, e == cudaSuccess ? "" : cudaGetErrorString(e));
  fflush(stdout);
}
int main() {
  const uint64_t R = 5ull << 20, regions = 512, n = R * regions;      // 2.7 G records, 10.7 GB in + out
  uint32_t *rec, *out, *counts, *cursor;
  cudaMalloc(&rec, n * 4); cudaMalloc(&out, n * 4); cudaMalloc(&counts, regions * 1024 * 4); cudaMalloc(&cursor, regions * 1024 * 4);
  gen<<<148 * 8, 256>>>(rec, n);
  run<8, 15, 512, 16>(rec, out, counts, cursor, n, R);
  run<9, 14, 512, 16>(rec, out, counts, cursor, n, R);
  run<9, 14, 512, 32>(rec, out, counts, cursor, n, R);
  run<9, 14, 1024, 16>(rec, out, counts, cursor, n, R);
  run<10, 13, 1024, 16>(rec, out, counts, cursor, n, R);
  return 0;
}
