// Microbenchmark for the round-2 K2 candidate ("window insert"): records already grouped by a
// shared-memory-sized window of the table; one CTA per window loads the window's slots into
// shared memory, applies its records with shared-memory CAS (quadratic reprobe inside the window,
// probes leaving the window are deferred), and stores the window back. Models 32-bit slots:
// [counter:7 | key field:25], 0 = empty. Reports G records/s including the window load/store and
// the record stream, i.e. what K2b would cost per drained record.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_window smem_window.cu && ./smem_window
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }

// records of window w: rec[w * per_win + i] = [key field : 32-LG bits | home slot in window : LG bits], 4 B like the engine's 32-bit records
template<int LG>
__global__ void gen(uint32_t* rec, uint64_t n, uint64_t seed, uint32_t distinct_per_win, uint64_t per_win) {
  for(uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t w = i / per_win;
    uint64_t id = mix(seed + i) % distinct_per_win;          // which distinct key of this window
    uint64_t h = mix(w * 0x100000001B3ull + id);
    uint32_t home = (uint32_t)h & ((1u << LG) - 1);
    uint32_t kf = (uint32_t)(h >> 32) & ((1u << (32 - LG)) - 1);   // record = [kf | home] in 32 bits
    rec[i] = kf << LG | home;
  }
}

template<int LG, int NTH>
__global__ void __launch_bounds__(NTH) window_insert(uint32_t* __restrict__ table, const uint32_t* __restrict__ rec, uint64_t per_win, uint32_t n_win, unsigned long long* deferred) {
  extern __shared__ uint32_t win[];
  constexpr uint32_t SLOTS = 1u << LG;
  constexpr uint32_t KF_BITS = 25, KF_MASK = (1u << KF_BITS) - 1, ONE = 1u << KF_BITS;
  for(uint32_t w = blockIdx.x; w < n_win; w += gridDim.x) {
    uint4* gw = reinterpret_cast<uint4*>(table + (uint64_t)w * SLOTS);
    uint4* sw = reinterpret_cast<uint4*>(win);
    for(uint32_t i = threadIdx.x; i < SLOTS / 4; i += NTH) sw[i] = __ldcs(gw + i);
    __syncthreads();
    const uint4* r4 = reinterpret_cast<const uint4*>(rec + (uint64_t)w * per_win);
    uint32_t ndef = 0;
    for(uint64_t i = threadIdx.x; i < per_win / 4; i += NTH) {
      uint4 v = __ldcs(r4 + i);
      uint32_t rr[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for(int j = 0; j < 4; ++j) {
        uint32_t home = rr[j] & (SLOTS - 1);
        uint32_t kf = ((rr[j] >> LG) | 1u) & KF_MASK;       // non-zero key field (engine: reprobe+1 in the low bits)
        uint32_t p = home; bool done = false;
        for(uint32_t t = 1; t <= 62 && !done; ++t) {
          if(p >= SLOTS) { ++ndef; done = true; break; }     // probe leaves the window: deferred to the global pass
          uint32_t old = atomicCAS(&win[p], 0u, kf | ONE);
          if(old == 0) done = true;
          else if((old & KF_MASK) == kf) { atomicAdd(&win[p], ONE); done = true; }   // (engine: carry into the overflow table)
          else p = home + t * (t + 1) / 2;
        }
        if(!done) ++ndef;
      }
    }
    __syncthreads();
    for(uint32_t i = threadIdx.x; i < SLOTS / 4; i += NTH) __stcs(gw + i, sw[i]);
    __syncthreads();
    if(ndef) atomicAdd(deferred, (unsigned long long)ndef);
  }
}

template<int LG, int NTH>
void run(uint32_t* table, uint32_t* rec, uint32_t n_win, double load, double dup, unsigned long long* d_def) {
  constexpr uint32_t SLOTS = 1u << LG;
  uint32_t distinct = (uint32_t)(SLOTS * load);
  uint64_t per_win = ((uint64_t)(distinct * dup) + 3) & ~3ull;       // records per window (dup = occurrences per distinct key)
  uint64_t n = per_win * n_win;
  gen<LG><<<148 * 8, 256>>>(rec, n, 777, distinct, per_win);
  cudaMemset(table, 0, (uint64_t)n_win * SLOTS * 4); cudaMemset(d_def, 0, 8);
  size_t smem = SLOTS * 4;
  cudaFuncSetAttribute(window_insert<LG, NTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, window_insert<LG, NTH>, NTH, smem);
  if(per_sm < 1) { printf("window 2^%d slots: does not fit\n", LG); return; }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  window_insert<LG, NTH><<<148 * per_sm, NTH, smem>>>(table, rec, per_win, n_win, d_def);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  unsigned long long def; cudaMemcpy(&def, d_def, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  double bytes = (double)n_win * SLOTS * 8 + (double)n * 4;
  printf("window %3u KB x %u, %4d thr x %d CTA/SM, load %.2f dup %.1f: %7.1f G rec/s  %6.2f ms  %5.0f GB/s  deferred %.4f%%  %s\n", SLOTS * 4 / 1024, n_win, NTH, per_sm, load, dup,
         n / ms / 1e6, ms, bytes / ms / 1e6, 100.0 * def / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
  fflush(stdout);
}

int main() {
  const uint64_t table_bytes = 16ull << 30;                 // a 16 GiB slice of table (> L2, so window traffic is HBM)
  uint32_t* table; cudaMalloc(&table, table_bytes);
  uint32_t* rec; cudaMalloc(&rec, table_bytes);             // up to 1 record per slot
  unsigned long long* d_def; cudaMalloc(&d_def, 8);
  for(double load : {0.3, 0.6}) for(double dup : {1.0, 3.0}) {
    if(load * dup > 1.0) continue;
    run<13, 256>(table, rec, (uint32_t)(table_bytes >> 15), load, dup, d_def);
    run<14, 256>(table, rec, (uint32_t)(table_bytes >> 16), load, dup, d_def);
    run<14, 512>(table, rec, (uint32_t)(table_bytes >> 16), load, dup, d_def);
    run<15, 1024>(table, rec, (uint32_t)(table_bytes >> 17), load, dup, d_def);
  }
  return 0;
}
