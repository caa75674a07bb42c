// What do K1's record stores cost?  Every CTA appends to 2048 open chunks of its own (8 KB each, like the record pool):
// a thread picks a random chunk and writes the next 4 / 16 / 32 bytes of it.  Reports stores/s and the rate per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template<int BYTES>
__global__ void __launch_bounds__(1024, 1) k(uint8_t* pool, uint32_t iters, uint32_t n_chunks) {
  __shared__ uint32_t fill[2048];
  for(uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) fill[i] = 0;
  __syncthreads();
  uint32_t x = (blockIdx.x * 1024u + threadIdx.x) * 2654435761u + 12345u;
  uint8_t* arena = pool + (size_t)blockIdx.x * n_chunks * 8192;
  for(uint32_t it = 0; it < iters; ++it) {
    x = x * 1664525u + 1013904223u;
    const uint32_t p = (x >> 9) & 2047u;
    const uint32_t slot = atomicAdd(&fill[p], 1u);
    uint8_t* dst = arena + ((size_t)(p + 2048u * ((slot * BYTES / 8192u) % (n_chunks / 2048u))) * 8192u) + (slot * BYTES) % 8192u;
    if(BYTES == 4) *reinterpret_cast<uint32_t*>(dst) = x;
    else if(BYTES == 16) *reinterpret_cast<uint4*>(dst) = make_uint4(x, x, x, x);
    else { reinterpret_cast<uint4*>(dst)[0] = make_uint4(x, x, x, x); reinterpret_cast<uint4*>(dst)[1] = make_uint4(x, x, x, x); }
  }
}
template<int BYTES> void run(uint8_t* pool, uint32_t n_chunks) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const uint32_t iters = 8192;
  k<BYTES><<<148, 1024>>>(pool, 256, n_chunks);
  cudaEventRecord(a);
  k<BYTES><<<148, 1024>>>(pool, iters, n_chunks);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double n = 148.0 * 1024 * iters;
  printf("%2d-byte appends to 2048 chunks per CTA: %7.2f G stores/s = %5.3f stores/clk/SM, %7.1f GB/s  [%s]\n", BYTES, n / ms / 1e6, n / ms / 1e6 / 148 / 1.95,
         n * BYTES / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const uint32_t n_chunks = 2048 * 8;                 // 128 MB per CTA, 19 GB in all
  uint8_t* pool; cudaMalloc(&pool, (size_t)148 * n_chunks * 8192);
  run<4>(pool, n_chunks); run<16>(pool, n_chunks); run<32>(pool, n_chunks);
  return 0;
}
