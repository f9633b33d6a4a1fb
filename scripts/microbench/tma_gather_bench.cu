// microbenchmark: per-SM rate of 16-byte cp.async.bulk gathers vs LDG.64 gathers from a 2 MB L2-resident table
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(512,1) k_tma(const uint4* tab, uint32_t mask, int iters, unsigned long long* out){
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = (uint64_t*)smem;                // one mbarrier per warp
  unsigned char* land = smem + 256;               // 512 threads x 8 x 16 B = 64 KB
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(bar+warp))); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1;
  uint32_t parity = 0;
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(bar+warp)), "r"(32*8*16) : "memory");
    __syncwarp();
    #pragma unroll
    for (int j = 0; j < 8; j++) {
      x = x * 1664525u + 1013904223u;
      const uint4* src = tab + ((x >> 8) & mask);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" :: "r"(s32(land + (threadIdx.x*8 + j)*16)), "l"(src), "r"(s32(bar+warp)) : "memory");
    }
    asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" :: "r"(s32(bar+warp)), "r"(parity) : "memory");
    parity ^= 1;
  }
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (land[threadIdx.x] == 123 && iters < 0) out[0] = 0;
}
__global__ void __launch_bounds__(512,1) k_ldg(const uint2* tab, uint32_t mask, int iters, unsigned long long* out){
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1;
  unsigned long long t0 = clock64();
  float acc = 0;
  for (int it = 0; it < iters; it++) {
    uint2 v[8];
    #pragma unroll
    for (int j = 0; j < 8; j++) { x = x * 1664525u + 1013904223u; v[j] = __ldg(tab + ((x >> 8) & mask)); }
    #pragma unroll
    for (int j = 0; j < 8; j++) acc += __uint_as_float(v[j].x) + __uint_as_float(v[j].y);
  }
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 1.2345f) out[0] = 0;
}
int main(){
  const size_t bytes = 2u<<20; void* tab; cudaMalloc(&tab, bytes); cudaMemset(tab, 0, bytes);
  unsigned long long* out; cudaMallocManaged(&out, 148*8);
  cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  int iters = 200;
  for (int rep = 0; rep < 2; rep++) {
    k_tma<<<148,512,70000>>>((const uint4*)tab, (uint32_t)(bytes/16-1), iters, out); cudaError_t e = cudaDeviceSynchronize();
    double c = 0; for (int i=0;i<148;i++) c += out[i]; c /= 148;
    printf("TMA 16B gathers: %s  %.0f cycles for %d x 4096 copies per SM -> %.2f cycles per copy per SM\n", cudaGetErrorString(e), c, iters, c/(iters*4096.0));
    k_ldg<<<148,512>>>((const uint2*)tab, (uint32_t)(bytes/8-1), iters, out); e = cudaDeviceSynchronize();
    c = 0; for (int i=0;i<148;i++) c += out[i]; c /= 148;
    printf("LDG.64 gathers : %s  %.0f cycles for %d x 4096 loads per SM  -> %.2f cycles per load per SM\n", cudaGetErrorString(e), c, iters, c/(iters*4096.0));
  }
  return 0;
}
