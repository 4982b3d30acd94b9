#include <cstdio>
#include <stdint.h>
#include <cuda_runtime.h>
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
               "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__global__ void k(float* out, int iters) {
  __shared__ uint32_t base_s;
  const int w = threadIdx.x >> 5;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = base_s;
  const uint32_t taddr = base + ((uint32_t)((w & 3) * 32) << 16) + (uint32_t)((w >> 2) * 64);
  uint32_t r[16];
  for (int j = 0; j < 16; ++j) r[j] = __float_as_uint((float)(threadIdx.x * 100 + j));
  tmem_st16(taddr, r);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  for (int it = 0; it < iters; ++it) {
    uint32_t q[16];
    tmem_ld16(taddr, q);
    for (int j = 0; j < 16; ++j) q[j] = __float_as_uint(__uint_as_float(q[j]) + 1.0f);
    tmem_st16(taddr, q);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  uint32_t q[16];
  tmem_ld16(taddr, q);
  for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = __uint_as_float(q[j]);
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(128));
}
int main() {
  float* d; cudaMalloc(&d, 256 * 16 * 4);
  k<<<2, 256>>>(d, 3);
  cudaError_t e = cudaDeviceSynchronize();
  float h[256 * 16]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int t = 0; t < 256; ++t) for (int j = 0; j < 16; ++j) if (h[t * 16 + j] != (float)(t * 100 + j) + 3.0f) ++bad;
  printf("tmem test: %s bad=%d h[0]=%f h[4095]=%f\n", cudaGetErrorString(e), bad, h[0], h[4095]);
  return bad != 0;
}
