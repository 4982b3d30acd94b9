// Does a tiled tensor map with elementStrides = {2, 1, 1} deliver every other 8-byte element (complex64 sample) of a
// row, with the 128-byte swizzle, and does the start coordinate 1 select the odd samples?  (Plan for N = 8192 as two
// decimation-in-time halves through the one-engine 4096-point kernel.)  Prints, per variant, whether encoding worked and
// what landed in shared memory.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_stride_probe tools/tma_stride_probe.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int c0, int frame, uint64_t* out, int n_out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t bar_u32 = smem_u32(&bar), dst = smem_u32(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u32));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) reinterpret_cast<uint64_t*>(smem)[i] = 0xdeadbeefdeadbeefull;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_u32), "r"(n_out * 8) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"((uint64_t)&tmap), "r"(c0), "r"(0), "r"(frame), "r"(bar_u32)
                 : "memory");
  }
  // bounded wait: a wrong byte count must not hang the probe
  bool ok = false;
  for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_u32), "r"(0) : "memory");
    ok = done != 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) out[i] = reinterpret_cast<uint64_t*>(smem)[i];
  if (threadIdx.x == 0) out[n_out] = ok ? 1 : 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encode fn\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int frames = 3, rows = 256, per_row = 32;            // 8192 samples per frame as [256 rows][32 samples]
  const size_t n = (size_t)frames * rows * per_row;
  uint64_t* h = new uint64_t[n];
  for (size_t i = 0; i < n; ++i) h[i] = i;                   // value = global sample index
  uint64_t* d; cudaMalloc(&d, n * 8); cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
  const int n_out = 4096;
  uint64_t* dout; cudaMalloc(&dout, (n_out + 1) * 8);
  uint64_t* hout = new uint64_t[n_out + 1];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  struct Variant { const char* name; CUtensorMapDataType dt; cuuint64_t dim0; cuuint32_t box0, estr0; int elem_bytes; };
  const Variant vs[] = {
      {"uint64, box0=32, stride 2", CU_TENSOR_MAP_DATA_TYPE_UINT64, 32, 32, 2, 8},
      {"uint64, box0=16, stride 2", CU_TENSOR_MAP_DATA_TYPE_UINT64, 32, 16, 2, 8},
      {"float64, box0=32, stride 2", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 32, 32, 2, 8},
  };
  for (const Variant& v : vs) {
    CUtensorMap tm;
    const cuuint64_t dims[3] = {v.dim0, (cuuint64_t)rows, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)per_row * 8, (cuuint64_t)rows * per_row * 8};
    const cuuint32_t box[3] = {v.box0, 256, 1};
    const cuuint32_t estr[3] = {v.estr0, 1, 1};
    const CUresult r = enc(&tm, v.dt, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("%s: encode rc=%d\n", v.name, (int)r);
    if (r != CUDA_SUCCESS) continue;
    for (int c0 = 0; c0 < 2; ++c0) {
      cudaMemset(dout, 0, (n_out + 1) * 8);
      probe<<<1, 256, 40 * 1024>>>(tm, c0, 1, dout, n_out);
      const cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(hout, dout, (n_out + 1) * 8, cudaMemcpyDeviceToHost);
      printf("  c0=%d: %s, barrier completed=%llu\n", c0, cudaGetErrorString(e), (unsigned long long)hout[n_out]);
      if (e != cudaSuccess) return 2;
      // expected with swizzle 128B: smem row m (128 B = 16 elements), 16-byte chunk index XOR (m & 7); element (m, col)
      // should be global sample frame*8192 + 32 m + 2 col + c0
      int good = 0, filled = 0;
      for (int m = 0; m < 256; ++m)
        for (int col = 0; col < 16; ++col) {
          const int chunk = (col >> 1) ^ (m & 7);
          const uint64_t got = hout[m * 16 + chunk * 2 + (col & 1)];
          if (got != 0xdeadbeefdeadbeefull) ++filled;
          if (got == (uint64_t)(1 * 8192 + 32 * m + 2 * col + c0)) ++good;
        }
      printf("    filled %d / 4096, as expected %d / 4096; row 0: ", filled, good);
      for (int i = 0; i < 16; ++i) printf("%llu ", (unsigned long long)hout[i]);
      printf("| row 1: ");
      for (int i = 16; i < 32; ++i) printf("%llu ", (unsigned long long)hout[i]);
      printf("\n");
    }
  }
  return 0;
}
