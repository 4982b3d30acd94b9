// ubench.cu — pipe-level microbenchmarks behind the kernel design decisions (developer tool, not product).
//   A: register-only radix-16 butterflies (dft16_pretw) at 1..4 warps per SM sub-partition: how many warps
//      does it take to fill the FP64 / FP32 pipe?
//   B: the shared-memory exchange alone (16 stores, barrier, 16 loads per thread) at 1 and 2 CTAs per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I topdogspectrumanalyser_b200/csrc tools/ubench.cu -o tools/bin/ubench
#include <cstdio>
#include <vector>

#include "tdsa_fft.cuh"

using namespace tdsa;

template <typename T, int ITER>
__global__ void __launch_bounds__(512, 1) k_dft(T* out, long long* cyc) {
  T re[16], im[16], wr[16], wi[16], br[4], bi[4];   // four distinct twiddles keep the register budget of the real kernel
#pragma unroll
  for (int j = 0; j < 4; ++j) { br[j] = T(0.05) + T(1e-4) * T((threadIdx.x + j) & 7); bi[j] = T(0.03) + T(1e-3) * T(j); }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    re[j] = T(threadIdx.x + j) * T(1e-3); im[j] = T(j) * T(2e-3);
    wr[j] = br[j & 3]; wi[j] = bi[j & 3];
  }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) dft16_pretw<T>(re, im, wr, wi);
  const long long t1 = clock64();
  T s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += re[j] + im[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// exchange: thread t writes 16 elements at t + 256*q (pass-0 pattern, padded), barrier, reads c + 16*j of block s
template <typename T, int ITER, bool WARP_LOCAL>
__global__ void __launch_bounds__(256, 2) k_exch(T* out, long long* cyc) {
  using P = Plan<T, 12, 4>;
  using CT = typename CplxOf<T>::type;
  extern __shared__ __align__(128) unsigned char sm[];
  CT* ex = reinterpret_cast<CT*>(sm);
  const int t = threadIdx.x;
  T re[16], im[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { re[j] = T(t + j); im[j] = T(j); }
  const int c = t & 15, s = t >> 4;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if constexpr (!WARP_LOCAL) {
      const int pb = P::phys(t);
#pragma unroll
      for (int q = 0; q < 16; ++q) ex[pb + P::phys(q * 256)] = mk<T>(re[q], im[q]);
      __syncthreads();
      const int pr = P::phys(s * 256 + c);
#pragma unroll
      for (int j = 0; j < 16; ++j) { CT x = ex[pr + P::phys(j * 16)]; re[j] += x.x; im[j] += x.y; }
      __syncthreads();
    } else {
      // half-warp private 16x16 transpose, row pitch 17 elements
      CT* my = ex + s * (17 * 16);
#pragma unroll
      for (int q = 0; q < 16; ++q) my[c * 17 + q] = mk<T>(re[q], im[q]);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 16; ++j) { CT x = my[j * 17 + c]; re[j] += x.x; im[j] += x.y; }
      __syncwarp();
    }
  }
  const long long t1 = clock64();
  T sacc = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) sacc += re[j] + im[j];
  out[blockIdx.x * blockDim.x + t] = sacc;
  if (t == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename T> void run_dft(const char* name) {
  constexpr int ITER = 64;
  T* out; long long* cyc;
  cudaMalloc(&out, sizeof(T) * 148 * 2 * 512);
  cudaMalloc(&cyc, sizeof(long long) * 148 * 2);
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; ++warps_per_smsp) {
    const int threads = 128 * warps_per_smsp;
    k_dft<T, ITER><<<148, threads>>>(out, cyc);
    cudaDeviceSynchronize();
    k_dft<T, ITER><<<148, threads>>>(out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * 148, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += (double)v; avg /= 148;
    printf("A %s dft16_pretw: %d warps/SMSP: %.0f cycles per dft16 per warp, %.1f cycles per warp-dft16 of SMSP pipe time (%s)\n",
           name, warps_per_smsp, avg / ITER, avg / ITER / warps_per_smsp, cudaGetErrorString(e));
  }
  cudaFree(out); cudaFree(cyc);
}

template <typename T, bool WL> void run_exch(const char* name) {
  constexpr int ITER = 64;
  using P = Plan<T, 12, 4>;
  T* out; long long* cyc;
  cudaMalloc(&out, sizeof(T) * 148 * 2 * 256);
  cudaMalloc(&cyc, sizeof(long long) * 148 * 2);
  const size_t smem = (size_t)P::PHYS_SIZE * 2 * sizeof(T) + 4096;
  auto kern = k_exch<T, ITER, WL>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int ctas = 1; ctas <= 2; ++ctas) {
    kern<<<148 * ctas, 256, smem>>>(out, cyc);
    cudaDeviceSynchronize();
    kern<<<148 * ctas, 256, smem>>>(out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148 * ctas);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * 148 * ctas, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += (double)v; avg /= (148 * ctas);
    printf("B %s exchange (%s): %d CTA/SM: %.0f cycles per exchange per CTA -> %.0f SM cycles per frame-exchange (%s)\n", name,
           WL ? "warp-local, syncwarp" : "CTA-wide, syncthreads", ctas, avg / ITER, avg / ITER / ctas, cudaGetErrorString(e));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run_dft<double>("f64");
  run_dft<float>("f32");
  run_exch<double, false>("f64");
  run_exch<double, true>("f64");
  run_exch<float, false>("f32");
  run_exch<float, true>("f32");
  return 0;
}
