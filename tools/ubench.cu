// ubench.cu — pipe-level microbenchmarks behind the kernel design decisions (developer tool, not product).
//   A: register-only radix-16 butterflies (dft16_pretw) at 1..4 warps per SM sub-partition: how many warps
//      does it take to fill the FP64 / FP32 pipe?
//   B: the shared-memory exchange alone (16 stores, barrier, 16 loads per thread) at 1 and 2 CTAs per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I topdogspectrumanalyser_b200/csrc tools/ubench.cu -o tools/bin/ubench
#include <algorithm>
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

// C: do FP sections and shared-memory traffic of DIFFERENT warps overlap?  512 threads: warps 0-7 run the register-only
// radix-16 loop, warps 8-15 run the warp-local exchange loop (mode 3), or only one half works (modes 1, 2).
template <typename T, int ITER>
__global__ void __launch_bounds__(512, 1) k_mix(T* out, long long* cyc, int mode) {
  using CT = typename CplxOf<T>::type;
  extern __shared__ __align__(128) unsigned char sm[];
  CT* ex = reinterpret_cast<CT*>(sm);
  const int t = threadIdx.x, half = t >> 8, tt = t & 255;
  T re[16], im[16], wr[16], wi[16], br[4], bi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { br[j] = T(0.05) + T(1e-4) * T((t + j) & 7); bi[j] = T(0.03) + T(1e-3) * T(j); }
#pragma unroll
  for (int j = 0; j < 16; ++j) { re[j] = T(t + j) * T(1e-3); im[j] = T(j) * T(2e-3); wr[j] = br[j & 3]; wi[j] = bi[j & 3]; }
  __syncthreads();
  const long long t0 = clock64();
  if (half == 0) {
    if (mode & 1) {
#pragma unroll 1
      for (int it = 0; it < ITER; ++it) dft16_pretw<T>(re, im, wr, wi);
    }
  } else {
    if (mode & 2) {
      const int c = tt & 15, s = tt >> 4;
      CT* my = ex + s * (17 * 16);
#pragma unroll 1
      for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) my[c * 17 + q] = mk<T>(re[q], im[q]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; ++j) { CT x = my[j * 17 + c]; re[j] += x.x; im[j] += x.y; }
        __syncwarp();
      }
    }
  }
  const long long t1 = clock64();
  T sacc = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) sacc += re[j] + im[j];
  out[blockIdx.x * blockDim.x + t] = sacc;
  if ((t & 31) == 0) cyc[blockIdx.x * 16 + (t >> 5)] = t1 - t0;
}

template <typename T> void run_mix(const char* name) {
  constexpr int ITER = 64;
  T* out; long long* cyc;
  cudaMalloc(&out, sizeof(T) * 148 * 512);
  cudaMalloc(&cyc, sizeof(long long) * 148 * 16);
  const size_t smem = (size_t)16 * 17 * 16 * 2 * sizeof(T) + 1024;
  auto kern = k_mix<T, ITER>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 1; mode <= 3; ++mode) {
    kern<<<148, 512, smem>>>(out, cyc, mode);
    cudaDeviceSynchronize();
    kern<<<148, 512, smem>>>(out, cyc, mode);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148 * 16);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * 148 * 16, cudaMemcpyDeviceToHost);
    double fp = 0, ex = 0;
    for (int b = 0; b < 148; ++b) {
      long long mf = 0, mx = 0;
      for (int w = 0; w < 8; ++w) { mf = std::max(mf, h[b * 16 + w]); mx = std::max(mx, h[b * 16 + 8 + w]); }
      fp += (double)mf; ex += (double)mx;
    }
    printf("C %s mode %d (%s): FP warps %.0f cycles per radix-16 (8 warps, slowest), exchange warps %.0f cycles per exchange (%s)\n", name, mode,
           mode == 1 ? "FP only" : mode == 2 ? "exchange only" : "both", fp / 148 / ITER, ex / 148 / ITER, cudaGetErrorString(e));
  }
  cudaFree(out); cudaFree(cyc);
}

// D: what do float<->double conversions and MUFU cost next to FP64 work?  One radix-16 (192 FP64 ops) per iteration plus
// 32 DADDs whose operands are (0) doubles, (1) converted floats (32 F2F.F64.F32), (2) as 0 plus 16 F2F.F32.F64 + 16 MUFU.LG2.
template <int MODE, int ITER>
__global__ void __launch_bounds__(512, 1) k_cvt(double* out, long long* cyc) {
  double re[16], im[16], wr[16], wi[16], br[4], bi[4];
  float v[32];
#pragma unroll
  for (int j = 0; j < 4; ++j) { br[j] = 0.05 + 1e-4 * ((threadIdx.x + j) & 7); bi[j] = 0.03 + 1e-3 * j; }
#pragma unroll
  for (int j = 0; j < 16; ++j) { re[j] = (threadIdx.x + j) * 1e-3; im[j] = j * 2e-3; wr[j] = br[j & 3]; wi[j] = bi[j & 3]; }
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = 1e-3f * (float)(threadIdx.x + j);
  float facc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (MODE == 1) { re[j] += (double)v[j]; im[j] += (double)v[16 + j]; }
      else { re[j] += wr[(j + 1) & 15]; im[j] += wi[(j + 1) & 15]; }
    }
    dft16_pretw<double>(re, im, wr, wi);
    if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < 16; ++j) facc += lg2_approx((float)re[j]);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += 1.0f;
  }
  const long long t1 = clock64();
  double s = facc;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += re[j] + im[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run_cvt(const char* what) {
  constexpr int ITER = 64;
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 512);
  cudaMalloc(&cyc, sizeof(long long) * 148);
  for (int wps = 1; wps <= 2; ++wps) {
    k_cvt<MODE, ITER><<<148, 128 * wps>>>(out, cyc);
    cudaDeviceSynchronize();
    k_cvt<MODE, ITER><<<148, 128 * wps>>>(out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * 148, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += (double)v; avg /= 148;
    printf("D %s: %d warps/SMSP: %.0f cycles per iteration per warp (%s)\n", what, wps, avg / ITER, cudaGetErrorString(e));
  }
  cudaFree(out); cudaFree(cyc);
}

// E: packed float32 (FFMA2 / FADD2, sm_100) against scalar FFMA: 16 independent accumulator chains per thread
template <int MODE, int ITER>
__global__ void __launch_bounds__(512, 1) k_pack(float* out, long long* cyc) {
  float2 acc[16], x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { acc[j] = make_float2(threadIdx.x * 1e-3f + j, j * 1e-2f); x[j] = make_float2(1.0f + 1e-6f * j, 1.0f - 1e-6f * j); }
  const float2 y = make_float2(1e-3f * (threadIdx.x & 3), 2e-3f);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (MODE == 0) { acc[j].x = __fmaf_rn(acc[j].x, x[j].x, y.x); acc[j].y = __fmaf_rn(acc[j].y, x[j].y, y.y); }   // 2 scalar FFMA
        else if (MODE == 1) acc[j] = __ffma2_rn(acc[j], x[j], y);                                                       // 1 FFMA2
        else acc[j] = __fadd2_rn(acc[j], x[j]);                                                                         // 1 FADD2
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += acc[j].x + acc[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + (threadIdx.x >> 5)] = t1 - t0;
}

template <int MODE> void run_pack(const char* what) {
  constexpr int ITER = 32;
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * 148 * 512);
  cudaMalloc(&cyc, sizeof(long long) * 148 * 16);
  for (int wps = 1; wps <= 4; wps *= 2) {
    k_pack<MODE, ITER><<<148, 128 * wps>>>(out, cyc);
    cudaDeviceSynchronize();
    k_pack<MODE, ITER><<<148, 128 * wps>>>(out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148 * 16);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * 148 * 16, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int b = 0; b < 148; ++b) { long long m = 0; for (int w = 0; w < 4 * wps; ++w) m = std::max(m, h[b * 16 + w]); mx += (double)m; }
    mx /= 148;
    // 128 complex (= 256 scalar) multiply-adds per thread and iteration
    printf("E %s: %d warps/SMSP: %.2f cycles per 32 lanes x 2 floats (slowest warp; %.0f cycles per iteration) (%s)\n", what, wps,
           mx / ITER / 128.0 / wps, mx / ITER, cudaGetErrorString(e));
  }
  cudaFree(out); cudaFree(cyc);
}

// F: the ping-pong skeleton alone.  512 threads = two groups of 8 warps; a group holds the math token (a pair of named
// barriers, as fft_wlpp_kernel did) only for its FP section (2 x radix-16) and runs its exchange without it.
template <int ITER, bool TOKEN>
__global__ void __launch_bounds__(512, 1) k_pingpong(double* out, long long* cyc) {
  extern __shared__ __align__(128) unsigned char sm[];
  double2* ex = reinterpret_cast<double2*>(sm);
  const int tall = threadIdx.x, g = tall >> 8, t = tall & 255;
  double re[16], im[16], wr[16], wi[16], br[4], bi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { br[j] = 0.05 + 1e-4 * ((t + j) & 7); bi[j] = 0.03 + 1e-3 * j; }
#pragma unroll
  for (int j = 0; j < 16; ++j) { re[j] = (t + j) * 1e-3; im[j] = j * 2e-3; wr[j] = br[j & 3]; wi[j] = bi[j & 3]; }
  const int c = t & 15, s = t >> 4;
  double2* my = ex + (g * 16 + s) * (17 * 16);
  long long fp_cycles = 0;
  __syncthreads();
  if (TOKEN && g == 1) bar_arrive(3, 512);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if (TOKEN) bar_sync(3 + g, 512);
    const long long f0 = clock64();
    dft16_pretw<double>(re, im, wr, wi);
    dft16_pretw<double>(re, im, wr, wi);
    fp_cycles += clock64() - f0;
    if (TOKEN) bar_arrive(3 + (1 - g), 512);
#pragma unroll
    for (int q = 0; q < 16; ++q) my[c * 17 + q] = make_double2(re[q], im[q]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) { double2 x = my[j * 17 + c]; re[j] = x.x * 1e-3; im[j] = x.y * 1e-3; }
    __syncwarp();
  }
  const long long t1 = clock64();
  double sacc = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) sacc += re[j] + im[j];
  out[blockIdx.x * 512 + tall] = sacc;
  if ((tall & 31) == 0) { cyc[(blockIdx.x * 16 + (tall >> 5)) * 2] = t1 - t0; cyc[(blockIdx.x * 16 + (tall >> 5)) * 2 + 1] = fp_cycles; }
}

template <bool TOKEN> void run_pingpong() {
  constexpr int ITER = 64;
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 512);
  cudaMalloc(&cyc, sizeof(long long) * 148 * 32);
  const size_t smem = (size_t)32 * 17 * 16 * 16 + 1024;
  auto kern = k_pingpong<ITER, TOKEN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<148, 512, smem>>>(out, cyc);
  cudaDeviceSynchronize();
  kern<<<148, 512, smem>>>(out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148 * 32);
  cudaMemcpy(h.data(), cyc, sizeof(long long) * 148 * 32, cudaMemcpyDeviceToHost);
  double tot = 0, fp = 0;
  for (int i = 0; i < 148 * 16; ++i) { tot += (double)h[2 * i]; fp += (double)h[2 * i + 1]; }
  printf("F ping-pong skeleton (%s): %.0f cycles per iteration per warp, of which %.0f in the FP section (2 x radix-16; 1600 = pipe-bound for a group alone) (%s)\n",
         TOKEN ? "math token" : "no token", tot / (148 * 16) / ITER, fp / (148 * 16) / ITER, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
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

// G: does SHFL share the shared-memory data path?  mode 0: all 16 warps shuffle (64 SHFL.BFLY per iteration);
// mode 1: all 16 warps do conflict-free LDS.128 + STS.128 (16 each); mode 2: warps 0-7 shuffle, warps 8-15 LDS/STS.
template <int ITER>
__global__ void __launch_bounds__(512, 1) k_shfl(float* out, long long* cyc, int mode) {
  extern __shared__ __align__(128) unsigned char sm[];
  float4* ex = reinterpret_cast<float4*>(sm);
  const int t = threadIdx.x, w = t >> 5;
  float v[16];
  float4 q[4];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = (float)(t + j);
#pragma unroll
  for (int j = 0; j < 4; ++j) q[j] = make_float4(t, j, 1, 2);
  const bool do_shfl = mode == 0 || (mode == 2 && w < 8);
  __syncthreads();
  const long long t0 = clock64();
  if (do_shfl) {
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], 1 << k);
    }
  } else {
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ex[t + 512 * j] = q[j];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float4 x = ex[(t ^ 1) + 512 * j]; q[j].x += x.x; q[j].y += x.w; }
        __syncwarp();
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += v[j];
#pragma unroll
  for (int j = 0; j < 4; ++j) s += q[j].x + q[j].y;
  out[blockIdx.x * blockDim.x + t] = s;
  if ((t & 255) == 0) cyc[blockIdx.x * 2 + (t >> 8)] = t1 - t0;
}

static void run_shfl() {
  constexpr int ITER = 64;
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * sizeof(float)); cudaMalloc(&cyc, 148 * 2 * sizeof(long long));
  cudaFuncSetAttribute(k_shfl<ITER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) k_shfl<ITER><<<148, 512, 64 * 1024>>>(out, cyc, mode);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148 * 2);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * 148 * 2, cudaMemcpyDeviceToHost);
    double a0 = 0, a1 = 0; for (int i = 0; i < 148; ++i) { a0 += h[2 * i]; a1 += h[2 * i + 1]; } a0 /= 148; a1 /= 148;
    // per iteration: shuffling warps issue 64 SHFL each; LDS/STS warps 16 STS.128 + 16 LDS.128 each (= 128 wavefronts per warp)
    printf("G mode %d (%s): warps 0-7 %.0f cycles/iter, warps 8-15 %.0f cycles/iter (%s)\n", mode,
           mode == 0 ? "16 warps SHFL: 1024 SHFL per iter and SM" : mode == 1 ? "16 warps LDS/STS.128: 2048 wavefronts per iter and SM"
                                                                         : "8 warps SHFL (512) + 8 warps LDS/STS (1024 wavefronts)",
           a0 / ITER, a1 / ITER, cudaGetErrorString(e));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run_shfl();
  run_pingpong<false>();
  run_pingpong<true>();
  run_pack<0>("2 x scalar FFMA");
  run_pack<1>("FFMA2");
  run_pack<2>("FADD2");
  run_cvt<0>("radix-16 + 32 DADD (224 FP64 ops)");
  run_cvt<1>("same + 32 F2F.F64.F32");
  run_cvt<2>("same as first + 16 F2F.F32.F64 + 16 MUFU.LG2");
  run_mix<double>("f64");
  run_mix<float>("f32");
  run_dft<double>("f64");
  run_dft<float>("f32");
  run_exch<double, false>("f64");
  run_exch<double, true>("f64");
  run_exch<float, false>("f32");
  run_exch<float, true>("f32");
  return 0;
}
