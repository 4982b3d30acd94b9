// tdsa_big.cuh — head kernel of the two-kernel path for N = 256*M > one CTA's shared memory
// (config 3: 65536-point Welch).  The N-point in-place DIF is split after its first two
// radix-16 passes: this kernel runs passes 0 and 1 on a tile of 16 adjacent columns
// (4096 points, 256 threads x 16) and leaves 256 independent M-point sub-transforms in a
// scratch buffer that stays in L2; fft_fused_kernel<..., TAIL> finishes them.
#pragma once
#include "tdsa_fft.cuh"

namespace tdsa {

template <typename T> struct BigArgs {
  const float2* iq;
  int64_t n_frames;
  int64_t frame_stride;
  const T* window;                              // T[N] with (-1)^n folded in
  const typename CplxOf<T>::type* tw;           // N-point plan tables: pass 0 at [0, N), pass 1 at [N, N + N/16)
  const double2* dc;
  typename CplxOf<T>::type* y;                  // scratch [n_frames][N]
  int log2n;
};

template <typename T>
__global__ void __launch_bounds__(256, 2) big_head_kernel(const BigArgs<T> a) {
  using CT = typename CplxOf<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];   // 4096 complex T
  CT* ex = reinterpret_cast<CT*>(smem_raw);
  const int t = threadIdx.x;
  const int cl = t & 15, hi = t >> 4;
  const int64_t n = (int64_t)1 << a.log2n;
  const int64_t m = n >> 8;                     // inner transform size
  const int64_t s0 = n >> 4;                    // pass-0 stride = 16*M
  const int64_t tiles_per_frame = m >> 4;
  const int64_t total = a.n_frames * tiles_per_frame;
  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int64_t f = w / tiles_per_frame;
    const int64_t c1 = (w - f * tiles_per_frame) * 16 + cl;     // column within [0, M)
    T re[16], im[16];
    // ---- pass 0: column c = hi*M + c1 of the N-point frame ----
    const int64_t c = (int64_t)hi * m + c1;
    {
      const float2* src = a.iq + f * a.frame_stride + c;
      float2 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = ldg_stream(src + j * s0);
      T dcr = T(0), dci = T(0);
      if (a.dc != nullptr) { double2 d = a.dc[f]; dcr = (T)d.x; dci = (T)d.y; }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const T wv = a.window[c + j * s0];
        re[j] = ((T)v[j].x - dcr) * wv;
        im[j] = ((T)v[j].y - dci) * wv;
      }
    }
    dft16<T>(re, im);
#pragma unroll
    for (int q = 1; q < 16; ++q) { const CT wq = a.tw[q * s0 + c]; cmul<T>(re[q], im[q], wq.x, wq.y); }
#pragma unroll
    for (int q = 0; q < 16; ++q) ex[(q * 16 + hi) * 16 + cl] = mk<T>(re[q], im[q]);
    __syncthreads();
    // ---- pass 1: sub-transform s = hi (pass-0 output digit), column c1, stride M ----
#pragma unroll
    for (int j = 0; j < 16; ++j) { const CT x = ex[(hi * 16 + j) * 16 + cl]; re[j] = x.x; im[j] = x.y; }
    __syncthreads();
    dft16<T>(re, im);
    const CT* tw1 = a.tw + n;
#pragma unroll
    for (int q = 1; q < 16; ++q) { const CT wq = tw1[q * m + c1]; cmul<T>(re[q], im[q], wq.x, wq.y); }
    CT* dst = a.y + f * n + (int64_t)hi * s0 + c1;
#pragma unroll
    for (int q = 0; q < 16; ++q) dst[q * m] = mk<T>(re[q], im[q]);
  }
}

// One-pass head for N = 16*M (M <= 4096, e.g. 65536 = 16 x 4096): every thread owns one column c of the
// N-point frame, reads x[c + j*N/16] (coalesced across threads), applies the window, runs one radix-16 and the
// post-twiddle W_N^(c*q), and writes sub-transform q to Y[q*N/16 + c] (coalesced). No shared memory at all;
// the M-point sub-transforms are finished by fft_fused_kernel<TAIL = 3>, which for M = 4096 is the tuned kernel.
template <typename T>
__global__ void __launch_bounds__(256) big_head1_kernel(const BigArgs<T> a) {
  using CT = typename CplxOf<T>::type;
  const int64_t n = (int64_t)1 << a.log2n;
  const int64_t s0 = n >> 4;                               // columns per frame = M
  const int64_t total = a.n_frames * s0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = w / s0, c = w - f * s0;
    T re[16], im[16];
    {
      const float2* src = a.iq + f * a.frame_stride + c;
      float2 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = ldg_stream(src + j * s0);
      T dcr = T(0), dci = T(0);
      if (a.dc != nullptr) { double2 d = a.dc[f]; dcr = (T)d.x; dci = (T)d.y; }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const T wv = a.window[c + j * s0];
        re[j] = ((T)v[j].x - dcr) * wv;
        im[j] = ((T)v[j].y - dci) * wv;
      }
    }
    dft16<T>(re, im);
#pragma unroll
    for (int q = 1; q < 16; ++q) { const CT wq = a.tw[q * s0 + c]; cmul<T>(re[q], im[q], wq.x, wq.y); }
    CT* dst = a.y + f * n + c;
#pragma unroll
    for (int q = 0; q < 16; ++q) dst[q * s0] = mk<T>(re[q], im[q]);
  }
}

}  // namespace tdsa
