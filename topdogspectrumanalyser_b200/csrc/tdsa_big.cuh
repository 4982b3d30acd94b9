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
  extern __shared__ __align__(16) unsigned char big_smem[];   // 4096 complex T
  CT* ex = reinterpret_cast<CT*>(big_smem);
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

// 8-byte asynchronous copy global -> shared (LDGSTS): the prefetch of the next segment needs no registers
__device__ __forceinline__ void cp_async_8(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// shared memory of big_head_wl_kernel: two sample buffers [16 j][16 groups][18 float2] (sixteen columns per group, pitch 18:
// 144-byte groups keep bulk-copy destinations 16-byte aligned), the CTA's window slice [16 j][256] in T, two mbarriers
constexpr int kHeadPitch = 18, kHeadRow = 16 * kHeadPitch, kHeadBuf = 16 * kHeadRow * 8;
template <typename T> constexpr int head_wl_smem() { return 2 * kHeadBuf + 16 * 256 * (int)sizeof(T) + 16; }

// Head of the 65536-point Welch path whose tails run through fft_wl_kernel (kAccSub): like big_head1_kernel, one radix-16
// over the samples c + 4096 j of a windowed segment and the post-twiddle W_N^(c q), but
//   * a CTA keeps one block of 256 columns (b = blockIdx.x % 16) for all the segments it processes: its window slice
//     lives in shared memory and its base twiddles W_N^(c q) in registers (the other twiddles are products), so per
//     segment only the samples are read;
//   * sub-transform q of segment f is written in the order fft_wl_kernel's threads consume it: y[((f*16 + q)*16 + j)*256
//     + tid] is sample n0 = r + 16 c' + 256 j of the 4096-point input, (r, c') = wl_thread_identity(tid).  The 256
//     values of one (q, j) are exactly one CTA's columns (j = b), so head thread tid computes the column the tail's
//     thread tid will want and stores are fully coalesced (dst[tid]);
//   * the samples of the NEXT segment travel global -> shared in natural column order while this segment is
//     transformed; the permutation to the tail's order happens on the shared-memory read (groups of sixteen columns at
//     pitch 18).  BULK: one 128-byte cp.async.bulk per thread and segment (sixteen columns of one row j), completion on an
//     mbarrier (needs 16-byte aligned segment starts: even hop); otherwise sixteen 8-byte cp.async per thread.
// Measured (round 2, 2047 segments): storing in permuted order cost 1.07 GB of DRAM fill reads in float32 (half
// sectors); permuted global loads of samples and window doubled the L2 read traffic (7.4 GB, the L2 limit) at 704 us
// (float64); see profiles/r02_experiments.md.
template <typename T, bool BULK>
__global__ void __launch_bounds__(256, 2) big_head_wl_kernel(const BigArgs<T> a) {
  using CT = typename CplxOf<T>::type;
  constexpr int64_t s0 = 4096;                              // columns per segment
  const int tid = threadIdx.x;
  const int b = (int)(blockIdx.x & 15);
  // the tail's thread identity (wl_thread_identity): warp w, lane l -> r = 2 w + ((l >> 3) & 1), c' = (l & 7) + 8 (l >> 4)
  const int wl_w = tid >> 5, wl_l = tid & 31;
  const int r = 2 * wl_w + ((wl_l >> 3) & 1), cp = (wl_l & 7) + 8 * (wl_l >> 4);
  const int c = 256 * b + r + 16 * cp;                      // this thread's column of the segment
  extern __shared__ __align__(16) unsigned char head_smem[];
  float2* sbuf = reinterpret_cast<float2*>(head_smem);       // [2][16][16][18]
  T* swin = reinterpret_cast<T*>(head_smem + 2 * kHeadBuf);   // [16][256], thread order
  const uint32_t bar_u32 = smem_u32(head_smem) + (uint32_t)(2 * kHeadBuf + 16 * 256 * (int)sizeof(T));   // full[2]
  if constexpr (BULK) {
    if (tid == 0) { mbar_init(bar_u32, 1); mbar_init(bar_u32 + 8, 1); fence_mbar_init(); }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) swin[j * 256 + tid] = a.window[c + j * s0];
  __syncthreads();
  // base twiddles kept in registers: float64 w^1 and w^4, float32 w^1..w^3, w^4, w^8, w^12
  constexpr bool kTwoBases = sizeof(T) == 8;
  T bwr[16], bwi[16];
#pragma unroll
  for (int q = 1; q < 16; ++q) {
    if (kTwoBases ? (q == 1 || q == 4) : (q < 4 || (q & 3) == 0)) { const CT w = a.tw[q * s0 + c]; bwr[q] = w.x; bwi[q] = w.y; }
  }
  const int64_t f_step = gridDim.x >> 4;
  int64_t f = blockIdx.x >> 4;
  // copy side: natural column t of row j goes to slot (t & 15) + 18 (t >> 4) of row j; read side: slot r + 18 c'
  const float2* get = sbuf + r + kHeadPitch * cp;
  auto prefetch = [&](int64_t seg, int buf) {
    if constexpr (BULK) {                                    // thread (j = tid >> 4, g = tid & 15): columns 16 g .. 16 g + 15 of row j
      const uint32_t bar = bar_u32 + 8 * buf;
      if (tid == 0) mbar_arrive_expect_tx(bar, 16 * 256 * 8);
      const float2* src = a.iq + seg * a.frame_stride + s0 * (tid >> 4) + 256 * b + 16 * (tid & 15);
      bulk_g2s(smem_u32(head_smem) + (uint32_t)(buf * kHeadBuf + (tid >> 4) * kHeadRow * 8 + (tid & 15) * kHeadPitch * 8), src, 128, bar);
    } else {
      const uint32_t put = smem_u32(head_smem) + (uint32_t)(buf * kHeadBuf + ((tid & 15) + kHeadPitch * (tid >> 4)) * 8);
      const float2* src = a.iq + seg * a.frame_stride + 256 * b + tid;
#pragma unroll
      for (int j = 0; j < 16; ++j) cp_async_8(put + (uint32_t)(j * kHeadRow * 8), src + j * s0);
    }
  };
  if (f < a.n_frames) prefetch(f, 0);
  if constexpr (!BULK) cp_async_commit();
  for (int it = 0; f < a.n_frames; f += f_step, ++it) {
    if constexpr (BULK) mbar_wait(bar_u32 + 8 * (it & 1), (uint32_t)((it >> 1) & 1));   // this segment's samples have landed
    else cp_async_wait<0>();                                 // my copies of this segment have landed ...
    __syncthreads();                                         // ... everybody's have, and the other buffer is no longer read
    if constexpr (BULK) fence_proxy_async();                 // generic-proxy reads of that buffer before the async-proxy writes
    if (f + f_step < a.n_frames) prefetch(f + f_step, (it + 1) & 1);
    if constexpr (!BULK) cp_async_commit();
    T re[16], im[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float2 v = get[((it & 1) * 16 + j) * kHeadRow];
      const T wv = swin[j * 256 + tid];
      re[j] = (T)v.x * wv;
      im[j] = (T)v.y * wv;
    }
    dft16<T>(re, im);
    T wr[16], wi[16];
#pragma unroll
    for (int q = 1; q < 16; ++q) {
      if (kTwoBases ? (q == 1 || q == 4) : (q < 4 || (q & 3) == 0)) { wr[q] = bwr[q]; wi[q] = bwi[q]; }
    }
    if constexpr (kTwoBases) {
      wr[2] = wr[1]; wi[2] = wi[1]; cmul<T>(wr[2], wi[2], wr[1], wi[1]);
      wr[3] = wr[2]; wi[3] = wi[2]; cmul<T>(wr[3], wi[3], wr[1], wi[1]);
      wr[8] = wr[4]; wi[8] = wi[4]; cmul<T>(wr[8], wi[8], wr[4], wi[4]);
      wr[12] = wr[8]; wi[12] = wi[8]; cmul<T>(wr[12], wi[12], wr[4], wi[4]);
    }
    CT* dst = a.y + ((f * 16) * 16 + b) * 256 + tid;
    dst[0] = mk<T>(re[0], im[0]);
#pragma unroll
    for (int q = 1; q < 16; ++q) {
      T xr = wr[q & 3], xi = wi[q & 3];
      if ((q & 3) == 0) { xr = wr[q]; xi = wi[q]; }
      else if (q >= 4) cmul<T>(xr, xi, wr[q & ~3], wi[q & ~3]);
      cmul<T>(re[q], im[q], xr, xi);
      dst[(int64_t)q * 4096] = mk<T>(re[q], im[q]);
    }
  }
}

}  // namespace tdsa
