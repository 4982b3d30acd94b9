// tdsa_fft.cuh — fused window + FFT + |.|^2 + dB + fftshift kernels for sm_100a.
//
// Replaces, per frame, datasources/rtl_samples.py:169-184 of the reference
// (x*w -> scipy.fft.fft -> fftshift -> abs**2 -> 10*log10(. + floor)) and the
// hackrf variant datasources/hackrf_samples.py:365-383, over a batch of frames.
//
// Decomposition (mirrored in tools/fft_plan_model.py, which checks it against numpy):
//   * 16 points per thread, T = N/16 threads per frame, in-place DIF.
//   * passes are radix 16 while >= 4 bits remain, then one final radix 2/4/8 pass.
//   * pass 0 reads x[t + j*T] straight from global memory (coalesced), multiplies by
//     the window (with (-1)^n folded in, which makes the output fftshift-ed for free)
//     and runs its butterflies in registers; later passes exchange through a padded
//     shared-memory buffer laid out so that every access pattern is bank-conflict free.
//   * the final pass gives thread t the sub-transform whose outputs are bins
//     k = t + T*u + (N/R)*q, so dB stores are coalesced; no bit-reversal pass.
//   * twiddles are exact-rounded tables built on the host in float64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace tdsa {

// radix-8 padding table (filled from tools/fft_plan_model.py r8 banks); generic fallback otherwise
// frames prefetched into L2 ahead of the shared-memory staging ring (cp.async.bulk.prefetch.L2)
#ifndef TDSA_L2_AHEAD
#define TDSA_L2_AHEAD 0   // measured: 0, 2 and 4 frames ahead are within noise of each other at N=4096
#endif
// programmatic dependent launch (griddepcontrol) on the fused kernel
#ifndef TDSA_PDL
#define TDSA_PDL 0   // measured: no gain at 80-150 us per launch (81.9 vs 82.0 us); kept as an option
#endif
#ifndef TDSA_STREAM_STORES
#define TDSA_STREAM_STORES 0
#endif
#ifndef TDSA_PADS_R8
#define TDSA_PADS_R8(LOG2N, WIDE) pads_r8(LOG2N, WIDE)
#endif

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
template <typename T> struct CplxOf;
template <> struct CplxOf<float> { using type = float2; };
template <> struct CplxOf<double> { using type = double2; };

template <typename T> __device__ __forceinline__ typename CplxOf<T>::type mk(T a, T b);
template <> __device__ __forceinline__ float2 mk<float>(float a, float b) { return make_float2(a, b); }
template <> __device__ __forceinline__ double2 mk<double>(double a, double b) { return make_double2(a, b); }

__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float2 ldg_stream(const float2* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

// ---- bulk-copy (TMA) staging: cp.async.bulk global -> shared, completion on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}
// L2 prefetch of a whole frame: deepens the prefetch distance beyond what fits in shared memory
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
// programmatic dependent launch: let the next launch on the stream start its prologue early, and make this
// grid wait for the previous one (completion + memory flush) before it touches any frame data
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// float -> T.  For T = double the hardware conversion (F2F.F64.F32) runs on the FP64 pipe at about four pipe cycles
// per warp instruction (tools/ubench.cu, test D): 32 of them per thread and frame are ~9 % of the pipe's work.
// widen_fast does the same job on the integer pipe: exact for every normal float; zero and denormals come out as
// +-2^-127 * (1 + m) (absolute error < 1.2e-38, far below the log floor); Inf/NaN do NOT survive, so the caller
// tracks them separately (see frame_special) and redoes the frame's conversion with the hardware path if any appear.
#ifndef TDSA_INT_WIDEN
#define TDSA_INT_WIDEN 0   // measured: 141.4 us with the integer path vs 135.7 us with F2F (extra instructions and spills cost more than the pipe time saved)
#endif
template <typename T> __device__ __forceinline__ T widen_fast(float x);
template <> __device__ __forceinline__ float widen_fast<float>(float x) { return x; }
template <> __device__ __forceinline__ double widen_fast<double>(float x) {
#if TDSA_INT_WIDEN
  const uint32_t u = __float_as_uint(x);
  const uint32_t hi = ((u >> 3) & 0x0fffffffu) + 0x38000000u + (u & 0x80000000u);
  return __hiloint2double((int)hi, (int)(u << 29));
#else
  return (double)x;
#endif
}

// Lean integer widening (TDSA_INT_WIDEN == 2): arithmetic shift keeps the sign in bit 31, the mask clears the three
// sign copies below it, the add re-biases the exponent (127 -> 1023): SHF.R.S32 + LOP3 + IADD3 + SHL per value.
// Inf/NaN inputs are caught separately by the caller (nonfinite_probe on the idle FP32 pipe).
__device__ __forceinline__ double widen_lean(float x) {
  const int u = __float_as_int(x);
  const uint32_t hi = ((uint32_t)(u >> 3) & 0x8fffffffu) + 0x38000000u;
  return __hiloint2double((int)hi, (int)((uint32_t)u << 29));
}
// acc stays 0 while every x seen is finite; 0 * Inf = 0 * NaN = NaN poisons it otherwise (one FFMA on the FP32 pipe)
__device__ __forceinline__ void nonfinite_probe(float x, float& acc) { acc = __fmaf_rn(x, 0.0f, acc); }

// float64 power -> float32 on the integer pipe (TDSA_INT_NARROW): clamp the high word to [2^-126, NaN pattern]
// (unsigned, so a NaN with the sign bit set stays NaN), re-bias the exponent, funnel-shift the top 23 mantissa bits
// in. Truncation instead of rounding: relative error < 2^-23, i.e. < 5.2e-7 dB after the log.
// Values >= 2^128 (|X| > 1.8e19) and NaN both come out as NaN.
#ifndef TDSA_INT_NARROW
#define TDSA_INT_NARROW 0
#endif
__device__ __forceinline__ float narrow_trunc(double p) {
  uint32_t hi = (uint32_t)__double2hiint(p);
  const uint32_t lo = (uint32_t)__double2loint(p);
  hi = min(max(hi, 0x38100000u), 0x47f80000u) - 0x38000000u;
  return __uint_as_float(__funnelshift_l(lo, hi, 3));
}
template <typename T> __device__ __forceinline__ float narrow_power(T p);
template <> __device__ __forceinline__ float narrow_power<float>(float p) { return p; }
template <> __device__ __forceinline__ float narrow_power<double>(double p) {
#if TDSA_INT_NARROW
  return narrow_trunc(p);
#else
  return (float)p;
#endif
}

// widen_fast that also keeps the running maximum of the magnitude bits (>= 0x7f800000 <=> an Inf or NaN was seen)
__device__ __forceinline__ double widen_track(float x, uint32_t& top) {
  const uint32_t u = __float_as_uint(x);
  const uint32_t a = u & 0x7fffffffu;
  top = max(top, a);
  const uint32_t hi = (a >> 3) + 0x38000000u + (u & 0x80000000u);
  return __hiloint2double((int)hi, (int)(u << 29));
}

template <typename T> __device__ __forceinline__ void cmul(T& xr, T& xi, T wr, T wi) {
  T r = xr * wr - xi * wi;
  T i = xr * wi + xi * wr;
  xr = r;
  xi = i;
}

// ---------------------------------------------------------------------------------------
// butterflies (forward DFT, W = exp(-2*pi*i/R)); all indices are compile-time after unrolling,
// so the re[]/im[] arrays live in registers
// ---------------------------------------------------------------------------------------
template <typename T, int I0, int I1, int I2, int I3>
__device__ __forceinline__ void r4(T* re, T* im) {
  T t0r = re[I0] + re[I2], t0i = im[I0] + im[I2];
  T t1r = re[I0] - re[I2], t1i = im[I0] - im[I2];
  T t2r = re[I1] + re[I3], t2i = im[I1] + im[I3];
  T t3r = re[I1] - re[I3], t3i = im[I1] - im[I3];
  re[I0] = t0r + t2r; im[I0] = t0i + t2i;
  re[I2] = t0r - t2r; im[I2] = t0i - t2i;
  re[I1] = t1r + t3i; im[I1] = t1i - t3r;   // t1 - i*t3
  re[I3] = t1r - t3i; im[I3] = t1i + t3r;   // t1 + i*t3
}

template <typename T, int A, int B> __device__ __forceinline__ void swp(T* re, T* im) {
  T r = re[A]; re[A] = re[B]; re[B] = r;
  T i = im[A]; im[A] = im[B]; im[B] = i;
}

// x *= (1 - i)/sqrt(2)   (W8^1 = W16^2)
template <typename T> __device__ __forceinline__ void mul_w8_1(T& xr, T& xi) {
  const T h = T(0.70710678118654752440);
  T r = (xr + xi) * h, i = (xi - xr) * h;
  xr = r; xi = i;
}
// x *= (-1 - i)/sqrt(2)  (W8^3 = W16^6)
template <typename T> __device__ __forceinline__ void mul_w8_3(T& xr, T& xi) {
  const T h = T(0.70710678118654752440);
  T r = (xi - xr) * h, i = -(xr + xi) * h;
  xr = r; xi = i;
}
// x *= -i
template <typename T> __device__ __forceinline__ void mul_mi(T& xr, T& xi) {
  T r = xi, i = -xr;
  xr = r; xi = i;
}

template <typename T> __device__ __forceinline__ T fm(T a, T b, T c);
template <> __device__ __forceinline__ float fm<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <> __device__ __forceinline__ double fm<double>(double a, double b, double c) { return __fma_rn(a, b, c); }

// Radix-4 butterfly whose inputs carry twiddles: a[Ik] stands for w[Ik]*a[Ik] (w[I0] == 1 when W0_ONE).
// The twiddle multiplies are folded into the first radix-2 level:
//   z0 = w0*a0;  t0 = z0 + w2*a2 (4 FMA);  t1 = 2*z0 - t0 (2 FMA)   — 10 ops instead of 12 per pair.
template <typename T, int I0, int I1, int I2, int I3, bool W0_ONE>
__device__ __forceinline__ void r4_tw(T* re, T* im, const T* wr, const T* wi) {
  T z0r, z0i;
  if constexpr (W0_ONE) { z0r = re[I0]; z0i = im[I0]; }
  else { z0r = fm<T>(re[I0], wr[I0], -(im[I0] * wi[I0])); z0i = fm<T>(re[I0], wi[I0], im[I0] * wr[I0]); }
  const T t0r = fm<T>(re[I2], wr[I2], fm<T>(-im[I2], wi[I2], z0r));
  const T t0i = fm<T>(re[I2], wi[I2], fm<T>(im[I2], wr[I2], z0i));
  const T t1r = fm<T>(T(2), z0r, -t0r), t1i = fm<T>(T(2), z0i, -t0i);
  const T z1r = fm<T>(re[I1], wr[I1], -(im[I1] * wi[I1])), z1i = fm<T>(re[I1], wi[I1], im[I1] * wr[I1]);
  const T t2r = fm<T>(re[I3], wr[I3], fm<T>(-im[I3], wi[I3], z1r));
  const T t2i = fm<T>(re[I3], wi[I3], fm<T>(im[I3], wr[I3], z1i));
  const T t3r = fm<T>(T(2), z1r, -t2r), t3i = fm<T>(T(2), z1i, -t2i);
  re[I0] = t0r + t2r; im[I0] = t0i + t2i;
  re[I2] = t0r - t2r; im[I2] = t0i - t2i;
  re[I1] = t1r + t3i; im[I1] = t1i - t3r;   // t1 - i*t3
  re[I3] = t1r - t3i; im[I3] = t1i + t3r;   // t1 + i*t3
}

// Radix-4 butterfly on windowed samples a[Ik] = x[Ik]*w[Ik] with REAL weights (pass 0): the window multiply is
// folded into the first radix-2 level (p0 = x0*w0; t0 = p0 + x2*w2; t1 = p0 - x2*w2: 3 operations per component
// instead of 4). xr/xi hold the raw (converted) samples, w the window values.
template <typename T, int I0, int I1, int I2, int I3>
__device__ __forceinline__ void r4_win(T* re, T* im, const T* w) {
  const T p0r = re[I0] * w[I0], p0i = im[I0] * w[I0];
  const T t0r = fm<T>(re[I2], w[I2], p0r), t0i = fm<T>(im[I2], w[I2], p0i);
  const T t1r = fm<T>(-re[I2], w[I2], p0r), t1i = fm<T>(-im[I2], w[I2], p0i);
  const T p1r = re[I1] * w[I1], p1i = im[I1] * w[I1];
  const T t2r = fm<T>(re[I3], w[I3], p1r), t2i = fm<T>(im[I3], w[I3], p1i);
  const T t3r = fm<T>(-re[I3], w[I3], p1r), t3i = fm<T>(-im[I3], w[I3], p1i);
  re[I0] = t0r + t2r; im[I0] = t0i + t2i;
  re[I2] = t0r - t2r; im[I2] = t0i - t2i;
  re[I1] = t1r + t3i; im[I1] = t1i - t3r;
  re[I3] = t1r - t3i; im[I3] = t1i + t3r;
}

// Second half of the 16-point DFT: four radix-4 butterflies over j0 whose inputs B[j0][q0] = a[4*q0 + j0]
// carry the constant twiddles W16^(j0*q0), folded in the same way; then the 4x4 transpose to natural order.
template <typename T> __device__ __forceinline__ void dft16_stage_b(T* re, T* im) {
  constexpr T c1 = T(0.92387953251128675613);   // cos(pi/8)
  constexpr T s1 = T(0.38268343236508977173);   // sin(pi/8)
  constexpr T h = T(0.70710678118654752440);    // sqrt(1/2)
  r4<T, 0, 1, 2, 3>(re, im);                                      // q0 = 0: no twiddles
  {                                                               // q0 = 1: 1, W16^1, W16^2, W16^3
    const T wr[4] = {T(1), c1, h, s1}, wi[4] = {T(0), -s1, -h, -c1};
    r4_tw<T, 0, 1, 2, 3, true>(re + 4, im + 4, wr, wi);
  }
  {                                                               // q0 = 2: 1, W16^2, -i, W16^6
    const T t0r = re[8] + im[10], t0i = im[8] - re[10];           // a8 + (-i)*a10
    const T t1r = re[8] - im[10], t1i = im[8] + re[10];
    const T u9r = re[9] + im[9], u9i = im[9] - re[9];             // (1 - i)*a9   (times h below)
    const T u11r = im[11] - re[11], u11i = -(re[11] + im[11]);    // (-1 - i)*a11 (times h below)
    const T sr = u9r + u11r, si = u9i + u11i, dr = u9r - u11r, di = u9i - u11i;
    re[8] = fm<T>(h, sr, t0r);   im[8] = fm<T>(h, si, t0i);       // t0 + t2
    re[10] = fm<T>(-h, sr, t0r); im[10] = fm<T>(-h, si, t0i);     // t0 - t2
    re[9] = fm<T>(h, di, t1r);   im[9] = fm<T>(-h, dr, t1i);      // t1 - i*t3
    re[11] = fm<T>(-h, di, t1r); im[11] = fm<T>(h, dr, t1i);      // t1 + i*t3
  }
  {                                                               // q0 = 3: 1, W16^3, W16^6, W16^9
    const T wr[4] = {T(1), s1, -h, -c1}, wi[4] = {T(0), -c1, -h, s1};
    r4_tw<T, 0, 1, 2, 3, true>(re + 12, im + 12, wr, wi);
  }
  swp<T, 1, 4>(re, im);  swp<T, 2, 8>(re, im);  swp<T, 3, 12>(re, im);
  swp<T, 6, 9>(re, im);  swp<T, 7, 13>(re, im); swp<T, 11, 14>(re, im);
}

// 16-point DFT over registers [0, 16), natural order in and out: a[q] = sum_j a[j] W16^(jq).
template <typename T> __device__ __forceinline__ void dft16(T* re, T* im) {
  // stage A: over j1 (stride 4); a[j0 + 4*q0] = B[j0][q0]
  r4<T, 0, 4, 8, 12>(re, im);
  r4<T, 1, 5, 9, 13>(re, im);
  r4<T, 2, 6, 10, 14>(re, im);
  r4<T, 3, 7, 11, 15>(re, im);
  dft16_stage_b<T>(re, im);
}

// 16-point DFT of x[j]*w[j] with real window weights w, the multiply folded into stage A (pass 0).
template <typename T> __device__ __forceinline__ void dft16_win(T* re, T* im, const T* w) {
  r4_win<T, 0, 4, 8, 12>(re, im, w);
  r4_win<T, 1, 5, 9, 13>(re, im, w);
  r4_win<T, 2, 6, 10, 14>(re, im, w);
  r4_win<T, 3, 7, 11, 15>(re, im, w);
  dft16_stage_b<T>(re, im);
}

// Same transform of the twiddled inputs w[j]*a[j] (w[0] == 1), the twiddles folded into stage A.
template <typename T> __device__ __forceinline__ void dft16_pretw(T* re, T* im, const T* wr, const T* wi) {
  r4_tw<T, 0, 4, 8, 12, true>(re, im, wr, wi);
  r4_tw<T, 1, 5, 9, 13, false>(re, im, wr, wi);
  r4_tw<T, 2, 6, 10, 14, false>(re, im, wr, wi);
  r4_tw<T, 3, 7, 11, 15, false>(re, im, wr, wi);
  dft16_stage_b<T>(re, im);
}

// R-point DFTs (R = 2, 4, 8) over registers [O, O+R), natural order in and out.
template <typename T, int R, int O> struct DftSmall;
template <typename T, int O> struct DftSmall<T, 2, O> {
  static __device__ __forceinline__ void run(T* re, T* im) {
    T ar = re[O] + re[O + 1], ai = im[O] + im[O + 1];
    T br = re[O] - re[O + 1], bi = im[O] - im[O + 1];
    re[O] = ar; im[O] = ai; re[O + 1] = br; im[O + 1] = bi;
  }
};
template <typename T, int O> struct DftSmall<T, 4, O> {
  static __device__ __forceinline__ void run(T* re, T* im) { r4<T, O, O + 1, O + 2, O + 3>(re, im); }
};
template <typename T, int O> struct DftSmall<T, 8, O> {
  static __device__ __forceinline__ void run(T* re, T* im) {
    // DIF radix-2 split: u_j = a_j + a_{j+4} (even outputs), v_j = (a_j - a_{j+4}) W8^j (odd outputs)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      T ur = re[O + j] + re[O + j + 4], ui = im[O + j] + im[O + j + 4];
      T vr = re[O + j] - re[O + j + 4], vi = im[O + j] - im[O + j + 4];
      re[O + j] = ur; im[O + j] = ui; re[O + j + 4] = vr; im[O + j + 4] = vi;
    }
    mul_w8_1<T>(re[O + 5], im[O + 5]);
    mul_mi<T>(re[O + 6], im[O + 6]);
    mul_w8_3<T>(re[O + 7], im[O + 7]);
    r4<T, O, O + 1, O + 2, O + 3>(re, im);          // X[2m]   at O+m
    r4<T, O + 4, O + 5, O + 6, O + 7>(re, im);      // X[2m+1] at O+4+m
    // interleave: out[2m] = a[m], out[2m+1] = a[4+m]
    T r1 = re[O + 1], r2 = re[O + 2], r3 = re[O + 3], r4_ = re[O + 4], r5 = re[O + 5], r6 = re[O + 6];
    T i1 = im[O + 1], i2 = im[O + 2], i3 = im[O + 3], i4_ = im[O + 4], i5 = im[O + 5], i6 = im[O + 6];
    re[O + 1] = r4_; im[O + 1] = i4_;
    re[O + 2] = r1;  im[O + 2] = i1;
    re[O + 3] = r5;  im[O + 3] = i5;
    re[O + 4] = r2;  im[O + 4] = i2;
    re[O + 5] = r6;  im[O + 5] = i6;
    re[O + 6] = r3;  im[O + 6] = i3;
  }
};

// P-point DFT over all P registers of a thread (P = 16 or 8)
template <typename T, int P> __device__ __forceinline__ void dft_full(T* re, T* im) {
  if constexpr (P == 16) dft16<T>(re, im); else DftSmall<T, 8, 0>::run(re, im);
}

// P-point DFT of pre-twiddled inputs (w[0] == 1): fused for P == 16, plain multiplies otherwise
template <typename T, int P> __device__ __forceinline__ void dft_full_pretw(T* re, T* im, const T* wr, const T* wi) {
  if constexpr (P == 16) {
    dft16_pretw<T>(re, im, wr, wi);
  } else {
#pragma unroll
    for (int j = 1; j < P; ++j) cmul<T>(re[j], im[j], wr[j], wi[j]);
    dft_full<T, P>(re, im);
  }
}

// last pass: P/R independent R-point DFTs over consecutive register groups
template <typename T, int P, int R> __device__ __forceinline__ void dft_last(T* re, T* im) {
  if constexpr (R == P) {
    dft_full<T, P>(re, im);
  } else if constexpr (R == 8) {       // P == 16
    DftSmall<T, 8, 0>::run(re, im); DftSmall<T, 8, 8>::run(re, im);
  } else if constexpr (R == 4) {
    DftSmall<T, 4, 0>::run(re, im); DftSmall<T, 4, 4>::run(re, im);
    if constexpr (P == 16) { DftSmall<T, 4, 8>::run(re, im); DftSmall<T, 4, 12>::run(re, im); }
  } else {
    DftSmall<T, 2, 0>::run(re, im);  DftSmall<T, 2, 2>::run(re, im);
    DftSmall<T, 2, 4>::run(re, im);  DftSmall<T, 2, 6>::run(re, im);
    if constexpr (P == 16) {
      DftSmall<T, 2, 8>::run(re, im);  DftSmall<T, 2, 10>::run(re, im);
      DftSmall<T, 2, 12>::run(re, im); DftSmall<T, 2, 14>::run(re, im);
    }
  }
}

// ---------------------------------------------------------------------------------------
// compile-time plan for one (element type, size, digit width)
//   LOGR = 4: radix 16, 16 points per thread (N/16 threads per frame)
//   LOGR = 3: radix 8,   8 points per thread (N/8 threads per frame): half the registers per thread,
//             twice the warps per frame, one more exchange pass
// ---------------------------------------------------------------------------------------
struct PadSpec { int s0, c0, s1, c1, s2, c2; };   // phys(p) = p + c0*(p>>s0) + c1*(p>>s1) + c2*(p>>s2)

constexpr PadSpec pads_r8(int log2n, bool wide) {
  if (log2n == 9) return {6, 1, 0, 0, 0, 0};
  if (log2n == 10) return wide ? PadSpec{5, 2, 7, 1, 0, 0} : PadSpec{4, 1, 8, 2, 0, 0};
  if (log2n == 11) return wide ? PadSpec{8, 1, 0, 0, 0, 0} : PadSpec{5, 4, 9, 1, 0, 0};
  if (log2n == 12) return wide ? PadSpec{9, 1, 0, 0, 0, 0} : PadSpec{6, 8, 9, 1, 0, 0};   // conflict-free
  return {3, 1, log2n - 3, 1, 0, 0};
}

// Padded exchange layouts found by tools/fft_plan_model.py ("banks" / "r8 banks"): zero shared-memory
// bank conflicts for every pass at the listed sizes; other sizes get a generic (correct) padding.
template <int LOG2N, bool WIDE, int LOGR> constexpr PadSpec pad_spec() {
  if (LOGR == 4) {
    if (LOG2N == 12) return {8, 1, 0, 0, 0, 0};
    if (LOG2N == 11 && WIDE) return {7, 1, 0, 0, 0, 0};
    if (LOG2N <= 11) return {4, 1, WIDE ? 7 : 8, 1, 0, 0};
    return {4, 1, LOG2N - 4, 1, 0, 0};
  }
  return TDSA_PADS_R8(LOG2N, WIDE);
}

template <typename T, int LOG2N, int LOGR = 4> struct Plan {
  static constexpr int N = 1 << LOG2N;
  static constexpr int P = 1 << LOGR;                  // points per thread = radix of the full passes
  static constexpr int THREADS = N >> LOGR;
  static constexpr int NPASS = (LOG2N + LOGR - 1) / LOGR;
  static constexpr int R_LAST = (LOG2N % LOGR) ? (1 << (LOG2N % LOGR)) : P;
  static constexpr int NB_LAST = P / R_LAST;           // butterflies per thread in the last pass
  static constexpr int REV_DIGITS = NPASS - 1;         // digits reversed in the last pass
  static constexpr bool WIDE = sizeof(T) == 8;         // 16-byte exchange elements
  static constexpr PadSpec PAD = pad_spec<LOG2N, WIDE, LOGR>();
  static __host__ __device__ constexpr int phys(int p) {
    return p + PAD.c0 * (p >> PAD.s0) + (PAD.c1 ? PAD.c1 * (p >> PAD.s1) : 0) + (PAD.c2 ? PAD.c2 * (p >> PAD.s2) : 0);
  }
  static constexpr int PHYS_SIZE = phys(N - 1) + 1;
  static __host__ __device__ constexpr int len(int i) { return N >> (LOGR * i); }             // L_i
  static __host__ __device__ constexpr int stride(int i) { return N >> (LOGR * (i + 1)); }    // S_i
  // Twiddles sit on the INPUT side of passes 1..NPASS-1 (decimation-in-time placement): input j of a
  // butterfly whose already-known output digits are K = k0 + P*k1 + ... is multiplied by W_N^(j*S_i*K).
  // Table of pass i: [j < R_i][K < P^i]; tables are concatenated in pass order starting at pass 1.
  static __host__ __device__ constexpr int radix(int i) { return i == NPASS - 1 ? R_LAST : P; }
  static __host__ __device__ constexpr int kcount(int i) { return 1 << (LOGR * i); }            // P^i
  static __host__ __device__ constexpr int tw_offset(int i) {   // entries before pass i's table (i >= 1)
    int o = 0;
    for (int k = 1; k < i; ++k) o += radix(k) * kcount(k);
    return o;
  }
  // Plans of more than three passes keep the decimation-in-frequency placement instead (twiddle
  // W_{L_i}^(c*q) on OUTPUT q of pass i, table i = [q][c], c < S_i): there the DIT tables of the late
  // passes would be too large for shared memory, while the DIF tables of passes >= 1 are small.
  static constexpr bool DIT = NPASS <= 3;
  static __host__ __device__ constexpr int dif_offset(int i) {
    int o = 0;
    for (int k = 0; k < i; ++k) o += len(k);
    return o;
  }
  static constexpr int TW_TOTAL = DIT ? tw_offset(NPASS) : dif_offset(NPASS - 1);
  // shared-memory resident tables: DIT: pass 1's P*P table when it is a middle pass; DIF: passes >= 1
  static constexpr int TW_SMEM = DIT ? ((NPASS >= 3) ? P * P : 0) : (dif_offset(NPASS - 1) - N);
  static constexpr size_t SMEM_BYTES = (size_t)(PHYS_SIZE + TW_SMEM) * 2 * sizeof(T);
  // bulk-copy staging: NSTAGE buffers of one complex64 frame each + one mbarrier per stage
  static constexpr size_t STAGE_BYTES = (size_t)N * 8;
  static constexpr size_t STAGE_OFFSET = (SMEM_BYTES + 127) & ~(size_t)127;
  static __host__ __device__ constexpr size_t smem_staged(int nstage) {
    return STAGE_OFFSET + (size_t)nstage * STAGE_BYTES + 64;
  }
};

template <int LOGR> __host__ __device__ constexpr int digitrev(int v, int digits) {
  int o = 0;
  for (int d = 0; d < digits; ++d) { o = (o << LOGR) | (v & ((1 << LOGR) - 1)); v >>= LOGR; }
  return o;
}

// ---------------------------------------------------------------------------------------
// epilogues: what happens to |X[k]|^2 of frame f, bin k (k already fftshift-ed)
// ---------------------------------------------------------------------------------------
enum : int { kModePower = 0, kModePsd = 1, kModeMag20 = 2 };

struct EpiParams {
  float* db_out;       // float32 [F][N]   (dB epilogue)
  double* lin_out;     // float64 [F][N]   (linear epilogue)
  double scale;        // 1 for power/mag20, 1/(fs*N) for psd
  double floor;        // log floor
  int mode;
};

// dB from linear power; MAG20 is a compile-time flag so the per-element path has no branch.
template <typename T, bool MAG20> __device__ __forceinline__ float to_db_m(T p, const EpiParams& ep) {
  const float kDbPerLog2 = 3.01029995663981195f;   // 10*log10(2)
  if constexpr (MAG20) {
    const float m = sqrtf((float)p) + (float)ep.floor;
    return 2.0f * kDbPerLog2 * lg2_approx(m);
  } else {
    // narrow once, then scale and add the floor (1e-10 / 1e-12) with one float32 FMA: 6e-8 relative on the
    // argument of the log is 2.6e-7 dB, and it keeps 16 multiplies per thread off the FP64 pipe
    const float v = __fmaf_rn(narrow_power<T>(p), (float)ep.scale, (float)ep.floor);
    return kDbPerLog2 * lg2_approx(v);
  }
}
template <typename T> __device__ __forceinline__ float to_db(T p, const EpiParams& ep) {
  return ep.mode == kModeMag20 ? to_db_m<T, true>(p, ep) : to_db_m<T, false>(p, ep);
}

struct EpiDb {
  template <typename T, bool MAG20>
  static __device__ __forceinline__ void store(const EpiParams& ep, int64_t f, int n, int k, T p) {
#if TDSA_STREAM_STORES
    __stcs(ep.db_out + f * n + k, to_db_m<T, MAG20>(p, ep));     // written once, never re-read by this kernel
#else
    ep.db_out[f * n + k] = to_db_m<T, MAG20>(p, ep);
#endif
  }
};
struct EpiLinear {
  template <typename T, bool MAG20>
  static __device__ __forceinline__ void store(const EpiParams& ep, int64_t f, int n, int k, T p) {
    ep.lin_out[f * n + k] = (double)p * ep.scale;
  }
};

// ---------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------
template <typename T> struct FftArgs {
  const float2* iq;            // complex64 input
  int64_t n_frames;
  int64_t frame_stride;        // in complex samples
  const T* window;             // T[N], includes the (-1)^n fftshift factor
  const typename CplxOf<T>::type* tw;   // pre-twiddle tables of passes 1..NPASS-1, each [j][K] (see Plan)
  const double2* dc;           // optional per-frame DC estimate to subtract (hackrf path) or nullptr
  const typename CplxOf<T>::type* in_ct;   // TAIL kernels: complex T input [n_frames][N] (no window)
  EpiParams ep;
  int* sched = nullptr;        // dynamic frame scheduling: {next unclaimed frame, CTAs that have left}; nullptr = static
  int stagger = 0;             // diagnostic (TDSA_DEBUG_STAGGER): cycles the second half of the grid waits before its first frame
  long long* dbg = nullptr;    // diagnostic (-DTDSA_DEBUG_TIMING): per-warp phase time stamps
};

// Phase time stamps of the first 32 frames of every warp: dbg[((block*8 + warp)*32 + it)*16 + i] = clock64,
// and the SM id of each block behind them (tools/phase_timing.py reads the dump).
#ifdef TDSA_DEBUG_TIMING
#define TDSA_STAMP_AT(i, it_)                                                                                  \
  do {                                                                                                        \
    if ((threadIdx.x & 31) == 0 && a.dbg != nullptr && (it_) < 32 && blockDim.x == 256) {                     \
      long long c_;                                                                                           \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_)::"memory");                                            \
      a.dbg[(((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 32 + (it_)) * 16 + (i)] = c_;                   \
    }                                                                                                         \
  } while (0)
#define TDSA_STAMP(i) TDSA_STAMP_AT(i, it)
#else
#define TDSA_STAMP(i) do {} while (0)
#define TDSA_STAMP_AT(i, it_) do {} while (0)
#endif

// ---- named barriers (ids 1..4; id 0 is __syncthreads) ---------------------------------------
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// TAIL = 0: the whole transform (complex64 frames in, window applied).
// TAIL = 1: second kernel of the two-kernel large-FFT path: "frame" g is sub-transform
//           s2 = g & 255 of big frame g >> 8 (input complex T from big_head_kernel's scratch);
//           bin klow of it is bin (s2 >> 4) + 16*(s2 & 15) + 256*klow of the 256*N-point frame.
// TAIL = 2: same input, rows left in the permuted [g][klow] order (coalesced stores).
// TAIL = 3: second kernel after the ONE-pass head (N_big = 16*N): "frame" g is sub-transform s = g & 15 of big
//           frame g >> 4; bin klow of it is bin s + 16*klow.
// TWMODE: where the per-thread pass-0 constants (15 twiddles W_N^(t*q), 16 window values) live.
//   0 = read from the tables every frame (L1/L2);
//   1 = all of them in registers for the whole kernel (float32: 46 registers);
//   2 = six base twiddles W^(t*{1,2,3,4,8,12}) in registers, the other nine formed per frame as
//       W^(t*q0) * W^(t*4*q1) (one complex multiply each); window still read per frame (float64).
// NSTAGE > 0 (TAIL == 0 only): frames are staged by cp.async.bulk into an NSTAGE-deep ring of
//           shared-memory buffers, so frame i+NSTAGE streams in from HBM while frame i computes.
//           Needs 16-byte aligned frames (even frame_stride); the launcher falls back to NSTAGE = 0.
// GROUPS = 2: the CTA holds two independent frame engines of THREADS threads each (own exchange
//           buffer, staging ring and mbarriers).  Their butterfly phases are serialised against each
//           other with a pair of named barriers, so while one group is in a register/FP phase the
//           other is in a shared-memory exchange phase ("ping-pong"): the arithmetic pipe and the
//           shared-memory pipe overlap instead of both groups hitting the same pipe in lock-step.
template <typename T, int LOG2N, typename Epi, int TWMODE, int MIN_CTAS, int TAIL, int NSTAGE, int GROUPS, bool HAS_DC,
          int LOGR>
__global__ void __launch_bounds__(Plan<T, LOG2N, LOGR>::THREADS * GROUPS, MIN_CTAS)
fft_fused_kernel(const FftArgs<T> a) {
  using P = Plan<T, LOG2N, LOGR>;
  using CT = typename CplxOf<T>::type;
  constexpr int N = P::N, TH = P::THREADS, NPASS = P::NPASS, PP = P::P;
  constexpr size_t GROUP_BYTES = ((NSTAGE > 0 ? P::smem_staged(NSTAGE) : P::SMEM_BYTES) + 127) & ~(size_t)127;
  constexpr int kL2Ahead = TDSA_L2_AHEAD;   // frames prefetched into L2 beyond the shared-memory ring
  extern __shared__ __align__(128) unsigned char smem_all[];

  const int g = (GROUPS > 1) ? (int)(threadIdx.x / TH) : 0;
  const int t = (int)threadIdx.x - g * TH;
  unsigned char* smem_raw = smem_all + (size_t)g * GROUP_BYTES;
  CT* ex = reinterpret_cast<CT*>(smem_raw);
  CT* tws = ex + P::PHYS_SIZE;   // pass >= 1 twiddles
  const float2* stage0 = reinterpret_cast<const float2*>(smem_raw + P::STAGE_OFFSET);
  const uint32_t stage_u32 = smem_u32(smem_raw + P::STAGE_OFFSET);
  const uint32_t bar_u32 = stage_u32 + (uint32_t)(NSTAGE * P::STAGE_BYTES);   // NSTAGE mbarriers behind the ring
  (void)stage0; (void)bar_u32;

  // frames are dealt round-robin to (CTA, group) units
  const int64_t unit = (int64_t)blockIdx.x * GROUPS + g;
  const int64_t unit_stride = (int64_t)gridDim.x * GROUPS;

  auto group_sync = [&]() {
    if constexpr (GROUPS > 1) bar_sync(1 + g, TH); else __syncthreads();
  };
  // compute token: acquire before a register/FP phase, hand to the other group after it
  auto acquire = [&]() { if constexpr (GROUPS > 1) bar_sync(3 + g, 2 * TH); };
  auto release = [&]() { if constexpr (GROUPS > 1) bar_arrive(3 + (1 - g), 2 * TH); };

#if TDSA_PDL
  pdl_launch_dependents();
#endif
  if constexpr (NSTAGE > 0) {
    if (t == 0) {
#pragma unroll
      for (int s = 0; s < NSTAGE; ++s) mbar_init(bar_u32 + 8 * s, 1);
      fence_mbar_init();
    }
  }

  // one-time: the small tables into shared memory (DIT: pass 1's [j][K]; DIF: the tables of passes >= 1)
  constexpr bool DIT = P::DIT;
  if constexpr (P::TW_SMEM > 0) {
    for (int i = t; i < P::TW_SMEM; i += TH) tws[i] = a.tw[(DIT ? 0 : N) + i];
  }
  // one-time: per-thread constants that do not change from frame to frame: the window values of
  // pass 0 and the pre-twiddles of the LAST pass, W_N^(j * b) with b = t + TH*u (register index u*R + j)
  constexpr int R = P::R_LAST, NB = P::NB_LAST;
  constexpr int KLAST = N / R;                                   // entries per j in the last pass's table
  // the per-thread table: DIT = last pass's pre-twiddles [j][b]; DIF = pass 0's post-twiddles [q][t]
  const CT* tw_last = DIT ? a.tw + P::tw_offset(NPASS - 1) : a.tw;
  // base-twiddle mode needs 16 twiddles W^(j*x) per thread: a radix-16 last pass (DIT) or pass 0 (DIF);
  // otherwise fall back to registers for small CTAs (registers to spare) or per-frame loads
  constexpr int TWM = (TWMODE == 2 && (DIT ? R != 16 : PP != 16)) ? (TH <= 128 ? 1 : 0) : TWMODE;
  T win[PP];
  T twlr[TWM != 0 ? PP : 1], twli[TWM != 0 ? PP : 1];
  if constexpr (TWM == 1) {
    if constexpr (TAIL == 0) {
#pragma unroll
      for (int j = 0; j < PP; ++j) win[j] = a.window[t + j * TH];
    }
    if constexpr (DIT) {
#pragma unroll
      for (int u = 0; u < NB; ++u)
#pragma unroll
        for (int j = 1; j < R; ++j) { const CT w = tw_last[j * KLAST + t + TH * u]; twlr[u * R + j] = w.x; twli[u * R + j] = w.y; }
    } else {
#pragma unroll
      for (int q = 1; q < PP; ++q) { const CT w = tw_last[q * TH + t]; twlr[q] = w.x; twli[q] = w.y; }
    }
  } else if constexpr (TWM == 2) {
#pragma unroll
    for (int j = 1; j < 16; ++j) {
      if (j < 4 || (j & 3) == 0) { const CT w = tw_last[j * (DIT ? KLAST : TH) + t]; twlr[j] = w.x; twli[j] = w.y; }
    }
  }
#if TDSA_PDL
  pdl_wait_prior_grid();            // everything above only touched plan tables; frame data may come from the prior grid
#endif
  // Dynamic scheduling (staged single-group kernels with a scheduler): frame indices come from a global counter,
  // so the faster of the CTAs sharing an SM takes more frames instead of idling at the end; slot[s] holds the
  // frame that stage s carries.  The index for a refill is claimed one iteration ahead (no wait on the atomic).
  const bool dyn = NSTAGE > 0 && GROUPS == 1 && a.sched != nullptr;
  volatile int* slot = reinterpret_cast<volatile int*>(smem_raw + P::STAGE_OFFSET + (size_t)NSTAGE * P::STAGE_BYTES + 32);
  (void)slot;
  if constexpr (NSTAGE > 0) {
    if (t == 0) {
#pragma unroll
      for (int s = 0; s < NSTAGE; ++s) {
        const int64_t fs = dyn ? (int64_t)atomicAdd(a.sched, 1) : unit + (int64_t)s * unit_stride;
        if (dyn) slot[s] = (int)fs;
        if (fs < a.n_frames) {
          mbar_arrive_expect_tx(bar_u32 + 8 * s, (uint32_t)P::STAGE_BYTES);
          bulk_g2s(stage_u32 + (uint32_t)(s * P::STAGE_BYTES), a.iq + fs * a.frame_stride, (uint32_t)P::STAGE_BYTES,
                   bar_u32 + 8 * s);
        }
      }
#pragma unroll
      for (int s = NSTAGE; s < NSTAGE + kL2Ahead; ++s) {
        const int64_t fs = unit + (int64_t)s * unit_stride;
        if (fs < a.n_frames) bulk_prefetch_l2(a.iq + fs * a.frame_stride, (uint32_t)P::STAGE_BYTES);
      }
    }
  }
  if constexpr (P::TW_SMEM > 0 || NSTAGE > 0) __syncthreads();   // tables staged, mbarriers initialised
  if constexpr (GROUPS > 1) {
    if (g == 1) bar_arrive(3, 2 * TH);                           // group 0 owns the first compute phase
  }
  const bool mag20 = a.ep.mode == kModeMag20;
  if (a.stagger > 0 && blockIdx.x >= gridDim.x / 2) {
    const long long c0 = clock64();
    while (clock64() - c0 < a.stagger) {}
  }

  // Both groups run the same number of iterations so that the token hand-offs always pair up;
  // a group without a frame in the last iteration only passes the token on.
  const int64_t first_unit = (int64_t)blockIdx.x * GROUPS;
  const int64_t iters = dyn ? ((int64_t)1 << 40)
                            : (first_unit < a.n_frames ? (a.n_frames - first_unit + unit_stride - 1) / unit_stride : 0);

  for (int64_t it64 = 0; it64 < iters; ++it64) {
    const int it = (int)it64;
    int64_t f = unit + it64 * unit_stride;
    int fnext = 0;
    if constexpr (NSTAGE > 0) {
      if (dyn) {
        f = slot[it % NSTAGE];
        if (f >= a.n_frames) break;
        if (t == 0) fnext = atomicAdd(a.sched, 1);           // consumed at the refill below
      }
    }
    (void)fnext;
    if (GROUPS > 1 && f >= a.n_frames) {
#pragma unroll
      for (int ph = 0; ph < NPASS; ++ph) { acquire(); release(); }
      continue;
    }
    T re[PP], im[PP];
#ifdef TDSA_DEBUG_TIMING
    if (it == 0 && threadIdx.x == 0 && a.dbg != nullptr) {
      unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      a.dbg[(int64_t)gridDim.x * 8 * 32 * 16 + blockIdx.x] = smid;
    }
#endif
    TDSA_STAMP(0);
    // ---- pass 0: global/staged -> registers, window, full-radix DFT (no twiddles) -------------
    if constexpr (TAIL != 0) {
      const CT* src = a.in_ct + f * N + t;
#pragma unroll
      for (int j = 0; j < PP; ++j) { const CT x = src[j * TH]; re[j] = x.x; im[j] = x.y; }
      acquire();
    } else {
      if constexpr (TWM != 1) {             // issue the table loads before waiting on the frame
#pragma unroll
        for (int j = 0; j < PP; ++j) win[j] = a.window[t + j * TH];
      }
      T dcr = T(0), dci = T(0);
      if constexpr (HAS_DC) { double2 d = a.dc[f]; dcr = (T)d.x; dci = (T)d.y; }
      float2 v[PP];
      if constexpr (NSTAGE > 0) {
        const int stg = it % NSTAGE;
#if defined(TDSA_DEBUG_SKIP_MEM) || defined(TDSA_DEBUG_SKIP_LOAD)   // diagnostic: ring filled once, never refilled
        if (it < NSTAGE)
#endif
        mbar_wait(bar_u32 + 8 * stg, (uint32_t)((it / NSTAGE) & 1));
        TDSA_STAMP(1);
        const float2* src = stage0 + (size_t)stg * N + t;
#pragma unroll
        for (int j = 0; j < PP; ++j) v[j] = src[j * TH];
      } else {
        const float2* src = a.iq + f * a.frame_stride + t;
#pragma unroll
        for (int j = 0; j < PP; ++j) v[j] = ldg_stream(src + j * TH);
      }
      acquire();
      if constexpr (PP == 16) {            // window folded into the first butterfly level
#pragma unroll
        for (int j = 0; j < PP; ++j) {
          if constexpr (HAS_DC) { re[j] = (T)v[j].x - dcr; im[j] = (T)v[j].y - dci; }
          else { re[j] = (T)v[j].x; im[j] = (T)v[j].y; }
        }
        dft16_win<T>(re, im, win);
      } else {
#pragma unroll
        for (int j = 0; j < PP; ++j) {
          if constexpr (HAS_DC) {
            re[j] = ((T)v[j].x - dcr) * win[j];
            im[j] = ((T)v[j].y - dci) * win[j];
          } else {
            re[j] = (T)v[j].x * win[j];
            im[j] = (T)v[j].y * win[j];
          }
        }
      }
    }
    if constexpr (TAIL != 0 || PP != 16) dft_full<T, PP>(re, im);
    if constexpr (!DIT) {                    // DIF: post-twiddle W_N^(t*q)
#pragma unroll
      for (int q = 1; q < PP; ++q) {
        T wr, wi;
        if constexpr (TWM == 1) { wr = twlr[q]; wi = twli[q]; }
        else if constexpr (TWM == 2) {
          if (q < 4 || (q & 3) == 0) { wr = twlr[q]; wi = twli[q]; }
          else { wr = twlr[q & 3]; wi = twli[q & 3]; cmul<T>(wr, wi, twlr[q & ~3], twli[q & ~3]); }
        } else { const CT w = tw_last[q * TH + t]; wr = w.x; wi = w.y; }
        cmul<T>(re[q], im[q], wr, wi);
      }
    }
    release();
    TDSA_STAMP(2);
    {
      const int pb = P::phys(t);
#pragma unroll
      for (int q = 0; q < PP; ++q) ex[pb + P::phys(q * TH)] = mk<T>(re[q], im[q]);
    }
    TDSA_STAMP(3);
    group_sync();
    TDSA_STAMP(4);
    if constexpr (NSTAGE > 0) {
      // every thread of the group has consumed this stage (its reads precede the barrier): refill it
#if !defined(TDSA_DEBUG_SKIP_MEM) && !defined(TDSA_DEBUG_SKIP_LOAD)
      if (t == 0) {
        const int64_t fn = dyn ? (int64_t)fnext : f + (int64_t)NSTAGE * unit_stride;
        if (dyn) slot[it % NSTAGE] = fnext;
        if (fn < a.n_frames) {
          const int stg = it % NSTAGE;
          fence_proxy_async();
          mbar_arrive_expect_tx(bar_u32 + 8 * stg, (uint32_t)P::STAGE_BYTES);
          bulk_g2s(stage_u32 + (uint32_t)(stg * P::STAGE_BYTES), a.iq + fn * a.frame_stride, (uint32_t)P::STAGE_BYTES,
                   bar_u32 + 8 * stg);
        }
        const int64_t fp = f + (int64_t)(NSTAGE + kL2Ahead) * unit_stride;   // one more frame into L2
        if (kL2Ahead > 0 && fp < a.n_frames) bulk_prefetch_l2(a.iq + fp * a.frame_stride, (uint32_t)P::STAGE_BYTES);
      }
#endif
    }
    // ---- middle passes (full radix, in place, one butterfly per thread, pre-twiddled inputs) ---
#pragma unroll
    for (int i = 1; i < NPASS - 1; ++i) {
      const int L = N >> (LOGR * i), S = N >> (LOGR * (i + 1));
      const int c = t & (S - 1), s = t / S;
      const int pb = P::phys(s * L + c);
      if constexpr (DIT) {
        const int K = digitrev<LOGR>(s, i);                      // output digits fixed by passes 0..i-1
        const CT* tab = tws + K;                                 // DIT plans have one middle pass (i == 1)
        T wr[PP], wi[PP];
        wr[0] = T(1); wi[0] = T(0);
#pragma unroll
        for (int j = 1; j < PP; ++j) { const CT w = tab[j * P::kcount(i)]; wr[j] = w.x; wi[j] = w.y; }
#pragma unroll
        for (int j = 0; j < PP; ++j) { CT x = ex[pb + P::phys(j * S)]; re[j] = x.x; im[j] = x.y; }
        acquire();
        dft_full_pretw<T, PP>(re, im, wr, wi);
      } else {
#pragma unroll
        for (int j = 0; j < PP; ++j) { CT x = ex[pb + P::phys(j * S)]; re[j] = x.x; im[j] = x.y; }
        acquire();
        dft_full<T, PP>(re, im);
        const CT* twi_ = tws + (P::dif_offset(i) - N);
#pragma unroll
        for (int q = 1; q < PP; ++q) { const CT w = twi_[q * S + c]; cmul<T>(re[q], im[q], w.x, w.y); }
      }
      release();
      TDSA_STAMP(5);
#pragma unroll
      for (int q = 0; q < PP; ++q) ex[pb + P::phys(q * S)] = mk<T>(re[q], im[q]);
      TDSA_STAMP(6);
      group_sync();
      TDSA_STAMP(7);
    }
    // ---- last pass: radix R, NB butterflies per thread, digit-reversed reads, pre-twiddled ------
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int b = t + TH * u;
      const int pb = P::phys(digitrev<LOGR>(b, P::REV_DIGITS) * R);
#pragma unroll
      for (int j = 0; j < R; ++j) { CT x = ex[pb + P::phys(j)]; re[u * R + j] = x.x; im[u * R + j] = x.y; }
    }
    acquire();
    if constexpr (!DIT) {
      dft_last<T, PP, R>(re, im);
    } else {
      T wr[PP], wi[PP];
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        wr[u * R] = T(1); wi[u * R] = T(0);
#pragma unroll
        for (int j = 1; j < R; ++j) {
          const int e = u * R + j;
          if constexpr (TWM == 1) { wr[e] = twlr[e]; wi[e] = twli[e]; }
          else if constexpr (TWM == 2) {
            if (j < 4 || (j & 3) == 0) { wr[e] = twlr[j]; wi[e] = twli[j]; }
            else { wr[e] = twlr[j & 3]; wi[e] = twli[j & 3]; cmul<T>(wr[e], wi[e], twlr[j & ~3], twli[j & ~3]); }
          } else { const CT w = tw_last[j * KLAST + t + TH * u]; wr[e] = w.x; wi[e] = w.y; }
        }
      }
      if constexpr (R == 16 && PP == 16) {
        dft16_pretw<T>(re, im, wr, wi);
      } else {
#pragma unroll
        for (int e = 0; e < PP; ++e) { if (e % R != 0) cmul<T>(re[e], im[e], wr[e], wi[e]); }
        dft_last<T, PP, R>(re, im);
      }
    }
    auto emit = [&](auto mag_tag) {
      constexpr bool MAG = decltype(mag_tag)::value;
#pragma unroll
      for (int u = 0; u < NB; ++u) {
#pragma unroll
        for (int q = 0; q < R; ++q) {
          const int k = t + TH * u + (N / R) * q;
          const int e = u * R + q;
          const T pw = re[e] * re[e] + im[e] * im[e];
          if constexpr (TAIL == 1) {
            const int s2 = (int)(f & 255);
            Epi::template store<T, MAG>(a.ep, f >> 8, N * 256, (s2 >> 4) + 16 * (s2 & 15) + 256 * k, pw);
          } else if constexpr (TAIL == 3) {
            Epi::template store<T, MAG>(a.ep, f >> 4, N * 16, (int)(f & 15) + 16 * k, pw);
          } else {
            Epi::template store<T, MAG>(a.ep, f, N, k, pw);
          }
        }
      }
    };
    TDSA_STAMP(8);
    if (mag20) emit(std::true_type{}); else emit(std::false_type{});
    release();
    TDSA_STAMP(9);
    group_sync();   // exchange buffer is reused by the next frame's pass 0
    TDSA_STAMP(10);
  }
  if constexpr (NSTAGE > 0 && GROUPS == 1) {
    if (dyn && t == 0) {                                     // the last CTA out re-arms the scheduler for the next launch
      __threadfence();
      if (atomicAdd(a.sched + 1, 1) == (int)gridDim.x - 1) { a.sched[0] = 0; a.sched[1] = 0; __threadfence(); }
    }
  }
}

}  // namespace tdsa
