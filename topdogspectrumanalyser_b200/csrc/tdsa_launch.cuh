// tdsa_launch.cuh — host-side dispatch of fft_fused_kernel over (size, epilogue).
#pragma once
#include <algorithm>
#include <atomic>

#include "tdsa_fft.cuh"

namespace tdsa {

enum : int { kEpiDb = 0, kEpiLinear = 1, kEpiDbTail = 2, kEpiLinearTail = 3, kEpiLinearPermuted = 4 };

struct LaunchInfo {
  int threads = 0;
  int smem = 0;
  int ctas_per_sm = 0;
  int grid = 0;
  int stages = 0;
};

extern std::atomic<int64_t> g_launch_count;

// Largest single-CTA size per element type (shared-memory bound): f32 2^14, f64 2^13.
template <typename T> struct MaxLog2 { static constexpr int value = sizeof(T) == 4 ? 14 : 13; };
constexpr int kMinLog2 = 6;

// ---- tuning knobs (overridable at build time for A/B runs; defaults are the measured best) -------
#ifndef TDSA_F32_TWMODE
#define TDSA_F32_TWMODE 1      // 1: window + all pass-0 twiddles in registers, 2: six base twiddles
#endif
#ifndef TDSA_F32_CTAS
#define TDSA_F32_CTAS 2        // CTAs per SM the float32 256-thread kernels are register-bounded for
#endif
#ifndef TDSA_F64_TWMODE
#define TDSA_F64_TWMODE 2
#endif
#ifndef TDSA_MAX_STAGES_F32
#define TDSA_MAX_STAGES_F32 3
#endif

template <typename T, int LOG2N> constexpr int target_ctas() {
  return (Plan<T, LOG2N>::THREADS >= 512) ? 1 : ((sizeof(T) == 4 && Plan<T, LOG2N>::THREADS == 256) ? TDSA_F32_CTAS : 2);
}

// How many staging buffers fit while keeping the CTAs/SM the register budget allows.
template <typename T, int LOG2N> constexpr int pick_stages() {
  using P = Plan<T, LOG2N>;
  constexpr size_t kSmemPerSm = 227 * 1024;
  constexpr int ctas = target_ctas<T, LOG2N>();
  int best = 0;
  if (LOG2N >= 9) {                            // tiny frames keep direct loads and many CTAs per SM
    for (int st = 1; st <= (sizeof(T) == 4 ? TDSA_MAX_STAGES_F32 : 2); ++st)
      if ((P::smem_staged(st) + 1024) * ctas <= kSmemPerSm) best = st;
  }
  return best;
}

template <typename T, int LOG2N, typename Epi, int TAIL, int NSTAGE>
cudaError_t launch_staged(const FftArgs<T>& a, int sm_count, cudaStream_t stream, LaunchInfo* info, bool dry);

// float64 at 256 threads per frame: one 512-thread CTA per SM holding two ping-pong frame groups
// (see fft_fused_kernel, GROUPS); everything else keeps independent CTAs.
// MEASURED (round 1, N=4096 float64): the ping-pong CTA ran 180 us vs 167 us for two independent
// CTAs per SM, because one 8-warp group alone cannot keep the FP64 pipe full; so it stays off.
#ifndef TDSA_PINGPONG
#define TDSA_PINGPONG 0
#endif
template <typename T, int LOG2N, int TAIL> constexpr int pick_groups() {
  return (TDSA_PINGPONG && sizeof(T) == 8 && TAIL == 0 && Plan<T, LOG2N>::THREADS == 256) ? 2 : 1;
}

template <typename T, int LOG2N, typename Epi, int TAIL>
cudaError_t launch_one(const FftArgs<T>& a, int sm_count, cudaStream_t stream, LaunchInfo* info, bool dry) {
  constexpr int kStages = (TAIL == 0) ? pick_stages<T, LOG2N>() : 0;
  if constexpr (kStages > 0) {
    const bool aligned = (((uintptr_t)a.iq & 15) == 0) && ((a.frame_stride & 1) == 0);
    if (aligned || dry) return launch_staged<T, LOG2N, Epi, TAIL, kStages>(a, sm_count, stream, info, dry);
  }
  return launch_staged<T, LOG2N, Epi, TAIL, 0>(a, sm_count, stream, info, dry);
}

template <typename T, int LOG2N, typename Epi, int TAIL, int NSTAGE, bool HAS_DC>
cudaError_t launch_dc(const FftArgs<T>& a, int sm_count, cudaStream_t stream, LaunchInfo* info, bool dry);

template <typename T, int LOG2N, typename Epi, int TAIL, int NSTAGE>
cudaError_t launch_staged(const FftArgs<T>& a, int sm_count, cudaStream_t stream, LaunchInfo* info, bool dry) {
  // the DC-removal variant (hackrf front end) only exists for the dB / linear epilogues of whole transforms
  if constexpr (TAIL == 0) {
    if (a.dc != nullptr) return launch_dc<T, LOG2N, Epi, TAIL, NSTAGE, true>(a, sm_count, stream, info, dry);
  }
  return launch_dc<T, LOG2N, Epi, TAIL, NSTAGE, false>(a, sm_count, stream, info, dry);
}

template <typename T, int LOG2N, typename Epi, int TAIL, int NSTAGE, bool HAS_DC>
cudaError_t launch_dc(const FftArgs<T>& a, int sm_count, cudaStream_t stream, LaunchInfo* info, bool dry) {
  using P = Plan<T, LOG2N>;
  constexpr int kGroups = pick_groups<T, LOG2N, TAIL>();
  constexpr size_t kSmemGroup = ((NSTAGE > 0 ? P::smem_staged(NSTAGE) : P::SMEM_BYTES) + 127) & ~(size_t)127;
  constexpr size_t kSmem = kSmemGroup * kGroups;
  constexpr int kThreads = P::THREADS * kGroups;
  // float32: window and pass-0 twiddles stay in registers across frames (47 registers);
  // float64 would need 94, so that path reads them through L1/L2 instead.
  // float64 keeps six base twiddles (24 registers) and forms the rest; large CTAs read tables.
  constexpr int kPersist = (P::THREADS > 512) ? 0 : (sizeof(T) == 4 ? TDSA_F32_TWMODE : TDSA_F64_TWMODE);
  constexpr int kMinCtas = (kThreads >= 512) ? 1 : target_ctas<T, LOG2N>();
  auto kern = fft_fused_kernel<T, LOG2N, Epi, kPersist, kMinCtas, TAIL, NSTAGE, kGroups, HAS_DC>;
  static int occ = -1;          // per instantiation, per process (single device type)
  if (occ < 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
    int o = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kThreads, kSmem);
    if (e != cudaSuccess) return e;
    occ = std::max(o, 1);
  }
  const int64_t want = (int64_t)sm_count * occ;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((a.n_frames + kGroups - 1) / kGroups, want));
  if (info) { info->threads = kThreads; info->smem = (int)kSmem; info->ctas_per_sm = occ; info->grid = grid; info->stages = NSTAGE; }
  if (dry || a.n_frames <= 0) return cudaSuccess;
  kern<<<grid, kThreads, kSmem, stream>>>(a);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

// Tail kernels exist for inner sizes 2^6 .. 2^12 (N = 2^14 .. 2^20).
constexpr int kMaxTailLog2 = 12;

template <typename T, int LOG2N>
cudaError_t launch_epi(int epi, const FftArgs<T>& a, int sm_count, cudaStream_t stream, LaunchInfo* info, bool dry) {
  switch (epi) {
    case kEpiDb: return launch_one<T, LOG2N, EpiDb, 0>(a, sm_count, stream, info, dry);
    case kEpiLinear: return launch_one<T, LOG2N, EpiLinear, 0>(a, sm_count, stream, info, dry);
    default: break;
  }
  if constexpr (LOG2N <= kMaxTailLog2) {
    switch (epi) {
      case kEpiDbTail: return launch_one<T, LOG2N, EpiDb, 1>(a, sm_count, stream, info, dry);
      case kEpiLinearTail: return launch_one<T, LOG2N, EpiLinear, 1>(a, sm_count, stream, info, dry);
      case kEpiLinearPermuted: return launch_one<T, LOG2N, EpiLinear, 2>(a, sm_count, stream, info, dry);
      default: break;
    }
  }
  return cudaErrorInvalidValue;
}

template <typename T, int LOG2N> struct Dispatch {
  static cudaError_t run(int log2n, int epi, const FftArgs<T>& a, int sm, cudaStream_t s, LaunchInfo* info, bool dry) {
    if (log2n == LOG2N) return launch_epi<T, LOG2N>(epi, a, sm, s, info, dry);
    if constexpr (LOG2N < MaxLog2<T>::value) return Dispatch<T, LOG2N + 1>::run(log2n, epi, a, sm, s, info, dry);
    return cudaErrorInvalidValue;
  }
};

template <typename T>
cudaError_t launch_fft_impl(int log2n, int epi, const FftArgs<T>& a, int sm, cudaStream_t s, LaunchInfo* info, bool dry) {
  if (log2n < kMinLog2 || log2n > MaxLog2<T>::value) return cudaErrorInvalidValue;
  return Dispatch<T, kMinLog2>::run(log2n, epi, a, sm, s, info, dry);
}

// defined in tdsa_fft_f32.cu / tdsa_fft_f64.cu
cudaError_t launch_fft_f32(int log2n, int epi, const FftArgs<float>& a, int sm, cudaStream_t s, LaunchInfo* info, bool dry);
cudaError_t launch_fft_f64(int log2n, int epi, const FftArgs<double>& a, int sm, cudaStream_t s, LaunchInfo* info, bool dry);
// head kernel of the two-kernel path (tdsa_big.cuh), defined next to the matching precision
template <typename T> struct BigArgs;
cudaError_t launch_big_head_f32(const BigArgs<float>& a, int sm, cudaStream_t s);
cudaError_t launch_big_head_f64(const BigArgs<double>& a, int sm, cudaStream_t s);

}  // namespace tdsa
