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

// How many staging buffers fit while keeping the CTAs/SM the register budget allows.
template <typename T, int LOG2N> constexpr int pick_stages() {
  using P = Plan<T, LOG2N>;
  constexpr size_t kSmemPerSm = 227 * 1024;
  constexpr int ctas = (P::THREADS >= 512) ? 1 : 2;
  int best = 0;
  if (LOG2N >= 9) {                            // tiny frames keep direct loads and many CTAs per SM
    for (int st = 1; st <= (sizeof(T) == 4 ? 3 : 2); ++st)
      if ((P::smem_staged(st) + 1024) * ctas <= kSmemPerSm) best = st;
  }
  return best;
}

template <typename T, int LOG2N, typename Epi, int TAIL, int NSTAGE>
cudaError_t launch_staged(const FftArgs<T>& a, int sm_count, cudaStream_t stream, LaunchInfo* info, bool dry);

template <typename T, int LOG2N, typename Epi, int TAIL>
cudaError_t launch_one(const FftArgs<T>& a, int sm_count, cudaStream_t stream, LaunchInfo* info, bool dry) {
  constexpr int kStages = (TAIL == 0) ? pick_stages<T, LOG2N>() : 0;
  if constexpr (kStages > 0) {
    const bool aligned = (((uintptr_t)a.iq & 15) == 0) && ((a.frame_stride & 1) == 0);
    if (aligned || dry) return launch_staged<T, LOG2N, Epi, TAIL, kStages>(a, sm_count, stream, info, dry);
  }
  return launch_staged<T, LOG2N, Epi, TAIL, 0>(a, sm_count, stream, info, dry);
}

template <typename T, int LOG2N, typename Epi, int TAIL, int NSTAGE>
cudaError_t launch_staged(const FftArgs<T>& a, int sm_count, cudaStream_t stream, LaunchInfo* info, bool dry) {
  using P = Plan<T, LOG2N>;
  constexpr size_t kSmem = NSTAGE > 0 ? P::smem_staged(NSTAGE) : P::SMEM_BYTES;
  // float32: window and pass-0 twiddles stay in registers across frames (47 registers);
  // float64 would need 94, so that path reads them through L1/L2 instead.
  constexpr bool kPersist = sizeof(T) == 4 && P::THREADS <= 512;
  constexpr int kMinCtas = (P::THREADS >= 512) ? 1 : 2;
  auto kern = fft_fused_kernel<T, LOG2N, Epi, kPersist, kMinCtas, TAIL, NSTAGE>;
  static int occ = -1;          // per instantiation, per process (single device type)
  if (occ < 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
    int o = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, P::THREADS, kSmem);
    if (e != cudaSuccess) return e;
    occ = std::max(o, 1);
  }
  const int64_t want = (int64_t)sm_count * occ;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(a.n_frames, want));
  if (info) { info->threads = P::THREADS; info->smem = (int)kSmem; info->ctas_per_sm = occ; info->grid = grid; info->stages = NSTAGE; }
  if (dry || a.n_frames <= 0) return cudaSuccess;
  kern<<<grid, P::THREADS, kSmem, stream>>>(a);
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
