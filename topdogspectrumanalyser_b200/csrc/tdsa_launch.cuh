// tdsa_launch.cuh — host-side dispatch of fft_fused_kernel over (size, epilogue, staging, radix).
#pragma once
#include <stdlib.h>

#include <stdio.h>

#include <algorithm>
#include <vector>
#include <atomic>

#include "tdsa_fft.cuh"
#include "tdsa_fft_wl.cuh"
#include "tdsa_welch_cluster.cuh"

namespace tdsa {

enum : int { kEpiDb = 0, kEpiLinear = 1, kEpiDbTail = 2, kEpiLinearTail = 3, kEpiLinearPermuted = 4,
             kEpiDbTail16 = 5, kEpiLinearTail16 = 6 };   // Tail16: after the one-pass head (N_big = 16*M)

struct LaunchInfo {
  int threads = 0;
  int smem = 0;
  int ctas_per_sm = 0;
  int grid = 0;
  int stages = 0;
  int logr = 4;
};

extern std::atomic<int64_t> g_launch_count;

constexpr int kMaxDevices = 64;   // per-device caches of launch attributes (function attributes are per device)
inline int current_device() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < kMaxDevices) ? d : 0; }

// Largest single-CTA size per element type (shared-memory bound): f32 2^14, f64 2^13.
template <typename T> struct MaxLog2 { static constexpr int value = sizeof(T) == 4 ? 14 : 13; };
constexpr int kMinLog2 = 6;
// radix-8 (8 points per thread) instantiations exist for these sizes
constexpr int kMinLog2R8 = 9, kMaxLog2R8 = 13;

// ---- tuning knobs (overridable at build time for A/B runs; defaults are the measured best) -------
#ifndef TDSA_F32_TWMODE
#define TDSA_F32_TWMODE 1      // 1: window + all pass-0 twiddles in registers, 2: base twiddles only
#endif
#ifndef TDSA_F32_CTAS
#define TDSA_F32_CTAS 2        // CTAs per SM the float32 256-thread radix-16 kernels are register-bounded for
#endif
#ifndef TDSA_F64_TWMODE
#define TDSA_F64_TWMODE 2
#endif
#ifndef TDSA_MAX_STAGES_F32
#define TDSA_MAX_STAGES_F32 3
#endif
// MEASURED (round 1, N=4096 float64): the ping-pong CTA ran 180 us vs 167 us for two independent
// CTAs per SM, because one 8-warp group alone cannot keep the FP64 pipe full; so it stays off.
#ifndef TDSA_PINGPONG
#define TDSA_PINGPONG 0
#endif
// default digit width per precision for sizes where both exist (runtime override: TDSA_LOGR_F32/F64)
#ifndef TDSA_DEFAULT_LOGR_F32
#define TDSA_DEFAULT_LOGR_F32 4
#endif
#ifndef TDSA_DEFAULT_LOGR_F64
#define TDSA_DEFAULT_LOGR_F64 4
#endif

#ifndef TDSA_SMALL_CTAS_F64
#define TDSA_SMALL_CTAS_F64 6
#endif
template <typename T, int LOG2N, int LOGR> constexpr int target_ctas() {
  constexpr int th = Plan<T, LOG2N, LOGR>::THREADS;
  if (LOGR == 3) return th >= 1024 ? 1 : (1024 / th > 8 ? 8 : 1024 / th);      // 64 registers per thread
  // float64, N <= 2048 (CTAs of 32 .. 128 threads): ask for 12 warps per SM.  With the old bound of two CTAs the kernel took
  // 170 registers and ran 4 CTAs of 64 threads per SM at N = 1024; 6 CTAs (156 registers, one staging buffer) measured
  // 127.2 -> 115.3 us at N = 1024, 123.9 -> 115.8 us at N = 512, 133.9 -> 123.7 us at N = 2048 (round 2); 5 or 7 per SM are
  // both slower (130 us: 7 needs 128 registers and spills).
  if (sizeof(T) == 8 && th <= 128 && th >= 32) return TDSA_SMALL_CTAS_F64 * 64 / th;
  return (th >= 512) ? 1 : ((sizeof(T) == 4 && th == 256) ? TDSA_F32_CTAS : 2);
}

// How many staging buffers fit while keeping the CTAs/SM the register budget allows.
template <typename T, int LOG2N, int LOGR> constexpr int pick_stages() {
  using P = Plan<T, LOG2N, LOGR>;
  constexpr size_t kSmemPerSm = 227 * 1024;
  constexpr int ctas = target_ctas<T, LOG2N, LOGR>();
  int best = 0;
  if (LOG2N >= 9) {                            // tiny frames keep direct loads and many CTAs per SM
    for (int st = 1; st <= (sizeof(T) == 4 ? TDSA_MAX_STAGES_F32 : 2); ++st)
      if ((P::smem_staged(st) + 1024) * ctas <= kSmemPerSm) best = st;
  }
  return best;
}

// TDSA_PINGPONG: 0 = off, 1 = float64 only, 2 = float32 only, 3 = both (N/16 == 256 threads per frame only)
template <typename T, int LOG2N, int TAIL> constexpr int pick_groups() {
  constexpr bool on = (sizeof(T) == 8) ? (TDSA_PINGPONG & 1) : (TDSA_PINGPONG & 2);
  return (on && TAIL == 0 && Plan<T, LOG2N, 4>::THREADS == 256) ? 2 : 1;
}

template <typename T, int LOG2N, typename Epi, int TAIL, int NSTAGE, bool HAS_DC, int LOGR>
cudaError_t launch_final(const FftArgs<T>& a, int sm_count, cudaStream_t stream, LaunchInfo* info, bool dry) {
  using P = Plan<T, LOG2N, LOGR>;
  constexpr int kGroups = (LOGR == 4) ? pick_groups<T, LOG2N, TAIL>() : 1;
  constexpr size_t kSmemGroup = ((NSTAGE > 0 ? P::smem_staged(NSTAGE) : P::SMEM_BYTES) + 127) & ~(size_t)127;
  // diagnostic: TDSA_DEBUG_EXTRA_SMEM=<bytes> pads the dynamic shared memory to force fewer CTAs per SM
  static const size_t kExtraSmem = [] { const char* e = getenv("TDSA_DEBUG_EXTRA_SMEM"); return e ? (size_t)atol(e) : (size_t)0; }();
  const size_t kSmem = std::min<size_t>(kSmemGroup * kGroups + kExtraSmem, 227 * 1024);
  constexpr int kThreads = P::THREADS * kGroups;
  // float32: window and pass-0 twiddles stay in registers across frames; float64 keeps base twiddles
  // and forms the rest; CTAs of more than 512 threads read the tables (64-register budget).
  constexpr int kTwMode = (kThreads > 512) ? 0 : (sizeof(T) == 4 ? TDSA_F32_TWMODE : TDSA_F64_TWMODE);
  constexpr int kMinCtas = (kGroups > 1) ? 1 : target_ctas<T, LOG2N, LOGR>();
  auto kern = fft_fused_kernel<T, LOG2N, Epi, kTwMode, kMinCtas, TAIL, NSTAGE, kGroups, HAS_DC, LOGR>;
  static int occ_of[kMaxDevices] = {};          // per instantiation and device (function attributes are per device)
  const int dev = current_device();
  if (occ_of[dev] == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
    int o = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kThreads, kSmem);
    if (e != cudaSuccess) return e;
    occ_of[dev] = std::max(o, 1);
  }
  const int occ = occ_of[dev];
  const int64_t want = (int64_t)sm_count * occ;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((a.n_frames + kGroups - 1) / kGroups, want));
  if (info) {
    info->threads = kThreads; info->smem = (int)kSmem; info->ctas_per_sm = occ; info->grid = grid;
    info->stages = NSTAGE; info->logr = LOGR;
  }
  if (dry || a.n_frames <= 0) return cudaSuccess;
  static const int kStagger = [] { const char* e = getenv("TDSA_DEBUG_STAGGER"); return e ? atoi(e) : 0; }();
  FftArgs<T> b = a;
  b.stagger = kStagger;
#ifdef TDSA_DEBUG_TIMING
  static long long* d_dbg = nullptr;
  const size_t dbg_count = (size_t)grid * 8 * 32 * 16 + grid;
  const char* dbg_path = getenv("TDSA_DEBUG_TIMING_OUT");
  if (dbg_path && kThreads == 256 && LOG2N == 12 && TAIL == 0) {
    if (!d_dbg) cudaMalloc(&d_dbg, sizeof(long long) * ((size_t)1024 * 8 * 32 * 16 + 1024));
    cudaMemsetAsync(d_dbg, 0, sizeof(long long) * dbg_count, stream);
    b.dbg = d_dbg;
  }
#endif
#if TDSA_PDL
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, b);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return le != cudaSuccess ? le : cudaGetLastError();
#else
  kern<<<grid, kThreads, kSmem, stream>>>(b);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
#ifdef TDSA_DEBUG_TIMING
  if (b.dbg) {
    cudaStreamSynchronize(stream);
    std::vector<long long> h(dbg_count);
    cudaMemcpy(h.data(), d_dbg, sizeof(long long) * dbg_count, cudaMemcpyDeviceToHost);
    char path[512];
    snprintf(path, sizeof path, "%s_%s_g%d.bin", dbg_path, sizeof(T) == 4 ? "f32" : "f64", grid);
    if (FILE* fp = fopen(path, "wb")) { fwrite(h.data(), sizeof(long long), dbg_count, fp); fclose(fp); }
  }
#endif
  return cudaGetLastError();
#endif
}

// DC removal (hackrf front end) and bulk-copy staging are only compiled for whole transforms (TAIL == 0).
template <typename T, int LOG2N, typename Epi, int TAIL, int LOGR>
cudaError_t launch_one(const FftArgs<T>& a, int sm_count, cudaStream_t stream, LaunchInfo* info, bool dry) {
  if constexpr (TAIL == 0) {
    constexpr int kStages = pick_stages<T, LOG2N, LOGR>();
    const bool aligned = (((uintptr_t)a.iq & 15) == 0) && ((a.frame_stride & 1) == 0);
    if constexpr (kStages > 0) {
      if (aligned || dry) {
        if (a.dc != nullptr) return launch_final<T, LOG2N, Epi, 0, kStages, true, LOGR>(a, sm_count, stream, info, dry);
        return launch_final<T, LOG2N, Epi, 0, kStages, false, LOGR>(a, sm_count, stream, info, dry);
      }
    }
    if (a.dc != nullptr) return launch_final<T, LOG2N, Epi, 0, 0, true, LOGR>(a, sm_count, stream, info, dry);
    return launch_final<T, LOG2N, Epi, 0, 0, false, LOGR>(a, sm_count, stream, info, dry);
  } else {
    return launch_final<T, LOG2N, Epi, TAIL, 0, false, LOGR>(a, sm_count, stream, info, dry);
  }
}

// Tail kernels exist for inner sizes 2^6 .. 2^12 (N = 2^14 .. 2^20).
constexpr int kMaxTailLog2 = 12;

template <typename T> int runtime_logr() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv(sizeof(T) == 4 ? "TDSA_LOGR_F32" : "TDSA_LOGR_F64");
    v = e ? atoi(e) : (sizeof(T) == 4 ? TDSA_DEFAULT_LOGR_F32 : TDSA_DEFAULT_LOGR_F64);
    if (v != 3 && v != 4) v = 4;
  }
  return v;
}

template <typename T, int LOG2N>
cudaError_t launch_epi(int epi, const FftArgs<T>& a, int sm_count, cudaStream_t stream, LaunchInfo* info, bool dry) {
  if constexpr (LOG2N >= kMinLog2R8 && LOG2N <= kMaxLog2R8) {
    if (runtime_logr<T>() == 3) {
      if (epi == kEpiDb) return launch_one<T, LOG2N, EpiDb, 0, 3>(a, sm_count, stream, info, dry);
      if (epi == kEpiLinear) return launch_one<T, LOG2N, EpiLinear, 0, 3>(a, sm_count, stream, info, dry);
    }
  }
  switch (epi) {
    case kEpiDb: return launch_one<T, LOG2N, EpiDb, 0, 4>(a, sm_count, stream, info, dry);
    case kEpiLinear: return launch_one<T, LOG2N, EpiLinear, 0, 4>(a, sm_count, stream, info, dry);
    default: break;
  }
  if constexpr (LOG2N <= kMaxTailLog2) {
    switch (epi) {
      case kEpiDbTail: return launch_one<T, LOG2N, EpiDb, 1, 4>(a, sm_count, stream, info, dry);
      case kEpiLinearTail: return launch_one<T, LOG2N, EpiLinear, 1, 4>(a, sm_count, stream, info, dry);
      case kEpiLinearPermuted: return launch_one<T, LOG2N, EpiLinear, 2, 4>(a, sm_count, stream, info, dry);
      case kEpiDbTail16: return launch_one<T, LOG2N, EpiDb, 3, 4>(a, sm_count, stream, info, dry);
      case kEpiLinearTail16: return launch_one<T, LOG2N, EpiLinear, 3, 4>(a, sm_count, stream, info, dry);
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

// which digit width launch_fft_* will use for this size (the twiddle tables must match)
template <typename T> int effective_logr(int log2n) {
  return (log2n >= kMinLog2R8 && log2n <= kMaxLog2R8 && runtime_logr<T>() == 3) ? 3 : 4;
}

// ---- warp-local kernels (tdsa_fft_wl.cuh): N = 4096 on one engine, N = 8192 on two ---------------------------
#ifndef TDSA_WL_STAGES_F32
#define TDSA_WL_STAGES_F32 2
#endif
#ifndef TDSA_WL_TWB_BASE_F32    // float32: pass-B twiddles from six base values in registers instead of 15 LDS.64 per frame
#define TDSA_WL_TWB_BASE_F32 1   // measured: 80.7 -> 78.8 us
#endif
#ifndef TDSA_WL_CTAS_F32          // CTAs per SM the float32 one-engine kernel (plain rows) is built for
#define TDSA_WL_CTAS_F32 2
#endif
#ifndef TDSA_WL_STAGES_F64
#define TDSA_WL_STAGES_F64 1
#endif

// lane -> (sub-transform r, team lane c) of an engine of fft_wl_kernel; the host permutes the window with the same map
inline void wl_thread_identity(int tid, int* r, int* c) {
  const int w = (tid & 255) >> 5, l = tid & 31;
  *r = 2 * w + ((l >> 3) & 1);
  *c = (l & 7) + 8 * (l >> 4);
}


// n_units: claimable units = frames, or groups of acc.group frames (kAccGroup)
template <typename T, typename Epi, bool HAS_DC, int NB, int ACC>
cudaError_t launch_wl_final(const FftArgs<T>& a, const CUtensorMap& tmap, const T* wperm, WlSched sched, const WlAcc& acc,
                            int device, int sm_count, cudaStream_t stream, LaunchInfo* info, bool dry) {
  constexpr int kStages = (NB == 2 && sizeof(T) == 8) ? 1 : (sizeof(T) == 4 ? TDSA_WL_STAGES_F32 : TDSA_WL_STAGES_F64);
  constexpr int kTwMode = sizeof(T) == 4 ? TDSA_F32_TWMODE : TDSA_F64_TWMODE;
  constexpr int kThreads = 256 * NB;
  constexpr int kMinCtas = NB == 1 ? ((sizeof(T) == 4 && ACC == 0) ? TDSA_WL_CTAS_F32 : 2) : 1;
  static const size_t kExtraSmem = [] { const char* e = getenv("TDSA_DEBUG_EXTRA_SMEM"); return e ? (size_t)atol(e) : (size_t)0; }();
  constexpr bool kSub = (ACC & kAccSub) != 0;                  // direct loads: no staging buffers
  constexpr bool kFused = (ACC & kAccFused) != 0;              // + the head pass' sample staging [16][272] float2
  const size_t kSmem = std::min<size_t>(WlPlan<T, NB>::smem_bytes(kSub ? 0 : kStages) + (kFused ? 16 * 272 * 8 : 0) + kExtraSmem,
                                        227 * 1024);
  auto kern = fft_wl_kernel<T, Epi, kTwMode, kStages, HAS_DC, kMinCtas, (sizeof(T) == 4 ? TDSA_WL_TWB_BASE_F32 : 0) != 0, NB, ACC>;
  static int occ_of[kMaxDevices] = {};          // 0 = not queried on that device yet
  if (device < 0 || device >= kMaxDevices) return cudaErrorInvalidDevice;
  if (occ_of[device] == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
    int o = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kThreads, kSmem);
    if (e != cudaSuccess) return e;
    // Kernels that allocate tensor memory are reported as one CTA per SM by the occupancy calculator (ncu capture of
    // round 2: grid 148, limits by registers and shared memory both 2).  The accumulating epilogue allocates 128 of
    // the 512 TMEM columns per CTA (at most 256 with the window values parked there too), so two CTAs do fit; registers (launch bounds) and shared memory allow exactly two.
    constexpr bool kTmemUser = ACC != 0 || (TDSA_WL_WIN_TMEM && sizeof(T) == 8 && NB == 1);
    if (kTmemUser && NB == 1 && o < 2 && 2 * kSmem <= 227 * 1024) o = 2;
    occ_of[device] = std::max(o, 1);
  }
  const int occ = occ_of[device];
  const int group = (ACC & kAccGroup) ? std::max(acc.group, 1) : 1;
  const int64_t n_units = a.n_frames / group;
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_units, (int64_t)sm_count * occ));
  if constexpr (kSub)                                          // CTA b serves class b % 16: whole sets of sixteen classes
    grid = 16 * (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(a.n_frames / 16, (int64_t)sm_count * occ / 16),
                                                            kFused ? kFusedMaxGroups : (1 << 20)));
  if (info) {
    info->threads = kThreads; info->smem = (int)kSmem; info->ctas_per_sm = occ; info->grid = grid;
    info->stages = kStages; info->logr = 4;
  }
  if (dry || a.n_frames <= 0) return cudaSuccess;
  FftArgs<T> b = a;
#ifdef TDSA_DEBUG_TIMING
  static long long* d_dbg = nullptr;
  const size_t dbg_count = (size_t)grid * 8 * 32 * 16 + grid;
  const char* dbg_path = getenv("TDSA_DEBUG_TIMING_OUT");
  if (dbg_path && NB == 1) {
    if (!d_dbg) cudaMalloc(&d_dbg, sizeof(long long) * ((size_t)1024 * 8 * 32 * 16 + 1024));
    cudaMemsetAsync(d_dbg, 0, sizeof(long long) * dbg_count, stream);
    b.dbg = d_dbg;
  }
#endif
  kern<<<grid, kThreads, kSmem, stream>>>(b, tmap, wperm, sched, acc);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
#ifdef TDSA_DEBUG_TIMING
  if (b.dbg) {
    cudaStreamSynchronize(stream);
    std::vector<long long> h(dbg_count);
    cudaMemcpy(h.data(), d_dbg, sizeof(long long) * dbg_count, cudaMemcpyDeviceToHost);
    char path[512];
    snprintf(path, sizeof path, "%s_wl_%s_g%d.bin", dbg_path, sizeof(T) == 4 ? "f32" : "f64", grid);
    if (FILE* fp = fopen(path, "wb")) { fwrite(h.data(), sizeof(long long), dbg_count, fp); fclose(fp); }
  }
#endif
  return cudaGetLastError();
}

// what the warp-local launchers need besides FftArgs
struct WlLaunch {
  const CUtensorMap* tmap;
  const void* wperm;       // T[N]: window in thread order (two halves for two engines)
  WlSched sched;
  WlAcc acc;
  int nb;                  // engines: 1 (N = 4096) or 2 (N = 8192)
  int acc_flags;           // 0, or one of the instantiated kAcc* combinations below
  int device, sm_count;
};
constexpr int kAccAvg = kAccSum;                               // running average as a weighted sum (last row only)
constexpr int kAccHold = kAccMax | kAccMin | kAccRows;         // dB rows + max/min hold on un-averaged frames
constexpr int kAccWelch = kAccSum | kAccMax;                   // Welch mean + peak
constexpr int kAccGroupMean = kAccSum | kAccGroup;             // one dB row per group of frames
constexpr int kAccWelchSub = kAccSum | kAccMax | kAccSub;      // Welch mean + peak over the sub-transforms of 65536-point segments
constexpr int kAccWelchFused = kAccWelchSub | kAccFused;       // ... with the head pass in the same kernel (intermediate stays in L2)

template <typename T>
cudaError_t launch_wl_impl(int epi, const FftArgs<T>& a, const WlLaunch& L, cudaStream_t s, LaunchInfo* info, bool dry) {
  const bool dc = a.dc != nullptr;
  const T* wperm = (const T*)L.wperm;
#define TDSA_WL_GO(EPI, DC, NB, ACC) \
  return launch_wl_final<T, EPI, DC, NB, ACC>(a, *L.tmap, wperm, L.sched, L.acc, L.device, L.sm_count, s, info, dry)
  if (L.nb == 1) {
    switch (L.acc_flags) {
      case 0:
        if (epi == kEpiDb) { if (dc) TDSA_WL_GO(EpiDb, true, 1, 0); TDSA_WL_GO(EpiDb, false, 1, 0); }
        if (epi == kEpiLinear) { if (dc) TDSA_WL_GO(EpiLinear, true, 1, 0); TDSA_WL_GO(EpiLinear, false, 1, 0); }
        break;
      case kAccAvg: if (!dc) TDSA_WL_GO(EpiDb, false, 1, kAccAvg); break;
      case kAccHold: if (!dc) TDSA_WL_GO(EpiDb, false, 1, kAccHold); break;
      case kAccWelch: if (!dc) TDSA_WL_GO(EpiDb, false, 1, kAccWelch); break;
      case kAccGroupMean: if (!dc) TDSA_WL_GO(EpiDb, false, 1, kAccGroupMean); break;
      case kAccWelchSub: if (!dc) TDSA_WL_GO(EpiDb, false, 1, kAccWelchSub); break;
      case kAccWelchFused: if (!dc) TDSA_WL_GO(EpiDb, false, 1, kAccWelchFused); break;
      default: break;
    }
  } else if (L.nb == 2 && !dc) {
    // Two engines (N = 8192): the group-mean epilogue in both precisions, and plain dB rows in float64.  Plain rows,
    // round 2: at first no better than the four-pass classic kernel (f64 207 vs 199 us, f32 115 vs 105 us: the engines
    // share one stage and run in lock step); with the float64 window values in tensor memory 188 us through the
    // group epilogue with one frame per group, so float64 rows take this kernel and float32 rows stay on the classic one.
    if (L.acc_flags == kAccGroupMean) TDSA_WL_GO(EpiDb, false, 2, kAccGroupMean);
    if constexpr (sizeof(T) == 8) {
      if (L.acc_flags == 0 && epi == kEpiDb) TDSA_WL_GO(EpiDb, false, 2, 0);
    }
  }
#undef TDSA_WL_GO
  return cudaErrorInvalidValue;
}
// true when launch_wl_impl has an instantiation for this combination
inline bool wl_supported(int nb, int epi, int acc_flags, bool dc, bool f64 = false) {
  if (nb == 1) {
    if (acc_flags == 0) return epi == kEpiDb || epi == kEpiLinear;
    return !dc && (acc_flags == kAccAvg || acc_flags == kAccHold || acc_flags == kAccWelch || acc_flags == kAccGroupMean ||
                   acc_flags == kAccWelchSub || acc_flags == kAccWelchFused);
  }
  if (nb == 2 && !dc) return acc_flags == kAccGroupMean || (f64 && acc_flags == 0 && epi == kEpiDb);
  return false;
}
cudaError_t launch_wl_f32(int epi, const FftArgs<float>& a, const WlLaunch& L, cudaStream_t s, LaunchInfo* info, bool dry);
cudaError_t launch_wl_f64(int epi, const FftArgs<double>& a, const WlLaunch& L, cudaStream_t s, LaunchInfo* info, bool dry);

// ---- cluster Welch kernel (tdsa_welch_cluster.cuh): clusters of 16 CTAs, one CTA per SM -------------------------
// max_clusters != nullptr: only report how many clusters can be co-resident (0 = this device cannot run it)
template <typename T>
cudaError_t launch_welch_cluster_impl(const WelchClusterArgs<T>& a, int clusters, cudaStream_t stream, int* max_clusters) {
  auto kern = welch_cluster_kernel<T>;
  constexpr size_t kSmem = WelchClusterPlan<T>::SMEM_BYTES;
  static int max_active_of[kMaxDevices];
  static bool queried[kMaxDevices] = {};
  const int dev = current_device();
  int& max_active = max_active_of[dev];
  if (!queried[dev]) { max_active = -1; queried[dev] = true; }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kWcCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = kSmem; cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = 1;
  if (max_active < 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int n = 0;
    if (e == cudaSuccess) {
      cfg.gridDim = dim3(kWcCluster * 8);
      e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    }
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    max_active = n;
    if (getenv("TDSA_DEBUG_PRINT")) fprintf(stderr, "welch_cluster_kernel: %d clusters of %d CTAs can be co-resident\n", n, kWcCluster);
  }
  if (max_clusters) { *max_clusters = max_active; return cudaSuccess; }
  if (max_active <= 0) return cudaErrorNotSupported;
  cfg.gridDim = dim3(kWcCluster * clusters);
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, a);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return le != cudaSuccess ? le : cudaGetLastError();
}
cudaError_t launch_welch_cluster_f32(const WelchClusterArgs<float>& a, int clusters, cudaStream_t s, int* max_clusters);
cudaError_t launch_welch_cluster_f64(const WelchClusterArgs<double>& a, int clusters, cudaStream_t s, int* max_clusters);

// defined in tdsa_fft_f32.cu / tdsa_fft_f64.cu
cudaError_t launch_fft_f32(int log2n, int epi, const FftArgs<float>& a, int sm, cudaStream_t s, LaunchInfo* info, bool dry);
cudaError_t launch_fft_f64(int log2n, int epi, const FftArgs<double>& a, int sm, cudaStream_t s, LaunchInfo* info, bool dry);
int effective_logr_f32(int log2n);
int effective_logr_f64(int log2n);
// head kernel of the two-kernel path (tdsa_big.cuh), defined next to the matching precision
template <typename T> struct BigArgs;
// passes = 2: big_head_kernel (N = 256*M); passes = 1: big_head1_kernel (N = 16*M)
cudaError_t launch_big_head_f32(const BigArgs<float>& a, int sm, cudaStream_t s, int passes);
cudaError_t launch_big_head_f64(const BigArgs<double>& a, int sm, cudaStream_t s, int passes);
// passes = 0: big_head_wl_kernel (N = 65536, output in the thread order of fft_wl_kernel's kAccSub mode)

}  // namespace tdsa
