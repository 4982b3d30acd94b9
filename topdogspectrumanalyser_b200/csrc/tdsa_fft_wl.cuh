// tdsa_fft_wl.cuh — "warp-local" fused window + FFT + |.|^2 + dB kernel for N = 4096 (one engine) and N = 8192
// (two engines), with an optional accumulating epilogue whose per-bin state lives in tensor memory (TMEM).
//
// Same arithmetic as fft_fused_kernel (tdsa_fft.cuh; reference datasources/rtl_samples.py:169-184), different
// schedule.  Phase time stamps of fft_fused_kernel (profiles/r01_phase_timing.md) showed that its three CTA-wide
// barriers per frame keep all eight warps of a CTA in the same phase, so the arithmetic pipe and the
// shared-memory pipe are used one after the other instead of together, and that the two CTAs of an SM run at
// very different speeds, which a static frame assignment turns into an idle tail.  Here:
//
//   * an ENGINE is 256 threads computing one 4096-point transform by decimation in time over the first factor,
//     4096 = 16 x 256, n = r + 16 m.  The sixteen 256-point sub-transforms Y_r are each computed by a team of 16
//     lanes (two teams per warp): radix-16 pass A, a 16x16 transpose through the team's private shared-memory
//     region ordered by __syncwarp only, radix-16 pass B.  Warps are independent for two of the three passes.
//   * one engine-wide exchange: X[kk + 256 q] = sum_r W16^(r q) W4096^(r kk) Y_r[kk]; thread kk reads the sixteen
//     Y_r[kk] (stride = region pitch, conflict-free) and owns bins kk + 256 q (coalesced stores).
//   * the frame is staged by cp.async.bulk.tensor (TMA, 3-D tensor map [frame][rows][128 B]) with the 128-byte
//     swizzle, so that a team's stride-16 sample reads (row m, column r) are bank-conflict free.
//   * frames are claimed from a global counter (dynamic scheduling): the faster CTA of an SM simply takes
//     more frames; the counter re-arms itself when the last CTA leaves.
//
// N = 8192 (NB = 2): one radix-2 decimation-in-frequency step is folded into the staged read.  Both engines read
// x[n] w[n] and x[n + 4096] w[n + 4096] from the same 64 KB stage; engine 0 transforms their sum (even bins),
// engine 1 their difference with every twiddle table evaluated at the half-integer bin k + 1/2 (odd bins:
// X[2k+1] = sum_n d[n] W4096^(n (k + 1/2))), which costs fifteen constant W32^j multiplies in pass A and nothing
// else.  The engines synchronise internally with named barriers and meet only at the stage hand-back (mbarrier).
//
// Accumulating epilogue (ACC != 0): per-bin running state over the frames a CTA processes — weighted sum of
// |X|^2 (float64), max and min of |X|^2 (float32) — for order-free reductions: the running average of
// TraceAverager (utils/signal_processing.py:35-61) written as a weighted sum, max/min hold on un-averaged rows
// (core/display_data_processor.py:371-395), Welch mean + peak (config 3) and per-sub-band means (config 4).
// Sixteen bins x 16 bytes per thread do not fit the register budget next to a float64 radix-16 butterfly, so the
// state is kept in TMEM (tcgen05.alloc / tcgen05.ld / tcgen05.st, 64 columns per thread's lane; no tensor-core
// math is involved) and flushed to per-CTA partial rows at the end of the launch.
//
// Twiddle tables per engine: [0, 256): pass B, entry [j][ka]; [256, 4352): last pass, entry [j][kk].
#pragma once
#include <cuda.h>

#include "tdsa_fft.cuh"

namespace tdsa {

enum : int { kAccSum = 1, kAccMax = 2, kAccMin = 4, kAccGroup = 8, kAccRows = 16, kAccSub = 32, kAccFused = 64 };
// kAccSub (tail of the 65536-point Welch path, tdsa_big.cuh): the "frames" are the sixteen 4096-point sub-transforms
// of each segment, already windowed, in complex T and in thread order ([frame][j][tid], written by big_head_wl_kernel).
// No window, no TMA staging: every thread loads its sixteen values with coalesced 16-byte (float64) loads.  CTA b only
// takes sub-transforms of class b % 16 (sub-transform s holds the bins k = s mod 16), so that its TMEM accumulators
// see one set of bins for the whole launch; each class has its own claim counter (sched.next[2 + s]).
//
// kAccFused (with kAccSub): the head pass runs inside the same kernel, so that the intermediate never leaves the L2.
// The grid is cut into groups of sixteen CTAs (group g = blockIdx / 16, class s = blockIdx % 16); group g owns the
// segments g, g + G, g + 2G, ... completely: for a segment, CTA (g, s) first computes column block s of the head pass
// (samples 256 s + t + 4096 j: window, radix 16, twiddle W_65536^(c q)) and writes its 256 values of each of the sixteen
// sub-transforms into the group's ring of kFusedRing segments; once all sixteen CTAs of the group have done so it transforms
// sub-transform s.  The head of segment i + 1 is issued BEFORE the tail of segment i, so the group's two counters
// (heads done, tails done; release / acquire through global memory) are normally already satisfied when they are
// looked at.  All CTAs must be co-resident (two per SM); a spin that does not end sets an error word instead of hanging.
// Measured slower than head and tails as two launches (see welch_fused_enabled in tdsa_api.cu): opt-in.
constexpr int kMaxPeers = 8;

// arguments of the accumulating epilogue
struct WlAcc {
  const double* weight = nullptr;   // [n_frames] weight of frame f in the sum; nullptr = 1
  const int32_t* skip = nullptr;    // [n_frames] 1 = the frame leaves all state untouched (hackrf silence hold); nullptr = none
  double* part_sum = nullptr;       // [grid][N] per-CTA partial weighted sums (kAccSum without kAccGroup)
  float* part_max = nullptr;        // [grid][N] per-CTA running max of |X|^2 (-inf = no live frame)
  float* part_min = nullptr;        // [grid][N] per-CTA running min of |X|^2 (+inf = no live frame)
  float* group_db = nullptr;        // kAccGroup: [n_frames / group][N] dB of each group's mean
  double* unit_sum = nullptr;       // kAccGroup, split groups: [n_frames / group][N] float64 SUM of each unit's frames
                                    // (a unit is a part of a group; group_finish_kernel adds the parts and takes the dB)
  // kAccGroup, multi-GPU (config 4): instead of group_db the row of group g is stored straight into EVERY rank's copy of
  // the gathered row table (peer memory over NVLink), at row peer_row0 + g: compute and all-gather in one kernel
  float* peer_rows[kMaxPeers] = {};
  int n_peers = 0;
  int64_t peer_row0 = 0;
  // kAccFused: the head pass' inputs and the groups' rings / counters
  const float2* fused_iq = nullptr; // sample stream
  int64_t fused_hop = 0;            // samples between segment starts
  int64_t fused_nseg = 0;
  const void* fused_tw = nullptr;   // complex T [16][4096]: W_65536^(c q)
  void* fused_y = nullptr;          // complex T [groups][kFusedRing][16 q][16 j][256]
  int* fused_cnt = nullptr;         // [groups][2] {heads done, tails done}, then one error word at [2 * groups]
  int group = 1;                    // frames per claimed unit (kAccGroup: frames per group)
  int64_t only_row = -1;            // kAccRows: >= 0 stores the dB row of this frame only (at row 0), -1 stores every row
};

constexpr int kWlTwPerEngine = 256 + 4096;

template <typename T, int NB = 1> struct WlPlan {
  static constexpr int N = 4096 * NB, TH = 256 * NB;
  static constexpr int REGION = 272 + (sizeof(T) == 4 ? 8 : 0);      // elements per team region (pitch-17 rows + bank offset)
  static constexpr int EX_ELEMS = 16 * REGION;                        // per engine
  static constexpr int TW_SMEM = 256;                                 // pass-B table [j][ka], per engine
  static constexpr size_t ENGINE_BYTES = (size_t)(EX_ELEMS + TW_SMEM) * 2 * sizeof(T);
  static constexpr size_t EX_BYTES = ENGINE_BYTES * NB;
  static constexpr size_t STAGE_BYTES = (size_t)N * 8;
  static constexpr size_t CTRL_BYTES = 128;                           // mbarriers (full[4], empty[4]), frame slots, TMEM base
  static __host__ __device__ constexpr size_t smem_bytes(int nstage) {
    return ((EX_BYTES + CTRL_BYTES + 1023) & ~(size_t)1023) + 1024 + (size_t)nstage * STAGE_BYTES;
  }
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst_smem),
      "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---- tensor memory as per-thread scratch: 32x32b shape = lane <-> thread, consecutive columns <-> registers ------
template <int COLS> __device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
  static_assert(COLS == 32 || COLS == 64 || COLS == 128 || COLS == 256 || COLS == 512, "power of two >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
               "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct WlSched {
  int* next;       // next unclaimed unit (frame, or group of frames); kAccSub: next[2 + s] = next segment of class s
  int* done;       // CTAs that have left the frame loop
};
#ifndef TDSA_FUSED_RING
#define TDSA_FUSED_RING 3
#endif
constexpr int kFusedRing = TDSA_FUSED_RING;   // segments per group ring: the head runs one segment ahead, the ring gives the slack
constexpr int kFusedMaxGroups = 32;     // fused mode: group counters live in the scheduler block
constexpr int kWlSchedWords = 128;      // {next, done}, 16 class counters, fused mode: [18, 18 + 2 * groups] group counters + error word

// W32^j = exp(-2 pi i j / 32), j = 0..15 (pass-A pre-twiddles of the half-bin engine); j is a compile-time constant
// at every use (unrolled loops), so these fold to immediates
template <typename T> __device__ __forceinline__ T w32_cos(int j) {
  constexpr double c[16] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                            0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785,
                            0.0, -0.19509032201612826785, -0.38268343236508977173, -0.55557023301960222474,
                            -0.70710678118654752440, -0.83146961230254523708, -0.92387953251128675613, -0.98078528040323044913};
  return (T)c[j];
}
template <typename T> __device__ __forceinline__ T w32_msin(int j) {   // imaginary part: -sin(2 pi j / 32)
  constexpr double s[16] = {0.0, -0.19509032201612826785, -0.38268343236508977173, -0.55557023301960222474,
                            -0.70710678118654752440, -0.83146961230254523708, -0.92387953251128675613, -0.98078528040323044913,
                            -1.0, -0.98078528040323044913, -0.92387953251128675613, -0.83146961230254523708,
                            -0.70710678118654752440, -0.55557023301960222474, -0.38268343236508977173, -0.19509032201612826785};
  return (T)s[j];
}

// TDSA_WL_EARLY = 1: the stage is refilled as soon as every warp has reported its staged reads done (one
// mbarrier.arrive per warp; thread 0 polls without blocking after its own pass A and pass B, and waits after the Y
// barrier at the latest) instead of always after the Y barrier; the frame is claimed at the top of the iteration.
// TDSA_WL_WIN_TMEM = 1 (float64, one engine): the thread's sixteen window values live in tensor memory (32 columns of
// its lane) instead of being re-read from global memory through L1 for every frame (the register budget has no room
// for them): 16 LDG.64 per thread and frame leave the LSU / L1 path, which is the tighter side of this kernel.
#ifndef TDSA_WL_EPI_F32SQ
#define TDSA_WL_EPI_F32SQ 0
#endif
// TDSA_WL_WIDEN_INT = 1 (float64): the staged float32 samples are widened to float64 with integer instructions (exponent
// re-bias + mantissa shift; zero / denormal -> signed zero, inf / nan -> exponent all ones) instead of 32 F2F.F64.F32 per
// thread and frame, which cost ~4 cycles each on the FP64 pipe (8 % of the pipe's time per frame).
#ifndef TDSA_WL_WIDEN_INT
#define TDSA_WL_WIDEN_INT 0
#endif
template <typename T> __device__ __forceinline__ T widen_sample(float x) { return (T)x; }
#if TDSA_WL_WIDEN_INT
template <> __device__ __forceinline__ double widen_sample<double>(float x) {
  const uint32_t b = __float_as_uint(x);
  const uint32_t ex = b & 0x7f800000u;
  uint32_t hi = (b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u);
  uint32_t lo = b << 29;
  if (ex == 0u) { hi = b & 0x80000000u; lo = 0u; }
  if (ex == 0x7f800000u) hi |= 0x7ff00000u;
  return __hiloint2double((int)hi, (int)lo);
}
#endif
// last-pass base twiddles held in registers: 6 (w^1..w^3, w^4, w^8, w^12; 24 registers in float64) or 2 (w^1, w^4; the
// other four bases by multiplication every frame: 16 registers fewer, 16 DFMA more)
#ifndef TDSA_WL_TWL_BASE2
#define TDSA_WL_TWL_BASE2 0
#endif
#ifndef TDSA_WL_WIN_TMEM
#define TDSA_WL_WIN_TMEM 0
#endif
// TDSA_WL_WIN_RELOAD = 1 (float64, one engine): force the per-frame re-read of the sixteen window values.  The source asks
// for it, but ptxas hoists the loop-invariant loads and keeps the values in 32 registers (ncu: 54 k global load requests
// per launch instead of one million).  Measured: 127.9 -> 135.2 us with the forced re-read, so the hoisted form stays.
#ifndef TDSA_WL_WIN_RELOAD
#define TDSA_WL_WIN_RELOAD 0
#endif
#ifndef TDSA_WL_WIN2_TMEM
#define TDSA_WL_WIN2_TMEM 1
#endif
// TDSA_WL_TWB_SMEM_BASE (float64): pass-B twiddles W256^(j c) from a few base values of the shared table plus complex
// multiplies instead of fifteen LDS.128 per thread and frame.  Measured (8192 frames, on top of the split barrier):
// 0: 132.1 us; 1 (six bases, nine multiplies): 129.0 us; 2 (w^1 and w^4 only, thirteen multiplies): 128.0 us.  The same
// idea for the last pass' register-held bases (TDSA_WL_TWL_BASE2: 16 registers fewer) changes nothing (129.0 us), nor do
// squaring in float32 after narrowing re / im (TDSA_WL_EPI_F32SQ: 130.5 us) or integer-pipe narrowing (129.1 us).
// Tensor memory for the per-frame window values
// (TDSA_WL_WIN_TMEM) measured slower (135.2 / 131.1 us without / with this option): the blocking tcgen05.wait::ld costs
// more than sixteen L1-cached loads.
#ifndef TDSA_WL_TWB_SMEM_BASE
#define TDSA_WL_TWB_SMEM_BASE 2
#endif
#ifndef TDSA_WL_EARLY
#define TDSA_WL_EARLY 0
#endif
// TDSA_WL_SPLIT_B2 = 1 (one engine only): the second engine-wide barrier of a frame ("every warp has read its Y values,
// the regions may be overwritten") becomes a split barrier: one mbarrier.arrive per warp right after the loads, the wait
// right before the next frame's team stores, i.e. a whole last pass + epilogue + staged read + pass A later.
// Measured (round 2, 8192 frames): 135.2 -> 131.1 us in float64, 77.8 -> 77.8 us (best 77.6 -> 75.8) in float32; on.
// (Together with TDSA_WL_EARLY it is pathological: 184 us.)
#ifndef TDSA_WL_SPLIT_B2
#define TDSA_WL_SPLIT_B2 1
#endif
// The same split barrier per engine of the two-engine kernel: measured slower in float64 (group mean 236 -> 273 us, rows
// 182 -> 234 us; float32 160 -> 154 us), so off: with one CTA per SM the rendezvous keeps the two engines' phases apart.
#ifndef TDSA_WL_SPLIT_B2_NB2
#define TDSA_WL_SPLIT_B2_NB2 0
#endif
// TDSA_WL_SPLIT_ALL = 1: every thread arrives on the regions-free barrier; 0: lane 0 of each warp after a __syncwarp.  Same
// speed (round 2: 128.9 / 77.8 us against 129.0 / 78.4 us), but compute-sanitizer's racecheck credits an mbarrier arrival to
// the arriving thread only and reported the other lanes' last-pass reads against the next frame's team stores as WAR hazards
// with the per-warp arrival (2.7 million reports on the config-3 workload, none with one arrival per thread); on.
#ifndef TDSA_WL_SPLIT_ALL
#define TDSA_WL_SPLIT_ALL 1
#endif

template <typename T, typename Epi, int TWMODE, int NSTAGE, bool HAS_DC, int MIN_CTAS, bool TWB_BASE, int NB, int ACC>
__global__ void __launch_bounds__(256 * NB, MIN_CTAS)
fft_wl_kernel(const FftArgs<T> a, const __grid_constant__ CUtensorMap tmap, const T* __restrict__ wperm, WlSched sched,
              const WlAcc acc) {
  using W = WlPlan<T, NB>;
  using CT = typename CplxOf<T>::type;
  constexpr int N = W::N, TH = W::TH, REGION = W::REGION;
  // tensor memory per warp (one lane per thread): 64 accumulator columns (ACC) and / or 32 window columns (WIN_TMEM);
  // 2 * NB warps share a lane quarter, each with its own column range
  constexpr bool kWinTmem = TDSA_WL_WIN_TMEM && sizeof(T) == 8 && NB == 1 && TWMODE != 1;
  constexpr int kFusedWinCols = (ACC & kAccFused) ? 16 * (int)sizeof(T) / 4 : 0;     // fused head: its window values
  // two engines, float64: the thread's 32 window values (both halves of the frame) live in tensor memory.  The permuted
  // table is 128 KB for the CTA and does not stay in what is left of the L1 next to 212 KB of shared memory, so the
  // per-frame re-read went to L2: `long_scoreboard` was the top stall of the config-4 kernel (2.25 cycles per issued
  // instruction, r02_group8192_f64_before_tmem_window_ncu_full.txt; 266 -> 233 us for config 4's 300 x 16 frames).  64 columns per thread: with the accumulators all 512 columns of the SM.
  constexpr bool kWin2Tmem = TDSA_WL_WIN2_TMEM && sizeof(T) == 8 && NB == 2 && TWMODE != 1;
  constexpr int kWin2Col = ACC != 0 ? 64 : 0;                 // behind the accumulators when there are any
  constexpr int kTmemPerWarp = (ACC != 0 ? 64 : 0) + (kWinTmem ? 32 : 0) + kFusedWinCols + (kWin2Tmem ? 64 : 0);
  constexpr int kTmemNeed = kTmemPerWarp * 2 * NB;
  constexpr int kTmemCols = kTmemNeed <= 32 ? 32 : kTmemNeed <= 64 ? 64 : kTmemNeed <= 128 ? 128 : kTmemNeed <= 256 ? 256 : 512;
  constexpr bool kUseTmem = kTmemPerWarp != 0;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t base_u32 = smem_u32(smem_raw);
  const uint32_t ctrl_u32 = base_u32 + (uint32_t)W::EX_BYTES;              // full[4] at +0, empty[4] at +32
  volatile int* slot = reinterpret_cast<volatile int*>(smem_raw + W::EX_BYTES + 64);      // staged frame per stage
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + W::EX_BYTES + 96);
  const uint32_t stage_u32 = (base_u32 + (uint32_t)(W::EX_BYTES + W::CTRL_BYTES) + 1023u) & ~1023u;
  const unsigned char* stage_ptr = smem_raw + (stage_u32 - base_u32);

  const int tid = (int)threadIdx.x;
  const int e = NB == 1 ? 0 : (tid >> 8);                   // engine
  const int te = tid & 255;                                 // thread inside the engine = bin identity kk of the last pass
  const int w = te >> 5, l = tid & 31;
  // lane -> (team half h, team lane c): each half-warp of lanes covers all eight 16-byte swizzle chunks and both
  // 8-byte halves, so the 64-bit staged reads are conflict free
  const int h = (l >> 3) & 1, c = (l & 7) + 8 * (l >> 4);
  const int r = 2 * w + h;                                  // sub-transform (team) 0..15
  CT* ex = reinterpret_cast<CT*>(smem_raw + (size_t)e * W::ENGINE_BYTES);
  CT* tws = ex + W::EX_ELEMS;
  CT* reg = ex + r * REGION;
  const CT* twe = a.tw + e * kWlTwPerEngine;                // this engine's tables

  constexpr bool kSplitB2 = TDSA_WL_SPLIT_B2 && NSTAGE <= 2 && (NB == 1 || TDSA_WL_SPLIT_B2_NB2);
  constexpr bool SUB = (ACC & kAccSub) != 0;
  constexpr bool FUSED = (ACC & kAccFused) != 0;
  static_assert(!SUB || (NB == 1 && !HAS_DC && (ACC & (kAccGroup | kAccRows)) == 0), "sub-transform tail: one engine, accumulate only");
  static_assert(!FUSED || (SUB && TDSA_WL_SPLIT_B2), "the fused head needs the sub-transform tail and its split barrier");
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < (SUB ? 0 : NSTAGE); ++s) {
      mbar_init(ctrl_u32 + 8 * s, 1);
      if constexpr (NB > 1 || TDSA_WL_EARLY) mbar_init(ctrl_u32 + 32 + 8 * s, 8 * NB);     // one arrival per warp
    }
    if constexpr (kSplitB2) {                                 // "regions free", one per engine (the slots of stages 2 and 3)
#pragma unroll
      for (int en = 0; en < NB; ++en) mbar_init(ctrl_u32 + 16 + 8 * en, TDSA_WL_SPLIT_ALL ? 256 : 8);
    }
    fence_mbar_init();
  }
  for (int i = te; i < W::TW_SMEM; i += 256) tws[i] = twe[i];
  uint32_t tacc = 0;                                        // TMEM address of this thread's 64 accumulator columns
  uint32_t twin = 0;                                        // TMEM address of this thread's 32 window columns
  if constexpr (kUseTmem) {
    if ((tid >> 5) == 0) tmem_alloc<kTmemCols>(base_u32 + (uint32_t)W::EX_BYTES + 96);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }

  // per-thread constants: window values of pass A (team identity) and last-pass pre-twiddles (bin identity te)
  const CT* tw_last = twe + 256;
  constexpr int NWIN = 16 * NB;
  T win[NWIN];
  T twlr[16], twli[16];
  if constexpr (TWMODE == 1) {
    if constexpr (!SUB) {
#pragma unroll
      for (int j = 0; j < NWIN; ++j) win[j] = wperm[j * TH + tid];
    }
#pragma unroll
    for (int j = 1; j < 16; ++j) { const CT x = tw_last[j * 256 + te]; twlr[j] = x.x; twli[j] = x.y; }
  } else {
#pragma unroll
    for (int j = 1; j < 16; ++j) {
      if (TDSA_WL_TWL_BASE2 && sizeof(T) == 8 && j != 1 && j != 4) continue;
      if (j < 4 || (j & 3) == 0) { const CT x = tw_last[j * 256 + te]; twlr[j] = x.x; twli[j] = x.y; }
    }
  }

  // pass-B twiddles depend on the thread's team lane only: optionally six of them stay in registers
  T twbr[TWB_BASE ? 16 : 1], twbi[TWB_BASE ? 16 : 1];
  if constexpr (TWB_BASE) {
#pragma unroll
    for (int j = 1; j < 16; ++j) {
      if (j < 4 || (j & 3) == 0) { const CT x = twe[j * 16 + c]; twbr[j] = x.x; twbi[j] = x.y; }
    }
  }
  // frame claims are made in units of `group` consecutive frames (1 unless the epilogue averages groups)
  const int group = (ACC & kAccGroup) ? acc.group : 1;
  int unit_base = 0, unit_pos = 0;                          // thread 0: the unit being handed out, frames already taken from it
  auto next_frame = [&]() -> int {                          // thread 0 only
    if constexpr (SUB) {
      const int cls = (int)(blockIdx.x & 15);
      const int u = atomicAdd(sched.next + 2 + cls, 1);
      return ((int64_t)u * 16 >= a.n_frames) ? 0x7fffffff : u * 16 + cls;
    }
    if (unit_pos == 0 || unit_pos == group) {
      const int u = atomicAdd(sched.next, 1);
      if ((int64_t)u * group >= a.n_frames) { unit_pos = group; return 0x7fffffff; }
      unit_base = u * group; unit_pos = 0;
    }
    return unit_base + unit_pos++;
  };
  // thread 0 only: publish the frame staged in slot s and start its copy; with no frame left the stage's barrier is
  // completed by a plain arrival so that every waiter wakes up, reads the sentinel and leaves the loop
  auto issue_stage = [&](int s, int f) {
    slot[s] = f;
    if (f < a.n_frames) {
      mbar_arrive_expect_tx(ctrl_u32 + 8 * s, (uint32_t)W::STAGE_BYTES);
#pragma unroll
      for (int b = 0; b < NB; ++b)
        tma_load_3d(stage_u32 + (uint32_t)(s * W::STAGE_BYTES) + (uint32_t)b * 32768u, &tmap, 0, 256 * b, f, ctrl_u32 + 8 * s);
    } else {
      mbar_arrive(ctrl_u32 + 8 * s);
    }
  };
  if (tid == 0) {
    if constexpr (FUSED) {
    } else if constexpr (SUB) {
      slot[0] = next_frame();
    } else {
#pragma unroll
      for (int s = 0; s < NSTAGE; ++s) issue_stage(s, next_frame());
    }
  }
  __syncthreads();
  if constexpr (kUseTmem) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tacc = *tmem_slot + ((uint32_t)(((tid >> 5) & 3) * 32) << 16) + (uint32_t)((tid >> 7) * kTmemPerWarp);
    twin = tacc + (ACC != 0 ? 64 : 0);
  }
  if constexpr (kWin2Tmem) {                                 // columns [64, 128): win[0..31] as float64 pairs
    const uint32_t tw2 = tacc + kWin2Col;
#pragma unroll
    for (int part = 0; part < 4; ++part) {
      uint32_t u[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double wv = (double)wperm[(8 * part + i) * TH + tid];
        u[2 * i] = (uint32_t)__double2loint(wv); u[2 * i + 1] = (uint32_t)__double2hiint(wv);
      }
      tmem_st16(tw2 + 16 * part, u);
    }
    tmem_wait_st();
  }
  if constexpr (kWinTmem) {                                  // park the window values of pass A in tensor memory
    uint32_t u[16];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double wv = (double)wperm[(8 * half + i) * TH + tid];
        u[2 * i] = (uint32_t)__double2loint(wv); u[2 * i + 1] = (uint32_t)__double2hiint(wv);
      }
      tmem_st16(twin + 16 * half, u);
    }
    tmem_wait_st();
  }
  if constexpr (ACC != 0) {
    // state: columns [0, 32) sixteen float64 sums, [32, 48) sixteen float32 maxima, [48, 64) sixteen float32 minima
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
    tmem_st16(tacc, z); tmem_st16(tacc + 16, z);
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0xff800000u;        // -inf
    tmem_st16(tacc + 32, z);
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0x7f800000u;        // +inf
    tmem_st16(tacc + 48, z);
    tmem_wait_st();
  }
  // ---- fused head pass (kAccFused) -------------------------------------------------------------------------------
  const int f_cls = (int)(blockIdx.x & 15), f_grp = (int)(blockIdx.x >> 4), f_ngrp = (int)(gridDim.x >> 4);
  int f_nit = 0;                                            // segments this CTA's group owns
  CT* f_ring = nullptr;
  int *f_hcnt = nullptr, *f_tcnt = nullptr, *f_err = nullptr;
  volatile int* f_dead = reinterpret_cast<volatile int*>(smem_raw + W::EX_BYTES + 100);
  const int f_col = 256 * f_cls + r + 16 * c;               // this thread's column of the segment (tail thread order)
  // staging of the head's samples [16 j][272 float2] behind the tail's shared memory: natural column t at
  // (t & 15) + 17 (t >> 4), filled by cp.async one segment ahead
  float2* f_stage = reinterpret_cast<float2*>(smem_raw + W::smem_bytes(0));
  const uint32_t f_put = base_u32 + (uint32_t)W::smem_bytes(0) + (uint32_t)((tid & 15) + 17 * (tid >> 4)) * 8u;
  auto f_fetch = [&](int64_t seg) {                          // samples 256 f_cls + tid + 4096 j of segment seg, coalesced
    const float2* src = acc.fused_iq + seg * acc.fused_hop + 256 * f_cls + tid;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(f_put + (uint32_t)(j * 272 * 8)), "l"(src + 4096 * j) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if constexpr (FUSED) {
    f_nit = f_grp < acc.fused_nseg ? (int)((acc.fused_nseg - f_grp + f_ngrp - 1) / f_ngrp) : 0;
    if (f_nit > 0) f_fetch(f_grp);
    // the window values of this thread's column live in tensor memory (behind the accumulators): no registers, no
    // shared memory, and no sector-inefficient permuted global reads per segment
    const uint32_t twn = tacc + 64;
#pragma unroll
    for (int part = 0; part < kFusedWinCols / 16; ++part) {
      uint32_t u[16];
      if constexpr (sizeof(T) == 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const double wv = (double)a.window[f_col + 4096 * (8 * part + i)];
          u[2 * i] = (uint32_t)__double2loint(wv); u[2 * i + 1] = (uint32_t)__double2hiint(wv);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) u[i] = __float_as_uint((float)a.window[f_col + 4096 * i]);
      }
      tmem_st16(twn + 16 * part, u);
    }
    tmem_wait_st();
    f_ring = reinterpret_cast<CT*>(acc.fused_y) + (int64_t)f_grp * kFusedRing * 65536;
    f_hcnt = acc.fused_cnt + 2 * f_grp; f_tcnt = f_hcnt + 1; f_err = acc.fused_cnt + 2 * f_ngrp;
    if (tid == 0) *f_dead = 0;
    __syncthreads();
  }
  // thread 0 polls the group counter (acquire); an endless wait (CTAs not co-resident) ends in the error word
  auto f_wait = [&](const int* cnt, int target) {
    if (tid == 0 && *f_dead == 0) {
      long long polls = 0;
      for (;;) {
        int v, e;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
        if (v >= target) break;
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(e) : "l"(f_err) : "memory");
        if (e != 0 || ++polls > (1ll << 24)) { *f_err = 1; *f_dead = 1; break; }
      }
    }
    __syncthreads();
  };
  // head pass of the staged segment for column block f_cls into slot `buf` of the group's ring, then the fetch of
  // segment `seg_after` (< 0: none).  The stores are NOT fenced here: the count that publishes them is added at the Y
  // barrier of the following tail, when they have long been performed (a fence right here waited ~1 us for them).
  int f_it = 0;                                              // iteration the head pass is charged to (time stamps)
  auto f_head = [&](int buf, int64_t seg_after) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                         // everybody's copies of this segment have landed
    TDSA_STAMP_AT(11, f_it);
    T hr[16], hi[16];
    {
      T wv[16];
      const uint32_t twn = tacc + 64;
#pragma unroll
      for (int part = 0; part < kFusedWinCols / 16; ++part) {
        uint32_t u[16];
        tmem_ld16(twn + 16 * part, u);
        if constexpr (sizeof(T) == 8) {
#pragma unroll
          for (int i = 0; i < 8; ++i) wv[8 * part + i] = (T)__hiloint2double((int)u[2 * i + 1], (int)u[2 * i]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) wv[i] = (T)__uint_as_float(u[i]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 v = f_stage[j * 272 + r + 17 * c];
        hr[j] = (T)v.x * wv[j]; hi[j] = (T)v.y * wv[j];
      }
    }
    __syncthreads();                                         // the staging area is free again
    TDSA_STAMP_AT(12, f_it);
    if (seg_after >= 0) f_fetch(seg_after);
    dft16<T>(hr, hi);
    const CT* htw = reinterpret_cast<const CT*>(acc.fused_tw) + f_col;
    T wr[16], wi[16];
    { const CT x = htw[1 * 4096]; wr[1] = x.x; wi[1] = x.y; }
    { const CT x = htw[4 * 4096]; wr[4] = x.x; wi[4] = x.y; }
    wr[2] = wr[1]; wi[2] = wi[1]; cmul<T>(wr[2], wi[2], wr[1], wi[1]);
    wr[3] = wr[2]; wi[3] = wi[2]; cmul<T>(wr[3], wi[3], wr[1], wi[1]);
    wr[8] = wr[4]; wi[8] = wi[4]; cmul<T>(wr[8], wi[8], wr[4], wi[4]);
    wr[12] = wr[8]; wi[12] = wi[8]; cmul<T>(wr[12], wi[12], wr[4], wi[4]);
    CT* dst = f_ring + ((int64_t)(buf * 16) * 16 + f_cls) * 256 + tid;   // [buf][q][j = f_cls][tid]
    TDSA_STAMP_AT(13, f_it);
    dst[0] = mk<T>(hr[0], hi[0]);
#pragma unroll
    for (int q = 1; q < 16; ++q) {
      T xr = wr[q & 3], xi = wi[q & 3];
      if ((q & 3) == 0) { xr = wr[q]; xi = wi[q]; }
      else if (q >= 4) cmul<T>(xr, xi, wr[q & ~3], wi[q & ~3]);
      cmul<T>(hr[q], hi[q], xr, xi);
      dst[(int64_t)q * 4096] = mk<T>(hr[q], hi[q]);
    }
    TDSA_STAMP_AT(14, f_it);
  };
  bool f_head_pending = false;                               // a head pass whose count has not been added yet
  if constexpr (FUSED) {
    if (f_nit > 0) {
      f_head(0, f_nit > 1 ? (int64_t)f_grp + f_ngrp : (int64_t)-1);
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicAdd(f_hcnt, 1);
    }
  }
  const bool mag20 = a.ep.mode == kModeMag20;
  // swizzled offset of (row m = c + 16 j, column r) inside a staged 4096-sample block: 128-byte rows, 16-byte chunk
  // index XOR (row & 7); r >> 1 == w, r & 1 == h, (c + 16 j) & 7 == c & 7
  const int stage_off = c * 128 + (((w ^ c) & 7) << 4) + h * 8;
  // engine-internal barriers: NB == 1 uses barrier 0 (__syncthreads), NB == 2 named barriers 1 + e over 256 threads
  auto engine_sync = [&]() {
    if constexpr (NB == 1) __syncthreads(); else bar_sync(1 + e, 256);
  };

  for (int it = 0;; ++it) {
    const int stg = it % NSTAGE;
    uint32_t wraw[kWinTmem ? 32 : 1];
    if constexpr (kWinTmem) {                                // issued here, waited for after the stage barrier
      uint32_t (&lo16)[16] = *reinterpret_cast<uint32_t (*)[16]>(&wraw[0]);
      uint32_t (&hi16)[16] = *reinterpret_cast<uint32_t (*)[16]>(&wraw[16]);
      tmem_ld16_nowait(twin, lo16);
      tmem_ld16_nowait(twin + 16, hi16);
    } else if constexpr (TWMODE != 1 && !SUB && !kWin2Tmem) {   // window values for pass A, re-read every frame (register budget)
#pragma unroll
      for (int j = 0; j < NWIN; ++j) {
#if TDSA_WL_WIN_RELOAD
        if constexpr (sizeof(T) == 8 && NB == 1) {           // volatile: really re-read per frame (ptxas otherwise hoists 32 registers)
          double wv;
          asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(wv) : "l"(wperm + j * TH + tid));
          win[j] = (T)wv;
        } else
#endif
        win[j] = wperm[j * TH + tid];
      }
    }
    int fnext = 0;
    bool refilled = false;
    auto try_refill = [&](bool must) {                       // thread 0 only
      if (refilled) return;
      const uint32_t bar = ctrl_u32 + 32 + 8 * stg, par = (uint32_t)((it / NSTAGE) & 1);
      if (must) mbar_wait(bar, par); else if (!mbar_try_wait(bar, par)) return;
      fence_proxy_async();
      issue_stage(stg, fnext);
      refilled = true;
    };
    if constexpr (TDSA_WL_EARLY && NB == 1) { if (tid == 0) fnext = next_frame(); }
    TDSA_STAMP(0);
    if constexpr (FUSED) {
      if (it >= f_nit) break;
      f_it = it;
      TDSA_STAMP(10);
      if (it + 1 < f_nit) {
        // head of the next segment first; its ring slot was last read by the tails of segment it + 1 - kFusedRing
        if (it + 2 > kFusedRing) f_wait(f_tcnt, 16 * (it + 2 - kFusedRing));
        f_head((it + 1) % kFusedRing, it + 2 < f_nit ? (int64_t)f_grp + (int64_t)(it + 2) * f_ngrp : (int64_t)-1);
        f_head_pending = true;
      }
      TDSA_STAMP(15);
      f_wait(f_hcnt, 16 * (it + 1));                         // all sixteen column blocks of this segment are in the ring
    }
    if constexpr (!SUB) mbar_wait(ctrl_u32 + 8 * stg, (uint32_t)((it / NSTAGE) & 1));
    // written by thread 0 before it armed / completed the barrier (SUB: before the previous frame's Y barrier)
    const int f = FUSED ? 0 : (SUB ? slot[it & 1] : slot[stg]);
    if (!FUSED && f >= a.n_frames) break;
#ifdef TDSA_DEBUG_TIMING
    if (it == 0 && tid == 0 && a.dbg != nullptr) {
      unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      a.dbg[(int64_t)gridDim.x * 8 * 32 * 16 + blockIdx.x] = smid;
    }
#endif
    // per-frame scalars of the accumulating epilogue, requested now so that their latency is hidden by the transform
    // (142.1 -> 140.0 us for the running average over 8192 frames; the rest of its distance to the 126 us of a plain
    // accumulating launch are the three small kernels around it: weights, flag block, finish)
    double acc_wgt = 1.0;
    bool acc_live = true;
    if constexpr (ACC != 0 && !SUB) {
      if constexpr ((ACC & kAccSum) != 0 && (ACC & kAccGroup) == 0) { if (acc.weight != nullptr) acc_wgt = acc.weight[f]; }
      if (acc.skip != nullptr) acc_live = acc.skip[f] == 0;
    }
    T re[16], im[16];
    if constexpr (kWinTmem) {
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) win[j] = (T)__hiloint2double((int)wraw[2 * j + 1], (int)wraw[2 * j]);
    }
    // ---- pass A: staged samples -> registers, window, radix 16 over j (samples r + 16 c + 256 j) ----------
    {
      T dcr = T(0), dci = T(0);
      if constexpr (HAS_DC) { const double2 d = a.dc[f]; dcr = (T)d.x; dci = (T)d.y; }
      TDSA_STAMP(1);
      const unsigned char* src = stage_ptr + (size_t)stg * W::STAGE_BYTES + stage_off;
      float2 v[SUB ? 1 : 16];
      if constexpr (SUB) {
        const CT* y = FUSED ? f_ring + (int64_t)((it % kFusedRing) * 16 + f_cls) * 4096 + tid
                            : a.in_ct + (int64_t)f * 4096 + tid;     // [frame][j][tid]: 256 consecutive values per j
#pragma unroll
        for (int j = 0; j < 16; ++j) { const CT x = __ldcg(y + j * 256); re[j] = x.x; im[j] = x.y; }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = *reinterpret_cast<const float2*>(src + j * 2048);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if constexpr (HAS_DC) { re[j] = (T)v[j].x - dcr; im[j] = (T)v[j].y - dci; }
          else { re[j] = widen_sample<T>(v[j].x); im[j] = widen_sample<T>(v[j].y); }
        }
      }
      if constexpr (SUB) {
        dft16<T>(re, im);
      } else if constexpr (NB == 1) {
        dft16_win<T>(re, im, win);
      } else {
        // radix-2 DIF step on the staged read: s[n] = x[n] w[n] +- x[n + 4096] w[n + 4096] (the sign lives in the table)
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = *reinterpret_cast<const float2*>(src + 32768 + j * 2048);
        if constexpr (kWin2Tmem) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {               // eight samples of each frame half per step
            uint32_t ulo[16], uhi[16];
            tmem_ld16_nowait(tacc + kWin2Col + 16 * half, ulo);
            tmem_ld16_nowait(tacc + kWin2Col + 32 + 16 * half, uhi);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int j = 8 * half + i;
              const T wl = (T)__hiloint2double((int)ulo[2 * i + 1], (int)ulo[2 * i]);
              const T wh = (T)__hiloint2double((int)uhi[2 * i + 1], (int)uhi[2 * i]);
              const T hr = (T)v[j].x, hi = (T)v[j].y;
              re[j] = fm<T>(hr, wh, re[j] * wl);
              im[j] = fm<T>(hi, wh, im[j] * wl);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            T hr = (T)v[j].x, hi = (T)v[j].y;
            if constexpr (HAS_DC) { hr -= dcr; hi -= dci; }
            re[j] = fm<T>(hr, win[16 + j], re[j] * win[j]);
            im[j] = fm<T>(hi, win[16 + j], im[j] * win[j]);
          }
        }
        if (e == 0) {
          dft16<T>(re, im);
        } else {                                             // half-bin transform: inputs carry W32^j
          T cr[16], ci[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { cr[j] = w32_cos<T>(j); ci[j] = w32_msin<T>(j); }
          dft16_pretw<T>(re, im, cr, ci);
        }
      }
    }
    TDSA_STAMP(2);
    // ---- team-local 16x16 transpose: A_c[ka] at position c + 16 ka, row pitch 17 ---------------------------
    if constexpr (kSplitB2) {
      if (it > 0) mbar_wait(ctrl_u32 + 16 + 8 * e, (uint32_t)((it - 1) & 1));   // every warp of the engine has read the previous frame's Y
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) reg[c + 17 * q] = mk<T>(re[q], im[q]);
    __syncwarp();
    if constexpr (NB > 1 || TDSA_WL_EARLY) {                 // this warp's samples of the stage are consumed
      if (l == 0) mbar_arrive(ctrl_u32 + 32 + 8 * stg);
    }
    if constexpr (TDSA_WL_EARLY && NB == 1) { if (tid == 0) try_refill(false); }
    else if constexpr (FUSED) { }
    else if constexpr (SUB) { if (tid == 0) slot[(it + 1) & 1] = next_frame(); }   // read by everyone after the Y barrier
    else { if (tid == 0) fnext = next_frame(); }             // consumed after the barrier below
    TDSA_STAMP(3);
    // ---- pass B: thread ka = c reads A_j[ka] (j = 0..15), pre-twiddle [j][ka], radix 16 over j --------------
    {
      T wr[16], wi[16];
      wr[0] = T(1); wi[0] = T(0);
      if constexpr (TWB_BASE) {                              // powers of one root: six base twiddles held in registers
#pragma unroll
        for (int j = 1; j < 16; ++j) {
          if (j < 4 || (j & 3) == 0) { wr[j] = twbr[j]; wi[j] = twbi[j]; }
          else { wr[j] = twbr[j & 3]; wi[j] = twbi[j & 3]; cmul<T>(wr[j], wi[j], twbr[j & ~3], twbi[j & ~3]); }
        }
      } else if constexpr (TDSA_WL_TWB_SMEM_BASE && sizeof(T) == 8) {
        // six base twiddles from the shared table, the other nine by one complex multiply each (9 LDS.128 fewer, 36 DFMA more)
        if constexpr (TDSA_WL_TWB_SMEM_BASE == 2) {          // only w^1 and w^4 from the table, the other bases by squaring
          { const CT x = tws[1 * 16 + c]; wr[1] = x.x; wi[1] = x.y; }
          { const CT x = tws[4 * 16 + c]; wr[4] = x.x; wi[4] = x.y; }
          wr[2] = wr[1]; wi[2] = wi[1]; cmul<T>(wr[2], wi[2], wr[1], wi[1]);
          wr[3] = wr[2]; wi[3] = wi[2]; cmul<T>(wr[3], wi[3], wr[1], wi[1]);
          wr[8] = wr[4]; wi[8] = wi[4]; cmul<T>(wr[8], wi[8], wr[4], wi[4]);
          wr[12] = wr[8]; wi[12] = wi[8]; cmul<T>(wr[12], wi[12], wr[4], wi[4]);
        } else {
#pragma unroll
          for (int j = 1; j < 16; ++j) {
            if (j < 4 || (j & 3) == 0) { const CT x = tws[j * 16 + c]; wr[j] = x.x; wi[j] = x.y; }
          }
        }
#pragma unroll
        for (int j = 1; j < 16; ++j) {
          if (!(j < 4 || (j & 3) == 0)) { wr[j] = wr[j & 3]; wi[j] = wi[j & 3]; cmul<T>(wr[j], wi[j], wr[j & ~3], wi[j & ~3]); }
        }
      } else {
#pragma unroll
        for (int j = 1; j < 16; ++j) { const CT x = tws[j * 16 + c]; wr[j] = x.x; wi[j] = x.y; }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) { const CT x = reg[17 * c + j]; re[j] = x.x; im[j] = x.y; }
      __syncwarp();                                          // every lane of the team has read before anyone overwrites
      dft16_pretw<T>(re, im, wr, wi);
    }
    TDSA_STAMP(4);
    if constexpr (TDSA_WL_EARLY && NB == 1) { if (tid == 0) try_refill(false); }
    // Y_r[kk], kk = c + 16 kb, at position kk + (kk >> 4) = c + 17 kb of the team region
#pragma unroll
    for (int q = 0; q < 16; ++q) reg[c + 17 * q] = mk<T>(re[q], im[q]);
    TDSA_STAMP(5);
    if constexpr (FUSED) { if (f_head_pending) __threadfence(); }   // release the head pass' ring stores (performed by now)
    engine_sync();                                           // all sixteen Y_r complete; every warp has left this stage
    TDSA_STAMP(6);
    // Measured (round 2): letting the LAST warp to finish its staged reads refill the stage (shared-memory counter, frame
    // claimed at the top of the iteration) instead of thread 0 after this barrier was slower: 78.0 -> 83.9 us (f32),
    // 135.2 -> 139.4 us (f64) at N = 4096, 206.8 -> 219.1 us (f64) at N = 8192.
    if constexpr (FUSED) {                                    // every thread's ring loads were consumed before the barrier
      if (tid == 0) { if (f_head_pending) atomicAdd(f_hcnt, 1); atomicAdd(f_tcnt, 1); }
      f_head_pending = false;
    }
    if (tid == 0 && !SUB) {
      if constexpr (TDSA_WL_EARLY && NB == 1) {
        try_refill(true);
      } else {
        if constexpr (NB > 1) mbar_wait(ctrl_u32 + 32 + 8 * stg, (uint32_t)((it / NSTAGE) & 1));   // the other engine too
        fence_proxy_async();
        issue_stage(stg, fnext);
      }
    }
    // ---- last pass: thread kk = te reads Y_j[kk], pre-twiddle [j][kk], radix 16 over j ----------------------
    {
      const CT* col = ex + te + (te >> 4);
#pragma unroll
      for (int j = 0; j < 16; ++j) { const CT x = col[j * REGION]; re[j] = x.x; im[j] = x.y; }
    }
    if constexpr (kSplitB2) {
#if TDSA_WL_SPLIT_ALL
      mbar_arrive(ctrl_u32 + 16 + 8 * e);
#else
      __syncwarp();                                          // the warp's loads are ordered before its one arrival
      if (l == 0) mbar_arrive(ctrl_u32 + 16 + 8 * e);
#endif
    } else {
      engine_sync();                                         // regions may be overwritten by the next frame's pass A
    }
    TDSA_STAMP(7);
    {
      T wr[16], wi[16];
      wr[0] = T(1); wi[0] = T(0);
      if constexpr (TWMODE != 1 && TDSA_WL_TWL_BASE2 && sizeof(T) == 8) {
        wr[1] = twlr[1]; wi[1] = twli[1]; wr[4] = twlr[4]; wi[4] = twli[4];
        wr[2] = wr[1]; wi[2] = wi[1]; cmul<T>(wr[2], wi[2], wr[1], wi[1]);
        wr[3] = wr[2]; wi[3] = wi[2]; cmul<T>(wr[3], wi[3], wr[1], wi[1]);
        wr[8] = wr[4]; wi[8] = wi[4]; cmul<T>(wr[8], wi[8], wr[4], wi[4]);
        wr[12] = wr[8]; wi[12] = wi[8]; cmul<T>(wr[12], wi[12], wr[4], wi[4]);
#pragma unroll
        for (int j = 5; j < 16; ++j) {
          if ((j & 3) != 0) { wr[j] = wr[j & 3]; wi[j] = wi[j & 3]; cmul<T>(wr[j], wi[j], wr[j & ~3], wi[j & ~3]); }
        }
      } else {
#pragma unroll
        for (int j = 1; j < 16; ++j) {
          if constexpr (TWMODE == 1) { wr[j] = twlr[j]; wi[j] = twli[j]; }
          else {
            if (j < 4 || (j & 3) == 0) { wr[j] = twlr[j]; wi[j] = twli[j]; }
            else { wr[j] = twlr[j & 3]; wi[j] = twli[j & 3]; cmul<T>(wr[j], wi[j], twlr[j & ~3], twli[j & ~3]); }
          }
        }
      }
      dft16_pretw<T>(re, im, wr, wi);
    }
    TDSA_STAMP(8);
    // bin of output q: NB == 1: te + 256 q; NB == 2: engine e owns bins of parity e
    auto bin_of = [&](int q) { return NB == 1 ? te + 256 * q : 2 * (te + 256 * q) + e; };
    if constexpr (ACC == 0 || (ACC & kAccRows) != 0) {
      const bool store_row = (ACC == 0) || acc.only_row < 0 || acc.only_row == (int64_t)f;
      const int64_t row = (ACC != 0 && acc.only_row >= 0) ? 0 : (int64_t)f;
      auto emit = [&](auto mag_tag) {
        constexpr bool MAG = decltype(mag_tag)::value;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          if constexpr (TDSA_WL_EPI_F32SQ && ACC == 0 && sizeof(T) == 8 && !MAG && std::is_same<Epi, EpiDb>::value) {
            // narrow re and im first and square in float32: two conversions instead of DMUL + DFMA + one conversion on
            // the FP64 pipe; |X|^2 has no cancellation, so the result is within 2 ulp(float32) = 5e-7 dB
            const float fr = (float)re[q], fi = (float)im[q];
            Epi::template store<float, MAG>(a.ep, row, N, bin_of(q), __fmaf_rn(fr, fr, fi * fi));
          } else {
            const T pw = re[q] * re[q] + im[q] * im[q];
            if constexpr (ACC != 0) re[q] = pw;              // keep the power for the accumulators
            if (store_row) Epi::template store<T, MAG>(a.ep, row, N, bin_of(q), pw);
          }
        }
      };
      if (mag20) emit(std::true_type{}); else emit(std::false_type{});
    } else {
#pragma unroll
      for (int q = 0; q < 16; ++q) re[q] = re[q] * re[q] + im[q] * im[q];
    }
    if constexpr (ACC != 0) {
      // ---- accumulate |X|^2 of the thread's sixteen bins into its TMEM columns -------------------------------
      const bool live = acc_live;
      tmem_wait_st();                                        // the previous frame's updates have landed
      if (live) {
        if constexpr ((ACC & kAccSum) != 0) {
          const double wgt = acc_wgt;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t u[16];
            tmem_ld16(tacc + 16 * half, u);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const double s = __hiloint2double((int)u[2 * i + 1], (int)u[2 * i]);
              const double t = __fma_rn(wgt, (double)re[8 * half + i], s);
              u[2 * i] = (uint32_t)__double2loint(t); u[2 * i + 1] = (uint32_t)__double2hiint(t);
            }
            tmem_st16(tacc + 16 * half, u);
          }
        }
        if constexpr ((ACC & kAccMax) != 0) {
          uint32_t u[16];
          tmem_ld16(tacc + 32, u);
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] = __float_as_uint(fmaxf(__uint_as_float(u[i]), (float)re[i]));
          tmem_st16(tacc + 32, u);
        }
        if constexpr ((ACC & kAccMin) != 0) {
          uint32_t u[16];
          tmem_ld16(tacc + 48, u);
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] = __float_as_uint(fminf(__uint_as_float(u[i]), (float)re[i]));
          tmem_st16(tacc + 48, u);
        }
      }
      if constexpr ((ACC & kAccGroup) != 0) {
        if ((f + 1) % group == 0) {                          // last frame of its group: emit the mean as one dB row, clear
          tmem_wait_st();
          const int64_t g = (int64_t)(f / group);
          const double inv = 1.0 / (double)group;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t u[16];
            tmem_ld16(tacc + 16 * half, u);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const double s = __hiloint2double((int)u[2 * i + 1], (int)u[2 * i]);
              const int64_t at = bin_of(8 * half + i);
              u[2 * i] = 0u; u[2 * i + 1] = 0u;
              if (acc.unit_sum != nullptr) { acc.unit_sum[g * N + at] = s; continue; }
              const float db = to_db_m<double, false>(s * inv, a.ep);
              if (acc.n_peers == 0) {
                acc.group_db[g * N + at] = db;
              } else {
#pragma unroll 1
                for (int pr = 0; pr < acc.n_peers; ++pr) acc.peer_rows[pr][(acc.peer_row0 + g) * N + at] = db;
              }
            }
            tmem_st16(tacc + 16 * half, u);
          }
        }
      }
    }
    TDSA_STAMP(9);
  }
  if constexpr (ACC != 0) {
    // ---- flush the per-CTA partial rows, release the tensor memory ---------------------------------------------
    tmem_wait_st();
    auto bin_of = [&](int q) { return NB == 1 ? te + 256 * q : 2 * (te + 256 * q) + e; };
    const int64_t base = (int64_t)blockIdx.x * N;
    if constexpr ((ACC & kAccSum) != 0 && (ACC & kAccGroup) == 0) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t u[16];
        tmem_ld16(tacc + 16 * half, u);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          acc.part_sum[base + bin_of(8 * half + i)] = __hiloint2double((int)u[2 * i + 1], (int)u[2 * i]);
      }
    }
    if constexpr ((ACC & kAccMax) != 0) {
      uint32_t u[16];
      tmem_ld16(tacc + 32, u);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc.part_max[base + bin_of(i)] = __uint_as_float(u[i]);
    }
    if constexpr ((ACC & kAccMin) != 0) {
      uint32_t u[16];
      tmem_ld16(tacc + 48, u);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc.part_min[base + bin_of(i)] = __uint_as_float(u[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if constexpr (kUseTmem) {
    if constexpr (ACC == 0) { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); }
    if ((tid >> 5) == 0) tmem_dealloc<kTmemCols>(*tmem_slot);
  }
  // leave: the last CTA out re-arms the scheduler for the next launch
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(sched.done, 1) == (int)gridDim.x - 1) {
      *sched.next = 0;
      if constexpr (SUB) {
#pragma unroll
        for (int i = 0; i < 16; ++i) sched.next[2 + i] = 0;
      }
      if constexpr (FUSED) {                                 // the error word stays
        for (int i = 0; i < 2 * (int)(gridDim.x >> 4); ++i) acc.fused_cnt[i] = 0;
      }
      *sched.done = 0;
      __threadfence();
    }
  }
}

}  // namespace tdsa
