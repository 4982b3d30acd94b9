// tdsa_fft_wl.cuh — "warp-local" variant of the fused window + FFT + |.|^2 + dB kernel for N = 4096.
//
// Same arithmetic as fft_fused_kernel (tdsa_fft.cuh; reference datasources/rtl_samples.py:169-184), different
// schedule.  Phase time stamps of fft_fused_kernel (profiles/r01_phase_timing.md) showed that its three CTA-wide
// barriers per frame keep all eight warps of a CTA in the same phase, so the arithmetic pipe and the
// shared-memory pipe are used one after the other instead of together, and that the two CTAs of an SM run at
// very different speeds, which a static frame assignment turns into an idle tail.  Here:
//
//   * decimation in time over the FIRST factor: N = 16 x 256, n = r + 16 m.  The sixteen 256-point
//     sub-transforms Y_r = FFT256(x[r + 16 m] w[r + 16 m]) are each computed by a team of 16 lanes (two teams
//     per warp): radix-16 pass A, a 16x16 transpose through the team's private shared-memory region ordered by
//     __syncwarp only, radix-16 pass B.  Warps are independent of each other for two of the three passes.
//   * one CTA-wide exchange: X[kk + 256 q] = sum_r W16^(r q) W4096^(r kk) Y_r[kk]; thread kk reads the sixteen
//     Y_r[kk] (stride = region pitch, conflict-free) and stores bins kk + 256 q (coalesced), as before.
//   * the frame is staged by ONE cp.async.bulk.tensor (TMA, 3-D tensor map [frame][256 rows][128 B]) with the
//     128-byte swizzle, so that a team's stride-16 sample reads (row m, column r) are bank-conflict free.
//   * frames are claimed from a global counter (dynamic scheduling): the faster CTA of an SM simply takes
//     more frames; the counter re-arms itself when the last CTA leaves.
//
// Twiddle tables are the ones of the 4096-point DIT plan (pass 1: [j][K] = W256^(jK); last pass: [j][b]).
#pragma once
#include <cuda.h>

#include "tdsa_fft.cuh"

namespace tdsa {

template <typename T> struct WlPlan {
  static constexpr int N = 4096, TH = 256;
  static constexpr int REGION = 272 + (sizeof(T) == 4 ? 8 : 0);      // elements per team region (pitch-17 rows + bank offset)
  static constexpr int EX_ELEMS = 16 * REGION;
  static constexpr int TW_SMEM = 256;                                 // pass-B table [j][K]
  static constexpr size_t EX_BYTES = (size_t)(EX_ELEMS + TW_SMEM) * 2 * sizeof(T);
  static constexpr size_t STAGE_BYTES = (size_t)N * 8;
  static constexpr size_t CTRL_BYTES = 128;                           // mbarriers + frame slots
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

// L2 prefetch of one frame through the tensor map (no shared-memory destination, no completion)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"((uint64_t)map), "r"(c0), "r"(c1),
               "r"(c2)
               : "memory");
}

struct WlSched {
  int* next;       // next unclaimed frame
  int* done;       // CTAs that have left the frame loop
};

template <typename T, typename Epi, int TWMODE, int NSTAGE, bool HAS_DC, int MIN_CTAS, bool L2_AHEAD, bool TWB_BASE>
__global__ void __launch_bounds__(256, MIN_CTAS)
fft_wl_kernel(const FftArgs<T> a, const __grid_constant__ CUtensorMap tmap, const T* __restrict__ wperm, WlSched sched) {
  using W = WlPlan<T>;
  using CT = typename CplxOf<T>::type;
  constexpr int N = W::N, TH = W::TH, REGION = W::REGION;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  CT* ex = reinterpret_cast<CT*>(smem_raw);
  CT* tws = ex + W::EX_ELEMS;
  const uint32_t base_u32 = smem_u32(smem_raw);
  const uint32_t ctrl_u32 = base_u32 + (uint32_t)W::EX_BYTES;                    // NSTAGE mbarriers, then frame slots
  volatile int* slot = reinterpret_cast<volatile int*>(smem_raw + W::EX_BYTES + 64);
  const uint32_t stage_u32 = (base_u32 + (uint32_t)(W::EX_BYTES + W::CTRL_BYTES) + 1023u) & ~1023u;
  const unsigned char* stage_ptr = smem_raw + (stage_u32 - base_u32);

  const int tid = (int)threadIdx.x;
  const int w = tid >> 5, l = tid & 31;
  // lane -> (team half h, team lane c): each half-warp of lanes covers all eight 16-byte swizzle chunks and both
  // 8-byte halves, so the 64-bit staged reads are conflict free
  const int h = (l >> 3) & 1, c = (l & 7) + 8 * (l >> 4);
  const int r = 2 * w + h;                                  // sub-transform (team) 0..15
  CT* reg = ex + r * REGION;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) mbar_init(ctrl_u32 + 8 * s, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < W::TW_SMEM; i += TH) tws[i] = a.tw[i];

  // per-thread constants: window values of pass A (team identity) and last-pass pre-twiddles (bin identity tid)
  const CT* tw_last = a.tw + 256;                           // Plan<T,12>::tw_offset(2)
  T win[16];
  T twlr[16], twli[16];
  if constexpr (TWMODE == 1) {
#pragma unroll
    for (int j = 0; j < 16; ++j) win[j] = wperm[j * TH + tid];
#pragma unroll
    for (int j = 1; j < 16; ++j) { const CT x = tw_last[j * 256 + tid]; twlr[j] = x.x; twli[j] = x.y; }
  } else {
#pragma unroll
    for (int j = 1; j < 16; ++j) {
      if (j < 4 || (j & 3) == 0) { const CT x = tw_last[j * 256 + tid]; twlr[j] = x.x; twli[j] = x.y; }
    }
  }

  // pass-B twiddles W256^(j c) depend on the thread's team lane only: optionally six of them stay in registers
  T twbr[TWB_BASE ? 16 : 1], twbi[TWB_BASE ? 16 : 1];
  if constexpr (TWB_BASE) {
#pragma unroll
    for (int j = 1; j < 16; ++j) {
      if (j < 4 || (j & 3) == 0) { const CT x = a.tw[j * 16 + c]; twbr[j] = x.x; twbi[j] = x.y; }
    }
  }
  // claim the first NSTAGE frames and start their copies
  int pend = 0;                                             // thread 0: frame claimed one refill ahead (L2_AHEAD)
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      const int fs = atomicAdd(sched.next, 1);
      slot[s] = fs;
      if (fs < a.n_frames) {
        mbar_arrive_expect_tx(ctrl_u32 + 8 * s, (uint32_t)W::STAGE_BYTES);
        tma_load_3d(stage_u32 + (uint32_t)(s * W::STAGE_BYTES), &tmap, 0, 0, fs, ctrl_u32 + 8 * s);
      }
    }
    if constexpr (L2_AHEAD) {                               // the frame after those: claimed now, copied one iteration later
      pend = atomicAdd(sched.next, 1);
      if (pend < a.n_frames) tma_prefetch_3d(&tmap, 0, 0, pend);
    }
  }
  __syncthreads();
  const bool mag20 = a.ep.mode == kModeMag20;
  // swizzled offset of (row m = c + 16 j, column r) inside a staged frame: 128-byte rows, 16-byte chunk index
  // XOR (row & 7); r >> 1 == w, r & 1 == h, (c + 16 j) & 7 == c & 7
  const int stage_off = c * 128 + (((w ^ c) & 7) << 4) + h * 8;

  for (int it = 0;; ++it) {
    const int stg = it % NSTAGE;
    const int f = slot[stg];
    if (f >= a.n_frames) break;
#ifdef TDSA_DEBUG_TIMING
    if (it == 0 && tid == 0 && a.dbg != nullptr) {
      unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      a.dbg[(int64_t)gridDim.x * 8 * 32 * 16 + blockIdx.x] = smid;
    }
#endif
    TDSA_STAMP(0);
    T re[16], im[16];
    // ---- pass A: staged samples -> registers, window, radix 16 over j (samples r + 16 c + 256 j) ----------
    {
      if constexpr (TWMODE != 1) {
#pragma unroll
        for (int j = 0; j < 16; ++j) win[j] = wperm[j * TH + tid];
      }
      T dcr = T(0), dci = T(0);
      if constexpr (HAS_DC) { const double2 d = a.dc[f]; dcr = (T)d.x; dci = (T)d.y; }
      mbar_wait(ctrl_u32 + 8 * stg, (uint32_t)((it / NSTAGE) & 1));
      TDSA_STAMP(1);
      const unsigned char* src = stage_ptr + (size_t)stg * W::STAGE_BYTES + stage_off;
      float2 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = *reinterpret_cast<const float2*>(src + j * 2048);
      if constexpr (sizeof(T) == 8 && TDSA_INT_WIDEN == 2) {
        float probe = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          nonfinite_probe(v[j].x, probe); nonfinite_probe(v[j].y, probe);
          re[j] = widen_lean(v[j].x); im[j] = widen_lean(v[j].y);
        }
        if (probe != probe) re[0] = (T)probe;                // an Inf/NaN sample: the whole frame becomes NaN, as in the reference
        if constexpr (HAS_DC) {
#pragma unroll
          for (int j = 0; j < 16; ++j) { re[j] -= dcr; im[j] -= dci; }
        }
      } else if constexpr (sizeof(T) == 8 && TDSA_INT_WIDEN == 1) {
        // integer-pipe widening; the largest |bits| seen tells whether an Inf/NaN went through (exponent 0xFF)
        uint32_t top = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          re[j] = widen_track(v[j].x, top);
          im[j] = widen_track(v[j].y, top);
        }
        if (top >= 0x7f800000u) {                            // rare: redo from the staged frame with the hardware conversion
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 x = *reinterpret_cast<const float2*>(src + j * 2048);
            re[j] = (T)x.x; im[j] = (T)x.y;
          }
        }
        if constexpr (HAS_DC) {
#pragma unroll
          for (int j = 0; j < 16; ++j) { re[j] -= dcr; im[j] -= dci; }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if constexpr (HAS_DC) { re[j] = (T)v[j].x - dcr; im[j] = (T)v[j].y - dci; }
          else { re[j] = (T)v[j].x; im[j] = (T)v[j].y; }
        }
      }
      dft16_win<T>(re, im, win);
    }
    TDSA_STAMP(2);
    // ---- team-local 16x16 transpose: A_c[ka] at position c + 16 ka, row pitch 17 ---------------------------
#pragma unroll
    for (int q = 0; q < 16; ++q) reg[c + 17 * q] = mk<T>(re[q], im[q]);
    __syncwarp();
    int fnext = 0;
    if (tid == 0) fnext = atomicAdd(sched.next, 1);          // consumed after the CTA-wide barrier below
    TDSA_STAMP(3);
    // ---- pass B: thread ka = c reads A_j[ka] (j = 0..15), pre-twiddle W256^(j ka), radix 16 over j ----------
    {
      T wr[16], wi[16];
      wr[0] = T(1); wi[0] = T(0);
      if constexpr (TWB_BASE) {                              // W256^(j c) from six base twiddles held in registers
#pragma unroll
        for (int j = 1; j < 16; ++j) {
          if (j < 4 || (j & 3) == 0) { wr[j] = twbr[j]; wi[j] = twbi[j]; }
          else { wr[j] = twbr[j & 3]; wi[j] = twbi[j & 3]; cmul<T>(wr[j], wi[j], twbr[j & ~3], twbi[j & ~3]); }
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
    // Y_r[kk], kk = c + 16 kb, at position kk + (kk >> 4) = c + 17 kb of the team region
#pragma unroll
    for (int q = 0; q < 16; ++q) reg[c + 17 * q] = mk<T>(re[q], im[q]);
    TDSA_STAMP(5);
    __syncthreads();                                         // all sixteen Y_r complete; every warp has left this stage
    TDSA_STAMP(6);
    if (tid == 0) {
      // L2_AHEAD: copy the frame claimed (and prefetched into L2) one iteration ago, prefetch the one claimed now
      const int fcopy = L2_AHEAD ? pend : fnext;
      slot[stg] = fcopy;
      if (fcopy < a.n_frames) {
        fence_proxy_async();
        mbar_arrive_expect_tx(ctrl_u32 + 8 * stg, (uint32_t)W::STAGE_BYTES);
        tma_load_3d(stage_u32 + (uint32_t)(stg * W::STAGE_BYTES), &tmap, 0, 0, fcopy, ctrl_u32 + 8 * stg);
      }
      if constexpr (L2_AHEAD) {
        pend = fnext;
        if (pend < a.n_frames) tma_prefetch_3d(&tmap, 0, 0, pend);
      }
    }
    // ---- last pass: thread kk = tid reads Y_j[kk], pre-twiddle W4096^(j kk), radix 16 over j ----------------
    {
      const CT* col = ex + tid + (tid >> 4);
#pragma unroll
      for (int j = 0; j < 16; ++j) { const CT x = col[j * REGION]; re[j] = x.x; im[j] = x.y; }
    }
    __syncthreads();                                         // regions may be overwritten by the next frame's pass A
    TDSA_STAMP(7);
    {
      T wr[16], wi[16];
      wr[0] = T(1); wi[0] = T(0);
#pragma unroll
      for (int j = 1; j < 16; ++j) {
        if constexpr (TWMODE == 1) { wr[j] = twlr[j]; wi[j] = twli[j]; }
        else {
          if (j < 4 || (j & 3) == 0) { wr[j] = twlr[j]; wi[j] = twli[j]; }
          else { wr[j] = twlr[j & 3]; wi[j] = twli[j & 3]; cmul<T>(wr[j], wi[j], twlr[j & ~3], twli[j & ~3]); }
        }
      }
      dft16_pretw<T>(re, im, wr, wi);
    }
    TDSA_STAMP(8);
    auto emit = [&](auto mag_tag) {
      constexpr bool MAG = decltype(mag_tag)::value;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const T pw = re[q] * re[q] + im[q] * im[q];
        Epi::template store<T, MAG>(a.ep, (int64_t)f, N, tid + 256 * q, pw);
      }
    };
    if (mag20) emit(std::true_type{}); else emit(std::false_type{});
    TDSA_STAMP(9);
  }
  // leave: the last CTA out re-arms the scheduler for the next launch
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(sched.done, 1) == (int)gridDim.x - 1) {
      *sched.next = 0;
      *sched.done = 0;
      __threadfence();
    }
  }
}

}  // namespace tdsa
