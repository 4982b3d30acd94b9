// tdsa_welch_cluster.cuh — Welch averaging + peak hold of 65536-point segments in ONE kernel (BASELINE config 3).
//
// Reference arithmetic per segment: datasources/rtl_samples.py:169-184 (window, FFT, fftshift, |.|^2), averaged
// as utils/signal_processing.py:56-59 ('lin' running mean over all segments) and peak-held as
// core/display_data_processor.py:371-382 (np.fmax in dB); tdsa_welch() in tdsa_api.cu is the entry point.
//
// The two-kernel path (big_head1_kernel -> scratch -> fft_fused_kernel<TAIL> -> linear rows -> welch_reduce_kernel)
// moves ~48 bytes per point through L2/HBM (complex scratch written and re-read, float64 rows written and re-read).
// Here a thread-block CLUSTER of 16 CTAs owns a segment and nothing but the IQ samples ever leaves the SMs:
//
//   * N = 16 x 4096 decimation in frequency: CTA rank rho owns columns c in [256 rho, 256 rho + 256); thread c reads
//     x[c + 4096 j] (j = 0..15, coalesced), applies the window, runs one radix-16 and the post-twiddle W_N^(c q),
//     and stores output q straight into the shared memory of CTA q (distributed shared memory, st.shared::cluster):
//     sub-transform q of the segment assembles in CTA q without touching global memory.
//   * one barrier.cluster per segment; the receive buffers are double buffered, so the head pass of segment i+1
//     overlaps the other CTAs' tail of segment i.
//   * every CTA then runs the 4096-point transform of its sub-sequence in place in shared memory (the three
//     radix-16 passes and tables of the 4096-point plan) and gets bins k = q + 16 kl.
//   * sum |X|^2 (float64) and max (float)|X|^2 per bin stay in REGISTERS for the whole kernel (16 bins per thread);
//     the clusters' partial rows are combined by welch_cluster_finish_kernel, which also takes the dB.
//     (10 log10 is monotonic, so the maximum of the dB values is the dB of the maximum.)
#pragma once
#include "tdsa_fft.cuh"

namespace tdsa {

#ifndef TDSA_WC_DIAG
#define TDSA_WC_DIAG 0
#endif
constexpr int kWcCluster = 16;                 // CTAs per cluster = sub-transforms per segment
constexpr int kWcLog2N = 16;

template <typename T> struct WelchClusterArgs {
  const float2* iq;                             // flat complex64 stream
  int64_t n_seg;
  int64_t hop;                                  // samples between segment starts
  const T* window;                              // T[65536], (-1)^n folded in
  const typename CplxOf<T>::type* tw_head;      // [q][c], c < 4096: W_65536^(c q)
  const typename CplxOf<T>::type* tw_inner;     // 4096-point DIT plan tables (pass 1 [j][K], last pass [j][b])
  double* part_sum;                             // [clusters][65536]
  float* part_peak;                             // [clusters][65536], max of (float)|X|^2
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float2 v) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_cluster(uint32_t addr, double2 v) {
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

template <typename T> struct WelchClusterPlan {
  using P = Plan<T, 12, 4>;
  static constexpr size_t BUF_BYTES = (((size_t)P::PHYS_SIZE * 2 * sizeof(T)) + 127) & ~(size_t)127;
  static constexpr size_t TW_BYTES = (size_t)256 * 2 * sizeof(T);
  static constexpr size_t WIN_BYTES = (size_t)16 * 256 * sizeof(T);   // this CTA's window columns, [j][t]
  static constexpr size_t SMEM_BYTES = 2 * BUF_BYTES + TW_BYTES + WIN_BYTES;
};

// base twiddles W^(x*{1,2,3,4,8,12}) -> all fifteen W^(x*q) (one complex multiply each for the other nine)
template <typename T> __device__ __forceinline__ void expand_base(const T* br, const T* bi, T* wr, T* wi) {
  wr[0] = T(1); wi[0] = T(0);
#pragma unroll
  for (int q = 1; q < 16; ++q) {
    if (q < 4 || (q & 3) == 0) { wr[q] = br[q]; wi[q] = bi[q]; }
    else { wr[q] = br[q & 3]; wi[q] = bi[q & 3]; cmul<T>(wr[q], wi[q], br[q & ~3], bi[q & ~3]); }
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 1) welch_cluster_kernel(const WelchClusterArgs<T> a) {
  using P = Plan<T, 12, 4>;
  using CT = typename CplxOf<T>::type;
  using W = WelchClusterPlan<T>;
  constexpr int M = 4096;
  extern __shared__ __align__(128) unsigned char wc_smem[];
  unsigned char* smem_raw = wc_smem;
  CT* tws = reinterpret_cast<CT*>(smem_raw + 2 * W::BUF_BYTES);
  T* wins = reinterpret_cast<T*>(smem_raw + 2 * W::BUF_BYTES + W::TW_BYTES);
  const uint32_t buf_u32 = smem_u32(smem_raw);
  const int t = (int)threadIdx.x;
  const uint32_t rho = cluster_ctarank();
  const int c = 256 * (int)rho + t;                        // this thread's column of the head pass

  for (int i = t; i < 256; i += 256) tws[i] = a.tw_inner[i];
#pragma unroll
  for (int j = 0; j < 16; ++j) wins[j * 256 + t] = a.window[c + j * M];     // constant across segments
  // base twiddles of the head post-multiply W_N^(c q) and of the last inner pass W_4096^(j t)
  T hbr[16], hbi[16], lbr[16], lbi[16];
  const CT* tw_last = a.tw_inner + 256;
#pragma unroll
  for (int q = 1; q < 16; ++q) {
    if (q < 4 || (q & 3) == 0) {
      const CT x = a.tw_head[q * M + c]; hbr[q] = x.x; hbi[q] = x.y;
      const CT y = tw_last[q * 256 + t]; lbr[q] = y.x; lbi[q] = y.y;
    }
  }
  // where output q of this thread lands in CTA q: position c of its receive buffer
  const uint32_t my_slot = (uint32_t)(P::phys(c) * sizeof(CT));
  uint32_t remote[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) remote[q] = map_to_cta(buf_u32 + my_slot, (uint32_t)q);

  double sum[16];
  float peak[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) { sum[q] = 0.0; peak[q] = -INFINITY; }

  const int64_t cid = cluster_id_x(), ncl = cluster_count_x();
  float2 v[16];
  if (cid < a.n_seg) {
    const float2* src = a.iq + cid * a.hop + c;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = ldg_stream(src + j * M);
  }
  cluster_arrive();                                        // every CTA of the cluster is resident before remote stores
  cluster_wait();
  __syncthreads();                                         // window columns and table staged

  // head of segment `seg` (samples already in v[]) into receive buffer `buf` of every CTA; then prefetch seg + ncl
  auto head = [&](int64_t seg, int buf) {
    const uint32_t boff = (uint32_t)(buf * W::BUF_BYTES);
    T re[16], im[16];
    {
      T win[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) win[j] = wins[j * 256 + t];
#pragma unroll
      for (int j = 0; j < 16; ++j) { re[j] = (T)v[j].x; im[j] = (T)v[j].y; }
      dft16_win<T>(re, im, win);
    }
    {
      T wr[16], wi[16];
      expand_base<T>(hbr, hbi, wr, wi);
#pragma unroll
      for (int q = 1; q < 16; ++q) cmul<T>(re[q], im[q], wr[q], wi[q]);
    }
#if TDSA_WC_DIAG == 1      // diagnostic (wrong results): keep the scatter local, to price the SM-to-SM traffic
    CT* exl = reinterpret_cast<CT*>(smem_raw + boff);
#pragma unroll
    for (int q = 0; q < 16; ++q) exl[P::phys(t) + P::phys(q * 256)] = mk<T>(re[q], im[q]);
#else
#pragma unroll
    for (int q = 0; q < 16; ++q) st_cluster(remote[q] + boff, mk<T>(re[q], im[q]));
#endif
    const int64_t nseg = seg + ncl;
    if (nseg < a.n_seg) {
      const float2* src = a.iq + nseg * a.hop + c;
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = ldg_stream(src + j * M);
    }
  };

  // Order per CTA: H(0) | barrier | { H(i+1) ; T(i) ; barrier }.  The head of the NEXT segment is issued before the
  // tail of the current one, so its SM-to-SM stores drain while the tail computes; the barrier after the tail both
  // publishes H(i+1) and tells everyone that buffer (i & 1) has been consumed (H(i+2) may overwrite it).
  if (cid < a.n_seg) head(cid, 0);
  cluster_arrive();
  cluster_wait();
  int it = 0;
  for (int64_t seg = cid; seg < a.n_seg; seg += ncl, ++it) {
    if (seg + ncl < a.n_seg) head(seg + ncl, (it + 1) & 1);
    CT* ex = reinterpret_cast<CT*>(smem_raw + (size_t)(it & 1) * W::BUF_BYTES);
    T re[16], im[16];
    // ---- tail: 4096-point transform in place (DIT plan of Plan<T,12>: passes 16 x 16 x 16) --------------------
#if TDSA_WC_DIAG == 2      // diagnostic (wrong results): no tail passes, to price head + exchange + barrier alone
    if (a.n_seg < 0)
#endif
    {
    {                                                      // pass 0: x[t + 256 j], no twiddles
      const int pb = P::phys(t);
#pragma unroll
      for (int j = 0; j < 16; ++j) { const CT x = ex[pb + P::phys(j * 256)]; re[j] = x.x; im[j] = x.y; }
      dft16<T>(re, im);
#pragma unroll
      for (int q = 0; q < 16; ++q) ex[pb + P::phys(q * 256)] = mk<T>(re[q], im[q]);
    }
    __syncthreads();
    {                                                      // pass 1: blocks of 256, stride 16, pre-twiddle [j][K]
      const int c1 = t & 15, s = t >> 4;
      const int pb = P::phys(s * 256 + c1);
      T wr[16], wi[16];
      wr[0] = T(1); wi[0] = T(0);
#pragma unroll
      for (int j = 1; j < 16; ++j) { const CT x = tws[j * 16 + s]; wr[j] = x.x; wi[j] = x.y; }
#pragma unroll
      for (int j = 0; j < 16; ++j) { const CT x = ex[pb + P::phys(j * 16)]; re[j] = x.x; im[j] = x.y; }
      dft16_pretw<T>(re, im, wr, wi);
#pragma unroll
      for (int q = 0; q < 16; ++q) ex[pb + P::phys(q * 16)] = mk<T>(re[q], im[q]);
    }
    __syncthreads();
    {                                                      // pass 2: digit-reversed reads, pre-twiddle W_4096^(j t)
      const int pb = P::phys(digitrev<4>(t, 2) * 16);
#pragma unroll
      for (int j = 0; j < 16; ++j) { const CT x = ex[pb + P::phys(j)]; re[j] = x.x; im[j] = x.y; }
      T wr[16], wi[16];
      expand_base<T>(lbr, lbi, wr, wi);
      dft16_pretw<T>(re, im, wr, wi);
    }
    }
    // bins kl = t + 256 q2 of sub-transform rho  ->  bin rho + 16 kl of the segment
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const T pw = re[q] * re[q] + im[q] * im[q];
      sum[q] += (double)pw;
      peak[q] = fmaxf(peak[q], (float)pw);                 // NaN-ignoring, like np.fmax
    }
    cluster_arrive();
    cluster_wait();
  }
  // partial rows of this cluster
  {
    double* ps = a.part_sum + cid * 65536;
    float* pp = a.part_peak + cid * 65536;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int k = (int)rho + 16 * (t + 256 * q);
      ps[k] = sum[q];
      pp[k] = peak[q];
    }
  }
  cluster_arrive();                                        // nobody leaves while a neighbour may still write here
  cluster_wait();
}

// avg_db[k] = dB(sum over clusters / n_seg), peak_db[k] = dB(max over clusters); rows are already fftshift-ed
static __global__ void __launch_bounds__(256) welch_cluster_finish_kernel(const double* __restrict__ part_sum,
                                                                  const float* __restrict__ part_peak, int clusters,
                                                                  int64_t n_seg, double scale, double floor,
                                                                  float* __restrict__ avg_db, float* __restrict__ peak_db) {
  const int k = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (k >= 65536) return;
  double s = 0.0;
  float pk = -INFINITY;
  for (int cl = 0; cl < clusters; ++cl) {
    s += part_sum[(size_t)cl * 65536 + k];
    pk = fmaxf(pk, part_peak[(size_t)cl * 65536 + k]);
  }
  EpiParams ep;
  ep.db_out = nullptr; ep.lin_out = nullptr; ep.scale = scale; ep.floor = floor; ep.mode = kModePower;
  avg_db[k] = to_db_m<double, false>(s / (double)n_seg, ep);
  const float kDbPerLog2 = 3.01029995663981195f;
  peak_db[k] = kDbPerLog2 * lg2_approx(__fmaf_rn(pk, (float)scale, (float)floor));
}

}  // namespace tdsa
