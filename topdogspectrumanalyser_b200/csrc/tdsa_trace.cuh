// tdsa_trace.cuh — trace state kept entirely on the device: TraceAverager count and the hold / tare flags live in a
// small int32 block next to the float64 buffer, so that no entry point has to read anything back to the host.
//
// Reference semantics: utils/signal_processing.py:35-61 (TraceAverager), core/display_data_processor.py:211-218
// (sweep-domain averaging), :317-369 (cal offset, tare), :371-395 + :473-480 (max/min hold, _nan_safe),
// datasources/hackrf_samples.py:351-355 (silent frames repeat the last good row).
#pragma once
#include "tdsa_aux.cuh"

namespace tdsa {

// layout of the device flag block (include/tdsa.h: TDSA_FLAG_*)
enum : int { kFlagCount = 0, kFlagMaxValid = 1, kFlagMinValid = 2, kFlagLive = 3, kFlagTareCollecting = 4,
             kFlagTareActive = 5, kFlagTareCount = 6, kFlagWords = 8 };

// ---------------------------------------------------------------------------------------------------------------
// Running average as a weighted sum (fused path).  TraceAverager folds frames x_1..x_L into its buffer s as
//   exp:  s <- (1 - a) s + a x_i, a = 1/n                 (first frame of an empty buffer: s <- x_1)
//   lin:  c_i = min(c_{i-1} + 1, n);  s <- s + (x_i - s)/c_i
// which is  s_L = carry * s_0 + sum_i w_i x_i  with weights that depend only on (mode, n, c_0, L, i):
//   exp:  w_i = a (1-a)^(L-i)   (w_1 = (1-a)^(L-1) when c_0 = 0);  carry = (1-a)^L  (0 when c_0 = 0)
//   lin:  P_i = prod_{u>i} (1 - 1/c_u) = [c_i / c_J for the uncapped steps, J = min(L, n - c_0)] * (1 - 1/n)^(capped steps),
//         w_i = P_i / c_i,  carry = P_0 (0 when c_0 = 0, because c_1 = 1)
// Frames can then be folded in any order by any CTA.  meta: {carry, c_0, new count}.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double lin_tail_product(int64_t i, int64_t L, int64_t c0, int64_t n) {
  // prod_{u = i+1 .. L} (1 - 1/c_u), c_u = min(c0 + u, n); i >= 0; for i = 0 with c0 = 0 the first factor is 0
  const int64_t U = n > c0 ? n - c0 : 0;          // steps u <= U are uncapped (c_u = c0 + u)
  const int64_t J = L < U ? L : U;
  double p = 1.0;
  if (i < J) {                                    // telescoping part: prod_{u=i+1..J} (c_u - 1)/c_u = c_i / c_J
    if (c0 + i == 0) return 0.0;
    p = (double)(c0 + i) / (double)(c0 + J);
  }
  const int64_t from = i > U ? i : U;             // capped steps u = from+1 .. L
  if (L > from) p *= pow(1.0 - 1.0 / (double)n, (double)(L - from));
  return p;
}

__global__ void __launch_bounds__(256) avg_weights_kernel(int avg_mode, int avg_n, int64_t n_frames,
                                                         int32_t* __restrict__ flags, double* __restrict__ weight,
                                                         double* __restrict__ meta) {
  const int64_t c0 = flags[kFlagCount];
  const int64_t L = n_frames, n = avg_n;
  const double a = 1.0 / (double)avg_n;
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_frames; f += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = f + 1;
    double w;
    if (avg_mode == 1) w = (c0 == 0 && i == 1) ? pow(1.0 - a, (double)(L - 1)) : a * pow(1.0 - a, (double)(L - i));
    else w = lin_tail_product(i, L, c0, n) / (double)((c0 + i) < n ? (c0 + i) : n);
    weight[f] = w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double carry;
    int64_t cnew;
    if (avg_mode == 1) { carry = c0 == 0 ? 0.0 : pow(1.0 - a, (double)L); cnew = c0 > 1 ? c0 : 1; }
    else { carry = lin_tail_product(0, L, c0, n); cnew = (c0 + L) < n ? (c0 + L) : n; }
    meta[0] = carry; meta[1] = (double)c0; meta[2] = (double)cnew;
  }
}

// snapshot of the flag block for kernels that follow, then the flags as they will be after `live` frames
__global__ void flags_prepare_kernel(int32_t* __restrict__ flags, double* __restrict__ meta, int avg_mode, int avg_n,
                                     int64_t live, int has_max, int has_min) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  meta[1] = (double)flags[kFlagCount];
  meta[3] = (double)flags[kFlagMaxValid];
  meta[4] = (double)flags[kFlagMinValid];
  flags[kFlagLive] = (int32_t)live;
  if (live > 0) {
    if (has_max) flags[kFlagMaxValid] = 1;
    if (has_min) flags[kFlagMinValid] = 1;
    (void)avg_mode; (void)avg_n;
  }
}

// Per-CTA partial rows [n_parts][width] are combined by blocks of 256 threads = 32 neighbouring bins x 8 slices of the
// partial rows (coalesced 32-bin reads, 8-way parallel walk down the rows, shared-memory combine in slice order, so
// the float64 sums are deterministic).  Launch with partial_grid(width) blocks; thread (slice 0, bin) finishes the bin.
constexpr int kPartBins = 32, kPartSlices = 8;
inline int partial_grid(int64_t width) { return (int)((width + kPartBins - 1) / kPartBins); }

template <typename V, typename Op>
__device__ __forceinline__ V reduce_partials(const V* __restrict__ part, int n_parts, int64_t width, V init, Op op, bool* owner,
                                             int64_t* bin) {
  __shared__ double s_buf[kPartSlices][kPartBins];
  const int kb = threadIdx.x & (kPartBins - 1), sl = threadIdx.x / kPartBins;
  const int64_t k = (int64_t)blockIdx.x * kPartBins + kb;
  V v = init;
  if (k < width)
    for (int c = sl; c < n_parts; c += kPartSlices) v = op(v, part[(int64_t)c * width + k]);
  __syncthreads();                                    // the buffer may still be read by a previous call's owner threads
  s_buf[sl][kb] = (double)v;
  __syncthreads();
  *owner = sl == 0 && k < width;
  *bin = k;
  if (sl == 0)
    for (int i = 1; i < kPartSlices; ++i) v = op(v, (V)s_buf[i][kb]);
  return v;
}

// fused running average, last step: state <- carry * state + scale * sum over CTAs; one dB row out
__global__ void __launch_bounds__(256) avg_finish_kernel(const double* __restrict__ part_sum, int n_parts, int64_t width,
                                                        const double* __restrict__ meta, double scale, double floor, int mode,
                                                        double* __restrict__ avg_state, int32_t* __restrict__ flags,
                                                        int64_t live, float* __restrict__ db_out, float* __restrict__ last_row) {
  bool owner; int64_t k;
  double s = reduce_partials<double>(part_sum, n_parts, width, 0.0, [](double a, double b) { return a + b; }, &owner, &k);
  if (!owner) return;
  s *= scale;
  const double carry = meta[0];
  const double v = carry != 0.0 ? __fma_rn(carry, avg_state[k], s) : s;
  avg_state[k] = v;
  EpiParams ep;
  ep.db_out = nullptr; ep.lin_out = nullptr; ep.scale = 1.0; ep.floor = floor; ep.mode = mode;
  const float db = to_db<double>(v, ep);
  db_out[k] = db;
  if (last_row) last_row[k] = db;
  if (k == 0) { flags[kFlagCount] = (int32_t)meta[2]; flags[kFlagLive] = (int32_t)live; }
}

// fused max/min hold on un-averaged rows: hold <- fmax(hold, dB(max_f |X_f|^2)) (dB is monotonic), first use initialises
__global__ void __launch_bounds__(256) hold_finish_kernel(const float* __restrict__ part_max, const float* __restrict__ part_min,
                                                         int n_parts, int64_t width, const double* __restrict__ meta,
                                                         double scale, double floor, int mode, float* __restrict__ max_hold,
                                                         float* __restrict__ min_hold) {
  EpiParams ep;
  ep.db_out = nullptr; ep.lin_out = nullptr; ep.scale = scale; ep.floor = floor; ep.mode = mode;
  if (max_hold) {
    bool owner; int64_t k;
    const float m = reduce_partials<float>(part_max, n_parts, width, -INFINITY, [](float a, float b) { return fmaxf(a, b); }, &owner, &k);
    if (owner) {
      const bool nan_only = (m == -INFINITY);                        // every live frame had NaN in this bin
      const float db = nan_only ? -500.0f : to_db<float>(m, ep);     // _nan_safe: NaN -> -500 on the initialising frame
      if (meta[3] != 0.0) { if (!nan_only) max_hold[k] = fmaxf(max_hold[k], db); }
      else max_hold[k] = db;
    }
  }
  if (min_hold) {
    bool owner; int64_t k;
    const float m = reduce_partials<float>(part_min, n_parts, width, INFINITY, [](float a, float b) { return fminf(a, b); }, &owner, &k);
    if (owner) {
      const bool nan_only = (m == INFINITY);
      const float db = nan_only ? 500.0f : to_db<float>(m, ep);
      if (meta[4] != 0.0) { if (!nan_only) min_hold[k] = fminf(min_hold[k], db); }
      else min_hold[k] = db;
    }
  }
}

// Welch (config 3) from the accumulating epilogue: avg = dB(scale * sum / n_seg), peak = dB(scale * max)
__global__ void __launch_bounds__(256) welch_acc_finish_kernel(const double* __restrict__ part_sum, const float* __restrict__ part_max,
                                                              int n_parts, int64_t width, int64_t n_seg, double scale,
                                                              double floor, int mode, float* __restrict__ avg_db,
                                                              float* __restrict__ peak_db) {
  bool owner, owner2; int64_t k, k2;
  const double s = reduce_partials<double>(part_sum, n_parts, width, 0.0, [](double a, double b) { return a + b; }, &owner, &k);
  const float m = reduce_partials<float>(part_max, n_parts, width, -INFINITY, [](float a, float b) { return fmaxf(a, b); }, &owner2, &k2);
  if (!owner) return;
  EpiParams ep;
  ep.db_out = nullptr; ep.lin_out = nullptr; ep.scale = 1.0; ep.floor = floor; ep.mode = mode;
  avg_db[k] = to_db<double>(s * scale / (double)n_seg, ep);
  ep.scale = scale;
  peak_db[k] = to_db<float>(m, ep);
}

__global__ void __launch_bounds__(256) welch_sub_clear_kernel(double* __restrict__ part_sum, float* __restrict__ part_max, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) { part_sum[i] = 0.0; part_max[i] = -INFINITY; }
}

// Welch over 65536-point segments whose 4096-point tails accumulated per class (fft_wl_kernel, kAccSub): partial row b
// holds sub-transform s = b % 16, whose local bin kl is bin s + 16 kl of the segment spectrum
__global__ void __launch_bounds__(256) welch_sub_finish_kernel(const double* __restrict__ part_sum, const float* __restrict__ part_max,
                                                              int n_parts, int64_t n_seg, double scale, double floor, int mode,
                                                              float* __restrict__ avg_db, float* __restrict__ peak_db,
                                                              const int* __restrict__ err) {
  const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (t >= 65536) return;
  if (err != nullptr && *err != 0) {   // the fused kernel gave up waiting for its group (CTAs not co-resident): no result
    avg_db[t] = __int_as_float(0x7fc00000); peak_db[t] = __int_as_float(0x7fc00000);
    return;
  }
  const int kl = t & 4095, s = t >> 12;
  double sum = 0.0;
  float m = -INFINITY;
  for (int b = s; b < n_parts; b += 16) {
    sum += part_sum[(int64_t)b * 4096 + kl];
    m = fmaxf(m, part_max[(int64_t)b * 4096 + kl]);
  }
  EpiParams ep;
  ep.db_out = nullptr; ep.lin_out = nullptr; ep.scale = 1.0; ep.floor = floor; ep.mode = mode;
  const int k = s + 16 * kl;
  avg_db[k] = to_db<double>(sum * scale / (double)n_seg, ep);
  ep.scale = scale;
  peak_db[k] = to_db<float>(m, ep);
}

// Config-4 rows from split groups: each group's `split` unit sums are added, divided by the frames of the group and
// converted to one dB row, stored locally (group_db) or straight into every rank's row table (peers, NVLink stores).
struct GroupFinishArgs {
  const double* unit_sum;   // [n_groups * split][W]
  int64_t n_groups, width;
  int split, frames;
  double scale, floor;
  int mode;
  float* group_db;          // [n_groups][W] or null
  float* peer_rows[8];
  int n_peers;
  int64_t peer_row0;
};
__global__ void __launch_bounds__(256) group_finish_kernel(const GroupFinishArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_groups * a.width) return;
  const int64_t g = i / a.width, k = i - g * a.width;
  double s = 0.0;
  for (int u = 0; u < a.split; ++u) s += a.unit_sum[(g * a.split + u) * a.width + k];
  EpiParams ep;
  ep.db_out = nullptr; ep.lin_out = nullptr; ep.scale = a.scale; ep.floor = a.floor; ep.mode = a.mode;
  const float db = to_db_m<double, false>(s / (double)a.frames, ep);
  if (a.n_peers == 0) a.group_db[i] = db;
  else
    for (int pr = 0; pr < a.n_peers; ++pr) a.peer_rows[pr][(a.peer_row0 + g) * a.width + k] = db;
}

// ---------------------------------------------------------------------------------------------------------------
// General path: frame-ordered scan over float64 linear rows (one thread per bin), flags read from the device.
// The rows of a chunk are written by the FFT kernel and consumed here while still in L2 (the caller sizes chunks).
// ---------------------------------------------------------------------------------------------------------------
struct TraceScanDevArgs {
  const double* lin;       // [F][W] linear power (already scaled for PSD)
  const int32_t* skip;     // [F] or null
  int64_t n_frames, width;
  int avg_mode, avg_n;
  const int32_t* flags;    // device flag block (read only here; trace_flags_after_scan_kernel updates it)
  double* avg_state;
  float* max_hold;
  float* min_hold;
  float* last_row;         // [W] or null: last good dB row, carried across launches (silent frames repeat it)
  int last_only;
  float* db_out;           // [F][W], or [W] when last_only
  double floor;
  int mode;
};

__global__ void __launch_bounds__(256) trace_scan_dev_kernel(const TraceScanDevArgs a) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.width) return;
  const bool averaging = a.avg_mode != 0 && a.avg_n > 1;
  int count = averaging ? a.flags[kFlagCount] : 0;
  double buf = (averaging && count > 0) ? a.avg_state[k] : 0.0;
  const double alpha = 1.0 / (double)a.avg_n;
  bool mxv = a.flags[kFlagMaxValid] != 0, mnv = a.flags[kFlagMinValid] != 0;
  float mx = (a.max_hold && mxv) ? a.max_hold[k] : 0.f;
  float mn = (a.min_hold && mnv) ? a.min_hold[k] : 0.f;
  EpiParams ep;
  ep.db_out = nullptr; ep.lin_out = nullptr; ep.scale = 1.0; ep.floor = a.floor; ep.mode = a.mode;
  float last_db = a.last_row ? a.last_row[k] : 0.0f;     // what a skipped frame repeats (hackrf_samples.py:351-355)
  bool any = false;
  constexpr int kAhead = 32;                             // rows fetched ahead of the recurrence (they come from L2; one
                                                         // warp per SM runs this kernel, so latency is hidden by depth only)
  for (int64_t f0 = 0; f0 < a.n_frames; f0 += kAhead) {
    double pre[kAhead];
#pragma unroll
    for (int u = 0; u < kAhead; ++u) pre[u] = (f0 + u < a.n_frames) ? a.lin[(f0 + u) * a.width + k] : 0.0;
#pragma unroll
    for (int u = 0; u < kAhead; ++u) {
      const int64_t f = f0 + u;
      if (f >= a.n_frames) break;
      if (a.skip != nullptr && a.skip[f]) {
        if (!a.last_only) a.db_out[f * a.width + k] = last_db;
        else if (f == a.n_frames - 1) a.db_out[k] = last_db;
        continue;
      }
      const double p = pre[u];
      double v = p;
      if (averaging) {
        if (count == 0) {
          buf = p;
          count = 1;
        } else if (a.avg_mode == 1) {
          buf = __dmul_rn(buf, 1.0 - alpha);              // buffer *= (1 - alpha)
          buf = __dadd_rn(buf, __dmul_rn(alpha, p));      // buffer += alpha * x
        } else {
          if (count < a.avg_n) ++count;
          buf = __dadd_rn(buf, __ddiv_rn(__dsub_rn(p, buf), (double)count));   // buffer += (x - buffer)/count
        }
        v = buf;
      }
      const float db = to_db<double>(v, ep);
      last_db = db;
      any = true;
      if (!a.last_only) a.db_out[f * a.width + k] = db;
      else if (f == a.n_frames - 1) a.db_out[k] = db;
      if (a.max_hold) {
        if (!mxv) { mx = isnan(db) ? -500.0f : db; mxv = true; }
        else mx = fmaxf(mx, db);
      }
      if (a.min_hold) {
        if (!mnv) { mn = isnan(db) ? 500.0f : db; mnv = true; }
        else mn = fminf(mn, db);
      }
    }
  }
  if (averaging && any) a.avg_state[k] = buf;
  if (a.max_hold && mxv) a.max_hold[k] = mx;
  if (a.min_hold && mnv) a.min_hold[k] = mn;
  if (a.last_row && any) a.last_row[k] = last_db;
}

// ---------------------------------------------------------------------------------------------------------------
// The same scan, parallel over frames as well (no skip flags): the averager's recurrence is affine in the buffer, so a
// block of kScanBlock frames maps its start state s to A s + B with A common to all bins.
//   1. scan_block_affine_kernel: (bin, block) -> B (and A once per block)            [reads the rows]
//   2. scan_chain_kernel:        bin -> state at the start of every block, final state
//   3. scan_block_emit_kernel:   (bin, block) -> the block's dB rows in the reference's operation order from the
//                                known start state, block max / min of the dB values   [reads the rows again, from L2]
//   4. scan_holds_kernel:        bin -> max / min hold from the block extrema
// One thread per bin walking 1024 frames is latency bound (175 us per 32 MB chunk on 16 SMs); this is ~6 x faster.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kScanBlock = 32;

struct TraceBlockArgs {
  const double* lin;       // [F][W]
  int64_t n_frames, width;
  int avg_mode, avg_n;
  const int32_t* flags;
  double* avg_state;
  float* max_hold;
  float* min_hold;
  float* last_row;
  int last_only;
  float* db_out;
  double floor;
  int mode;
  double* blk_b;           // [nblk][W]
  double* blk_a;           // [nblk]
  double* blk_start;       // [nblk][W]
  float* blk_max;          // [nblk][W]  (-inf: no value)
  float* blk_min;          // [nblk][W]  (+inf: no value)
};

// averager count BEFORE frame t of the call (every frame live): 0 means "the buffer is empty, this frame sets it"
__device__ __forceinline__ int count_before(int avg_mode, int avg_n, int count0, int64_t t) {
  if (avg_mode == 2) { const int64_t c = (int64_t)count0 + t; return (int)(c < avg_n ? c : avg_n); }
  return (count0 > 0 || t > 0) ? 1 : 0;
}

__global__ void __launch_bounds__(256) scan_block_affine_kernel(const TraceBlockArgs a) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.width) return;
  const int64_t t0 = (int64_t)blockIdx.y * kScanBlock, t1 = min(t0 + kScanBlock, a.n_frames);
  const int count0 = a.flags[kFlagCount];
  const double alpha = 1.0 / (double)a.avg_n;
  double A = 1.0, B = 0.0;
  double pre[kScanBlock];
#pragma unroll
  for (int u = 0; u < kScanBlock; ++u) pre[u] = (t0 + u < t1) ? a.lin[(t0 + u) * a.width + k] : 0.0;
#pragma unroll
  for (int u = 0; u < kScanBlock; ++u) {
    const int64_t t = t0 + u;
    if (t >= t1) break;
    const int c = count_before(a.avg_mode, a.avg_n, count0, t);
    if (c == 0) { A = 0.0; B = pre[u]; }
    else if (a.avg_mode == 1) { A *= (1.0 - alpha); B = __dadd_rn(__dmul_rn(B, 1.0 - alpha), __dmul_rn(alpha, pre[u])); }
    else {
      const double inv = 1.0 / (double)(c < a.avg_n ? c + 1 : a.avg_n);
      A *= (1.0 - inv); B = __fma_rn(B, 1.0 - inv, pre[u] * inv);
    }
  }
  a.blk_b[(int64_t)blockIdx.y * a.width + k] = B;
  if (k == 0) a.blk_a[blockIdx.y] = A;
}

__global__ void __launch_bounds__(256) scan_chain_kernel(const TraceBlockArgs a, int nblk) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.width) return;
  double s = a.flags[kFlagCount] > 0 ? a.avg_state[k] : 0.0;
  for (int b = 0; b < nblk; ++b) {
    a.blk_start[(int64_t)b * a.width + k] = s;
    s = __fma_rn(a.blk_a[b], s, a.blk_b[(int64_t)b * a.width + k]);
  }
  a.avg_state[k] = s;
}

__global__ void __launch_bounds__(256) scan_block_emit_kernel(const TraceBlockArgs a) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.width) return;
  const int64_t t0 = (int64_t)blockIdx.y * kScanBlock, t1 = min(t0 + kScanBlock, a.n_frames);
  const bool averaging = a.avg_mode != 0 && a.avg_n > 1;
  const int count0 = averaging ? a.flags[kFlagCount] : 0;
  const double alpha = 1.0 / (double)a.avg_n;
  double buf = averaging ? a.blk_start[(int64_t)blockIdx.y * a.width + k] : 0.0;
  EpiParams ep;
  ep.db_out = nullptr; ep.lin_out = nullptr; ep.scale = 1.0; ep.floor = a.floor; ep.mode = a.mode;
  float mx = -INFINITY, mn = INFINITY, db = 0.0f;
  double pre[kScanBlock];
#pragma unroll
  for (int u = 0; u < kScanBlock; ++u) pre[u] = (t0 + u < t1) ? a.lin[(t0 + u) * a.width + k] : 0.0;
#pragma unroll
  for (int u = 0; u < kScanBlock; ++u) {
    const int64_t t = t0 + u;
    if (t >= t1) break;
    double v = pre[u];
    if (averaging) {
      const int c = count_before(a.avg_mode, a.avg_n, count0, t);
      if (c == 0) buf = v;
      else if (a.avg_mode == 1) { buf = __dmul_rn(buf, 1.0 - alpha); buf = __dadd_rn(buf, __dmul_rn(alpha, v)); }
      else buf = __dadd_rn(buf, __ddiv_rn(__dsub_rn(v, buf), (double)(c < a.avg_n ? c + 1 : a.avg_n)));
      v = buf;
    }
    db = to_db<double>(v, ep);
    if (!a.last_only) a.db_out[t * a.width + k] = db;
    mx = fmaxf(mx, db); mn = fminf(mn, db);              // NaN-ignoring, like np.fmax / np.fmin
  }
  if (t1 == a.n_frames) {                                // the block that holds the call's last frame
    if (a.last_only) a.db_out[k] = db;
    if (a.last_row) a.last_row[k] = db;
  }
  if (a.max_hold) a.blk_max[(int64_t)blockIdx.y * a.width + k] = mx;
  if (a.min_hold) a.blk_min[(int64_t)blockIdx.y * a.width + k] = mn;
}

__global__ void __launch_bounds__(256) scan_holds_kernel(const TraceBlockArgs a, int nblk) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.width) return;
  if (a.max_hold) {
    float m = -INFINITY;
    for (int b = 0; b < nblk; ++b) m = fmaxf(m, a.blk_max[(int64_t)b * a.width + k]);
    if (a.flags[kFlagMaxValid] != 0) { if (m != -INFINITY) a.max_hold[k] = fmaxf(a.max_hold[k], m); }
    else a.max_hold[k] = m == -INFINITY ? -500.0f : m;   // _nan_safe on the initialising frame
  }
  if (a.min_hold) {
    float m = INFINITY;
    for (int b = 0; b < nblk; ++b) m = fminf(m, a.blk_min[(int64_t)b * a.width + k]);
    if (a.flags[kFlagMinValid] != 0) { if (m != INFINITY) a.min_hold[k] = fminf(a.min_hold[k], m); }
    else a.min_hold[k] = m == INFINITY ? 500.0f : m;
  }
}

// after a scan: fold the number of live (not skipped) frames into the flag block
__global__ void trace_flags_after_scan_kernel(int32_t* __restrict__ flags, const int32_t* __restrict__ skip, int64_t n_frames,
                                              int avg_mode, int avg_n, int has_max, int has_min, int first_chunk) {
  __shared__ int s_live;
  if (threadIdx.x == 0) s_live = 0;
  __syncthreads();
  int live = 0;
  for (int64_t f = threadIdx.x; f < n_frames; f += blockDim.x) live += (skip == nullptr || skip[f] == 0) ? 1 : 0;
  atomicAdd(&s_live, live);
  __syncthreads();
  if (threadIdx.x != 0) return;
  live = s_live;
  const bool averaging = avg_mode != 0 && avg_n > 1;
  if (live > 0) {
    if (averaging) {
      const int64_t c = flags[kFlagCount];
      flags[kFlagCount] = avg_mode == 2 ? (int32_t)((c + live) < avg_n ? (c + live) : avg_n) : (int32_t)(c > 1 ? c : 1);
    }
    if (has_max) flags[kFlagMaxValid] = 1;
    if (has_min) flags[kFlagMinValid] = 1;
  }
  flags[kFlagLive] = (first_chunk ? 0 : flags[kFlagLive]) + live;
}

// ---------------------------------------------------------------------------------------------------------------
// dB-row trace update (cal offset, tare, sweep-domain averaging, holds) with the flag block on the device.
// Same per-bin arithmetic as trace_update_kernel; the bookkeeping the host used to do is
// trace_flags_after_update_kernel.
// ---------------------------------------------------------------------------------------------------------------
struct TraceUpdateDevArgs {
  const float* rows;
  int64_t n_rows, width;
  double cal;
  int avg_mode, avg_n;
  const int32_t* flags;
  double* avg_state;
  float* max_hold;
  float* min_hold;
  const int32_t* has_value;   // per row
  float* rows_out;
  int tare_target;
  double* tare_buf;
  double* tare_baseline;
};

__global__ void __launch_bounds__(256) trace_update_dev_kernel(const TraceUpdateDevArgs a) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.width) return;
  const bool averaging = a.avg_mode != 0 && a.avg_n > 1;
  int count = averaging ? a.flags[kFlagCount] : 0;
  double buf = (averaging && count > 0) ? a.avg_state[k] : 0.0;
  const double alpha = 1.0 / (double)a.avg_n;
  bool mxv = a.flags[kFlagMaxValid] != 0, mnv = a.flags[kFlagMinValid] != 0;
  float mx = (a.max_hold && mxv) ? a.max_hold[k] : 0.f;
  float mn = (a.min_hold && mnv) ? a.min_hold[k] : 0.f;
  bool touched = false;
  const bool has_tare = a.tare_buf != nullptr && a.tare_baseline != nullptr;
  bool collecting = has_tare && a.flags[kFlagTareCollecting] != 0, tare_active = has_tare && a.flags[kFlagTareActive] != 0;
  bool captured = false;
  int tcount = collecting ? a.flags[kFlagTareCount] : 0;
  double tbuf = (collecting && tcount > 0) ? a.tare_buf[k] : 0.0;
  double base = tare_active ? a.tare_baseline[k] : 0.0;
  for (int64_t r = 0; r < a.n_rows; ++r) {
    double x = (double)a.rows[r * a.width + k];
    if (a.cal != 0.0) x += a.cal;
    if (!a.has_value[r]) {               // NaN-only frame: the reference returns before any state update (:211)
      if (a.rows_out) a.rows_out[r * a.width + k] = (float)x;
      continue;
    }
    if (collecting) {                    // accumulate the linear baseline, capture it after tare_target frames (:335-358)
      const double lin = pow(10.0, x / 10.0);
      if (tcount == 0) { tbuf = lin; tcount = 1; } else { tbuf += lin; ++tcount; }
      if (tcount >= a.tare_target) {
        base = 10.0 * log10(fmax(tbuf / (double)tcount, 1e-30));
        tare_active = true; collecting = false; captured = true;
      }
    }
    if (tare_active) x -= base;
    if (averaging) {
      const double lin = pow(10.0, x / 10.0);
      if (count == 0) {
        buf = lin;
        count = 1;
      } else if (a.avg_mode == 1) {
        buf = __dmul_rn(buf, 1.0 - alpha);
        buf = __dadd_rn(buf, __dmul_rn(alpha, lin));
      } else {
        if (count < a.avg_n) ++count;
        buf = __dadd_rn(buf, __ddiv_rn(__dsub_rn(lin, buf), (double)count));
      }
      touched = true;
      x = 10.0 * log10(fmax(buf, 1e-30));
    }
    const float db = (float)x;
    if (a.rows_out) a.rows_out[r * a.width + k] = db;
    if (a.max_hold) {
      if (!mxv) { mx = isnan(db) ? -500.0f : db; mxv = true; }
      else mx = fmaxf(mx, db);
    }
    if (a.min_hold) {
      if (!mnv) { mn = isnan(db) ? 500.0f : db; mnv = true; }
      else mn = fminf(mn, db);
    }
  }
  if (averaging && touched) a.avg_state[k] = buf;
  if (a.max_hold && mxv) a.max_hold[k] = mx;
  if (a.min_hold && mnv) a.min_hold[k] = mn;
  if (collecting) a.tare_buf[k] = tbuf;
  if (captured) a.tare_baseline[k] = base;
}

__global__ void trace_flags_after_update_kernel(int32_t* __restrict__ flags, const int32_t* __restrict__ has_value,
                                                int64_t n_rows, int avg_mode, int avg_n, int has_max, int has_min,
                                                int has_tare, int tare_target) {
  __shared__ int s_live;
  if (threadIdx.x == 0) s_live = 0;
  __syncthreads();
  int live = 0;
  for (int64_t r = threadIdx.x; r < n_rows; r += blockDim.x) live += has_value[r] != 0 ? 1 : 0;
  atomicAdd(&s_live, live);
  __syncthreads();
  if (threadIdx.x != 0) return;
  live = s_live;
  flags[kFlagLive] = live;
  if (live <= 0) return;
  if (has_tare && flags[kFlagTareCollecting]) {          // TareState bookkeeping, display_data_processor.py:335-358
    const int64_t c = (int64_t)flags[kFlagTareCount] + live;
    if (c >= tare_target) { flags[kFlagTareCollecting] = 0; flags[kFlagTareActive] = 1; flags[kFlagTareCount] = 0; }
    else flags[kFlagTareCount] = (int32_t)c;
  }
  if (avg_mode != 0 && avg_n > 1) {
    const int64_t c = flags[kFlagCount];
    flags[kFlagCount] = avg_mode == 2 ? (int32_t)((c + live) < avg_n ? (c + live) : avg_n) : (int32_t)(c > 1 ? c : 1);
  }
  if (has_max) flags[kFlagMaxValid] = 1;
  if (has_min) flags[kFlagMinValid] = 1;
}

}  // namespace tdsa
