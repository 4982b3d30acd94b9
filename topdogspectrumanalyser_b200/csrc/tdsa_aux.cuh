// tdsa_aux.cuh — trace-state, stitch, ring and front-end kernels around the fused FFT.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "tdsa_fft.cuh"

namespace tdsa {

// ---------------------------------------------------------------------------------------
// HackRF front end, datasources/hackrf_samples.py:351,360: per-frame mean(x), mean(|x|^2)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frame_stats_kernel(const float2* __restrict__ iq, int64_t n_frames,
                                                         int64_t frame_stride, int n, double2* __restrict__ mean_out,
                                                         double* __restrict__ pw_out) {
  __shared__ double s_r[8], s_i[8], s_p[8];
  for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
    const float2* x = iq + f * frame_stride;
    double sr = 0.0, si = 0.0, sp = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float2 v = x[i];
      sr += (double)v.x;
      si += (double)v.y;
      sp += (double)v.x * (double)v.x + (double)v.y * (double)v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, o);
      si += __shfl_xor_sync(0xffffffffu, si, o);
      sp += __shfl_xor_sync(0xffffffffu, sp, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s_r[w] = sr; s_i[w] = si; s_p[w] = sp; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0, c = 0.0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { a += s_r[k]; b += s_i[k]; c += s_p[k]; }
      mean_out[f] = make_double2(a / n, b / n);
      pw_out[f] = c / n;
    }
    __syncthreads();
  }
}

// dc_f = (1-alpha)*dc_{f-1} + alpha*mean_f for non-silent frames (hackrf_samples.py:351-365).
// Sequential in f by definition; one thread.
__global__ void dc_scan_kernel(const double2* __restrict__ mean, const double* __restrict__ pw, int64_t n_frames,
                               double alpha, double* __restrict__ dc_state, double2* __restrict__ dc_out,
                               int32_t* __restrict__ silent_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double dr = dc_state[0], di = dc_state[1];
  for (int64_t f = 0; f < n_frames; ++f) {
    const bool silent = pw[f] < 1e-20;
    if (!silent) {
      dr = (1.0 - alpha) * dr + alpha * mean[f].x;
      di = (1.0 - alpha) * di + alpha * mean[f].y;
    }
    dc_out[f] = make_double2(dr, di);
    if (silent_out) silent_out[f] = silent ? 1 : 0;
  }
  dc_state[0] = dr;
  dc_state[1] = di;
}

// ---------------------------------------------------------------------------------------
// TraceAverager + dB + max/min hold over consecutive frames, one thread per bin.
// utils/signal_processing.py:35-61; core/display_data_processor.py:371-395,473-480.
// ---------------------------------------------------------------------------------------
struct TraceScanArgs {
  const double* lin;       // [F][W] linear power (already scaled for PSD)
  const int32_t* skip;     // [F] or null: frames flagged 1 leave all state untouched (hackrf silence hold)
  int64_t n_frames;
  int64_t width;
  int avg_mode;            // 0 off, 1 exp, 2 lin
  int avg_n;
  int count0;              // averager count before this call (0 = empty buffer)
  double* avg_state;       // [W] float64 buffer (TraceAverager._buffer)
  float* max_hold;         // [W] or null
  float* min_hold;         // [W] or null
  int max_valid0, min_valid0;
  int last_only;
  float* db_out;           // [F][W] or [W] when last_only
  double floor;
  int mode;                // dB branch (mag20 never reaches here with averaging on)
};

__global__ void __launch_bounds__(256) trace_scan_kernel(const TraceScanArgs a) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.width) return;
  const bool averaging = a.avg_mode != 0 && a.avg_n > 1;
  double buf = (averaging && a.count0 > 0) ? a.avg_state[k] : 0.0;
  int count = a.count0;
  const double alpha = 1.0 / (double)a.avg_n;
  bool mxv = a.max_valid0 != 0, mnv = a.min_valid0 != 0;
  float mx = (a.max_hold && mxv) ? a.max_hold[k] : 0.f;
  float mn = (a.min_hold && mnv) ? a.min_hold[k] : 0.f;
  EpiParams ep;
  ep.db_out = nullptr; ep.lin_out = nullptr; ep.scale = 1.0; ep.floor = a.floor; ep.mode = a.mode;
  float last_db = 0.0f;       // what a skipped frame repeats (hackrf_samples.py:351-355)
  for (int64_t f = 0; f < a.n_frames; ++f) {
    if (a.skip != nullptr && a.skip[f]) {
      if (!a.last_only) a.db_out[f * a.width + k] = last_db;
      else if (f == a.n_frames - 1) a.db_out[k] = last_db;
      continue;
    }
    const double p = a.lin[f * a.width + k];
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
  if (averaging && a.n_frames > 0) a.avg_state[k] = buf;
  if (a.max_hold && mxv) a.max_hold[k] = mx;
  if (a.min_hold && mnv) a.min_hold[k] = mn;
}

// ---------------------------------------------------------------------------------------
// Trace update on dB rows: cal offset (display_data_processor.py:317-327), sweep-domain
// averaging (:211-218), holds (:371-395). Rows that are entirely NaN are skipped (:211).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row_allnan_kernel(const float* __restrict__ rows, int64_t n_rows, int64_t width,
                                                        int32_t* __restrict__ has_value) {
  // has_value[r] != 0 iff row r holds at least one non-NaN
  const int64_t r = blockIdx.y;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows || k >= width) return;
  if (!isnan(rows[r * width + k])) has_value[r] = 1;
}

struct TraceUpdateArgs {
  const float* rows;
  int64_t n_rows, width;
  double cal;
  int avg_mode, avg_n, count0;
  double* avg_state;
  float* max_hold;
  float* min_hold;
  int max_valid0, min_valid0;
  const int32_t* has_value;   // per row
  float* rows_out;            // [n_rows][width] or null
  // tare (display_data_processor.py:329-369): bit 0 of tare_mode = collecting, bit 1 = baseline active
  int tare_mode, tare_count0, tare_target;
  double* tare_buf;           // [width] running sum of 10^(dB/10) while collecting
  double* tare_baseline;      // [width] dB baseline once captured
};

__global__ void __launch_bounds__(256) trace_update_kernel(const TraceUpdateArgs a) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.width) return;
  const bool averaging = a.avg_mode != 0 && a.avg_n > 1;
  double buf = (averaging && a.count0 > 0) ? a.avg_state[k] : 0.0;
  int count = a.count0;
  const double alpha = 1.0 / (double)a.avg_n;
  bool mxv = a.max_valid0 != 0, mnv = a.min_valid0 != 0;
  float mx = (a.max_hold && mxv) ? a.max_hold[k] : 0.f;
  float mn = (a.min_hold && mnv) ? a.min_hold[k] : 0.f;
  bool touched = false;
  bool collecting = (a.tare_mode & 1) != 0, tare_active = (a.tare_mode & 2) != 0, captured = false;
  int tcount = a.tare_count0;
  double tbuf = (collecting && tcount > 0) ? a.tare_buf[k] : 0.0;
  double base = tare_active ? a.tare_baseline[k] : 0.0;
  for (int64_t r = 0; r < a.n_rows; ++r) {
    double x = (double)a.rows[r * a.width + k];
    if (a.cal != 0.0) x += a.cal;
    if (!a.has_value[r]) {               // NaN-only frame: the reference returns before any state update
      if (a.rows_out) a.rows_out[r * a.width + k] = (float)x;
      continue;
    }
    if (collecting) {                    // accumulate the linear baseline, capture it after tare_target frames
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

// ---------------------------------------------------------------------------------------
// Waterfall colour map, core/export_manager.py:72-79: norm = clip((x - lo)/max(hi - lo, 1e-9), 0, 1) in
// float32, index = uint8(norm*255) (truncation), RGBA = lut[index].  NaN rows map to index 0.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colormap_kernel(const float* __restrict__ rows, int64_t n, float lo, float inv_den_num,
                                                      const uchar4* __restrict__ lut, uchar4* __restrict__ rgba) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = __fdiv_rn(__fsub_rn(rows[i], lo), inv_den_num);   // inv_den_num holds max(hi - lo, 1e-9)
  v = fminf(fmaxf(v, 0.0f), 1.0f);                            // NaN -> 0 (np.clip keeps NaN; uint8(NaN) is 0 on x86)
  const int idx = (int)__fmul_rn(v, 255.0f);
  rgba[i] = lut[idx & 255];
}

// ---------------------------------------------------------------------------------------
// Density histogram with decay, displays/density_display.py:306-319: hist[W][512] float32;
// hist *= float32(decay) when decay < 1; bin = int32((db + 200)/300*512) in float64, +1 if in range.
// One thread per (frequency bin, 4 amplitude bins) for the decay, then the increment.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) density_update_kernel(const float* __restrict__ live_db, int64_t width, float decay,
                                                            int apply_decay, float* __restrict__ hist) {
  constexpr int kBins = 512;
  const int64_t f = blockIdx.x;                      // one CTA per frequency bin row: 512 floats
  if (f >= width) return;
  float* row = hist + f * kBins;
  if (apply_decay) {
    for (int b = threadIdx.x; b < kBins; b += blockDim.x) row[b] = __fmul_rn(row[b], decay);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float db = live_db[f];
    if (!isnan(db)) {
      const double v = __dmul_rn(__ddiv_rn(__dsub_rn((double)db, -200.0), 300.0), 512.0);
      // astype(int32) truncates toward zero, so (-1, 0) lands in bin 0; everything else outside [0, 512) is dropped
      if (v > -1.0 && v < 512.0) { const int idx = (int)v; row[idx] = __fadd_rn(row[idx], 1.0f); }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Band power, core/marker_manager.py:308-318: 10*log10(max(sum(10^(levels[mask]/10)) * bin_width, 1e-30))
// over bins with lo <= f <= hi. Single CTA; returns NaN when the mask is empty (reference returns None).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) band_power_kernel(const double* __restrict__ bins, const float* __restrict__ levels,
                                                        int64_t width, double lo, double hi, double* __restrict__ out) {
  __shared__ double s_sum[8];
  __shared__ int s_any[8];
  double acc = 0.0;
  int any = 0;
  for (int64_t i = threadIdx.x; i < width; i += blockDim.x) {
    const double f = bins[i];
    if (f >= lo && f <= hi) { acc += pow(10.0, (double)levels[i] / 10.0); any = 1; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); any |= __shfl_xor_sync(0xffffffffu, any, o); }
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = acc; s_any[threadIdx.x >> 5] = any; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0; int a = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t += s_sum[w]; a |= s_any[w]; }
    const double bw = (bins[width - 1] - bins[0]) / (double)(width > 1 ? width - 1 : 1);
    out[0] = a ? 10.0 * log10(fmax(t * bw, 1e-30)) : nan("");
  }
}

// ---------------------------------------------------------------------------------------
// Top-n peak search, core/display_data_processor.py:432-471: strict local maxima, visited in order of
// decreasing power; a candidate is rejected if it is closer than min_sep bins to a selected peak or if the
// valley between them is not at least min_exc dB below BOTH peaks. One CTA; width <= 16384.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) top_peaks_kernel(const float* __restrict__ power, int width, int n_want, int min_sep,
                                                        float min_exc, int32_t* __restrict__ out_idx, float* __restrict__ out_pwr,
                                                        int32_t* __restrict__ out_count) {
  extern __shared__ unsigned char raw[];
  float* key = reinterpret_cast<float*>(raw);                 // candidate powers, padded to cap with -inf
  int32_t* val = reinterpret_cast<int32_t*>(key + 8192);      // candidate bin indices
  __shared__ int s_n;
  __shared__ float s_red[32];
  __shared__ int s_sel_idx[16];
  __shared__ float s_sel_pwr[16];
  const int t = threadIdx.x, nt = blockDim.x;
  if (t == 0) s_n = 0;
  for (int i = t; i < 8192; i += nt) { key[i] = -INFINITY; val[i] = -1; }
  __syncthreads();
  for (int i = 1 + t; i < width - 1; i += nt) {
    const float p = power[i];
    if (p > power[i - 1] && p > power[i + 1]) {
      const int slot = atomicAdd(&s_n, 1);
      if (slot < 8192) { key[slot] = p; val[slot] = i; }
    }
  }
  __syncthreads();
  const int ncand = min(s_n, 8192);
  int cap = 1;
  while (cap < ncand) cap <<= 1;
  // bitonic sort, descending by (power, then lower index first for determinism)
  for (int k = 2; k <= cap; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < cap; i += nt) {
        const int l = i ^ j;
        if (l > i) {
          const bool desc = (i & k) == 0;
          const float a = key[i], b = key[l];
          const int ia = val[i], ib = val[l];
          const bool a_first = (a > b) || (a == b && ia < ib);     // should a precede b in descending order
          if (desc ? !a_first : a_first) { key[i] = b; key[l] = a; val[i] = ib; val[l] = ia; }
        }
      }
      __syncthreads();
    }
  }
  int nsel = 0;
  for (int ci = 0; ci < ncand && nsel < n_want; ++ci) {
    const int idx = val[ci];
    const float pw = key[ci];
    bool reject = false;
    for (int sidx = 0; sidx < nsel && !reject; ++sidx) {
      const int sel = s_sel_idx[sidx];
      const float spw = s_sel_pwr[sidx];
      if (abs(idx - sel) < min_sep) { reject = true; break; }
      const int lo = min(idx, sel), hi = max(idx, sel);
      float m = INFINITY;
      for (int i = lo + t; i <= hi; i += nt) m = fminf(m, power[i]);   // np.min propagates NaN; fminf ignores it
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if ((t & 31) == 0) s_red[t >> 5] = m;
      __syncthreads();
      float valley = INFINITY;
      for (int w = 0; w < (nt >> 5); ++w) valley = fminf(valley, s_red[w]);
      __syncthreads();
      if (pw - valley < min_exc || spw - valley < min_exc) reject = true;
    }
    if (!reject) {
      if (t == 0) { s_sel_idx[nsel] = idx; s_sel_pwr[nsel] = pw; out_idx[nsel] = idx; out_pwr[nsel] = pw; }
      ++nsel;
      __syncthreads();
    }
  }
  if (t == 0) out_count[0] = nsel;
}

// ---------------------------------------------------------------------------------------
// hackrf_sweep stitch, datasources/hackrf_sweep.py:150-166.
// rank[r] = position of row r when rows are ordered by hz_low (argsort; ties by index).
// ---------------------------------------------------------------------------------------
__global__ void stitch_rank_kernel(const double* __restrict__ lo, int64_t n_rows, int32_t* __restrict__ order) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const double v = lo[r];
  int32_t rank = 0;
  for (int64_t j = 0; j < n_rows; ++j) {
    const double u = lo[j];
    rank += (u < v) || (u == v && j < r);
  }
  order[rank] = (int32_t)r;   // order[i] = index of the row with the i-th lowest hz_low
}

struct StitchArgs {
  const float* rows;       // [R][K]
  const double* lo;        // [R]
  const int32_t* order;    // [R]
  double row_hz;           // hz_high - hz_low of every row
  int64_t n_rows, k;
  double start, stop;
  int64_t m;               // points of the whole grid
  int64_t g0 = 0;          // first grid point this launch computes (sharded stitch: each rank a slice)
  int64_t count = -1;      // how many (-1: all m)
  double* out;             // [count]: out[i] = grid point g0 + i
  // loop invariants, computed once on the host with the same IEEE operations numpy uses (a float64 division costs
  // ~40 FP64-pipe instructions on the device; ten of them per grid point were most of this kernel's 75 us)
  double step;             // (stop - start) / (m - 1)          np.linspace
  double bw;               // row_hz / k                        hackrf_sweep.py:159
  double half_bw;          // bw / 2
  double inv_bw, inv_row;  // reciprocals, used for index GUESSES only (every guess is verified by exact comparisons)
};

// x of sample i of row r, exactly as np.arange(lo + bw/2, hi, bw) produces it (start + i*bw).
__device__ __forceinline__ double stitch_x(const StitchArgs& a, int64_t sorted_idx, double bw) {
  const int64_t rr = sorted_idx / a.k, i = sorted_idx - rr * a.k;
  const double x0 = __dadd_rn(a.lo[a.order[rr]], a.half_bw);
  return __dadd_rn(x0, __dmul_rn((double)i, bw));
}
__device__ __forceinline__ double stitch_y(const StitchArgs& a, int64_t sorted_idx) {
  const int64_t rr = sorted_idx / a.k, i = sorted_idx - rr * a.k;
  return (double)a.rows[(int64_t)a.order[rr] * a.k + i];
}

// same values as stitch_x / stitch_y for sample i of sorted row rr, without the 64-bit division by k
__device__ __forceinline__ double stitch_x_ri(const StitchArgs& a, int64_t rr, int64_t i, double bw) {
  return __dadd_rn(__dadd_rn(a.lo[a.order[rr]], a.half_bw), __dmul_rn((double)i, bw));
}
__device__ __forceinline__ double stitch_y_ri(const StitchArgs& a, int64_t rr, int64_t i) {
  return (double)a.rows[(int64_t)a.order[rr] * a.k + i];
}

__global__ void __launch_bounds__(256) stitch_interp_kernel(const StitchArgs a) {
  const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t cnt = a.count < 0 ? a.m : a.count;
  if (gi >= cnt) return;
  const int64_t g = a.g0 + gi;
  // np.linspace(start, stop, m): arange(m)*step + start, last element forced to stop
  const double step = a.step;
  double xv = __dadd_rn(__dmul_rn((double)g, step), a.start);
  if (g == a.m - 1 && a.m > 1) xv = a.stop;
  const double bw = a.bw;
  const int64_t total = a.n_rows * a.k;
  // np.interp: left/right clamp, then j = largest index with xp[j] <= x
  const double x_first = stitch_x_ri(a, 0, 0, bw), x_last = stitch_x_ri(a, a.n_rows - 1, a.k - 1, bw);
  if (xv < x_first) { a.out[gi] = stitch_y_ri(a, 0, 0); return; }                         // default left = fp[0]
  if (xv > x_last) { a.out[gi] = stitch_y_ri(a, a.n_rows - 1, a.k - 1); return; }         // default right = fp[-1]
  // j = (row rlo, sample i) = largest index with x[j] <= xv.  hackrf_sweep's rows tile the span uniformly, so the row
  // and the sample follow from the spacing; exact comparisons (the ones np.interp's search makes) settle both, and
  // anything irregular falls back to binary searches.  Indices stay (row, sample) pairs: a 64-bit division by k per
  // coordinate evaluation used to be most of this kernel's time.
  int64_t rlo = (int64_t)floor((xv - x_first) * a.inv_row);
  rlo = rlo < 0 ? 0 : (rlo > a.n_rows - 1 ? a.n_rows - 1 : rlo);
  int steps = 0;
  while (steps < 4 && rlo < a.n_rows - 1 && stitch_x_ri(a, rlo + 1, 0, bw) <= xv) { ++rlo; ++steps; }
  while (steps < 4 && rlo > 0 && stitch_x_ri(a, rlo, 0, bw) > xv) { --rlo; ++steps; }
  if (steps >= 4) {
    rlo = 0;
    int64_t rhi = a.n_rows - 1;                      // x_first(0) <= xv (checked above)
    while (rlo < rhi) {
      const int64_t mid = (rlo + rhi + 1) >> 1;
      if (stitch_x_ri(a, mid, 0, bw) <= xv) rlo = mid; else rhi = mid - 1;
    }
  }
  const double x0 = stitch_x_ri(a, rlo, 0, bw);
  int64_t i = (int64_t)floor((xv - x0) * a.inv_bw);
  i = i < 0 ? 0 : (i > a.k - 1 ? a.k - 1 : i);
  while (i < a.k - 1 && stitch_x_ri(a, rlo, i + 1, bw) <= xv) ++i;
  while (i > 0 && stitch_x_ri(a, rlo, i, bw) > xv) --i;
  // successor of (rlo, i) in sorted order
  int64_t r1 = rlo, i1 = i + 1;
  if (i1 == a.k) { r1 = rlo + 1; i1 = 0; }
  const bool is_last = r1 >= a.n_rows;
  double xj = stitch_x_ri(a, rlo, i, bw);
  // overlapping rows break the ordering the shortcut relies on: verify, and fall back to the search over all samples
  if (!(xj <= xv) || (!is_last && stitch_x_ri(a, r1, i1, bw) <= xv)) {
    int64_t lo_i = 0, hi_i = total - 1;     // invariant: x[lo_i] <= xv, and (hi_i == total-1 or x[hi_i] > xv)
    while (hi_i - lo_i > 1) {
      const int64_t mid = (lo_i + hi_i) >> 1;
      if (stitch_x(a, mid, bw) <= xv) lo_i = mid; else hi_i = mid;
    }
    int64_t j = lo_i;
    if (stitch_x(a, hi_i, bw) <= xv) j = hi_i;
    const double xs = stitch_x(a, j, bw), ys = stitch_y(a, j);
    if (j == total - 1 || xs == xv) { a.out[gi] = ys; return; }
    const double xs1 = stitch_x(a, j + 1, bw), ys1 = stitch_y(a, j + 1);
    const double sl = __ddiv_rn(__dsub_rn(ys1, ys), __dsub_rn(xs1, xs));
    double rs = __dadd_rn(__dmul_rn(sl, __dsub_rn(xv, xs)), ys);
    if (isnan(rs)) {
      rs = __dadd_rn(__dmul_rn(sl, __dsub_rn(xv, xs1)), ys1);
      if (isnan(rs) && ys == ys1) rs = ys;
    }
    a.out[gi] = rs;
    return;
  }
  const double yj = stitch_y_ri(a, rlo, i);
  if (is_last || xj == xv) { a.out[gi] = yj; return; }
  const double xj1 = stitch_x_ri(a, r1, i1, bw), yj1 = stitch_y_ri(a, r1, i1);
  const double slope = __ddiv_rn(__dsub_rn(yj1, yj), __dsub_rn(xj1, xj));
  double res = __dadd_rn(__dmul_rn(slope, __dsub_rn(xv, xj)), yj);
  if (isnan(res)) {
    res = __dadd_rn(__dmul_rn(slope, __dsub_rn(xv, xj1)), yj1);
    if (isnan(res) && yj == yj1) res = yj;
  }
  a.out[gi] = res;
}

// ---------------------------------------------------------------------------------------
// waterfall ring, displays/waterfall.py:173-177: row r lands at (ptr0 - 1 - r) mod H and +H.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ring_push_kernel(const float* __restrict__ rows, int64_t first_row,
                                                       int64_t n_rows, float* __restrict__ ring, int64_t H, int64_t W,
                                                       int64_t ptr0) {
  const int64_t r = first_row + blockIdx.y;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows || k >= W) return;
  int64_t p = (ptr0 - 1 - r) % H;
  if (p < 0) p += H;
  const float v = rows[r * W + k];
  ring[p * W + k] = v;
  ring[(p + H) * W + k] = v;
}

// ---------------------------------------------------------------------------------------
// Welch reduce over a chunk of segments (config 3): sum of linear power, fmax of per-segment dB.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) welch_reduce_kernel(const double* __restrict__ lin, int64_t n_seg, int64_t width,
                                                          double* __restrict__ sum_state, float* __restrict__ peak_state,
                                                          int first_chunk, double floor, int mode) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= width) return;
  EpiParams ep;
  ep.db_out = nullptr; ep.lin_out = nullptr; ep.scale = 1.0; ep.floor = floor; ep.mode = mode;
  double s = first_chunk ? 0.0 : sum_state[k];
  float pk = first_chunk ? -INFINITY : peak_state[k];
  for (int64_t f = 0; f < n_seg; ++f) {
    const double p = lin[f * width + k];
    s += p;
    pk = fmaxf(pk, to_db<double>(p, ep));
  }
  sum_state[k] = s;
  peak_state[k] = pk;
}

// Config 4: each group of `frames` consecutive rows is averaged as TraceAverager('lin', n >= frames)
// does (signal_processing.py:56-59: buffer += (x - buffer)/count) and converted to one dB row.
__global__ void __launch_bounds__(256) group_mean_db_kernel(const double* __restrict__ lin, int64_t n_groups,
                                                           int64_t frames, int64_t width, double floor, int mode,
                                                           float* __restrict__ db_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_groups * width) return;
  const int64_t g = i / width, k = i - g * width;
  const double* src = lin + g * frames * width + k;
  double buf = src[0];
  for (int64_t f = 1; f < frames; ++f)
    buf = __dadd_rn(buf, __ddiv_rn(__dsub_rn(src[f * width], buf), (double)(f + 1)));
  EpiParams ep;
  ep.db_out = nullptr; ep.lin_out = nullptr; ep.scale = 1.0; ep.floor = floor; ep.mode = mode;
  db_out[i] = to_db<double>(buf, ep);
}

// final: avg_db[k] = dB(sum/nseg); optional un-permute for the two-kernel large-FFT path: permuted index
// i = s*M + klow;  two head passes: s = 16*k0 + k1 -> k = k0 + 16*k1 + 256*klow;  one head pass: k = s + 16*klow.
__global__ void __launch_bounds__(256) welch_finish_kernel(const double* __restrict__ sum_state,
                                                          const float* __restrict__ peak_state, int64_t width,
                                                          int64_t n_seg, int log2_m_perm, int head_passes, double floor,
                                                          int mode, float* __restrict__ avg_db, float* __restrict__ peak_db) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width) return;
  int64_t k = i;
  if (log2_m_perm > 0) {
    const int64_t m = (int64_t)1 << log2_m_perm;
    const int64_t s = i >> log2_m_perm, kl = i & (m - 1);
    k = (head_passes == 1) ? s + 16 * kl : (s >> 4) + 16 * (s & 15) + 256 * kl;
  }
  EpiParams ep;
  ep.db_out = nullptr; ep.lin_out = nullptr; ep.scale = 1.0; ep.floor = floor; ep.mode = mode;
  avg_db[k] = to_db<double>(sum_state[i] / (double)n_seg, ep);
  peak_db[k] = peak_state[i];
}

}  // namespace tdsa
