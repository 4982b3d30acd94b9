// tdsa_display.cuh — display-side kernels that sit directly behind the hot path's dB rows:
//   * waterfall ring push with the widget's identical-row dedupe (displays/waterfall.py:330-336, 173-177),
//     write pointer kept on the device because the number of new rows is only known there;
//   * the ring's display view colour-mapped to the RGBA image the exporter builds (core/export_manager.py:67-84);
//   * marker "snap to peak" = scipy.signal.find_peaks(levels, height, prominence, distance) + argmax of the peak
//     heights (core/marker_manager.py:74-99).
#pragma once
#include "tdsa_aux.cuh"

namespace tdsa {

// ring state block on the device: {ptr, has_last, rows pushed by the last call}
enum : int { kRingPtr = 0, kRingHasLast = 1, kRingPushed = 2, kRingWords = 4 };

// differs[r] = 1 unless row r equals its predecessor element for element (np.array_equal: NaN never equals);
// the predecessor of row 0 is the last row pushed by an earlier call, if there is one
__global__ void __launch_bounds__(256) ring_row_differs_kernel(const float* __restrict__ rows, int64_t n_rows, int64_t W,
                                                              const float* __restrict__ last_row,
                                                              const int64_t* __restrict__ state, int32_t* __restrict__ differs) {
  const int64_t r = blockIdx.x;
  if (r >= n_rows) return;
  const float* cur = rows + r * W;
  const float* prev = r > 0 ? rows + (r - 1) * W : last_row;
  __shared__ int s_diff;
  if (threadIdx.x == 0) s_diff = (r == 0 && state[kRingHasLast] == 0) ? 1 : 0;
  __syncthreads();
  int d = 0;
  for (int64_t k = threadIdx.x; k < W && !d; k += blockDim.x) d |= !(cur[k] == prev[k]);
  if (d) s_diff = 1;
  __syncthreads();
  if (threadIdx.x == 0) differs[r] = s_diff;
}

// slot[r] = ring row the r-th input row goes to, or -1 when it is dropped as a duplicate; updates the state block
__global__ void ring_assign_kernel(const int32_t* __restrict__ differs, int64_t n_rows, int64_t H, int64_t* __restrict__ state,
                                   int64_t* __restrict__ slot, int dedupe) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int64_t ptr = state[kRingPtr], pushed = 0;
  for (int64_t r = 0; r < n_rows; ++r) {
    if (!dedupe || differs[r]) {
      ptr = (ptr - 1) % H;                               // displays/waterfall.py:174
      if (ptr < 0) ptr += H;
      slot[r] = ptr;
      ++pushed;
    } else {
      slot[r] = -1;
    }
  }
  // a slot written more than once in this call (more new rows than H) keeps the LAST writer: clear earlier ones
  if (pushed > H) {
    int64_t seen = 0;
    for (int64_t r = n_rows - 1; r >= 0; --r) {
      if (slot[r] < 0) continue;
      if (seen >= H) slot[r] = -1;
      ++seen;
    }
  }
  state[kRingPtr] = ptr;
  state[kRingPushed] = pushed;
  if (pushed > 0) state[kRingHasLast] = 1;
}

__global__ void __launch_bounds__(256) ring_scatter_kernel(const float* __restrict__ rows, int64_t n_rows, int64_t W, int64_t H,
                                                          const int64_t* __restrict__ slot, const int64_t* __restrict__ state,
                                                          float* __restrict__ ring, float* __restrict__ last_row) {
  const int64_t r = blockIdx.y;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows || k >= W) return;
  const int64_t p = slot[r];
  if (p < 0) {
    // still the newest new row?  (a dropped duplicate never is; an overwritten row may be, only when pushed > H)
    return;
  }
  const float v = rows[r * W + k];
  ring[p * W + k] = v;                                    // displays/waterfall.py:175-176: both halves
  ring[(p + H) * W + k] = v;
  if (p == state[kRingPtr] && last_row != nullptr) last_row[k] = v;   // the newest row is the one at the write pointer
}

// RGBA image of the display view ring[ptr : ptr + H] (newest row first), core/export_manager.py:72-79
__global__ void __launch_bounds__(256) ring_image_kernel(const float* __restrict__ ring, int64_t H, int64_t W,
                                                        const int64_t* __restrict__ state, float lo, float den,
                                                        const uchar4* __restrict__ lut, uchar4* __restrict__ rgba) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int64_t ptr = state[kRingPtr];
  float v = __fdiv_rn(__fsub_rn(ring[ptr * W + i], lo), den);
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  rgba[i] = lut[(int)__fmul_rn(v, 255.0f) & 255];
}

// ---------------------------------------------------------------------------------------------------------------
// scipy.signal.find_peaks(x, height=h, prominence=p, distance=d), then the highest surviving peak
// (core/marker_manager.py:74-99).  One CTA, width <= 16384.
//   1. _local_maxima_1d: strict rise on the left, plateaus collapse to their midpoint, edges excluded;
//   2. height:     x[peak] >= h;
//   3. distance:   peaks visited by decreasing height, each kept peak deletes every other peak closer than d samples;
//   4. prominence: x[peak] - max(min to the left until a higher sample, same to the right) >= p;
//   5. result: first kept peak with the largest height, else argmax(x).
// out[0] = chosen index, out[1] = number of peaks that survived, out[2] = 1 if the fallback argmax was used.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) find_peaks_snap_kernel(const float* __restrict__ x, int width, float height, float prominence,
                                                              int distance, int32_t* __restrict__ out) {
  extern __shared__ unsigned char raw[];
  float* key = reinterpret_cast<float*>(raw);                 // candidate heights (sorted descending), -inf padding
  int32_t* val = reinterpret_cast<int32_t*>(key + 8192);      // candidate positions
  int32_t* ord = val + 8192;                                  // rank in position order -> candidate slot (for neighbours)
  unsigned char* keep = reinterpret_cast<unsigned char*>(ord + 8192);
  __shared__ int s_n, s_best, s_kept;
  const int t = threadIdx.x, nt = blockDim.x;
  if (t == 0) { s_n = 0; s_best = -1; s_kept = 0; }
  for (int i = t; i < 8192; i += nt) { key[i] = -INFINITY; val[i] = 0x7fffffff; }
  __syncthreads();
  // 1 + 2: local maxima (plateau midpoints) that reach the height
  for (int i = 1 + t; i < width - 1; i += nt) {
    const float v = x[i];
    if (!(x[i - 1] < v)) continue;                            // a plateau is reported by its left edge only
    int ahead = i + 1;
    while (ahead < width - 1 && x[ahead] == v) ++ahead;
    if (x[ahead] < v) {
      const int mid = (i + ahead - 1) / 2;
      if (v >= height) {
        const int slot = atomicAdd(&s_n, 1);
        if (slot < 8192) { key[slot] = v; val[slot] = mid; }
      }
    }
  }
  __syncthreads();
  const int ncand = min(s_n, 8192);
  int cap = 1;
  while (cap < ncand) cap <<= 1;
  // sort by position first (ascending) to get neighbour ranks, remember it, then by height (descending)
  for (int k = 2; k <= cap; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < cap; i += nt) {
        const int l = i ^ j;
        if (l > i) {
          const bool asc = (i & k) == 0;
          const int ia = val[i], ib = val[l];
          if (asc ? (ia > ib) : (ia < ib)) { val[i] = ib; val[l] = ia; const float a = key[i]; key[i] = key[l]; key[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  // candidates are now in position order: slot s = s-th peak from the left.  The greedy distance rule only needs, for
  // each peak, its height and its neighbours' positions; visit order = decreasing height (ties: lower position first).
  for (int i = t; i < cap; i += nt) { ord[i] = i; keep[i] = i < ncand ? 1 : 0; }
  __syncthreads();
  for (int k = 2; k <= cap; k <<= 1) {                        // ord sorted so that key[ord[.]] is descending
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < cap; i += nt) {
        const int l = i ^ j;
        if (l > i) {
          const bool desc = (i & k) == 0;
          const int oa = ord[i], ob = ord[l];
          const float a = key[oa], b = key[ob];
          const bool a_first = (a > b) || (a == b && oa < ob);
          if (desc ? !a_first : a_first) { ord[i] = ob; ord[l] = oa; }
        }
      }
      __syncthreads();
    }
  }
  if (t == 0 && distance > 1) {                               // scipy _select_by_peak_distance
    for (int i = 0; i < ncand; ++i) {
      const int j = ord[i];
      if (!keep[j]) continue;
      for (int k = j - 1; k >= 0 && val[j] - val[k] < distance; --k) keep[k] = 0;
      for (int k = j + 1; k < ncand && val[k] - val[j] < distance; ++k) keep[k] = 0;
    }
  }
  __syncthreads();
  // 4: prominence of the kept peaks (scipy _peak_prominences with wlen = -1)
  for (int s = t; s < ncand; s += nt) {
    if (!keep[s]) continue;
    const int pk = val[s];
    const float v = key[s];
    float lmin = v, rmin = v;
    for (int i = pk; i >= 0 && x[i] <= v; --i) lmin = fminf(lmin, x[i]);
    for (int i = pk; i < width && x[i] <= v; ++i) rmin = fminf(rmin, x[i]);
    if (!(v - fmaxf(lmin, rmin) >= prominence)) keep[s] = 0;
  }
  __syncthreads();
  // 5: the highest survivor (first one among equals)
  if (t == 0) {
    int best = -1, kept = 0;
    float bv = -INFINITY;
    for (int s = 0; s < ncand; ++s) {
      if (!keep[s]) continue;
      ++kept;
      if (key[s] > bv) { bv = key[s]; best = val[s]; }
    }
    s_best = best; s_kept = kept;
  }
  __syncthreads();
  if (s_best >= 0) {
    if (t == 0) { out[0] = s_best; out[1] = s_kept; out[2] = 0; }
    return;
  }
  // fallback: np.argmax(levels) (first maximum; a NaN anywhere wins, like numpy)
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = t; i < width; i += nt) {
    const float v = x[i];
    if (isnan(v)) { if (bi == 0x7fffffff || !isnan(bv) || i < bi) { bv = v; bi = i; } break; }
    if (v > bv) { bv = v; bi = i; }
  }
  // block reduce (NaN beats everything, then larger value, then lower index)
  __shared__ float r_v[1024];
  __shared__ int r_i[1024];
  r_v[t] = bv; r_i[t] = bi;
  __syncthreads();
  for (int o = nt >> 1; o > 0; o >>= 1) {
    if (t < o) {
      const float a = r_v[t], b = r_v[t + o];
      const int ia = r_i[t], ib = r_i[t + o];
      bool take_b;
      if (isnan(a) || isnan(b)) take_b = isnan(b) && (!isnan(a) || ib < ia);
      else take_b = (b > a) || (b == a && ib < ia);
      if (take_b) { r_v[t] = b; r_i[t] = ib; }
    }
    __syncthreads();
  }
  if (t == 0) { out[0] = r_i[0] == 0x7fffffff ? 0 : r_i[0]; out[1] = 0; out[2] = 1; }
}

}  // namespace tdsa
