// tdsa_api.cu — C ABI of libtdsa.so (see include/tdsa.h for the contract and the
// reference file:line each entry point replaces).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/tdsa.h"
#include "tdsa_aux.cuh"
#include "tdsa_trace.cuh"
#include "tdsa_display.cuh"
#include "tdsa_big.cuh"
#include "tdsa_launch.cuh"

namespace tdsa {
std::atomic<int64_t> g_launch_count{0};
}

using namespace tdsa;

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CK(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) return fail(TDSA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                             \
  } while (0)

static inline void count_launch() { g_launch_count.fetch_add(1, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------
struct tdsa_plan {
  int n = 0, log2n = 0;
  int window_id = 0, window_norm = 0, mode = 0, precision = 0;
  double floor = 1e-10, fs = 1.0;
  int device = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  std::vector<double> window_host;           // float64, no fftshift sign
  // device tables
  double* d_win64 = nullptr; float* d_win32 = nullptr;        // with (-1)^n folded in
  double2* d_tw64 = nullptr; float2* d_tw32 = nullptr;
  bool win_dirty = true;
  // warp-local kernels (tdsa_fft_wl.cuh; N = 4096: one engine, N = 8192: two): window permuted to thread order,
  // per-engine twiddle tables (N = 8192 only; N = 4096 uses d_tw*), frame scheduler words {next, done}, and the
  // tensor map of the last input batch
  int wl_nb = 0;                                              // engines, 0 = this size has no warp-local kernel
  double* d_wperm64 = nullptr; float* d_wperm32 = nullptr;
  double2* d_wltw64 = nullptr; float2* d_wltw32 = nullptr;
  int* d_sched = nullptr;
  CUtensorMap tmap; const void* tmap_ptr = nullptr; int64_t tmap_frames = -1, tmap_stride = -1;
  // accumulating epilogue: per-CTA partial rows, per-frame weights, {carry, count0, count1, max_valid0, min_valid0}
  void* acc_parts = nullptr; size_t acc_parts_bytes = 0;
  void* acc_weights = nullptr; size_t acc_weights_bytes = 0;
  double* d_meta = nullptr;
  // trace state owned by the plan for the host-scalar entry points: flag block and last good dB row
  int32_t* d_flags = nullptr; float* d_last_row = nullptr;
  // HackRF front end: per-frame mean / power / DC estimate (its own allocation: the large-FFT path uses scratch2)
  void* scratch_dc = nullptr; size_t scratch_dc_bytes = 0;
  // blocked scan of the general trace path: per-block affine terms, start states and extrema
  void* scan_scratch = nullptr; size_t scan_scratch_bytes = 0;
  // large-FFT (two-kernel) tables: inner plan size M = N/256
  double2* d_twin64 = nullptr; float2* d_twin32 = nullptr;    // twiddles of the M-point inner transform
  double2* d_twh64 = nullptr; float2* d_twh32 = nullptr;      // DIF tables of big_head_kernel (passes 0, 1)
  // scratch (grown on demand)
  void* scratch = nullptr; size_t scratch_bytes = 0;
  void* scratch2 = nullptr; size_t scratch2_bytes = 0;
  // host pipeline: H2D on `side`, kernels on the plan's stream, D2H on `back`
  cudaStream_t side = nullptr, back = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  void* d_in[2] = {nullptr, nullptr}; void* d_out[2] = {nullptr, nullptr};
  size_t d_in_bytes = 0, d_out_bytes = 0;
};

// Every entry point that takes a plan runs on the plan's device, whatever device is current in the calling thread.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

static bool is_big(const tdsa_plan* p) {
  return p->log2n > (p->precision == TDSA_PREC_F32 ? MaxLog2<float>::value : MaxLog2<double>::value);
}

// Large transforms: one radix-16 head pass when the remaining M = N/16 fits the single-CTA kernels best tuned
// range (M <= 4096), otherwise two head passes (N = 256*M).
static int big_head_passes(const tdsa_plan* p) { return (p->log2n - 4 <= 12) ? 1 : 2; }

static int ensure_scratch(void** ptr, size_t* have, size_t need) {
  if (*have >= need) return TDSA_OK;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr;
  *have = 0;
  cudaError_t e = cudaMalloc(ptr, need);
  if (e != cudaSuccess) return fail(TDSA_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
  *have = need;
  return TDSA_OK;
}

// np.hanning / np.hamming / np.blackman (numpy/lib/_function_base_impl.py): n = arange(1-M, M, 2)
static void build_window(int id, int norm, int n, std::vector<double>& w) {
  w.assign(n, 1.0);
  if (n == 1) return;
  const double pi = 3.141592653589793238462643383279502884;
  for (int i = 0; i < n; ++i) {
    const double nn = (double)(1 - n + 2 * i);
    const double x = pi * nn / (double)(n - 1);
    switch (id) {
      case TDSA_WINDOW_HANN: w[i] = 0.5 + 0.5 * cos(x); break;
      case TDSA_WINDOW_HAMMING: w[i] = 0.54 + 0.46 * cos(x); break;
      case TDSA_WINDOW_BLACKMAN: w[i] = 0.42 + 0.5 * cos(x) + 0.08 * cos(2.0 * x); break;
      default: w[i] = 1.0; break;
    }
  }
  if (norm == TDSA_NORM_RMS_F32) {   // hackrf_samples.py:314-316 (float32 arithmetic)
    std::vector<float> wf(n);
    double acc = 0.0;
    for (int i = 0; i < n; ++i) { wf[i] = (float)w[i]; acc += (double)(wf[i] * wf[i]); }
    const float rms = sqrtf((float)(acc / n));
    for (int i = 0; i < n; ++i) w[i] = (double)(wf[i] / rms);
  }
}

static int upload_window(tdsa_plan* p) {
  const int n = p->n;
  std::vector<double> w64(n);
  std::vector<float> w32(n);
  for (int i = 0; i < n; ++i) {
    const double s = (i & 1) ? -1.0 : 1.0;   // (-1)^n  <=>  fftshift of the spectrum (n even)
    w64[i] = s * p->window_host[i];
    w32[i] = (float)w64[i];
  }
  // launches that still read the old tables may be queued on the plan's (non-blocking) stream: drain it first
  CK(cudaStreamSynchronize(p->stream));
  CK(cudaMemcpy(p->d_win64, w64.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p->d_win32, w32.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
  if (p->d_wperm64) {
    // fft_wl_kernel: thread tid (engine e = tid >> 8) owns samples n0 = r + 16 c + 256 j of each 4096-sample half;
    // table entry [jj][tid], jj < 16: first half; jj >= 16 (two engines): second half, negated for engine 1
    // (radix-2 DIF: engine 0 transforms x[n] w[n] + x[n+4096] w[n+4096], engine 1 the difference)
    const int nb = p->wl_nb, th = 256 * nb;
    const size_t cnt = (size_t)n * nb;                       // every engine reads the whole window
    std::vector<double> p64(cnt);
    std::vector<float> p32(cnt);
    for (int tid = 0; tid < th; ++tid) {
      int r, c;
      wl_thread_identity(tid, &r, &c);
      const int e = tid >> 8;
      for (int jj = 0; jj < 16 * nb; ++jj) {
        const int half = jj >> 4, j = jj & 15;
        const int idx = 4096 * half + r + 16 * c + 256 * j;
        const double sg = (half == 1 && e == 1) ? -1.0 : 1.0;
        p64[jj * th + tid] = sg * w64[idx];
        p32[jj * th + tid] = (float)(sg * w64[idx]);
      }
    }
    CK(cudaMemcpy(p->d_wperm64, p64.data(), sizeof(double) * cnt, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->d_wperm32, p32.data(), sizeof(float) * cnt, cudaMemcpyHostToDevice));
  }
  p->win_dirty = false;
  return TDSA_OK;
}

// ---- tensor map of a batch of 4096-sample frames: [frame][256 rows][32 floats], 128-byte swizzle -------------
#ifndef TDSA_TMAP_L2_PROMOTION
#define TDSA_TMAP_L2_PROMOTION CU_TENSOR_MAP_L2_PROMOTION_L2_256B
#endif
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return (EncodeTiledFn)ptr;
  }();
  return fn;
}

// Dynamic frame scheduling inside fft_fused_kernel: measured (r01 experiments) 149.5 -> 139.3 us for float64 N=4096 but
// 69.6 -> 75.8 us for float32 N=1024 (more, shorter iterations per CTA), so it is opt-in (TDSA_DYNAMIC=1); N=4096
// takes the warp-local kernel, which always schedules dynamically.
static bool dyn_enabled() {
  static const bool on = [] { const char* e = getenv("TDSA_DYNAMIC"); return e && e[0] == '1'; }();
  return on;
}

static bool welch_cluster_enabled() {   // read at every call so that tests can compare both paths in one process
  const char* e = getenv("TDSA_WELCH_CLUSTER");
  return !(e && e[0] == '0');
}

static bool welch_sub_enabled() {       // TDSA_WELCH_SUB=0: the older cluster / two-kernel paths (kept for comparison)
  const char* e = getenv("TDSA_WELCH_SUB");
  return !(e && e[0] == '0');
}

// TDSA_WELCH_FUSED=1: head pass and tails in ONE kernel, intermediate in an L2-resident ring (0.6 GB of DRAM traffic
// instead of 4.9 GB).  Measured slower than the two launches (round 2, 2047 segments: 1.25 vs 1.09 ms in float64, 0.70 vs
// 0.61 ms in float32: every iteration of a CTA becomes a chain of latency-bound phases -- 64 KB of ring stores, then 64 KB
// of ring loads queued behind them -- with only two CTAs per SM to overlap; profiles/r02_experiments.md), so it is opt-in.
static bool welch_fused_enabled() {
  const char* e = getenv("TDSA_WELCH_FUSED");
  return e && e[0] == '1';
}

static bool wl_enabled() {
  static const bool on = [] { const char* e = getenv("TDSA_WL"); return !(e && e[0] == '0'); }();
  return on;
}

// true when this batch can take a warp-local kernel (epilogue / accumulator combination instantiated, frames
// describable by a tensor map); fills p->tmap: [frame][256 * nb rows][32 floats], 128-byte swizzle, box = 256 rows
static bool wl_prepare(tdsa_plan* p, const void* iq, int64_t n_frames, int64_t stride, int epi, int acc_flags, bool dc) {
  if (!wl_enabled() || p->wl_nb == 0 || !p->d_wperm64 || !wl_supported(p->wl_nb, epi, acc_flags, dc, p->precision == TDSA_PREC_F64)) return false;
  if (((uintptr_t)iq & 15) != 0 || (stride & 1) != 0 || stride <= 0 || n_frames <= 0 || n_frames >= (1 << 30)) return false;
  if (p->tmap_ptr == iq && p->tmap_frames == n_frames && p->tmap_stride == stride) return true;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[3] = {32, (cuuint64_t)256 * p->wl_nb, (cuuint64_t)n_frames};
  const cuuint64_t strides[2] = {128, (cuuint64_t)stride * 8};
  const cuuint32_t box[3] = {32, 256, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(&p->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(iq), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, TDSA_TMAP_L2_PROMOTION,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { p->tmap_ptr = nullptr; return false; }
  p->tmap_ptr = iq; p->tmap_frames = n_frames; p->tmap_stride = stride;
  return true;
}

// exp(-2*pi*i*m/L), folded to the first octant with exact integer arithmetic so that
// symmetric entries are bit-identical and the axes are exact.
static void twiddle(int64_t m, int64_t L, double* re, double* im) {
  const double half_pi = 1.570796326794896619231321691639751442;
  m %= L;
  const int64_t q = (4 * m) / L;          // quadrant
  const int64_t r = 4 * m - q * L;        // angle inside the quadrant = (pi/2) * r / L
  double c, s;
  if (2 * r <= L) { const double a = half_pi * ((double)r / (double)L); c = cos(a); s = sin(a); }
  else { const double a = half_pi * ((double)(L - r) / (double)L); c = sin(a); s = cos(a); }
  double C, S;
  switch (q) {
    case 0: C = c; S = s; break;
    case 1: C = -s; S = c; break;
    case 2: C = -c; S = -s; break;
    default: C = s; S = -c; break;
  }
  *re = C;
  *im = -S;
}

// Pre-twiddle tables of fft_fused_kernel (decimation-in-time placement, see Plan in tdsa_fft.cuh):
// for pass i = 1 .. npass-1 with radix R_i, P^i already-fixed output digits K and input stride
// S_i = N / (P^i * R_i): entry [j][K] = W_N^(j * S_i * K).  logr = 4: P = 16; logr = 3: P = 8.
static void build_twiddles(int log2n, int logr, std::vector<double2>& t64) {
  const int npass = (log2n + logr - 1) / logr;
  const int64_t n = (int64_t)1 << log2n;
  if (npass > 3) {   // decimation-in-frequency placement (see Plan::DIT): table i = [q][c], W_{L_i}^(c*q)
    t64.clear();
    for (int i = 0; i < npass - 1; ++i) {
      const int64_t L = (int64_t)1 << (log2n - logr * i), S = L >> logr;
      for (int q = 0; q < (1 << logr); ++q)
        for (int64_t c = 0; c < S; ++c) {
          double2 w;
          twiddle(c * q, L, &w.x, &w.y);
          t64.push_back(w);
        }
    }
    return;
  }
  const int r_last = (log2n % logr) ? (1 << (log2n % logr)) : (1 << logr);
  t64.clear();
  for (int i = 1; i < npass; ++i) {
    const int64_t r = (i == npass - 1) ? r_last : ((int64_t)1 << logr);
    const int64_t kc = (int64_t)1 << (logr * i);
    const int64_t s = n / (kc * r);
    for (int64_t j = 0; j < r; ++j)
      for (int64_t k = 0; k < kc; ++k) {
        double2 w;
        twiddle((j * s * k) % n, n, &w.x, &w.y);
        t64.push_back(w);
      }
  }
}

// Post-twiddle tables of big_head_kernel (passes 0 and 1 of the N-point in-place DIF):
// table 0: [q][c], c < N/16: W_N^(c*q); table 1: [q][c1], c1 < N/256: W_{N/16}^(c1*q).
static void build_twiddles_head(int log2n, std::vector<double2>& t64) {
  t64.clear();
  for (int i = 0; i < 2; ++i) {
    const int64_t L = (int64_t)1 << (log2n - 4 * i), S = L >> 4;
    for (int q = 0; q < 16; ++q)
      for (int64_t c = 0; c < S; ++c) {
        double2 w;
        twiddle(c * q, L, &w.x, &w.y);
        t64.push_back(w);
      }
  }
}

static int upload_pair(const std::vector<double2>& t64, double2** d64, float2** d32, bool want64, bool want32) {
  const size_t cnt = std::max<size_t>(t64.size(), 1);
  if (want64) {
    CK(cudaMalloc(d64, sizeof(double2) * cnt));
    if (!t64.empty()) CK(cudaMemcpy(*d64, t64.data(), sizeof(double2) * t64.size(), cudaMemcpyHostToDevice));
  }
  if (want32) {
    std::vector<float2> t32(cnt);
    for (size_t i = 0; i < t64.size(); ++i) t32[i] = make_float2((float)t64[i].x, (float)t64[i].y);
    CK(cudaMalloc(d32, sizeof(float2) * cnt));
    if (!t64.empty()) CK(cudaMemcpy(*d32, t32.data(), sizeof(float2) * t64.size(), cudaMemcpyHostToDevice));
  }
  return TDSA_OK;
}

// the float32 and float64 kernels of one size may use different digit widths, so each gets its own table
static int upload_twiddles(int log2n, int logr64, int logr32, double2** d64, float2** d32) {
  std::vector<double2> t64;
  build_twiddles(log2n, logr64, t64);
  int rc = upload_pair(t64, d64, d32, true, logr32 == logr64);
  if (rc || logr32 == logr64) return rc;
  build_twiddles(log2n, logr32, t64);
  return upload_pair(t64, d64, d32, false, true);
}

static EpiParams make_epi(const tdsa_plan* p, float* db, double* lin) {
  EpiParams ep;
  ep.db_out = db;
  ep.lin_out = lin;
  ep.mode = p->mode;
  ep.floor = p->floor;
  ep.scale = (p->mode == TDSA_MODE_PSD) ? 1.0 / (p->fs * (double)p->n) : 1.0;
  return ep;
}

static int run_big(tdsa_plan* p, const void* iq, int64_t n_frames, int64_t stride, const double2* dc, int epi, float* db,
                   double* lin, LaunchInfo* info, bool dry);

// Twiddle tables of the two-engine N = 8192 warp-local kernel (tdsa_fft_wl.cuh): per engine, pass B [j][ka] then the
// last pass [j][kk]; engine 1 evaluates both at the half-integer bin (ka + 1/2, kk + 1/2).
static void build_wl_tables_8192(std::vector<double2>& t64) {
  t64.clear();
  for (int e = 0; e < 2; ++e) {
    for (int j = 0; j < 16; ++j)
      for (int ka = 0; ka < 16; ++ka) {
        double2 w;
        if (e == 0) twiddle((int64_t)j * ka, 256, &w.x, &w.y);
        else twiddle((int64_t)j * (2 * ka + 1), 512, &w.x, &w.y);
        t64.push_back(w);
      }
    for (int j = 0; j < 16; ++j)
      for (int kk = 0; kk < 256; ++kk) {
        double2 w;
        if (e == 0) twiddle((int64_t)j * kk, 4096, &w.x, &w.y);
        else twiddle((int64_t)j * (2 * kk + 1), 8192, &w.x, &w.y);
        t64.push_back(w);
      }
  }
}

// one launch of a warp-local kernel (N = 4096 / 8192); the caller has called wl_prepare
static int run_wl(tdsa_plan* p, const void* iq, int64_t n_frames, int64_t stride, const double2* dc, int epi, float* db,
                  double* lin, int acc_flags, const WlAcc& acc, LaunchInfo* info, bool dry) {
  WlLaunch L;
  L.tmap = &p->tmap; L.sched = WlSched{p->d_sched, p->d_sched + 1}; L.acc = acc; L.nb = p->wl_nb; L.acc_flags = acc_flags;
  L.device = p->device; L.sm_count = p->sm_count;
  cudaError_t e;
  if (p->precision == TDSA_PREC_F32) {
    FftArgs<float> a;
    a.iq = (const float2*)iq; a.n_frames = n_frames; a.frame_stride = stride; a.window = p->d_win32;
    a.tw = p->wl_nb == 2 ? p->d_wltw32 : p->d_tw32; a.dc = dc; a.in_ct = nullptr; a.ep = make_epi(p, db, lin);
    L.wperm = p->d_wperm32;
    e = launch_wl_f32(epi, a, L, p->stream, info, dry);
  } else {
    FftArgs<double> a;
    a.iq = (const float2*)iq; a.n_frames = n_frames; a.frame_stride = stride; a.window = p->d_win64;
    a.tw = p->wl_nb == 2 ? p->d_wltw64 : p->d_tw64; a.dc = dc; a.in_ct = nullptr; a.ep = make_epi(p, db, lin);
    L.wperm = p->d_wperm64;
    e = launch_wl_f64(epi, a, L, p->stream, info, dry);
  }
  if (e != cudaSuccess) return fail(TDSA_ERR_CUDA, "warp-local FFT launch failed (N=%d, acc=%d): %s", p->n, acc_flags, cudaGetErrorString(e));
  return TDSA_OK;
}

// one launch of the fused path for frames already in device memory
static int run_fused(tdsa_plan* p, const void* iq, int64_t n_frames, int64_t stride, const double2* dc, int epi,
                     float* db, double* lin, LaunchInfo* info, bool dry) {
  if (p->win_dirty && !dry) { int rc = upload_window(p); if (rc) return rc; }
  if (is_big(p)) return run_big(p, iq, n_frames, stride, dc, epi, db, lin, info, dry);
  // dry runs (launch geometry queries) describe the warp-local kernel whenever the size has one
  const bool wl = dry ? (wl_enabled() && p->wl_nb != 0 && p->d_wperm64 && wl_supported(p->wl_nb, epi, 0, dc != nullptr, p->precision == TDSA_PREC_F64) && encode_tiled_fn())
                      : wl_prepare(p, iq, n_frames, stride, epi, 0, dc != nullptr);
  if (wl) return run_wl(p, iq, n_frames, stride, dc, epi, db, lin, 0, WlAcc(), info, dry);
  cudaError_t e;
  if (p->precision == TDSA_PREC_F32) {
    FftArgs<float> a;
    a.iq = (const float2*)iq; a.n_frames = n_frames; a.frame_stride = stride;
    a.window = p->d_win32; a.tw = p->d_tw32; a.dc = dc; a.in_ct = nullptr; a.ep = make_epi(p, db, lin);
    a.sched = (dyn_enabled() && n_frames < (1 << 30)) ? p->d_sched : nullptr;
    e = launch_fft_f32(p->log2n, epi, a, p->sm_count, p->stream, info, dry);
  } else {
    FftArgs<double> a;
    a.iq = (const float2*)iq; a.n_frames = n_frames; a.frame_stride = stride;
    a.window = p->d_win64; a.tw = p->d_tw64; a.dc = dc; a.in_ct = nullptr; a.ep = make_epi(p, db, lin);
    a.sched = (dyn_enabled() && n_frames < (1 << 30)) ? p->d_sched : nullptr;
    e = launch_fft_f64(p->log2n, epi, a, p->sm_count, p->stream, info, dry);
  }
  if (e != cudaSuccess) return fail(TDSA_ERR_CUDA, "fused FFT launch failed (N=%d): %s", p->n, cudaGetErrorString(e));
  return TDSA_OK;
}

// scratch for the accumulating epilogue: per-CTA partial rows {sum f64, max f32, min f32} for up to `ctas` CTAs
static int ensure_acc(tdsa_plan* p, int ctas, int64_t n_frames, double** part_sum, float** part_max, float** part_min,
                      double** weights) {
  const size_t per = (size_t)p->n * (sizeof(double) + 2 * sizeof(float));
  int rc = ensure_scratch(&p->acc_parts, &p->acc_parts_bytes, per * (size_t)ctas);
  if (rc) return rc;
  *part_sum = (double*)p->acc_parts;
  *part_max = (float*)(*part_sum + (size_t)ctas * p->n);
  *part_min = *part_max + (size_t)ctas * p->n;
  if (weights) {
    rc = ensure_scratch(&p->acc_weights, &p->acc_weights_bytes, sizeof(double) * (size_t)std::max<int64_t>(n_frames, 1));
    if (rc) return rc;
    *weights = (double*)p->acc_weights;
  }
  return TDSA_OK;
}

// grid the warp-local kernel will use for this many claimable units (= number of partial rows it writes)
static int wl_grid(tdsa_plan* p, int64_t n_units, int acc_flags) {
  LaunchInfo info;
  WlAcc acc;
  if (run_wl(p, nullptr, n_units, p->n, nullptr, kEpiDb, nullptr, nullptr, acc_flags, acc, &info, true) != TDSA_OK) return 0;
  return info.grid;
}

#include "tdsa_big_host.inl"

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

int tdsa_version(void) { return TDSA_VERSION; }
const char* tdsa_last_error(void) { return g_err.c_str(); }
int64_t tdsa_launch_count(void) { return g_launch_count.load(); }

int tdsa_create(int n_fft, int window_id, int window_norm, int mode, double log_floor, double fs, int precision,
                tdsa_handle_t* out) {
  if (!out) return fail(TDSA_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (n_fft < 64 || n_fft > (1 << 20) || (n_fft & (n_fft - 1)))
    return fail(TDSA_ERR_UNSUPPORTED, "n_fft=%d: only powers of two in [64, 1048576] are supported", n_fft);
  if (window_id < 0 || window_id > TDSA_WINDOW_CUSTOM) return fail(TDSA_ERR_INVALID, "bad window_id %d", window_id);
  if (mode < 0 || mode > TDSA_MODE_MAG20) return fail(TDSA_ERR_INVALID, "bad mode %d", mode);
  if (precision != TDSA_PREC_F64 && precision != TDSA_PREC_F32) return fail(TDSA_ERR_INVALID, "bad precision %d", precision);
  tdsa_plan* p = new tdsa_plan();
  p->n = n_fft;
  while ((1 << p->log2n) < n_fft) ++p->log2n;
  p->window_id = window_id; p->window_norm = window_norm; p->mode = mode; p->precision = precision;
  p->floor = log_floor; p->fs = fs;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, p->device);
  if (e != cudaSuccess) { delete p; return fail(TDSA_ERR_CUDA, "no CUDA device: %s", cudaGetErrorString(e)); }
  build_window(window_id == TDSA_WINDOW_CUSTOM ? TDSA_WINDOW_RECT : window_id, window_norm, n_fft, p->window_host);
  int rc = TDSA_OK;
  do {
    if (cudaMalloc(&p->d_win64, sizeof(double) * n_fft) != cudaSuccess ||
        cudaMalloc(&p->d_win32, sizeof(float) * n_fft) != cudaSuccess) { rc = fail(TDSA_ERR_NOMEM, "window alloc failed"); break; }
    rc = upload_twiddles(p->log2n, effective_logr_f64(p->log2n), effective_logr_f32(p->log2n), &p->d_tw64, &p->d_tw32);
    if (rc) break;
    if (cudaMalloc(&p->d_sched, kWlSchedWords * sizeof(int)) != cudaSuccess ||
        cudaMemset(p->d_sched, 0, kWlSchedWords * sizeof(int)) != cudaSuccess) { rc = fail(TDSA_ERR_NOMEM, "scheduler alloc failed"); break; }
    if ((p->log2n == 12 && effective_logr_f64(12) == 4 && effective_logr_f32(12) == 4) || p->log2n == 13) {   // warp-local kernels
      p->wl_nb = p->log2n == 12 ? 1 : 2;
      if (cudaMalloc(&p->d_wperm64, sizeof(double) * n_fft * p->wl_nb) != cudaSuccess ||
          cudaMalloc(&p->d_wperm32, sizeof(float) * n_fft * p->wl_nb) != cudaSuccess) { rc = fail(TDSA_ERR_NOMEM, "warp-local tables alloc failed"); break; }
      if (p->wl_nb == 2) {
        std::vector<double2> tw;
        build_wl_tables_8192(tw);
        rc = upload_pair(tw, &p->d_wltw64, &p->d_wltw32, true, true);
        if (rc) break;
      }
    }
    if (cudaMalloc(&p->d_meta, 8 * sizeof(double)) != cudaSuccess || cudaMalloc(&p->d_flags, kFlagWords * sizeof(int32_t)) != cudaSuccess ||
        cudaMalloc(&p->d_last_row, sizeof(float) * n_fft) != cudaSuccess ||
        cudaMemset(p->d_flags, 0, kFlagWords * sizeof(int32_t)) != cudaSuccess ||
        cudaMemset(p->d_last_row, 0, sizeof(float) * n_fft) != cudaSuccess) { rc = fail(TDSA_ERR_NOMEM, "trace state alloc failed"); break; }
    if (p->log2n > MaxLog2<float>::value || p->log2n > MaxLog2<double>::value) {
      // tables for the inner (N/256)-point transform of the two-kernel path
      rc = upload_twiddles(p->log2n - 4 * big_head_passes(p), 4, 4, &p->d_twin64, &p->d_twin32);
      if (rc) break;
      std::vector<double2> th;
      build_twiddles_head(p->log2n, th);
      rc = upload_pair(th, &p->d_twh64, &p->d_twh32, true, true);
      if (rc) break;
    }
  } while (0);
  if (rc) { tdsa_destroy(p); return rc; }
  *out = p;
  return TDSA_OK;
}

int tdsa_destroy(tdsa_handle_t p) {
  if (!p) return TDSA_OK;
  DeviceGuard guard(p->device);
  cudaFree(p->d_win64); cudaFree(p->d_win32); cudaFree(p->d_tw64); cudaFree(p->d_tw32);
  cudaFree(p->d_wperm64); cudaFree(p->d_wperm32); cudaFree(p->d_sched); cudaFree(p->d_wltw64); cudaFree(p->d_wltw32);
  cudaFree(p->acc_parts); cudaFree(p->acc_weights); cudaFree(p->d_meta); cudaFree(p->d_flags); cudaFree(p->d_last_row);
  cudaFree(p->scratch_dc); cudaFree(p->scan_scratch);
  cudaFree(p->d_twin64); cudaFree(p->d_twin32); cudaFree(p->d_twh64); cudaFree(p->d_twh32);
  cudaFree(p->scratch); cudaFree(p->scratch2);
  for (int i = 0; i < 2; ++i) {
    cudaFree(p->d_in[i]); cudaFree(p->d_out[i]);
    if (p->ev_h2d[i]) cudaEventDestroy(p->ev_h2d[i]);
    if (p->ev_k[i]) cudaEventDestroy(p->ev_k[i]);
    if (p->ev_d2h[i]) cudaEventDestroy(p->ev_d2h[i]);
  }
  if (p->side) cudaStreamDestroy(p->side);
  if (p->back) cudaStreamDestroy(p->back);
  delete p;
  return TDSA_OK;
}

int tdsa_set_stream(tdsa_handle_t p, void* s) {
  if (!p) return fail(TDSA_ERR_INVALID, "null plan");
  p->stream = (cudaStream_t)s;
  return TDSA_OK;
}

int tdsa_set_window(tdsa_handle_t p, int window_id, int window_norm) {
  if (!p) return fail(TDSA_ERR_INVALID, "null plan");
  if (window_id < 0 || window_id >= TDSA_WINDOW_CUSTOM) return fail(TDSA_ERR_INVALID, "bad window_id %d", window_id);
  p->window_id = window_id; p->window_norm = window_norm;
  build_window(window_id, window_norm, p->n, p->window_host);
  p->win_dirty = true;
  return TDSA_OK;
}

int tdsa_set_window_table_host(tdsa_handle_t p, const double* w) {
  if (!p || !w) return fail(TDSA_ERR_INVALID, "null argument");
  p->window_id = TDSA_WINDOW_CUSTOM;
  p->window_host.assign(w, w + p->n);
  p->win_dirty = true;
  return TDSA_OK;
}

int tdsa_get_window_table_host(tdsa_handle_t p, double* w) {
  if (!p || !w) return fail(TDSA_ERR_INVALID, "null argument");
  memcpy(w, p->window_host.data(), sizeof(double) * p->n);
  return TDSA_OK;
}

int tdsa_set_mode(tdsa_handle_t p, int mode, double log_floor, double fs) {
  if (!p) return fail(TDSA_ERR_INVALID, "null plan");
  if (mode < 0 || mode > TDSA_MODE_MAG20) return fail(TDSA_ERR_INVALID, "bad mode %d", mode);
  p->mode = mode; p->floor = log_floor; p->fs = fs;
  return TDSA_OK;
}

int tdsa_set_precision(tdsa_handle_t p, int precision) {
  if (!p) return fail(TDSA_ERR_INVALID, "null plan");
  if (precision != TDSA_PREC_F64 && precision != TDSA_PREC_F32) return fail(TDSA_ERR_INVALID, "bad precision %d", precision);
  p->precision = precision;
  return TDSA_OK;
}

static int check_batch(tdsa_handle_t p, const void* iq, int64_t n_frames, int64_t stride, const void* out) {
  if (!p) return fail(TDSA_ERR_INVALID, "null plan");
  if (n_frames < 0 || stride < 1) return fail(TDSA_ERR_INVALID, "n_frames=%lld frame_stride=%lld", (long long)n_frames, (long long)stride);
  if (n_frames > 0 && (!iq || !out)) return fail(TDSA_ERR_INVALID, "null buffer");
  if (((uintptr_t)iq & 7) != 0) return fail(TDSA_ERR_INVALID, "iq must be 8-byte aligned");
  return TDSA_OK;
}

int tdsa_psd_db_batch(tdsa_handle_t p, const void* iq, int64_t n_frames, int64_t stride, float* db_out) {
  int rc = check_batch(p, iq, n_frames, stride, db_out);
  if (rc) return rc;
  DeviceGuard guard(p->device);
  return run_fused(p, iq, n_frames, stride, nullptr, kEpiDb, db_out, nullptr, nullptr, false);
}

int tdsa_power_linear_batch(tdsa_handle_t p, const void* iq, int64_t n_frames, int64_t stride, double* lin_out) {
  int rc = check_batch(p, iq, n_frames, stride, lin_out);
  if (rc) return rc;
  DeviceGuard guard(p->device);
  return run_fused(p, iq, n_frames, stride, nullptr, kEpiLinear, nullptr, lin_out, nullptr, false);
}

// HackRF front end (hackrf_samples.py:351-365): per-frame mean and power, then the DC tracker; fills dc[] and silent[]
static int run_dc_front(tdsa_plan* p, const float2* iq, int64_t n_frames, int64_t stride, double dc_alpha, double* dc_state,
                        int32_t* silent_out, const double2** dc_out) {
  const size_t need = (size_t)n_frames * (sizeof(double2) * 2 + sizeof(double));
  int rc = ensure_scratch(&p->scratch_dc, &p->scratch_dc_bytes, need);
  if (rc) return rc;
  double2* mean = (double2*)p->scratch_dc;
  double2* dc = mean + n_frames;
  double* pw = (double*)(dc + n_frames);
  const int grid = (int)std::min<int64_t>(n_frames, (int64_t)p->sm_count * 8);
  frame_stats_kernel<<<grid, 256, 0, p->stream>>>(iq, n_frames, stride, p->n, mean, pw);
  count_launch();
  dc_scan_kernel<<<1, 32, 0, p->stream>>>(mean, pw, n_frames, dc_alpha, dc_state, dc, silent_out);
  count_launch();
  CK(cudaGetLastError());
  *dc_out = dc;
  return TDSA_OK;
}

int tdsa_psd_db_batch_dc(tdsa_handle_t p, const void* iq, int64_t n_frames, int64_t stride, double dc_alpha,
                         double* dc_state, int32_t* silent_out, float* db_out) {
  int rc = check_batch(p, iq, n_frames, stride, db_out);
  if (rc) return rc;
  if (!dc_state) return fail(TDSA_ERR_INVALID, "dc_state is NULL");
  if (n_frames == 0) return TDSA_OK;
  DeviceGuard guard(p->device);
  const double2* dc = nullptr;
  rc = run_dc_front(p, (const float2*)iq, n_frames, stride, dc_alpha, dc_state, silent_out, &dc);
  if (rc) return rc;
  return run_fused(p, iq, n_frames, stride, dc, kEpiDb, db_out, nullptr, nullptr, false);
}

// frames below this count are not worth the fused epilogue's extra launches (weights / finish kernels)
static int64_t fused_min_frames() {
  static const int64_t v = [] { const char* e = getenv("TDSA_FUSED_MIN_FRAMES"); return (int64_t)(e ? atoll(e) : 64); }();
  return v;
}
// linear rows of the general path are produced and consumed chunk by chunk so that they never leave the L2
// Float64 rows of the general trace path are produced and scanned in chunks that stay in the 126 MB L2.  Measured (round 2,
// 8192 x 4096 frames, averaging + hold + every row, f64 / f32 plan): 16 MB 718 / 647 us, 32 MB 532 / 484, 64 MB 440 / 383,
// 96 MB 425 / 371, 128 MB 407 / 353, one 256 MB chunk 395 / 340: six launches per chunk dominate, so the largest chunk that
// still fits the L2 comfortably is the default.
static int64_t scan_chunk_bytes() {
  static const int64_t v = [] { const char* e = getenv("TDSA_SCAN_CHUNK_MB"); return (int64_t)(e ? atoll(e) : 64) << 20; }();
  return v;
}

// Frames -> device-resident trace state. Three shapes:
//   fused running average (last row only, no holds): weighted sum in the FFT kernel's accumulating epilogue;
//   fused holds (no averaging): dB rows + running max/min of |X|^2 in the same epilogue;
//   general: float64 linear rows of an L2-sized chunk -> frame-ordered scan (averager + dB + holds per frame).
static int avg_hold_dev_impl(tdsa_handle_t p, const void* iq, int64_t n_frames, int64_t stride, bool with_dc,
                             double dc_alpha, double* dc_state, int32_t* silent_out, int avg_mode, int avg_n,
                             double* avg_state, float* max_hold, float* min_hold, int32_t* flags, float* last_row,
                             int last_only, float* db_out) {
  int rc = check_batch(p, iq, n_frames, stride, db_out);
  if (rc) return rc;
  const bool averaging = avg_mode != TDSA_AVG_OFF && avg_n > 1;
  if (!flags) return fail(TDSA_ERR_INVALID, "flags block (device int32[8]) is required");
  if (averaging && !avg_state) return fail(TDSA_ERR_INVALID, "averaging needs avg_state");
  if (p->mode == TDSA_MODE_MAG20 && averaging) return fail(TDSA_ERR_INVALID, "mag20 branch is never averaged (hackrf_samples.py:378-383)");
  if (with_dc && (!dc_state || !silent_out)) return fail(TDSA_ERR_INVALID, "dc path needs dc_state and silent_out");
  if (n_frames == 0) return TDSA_OK;
  DeviceGuard guard(p->device);
  if (p->win_dirty) { rc = upload_window(p); if (rc) return rc; }
  const double scale = (p->mode == TDSA_MODE_PSD) ? 1.0 / (p->fs * (double)p->n) : 1.0;
  const int fin_grid = (p->n + 255) / 256;
  const bool big_enough = n_frames >= fused_min_frames() && !with_dc && !is_big(p);
  // ---- fused running average -------------------------------------------------------------------------------
  if (big_enough && averaging && last_only && !max_hold && !min_hold &&
      wl_prepare(p, iq, n_frames, stride, kEpiDb, kAccAvg, false)) {
    const int grid = wl_grid(p, n_frames, kAccAvg);
    double *ps, *wts; float *pmx, *pmn;
    rc = ensure_acc(p, grid, n_frames, &ps, &pmx, &pmn, &wts);
    if (rc) return rc;
    avg_weights_kernel<<<(int)std::min<int64_t>((n_frames + 255) / 256, 1024), 256, 0, p->stream>>>(avg_mode, avg_n, n_frames, flags, wts, p->d_meta);
    count_launch();
    WlAcc acc;
    acc.weight = wts; acc.part_sum = ps;
    rc = run_wl(p, iq, n_frames, stride, nullptr, kEpiDb, nullptr, nullptr, kAccAvg, acc, nullptr, false);
    if (rc) return rc;
    avg_finish_kernel<<<partial_grid(p->n), 256, 0, p->stream>>>(ps, grid, p->n, p->d_meta, scale, p->floor, p->mode, avg_state, flags,
                                                      n_frames, db_out, last_row);
    count_launch();
    CK(cudaGetLastError());
    return TDSA_OK;
  }
  // ---- fused holds on un-averaged rows ------------------------------------------------------------------------
  if (big_enough && !averaging && (max_hold || min_hold) && wl_prepare(p, iq, n_frames, stride, kEpiDb, kAccHold, false)) {
    const int grid = wl_grid(p, n_frames, kAccHold);
    double* ps; float *pmx, *pmn;
    rc = ensure_acc(p, grid, n_frames, &ps, &pmx, &pmn, nullptr);
    if (rc) return rc;
    flags_prepare_kernel<<<1, 32, 0, p->stream>>>(flags, p->d_meta, avg_mode, avg_n, n_frames, max_hold != nullptr, min_hold != nullptr);
    count_launch();
    WlAcc acc;
    acc.part_max = pmx; acc.part_min = pmn; acc.only_row = last_only ? n_frames - 1 : -1;
    rc = run_wl(p, iq, n_frames, stride, nullptr, kEpiDb, db_out, nullptr, kAccHold, acc, nullptr, false);
    if (rc) return rc;
    hold_finish_kernel<<<partial_grid(p->n), 256, 0, p->stream>>>(pmx, pmn, grid, p->n, p->d_meta, scale, p->floor, p->mode, max_hold, min_hold);
    count_launch();
    if (last_row) {
      CK(cudaMemcpyAsync(last_row, last_only ? db_out : db_out + (n_frames - 1) * p->n, sizeof(float) * p->n,
                         cudaMemcpyDeviceToDevice, p->stream));
    }
    CK(cudaGetLastError());
    return TDSA_OK;
  }
  // ---- general path ---------------------------------------------------------------------------------------------
  const int64_t chunk_max = std::max<int64_t>(1, scan_chunk_bytes() / ((int64_t)p->n * 8));
  for (int64_t f0 = 0; f0 < n_frames; f0 += chunk_max) {
    const int64_t nf = std::min(chunk_max, n_frames - f0);
    rc = ensure_scratch(&p->scratch, &p->scratch_bytes, (size_t)std::min(chunk_max, n_frames) * p->n * sizeof(double));
    if (rc) return rc;
    double* lin = (double*)p->scratch;
    const float2* src = (const float2*)iq + f0 * stride;
    const double2* dc = nullptr;
    if (with_dc) {
      rc = run_dc_front(p, src, nf, stride, dc_alpha, dc_state, silent_out + f0, &dc);
      if (rc) return rc;
    }
    rc = run_fused(p, src, nf, stride, dc, kEpiLinear, nullptr, lin, nullptr, false);
    if (rc) return rc;
    if (!with_dc && nf >= 4 * kScanBlock) {        // every frame live: the scan runs parallel over frame blocks too
      const int nblk = (int)((nf + kScanBlock - 1) / kScanBlock);
      const size_t w = (size_t)p->n;
      rc = ensure_scratch(&p->scan_scratch, &p->scan_scratch_bytes,
                          (size_t)nblk * w * (2 * sizeof(double) + 2 * sizeof(float)) + (size_t)nblk * sizeof(double));
      if (rc) return rc;
      TraceBlockArgs b;
      b.lin = lin; b.n_frames = nf; b.width = p->n; b.avg_mode = avg_mode; b.avg_n = avg_n; b.flags = flags;
      b.avg_state = avg_state; b.max_hold = max_hold; b.min_hold = min_hold; b.last_row = last_row; b.last_only = last_only;
      b.db_out = last_only ? db_out : db_out + f0 * p->n; b.floor = p->floor; b.mode = p->mode;
      b.blk_b = (double*)p->scan_scratch; b.blk_start = b.blk_b + (size_t)nblk * w;
      b.blk_a = b.blk_start + (size_t)nblk * w;
      b.blk_max = (float*)(b.blk_a + nblk); b.blk_min = b.blk_max + (size_t)nblk * w;
      const dim3 g2((unsigned)fin_grid, (unsigned)nblk);
      if (averaging) {
        scan_block_affine_kernel<<<g2, 256, 0, p->stream>>>(b);
        count_launch();
        scan_chain_kernel<<<fin_grid, 256, 0, p->stream>>>(b, nblk);
        count_launch();
      }
      scan_block_emit_kernel<<<g2, 256, 0, p->stream>>>(b);
      count_launch();
      if (max_hold || min_hold) {
        scan_holds_kernel<<<fin_grid, 256, 0, p->stream>>>(b, nblk);
        count_launch();
      }
      trace_flags_after_scan_kernel<<<1, 256, 0, p->stream>>>(flags, nullptr, nf, avg_mode, avg_n, max_hold != nullptr,
                                                            min_hold != nullptr, f0 == 0);
      count_launch();
      CK(cudaGetLastError());
      continue;
    }
    TraceScanDevArgs a;
    a.lin = lin; a.skip = with_dc ? silent_out + f0 : nullptr;
    a.n_frames = nf; a.width = p->n; a.avg_mode = avg_mode; a.avg_n = avg_n; a.flags = flags;
    a.avg_state = avg_state; a.max_hold = max_hold; a.min_hold = min_hold; a.last_row = last_row;
    a.last_only = last_only;
    // last_only: every chunk writes its last row into db_out[0..N); the final chunk's survives
    a.db_out = last_only ? db_out : db_out + f0 * p->n;
    a.floor = p->floor; a.mode = p->mode;
    trace_scan_dev_kernel<<<fin_grid, 256, 0, p->stream>>>(a);
    count_launch();
    trace_flags_after_scan_kernel<<<1, 256, 0, p->stream>>>(flags, a.skip, nf, avg_mode, avg_n, max_hold != nullptr,
                                                          min_hold != nullptr, f0 == 0);
    count_launch();
    CK(cudaGetLastError());
  }
  return TDSA_OK;
}

int tdsa_psd_db_avg_hold_dev(tdsa_handle_t p, const void* iq, int64_t n_frames, int64_t stride, int use_dc, double dc_alpha,
                             double* dc_state, int32_t* silent_out, int avg_mode, int avg_n, double* avg_state,
                             float* max_hold, float* min_hold, int32_t* flags_dev, float* last_row_dev, int last_only,
                             float* db_out) {
  if (!p) return fail(TDSA_ERR_INVALID, "null plan");
  return avg_hold_dev_impl(p, iq, n_frames, stride, use_dc != 0, dc_alpha, dc_state, silent_out, avg_mode, avg_n, avg_state,
                           max_hold, min_hold, flags_dev, last_row_dev, last_only, db_out);
}

// Host-scalar flavour: the plan's own flag block is seeded from the host values, the device path runs, and the host
// values are brought up to date: arithmetically when every frame is live, by reading the block back (one stream
// synchronisation) on the HackRF path, where the number of live frames is only known on the device.
static int avg_hold_host_scalars(tdsa_handle_t p, const void* iq, int64_t n_frames, int64_t stride, bool with_dc,
                                 double dc_alpha, double* dc_state, int32_t* silent_out, int avg_mode, int avg_n,
                                 double* avg_state, int32_t* count_state_host, float* max_hold, float* min_hold,
                                 int32_t* hold_valid_host, int last_only, float* db_out) {
  if (!p) return fail(TDSA_ERR_INVALID, "null plan");
  const bool averaging = avg_mode != TDSA_AVG_OFF && avg_n > 1;
  if (averaging && !count_state_host) return fail(TDSA_ERR_INVALID, "averaging needs avg_state and count_state");
  if ((max_hold || min_hold) && !hold_valid_host) return fail(TDSA_ERR_INVALID, "holds need hold_valid");
  if (n_frames <= 0) return n_frames == 0 ? TDSA_OK : fail(TDSA_ERR_INVALID, "n_frames < 0");
  DeviceGuard guard(p->device);
  int32_t h[kFlagWords] = {};
  h[kFlagCount] = averaging ? *count_state_host : 0;
  h[kFlagMaxValid] = hold_valid_host ? hold_valid_host[0] : 0;
  h[kFlagMinValid] = hold_valid_host ? hold_valid_host[1] : 0;
  CK(cudaMemcpyAsync(p->d_flags, h, sizeof h, cudaMemcpyHostToDevice, p->stream));
  int rc = avg_hold_dev_impl(p, iq, n_frames, stride, with_dc, dc_alpha, dc_state, silent_out, avg_mode, avg_n, avg_state,
                             max_hold, min_hold, p->d_flags, p->d_last_row, last_only, db_out);
  if (rc) return rc;
  if (with_dc) {
    CK(cudaMemcpyAsync(h, p->d_flags, sizeof h, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
  } else {            // TraceAverager._count after n_frames more frames (signal_processing.py:46-58)
    if (averaging) h[kFlagCount] = avg_mode == TDSA_AVG_LIN ? (int32_t)std::min<int64_t>(avg_n, (int64_t)h[kFlagCount] + n_frames)
                                                            : std::max(h[kFlagCount], 1);
    if (max_hold) h[kFlagMaxValid] = 1;
    if (min_hold) h[kFlagMinValid] = 1;
  }
  if (averaging) *count_state_host = h[kFlagCount];
  if (hold_valid_host) { hold_valid_host[0] = h[kFlagMaxValid]; hold_valid_host[1] = h[kFlagMinValid]; }
  return TDSA_OK;
}

int tdsa_psd_db_avg_hold(tdsa_handle_t p, const void* iq, int64_t n_frames, int64_t stride, int avg_mode, int avg_n,
                         double* avg_state, int32_t* count_state_host, float* max_hold, float* min_hold,
                         int32_t* hold_valid_host, int last_only, float* db_out) {
  return avg_hold_host_scalars(p, iq, n_frames, stride, false, 1.0, nullptr, nullptr, avg_mode, avg_n, avg_state,
                               count_state_host, max_hold, min_hold, hold_valid_host, last_only, db_out);
}

int tdsa_psd_db_avg_hold_dc(tdsa_handle_t p, const void* iq, int64_t n_frames, int64_t stride, double dc_alpha,
                            double* dc_state, int32_t* silent_out, int avg_mode, int avg_n, double* avg_state,
                            int32_t* count_state_host, float* max_hold, float* min_hold, int32_t* hold_valid_host,
                            int last_only, float* db_out) {
  return avg_hold_host_scalars(p, iq, n_frames, stride, true, dc_alpha, dc_state, silent_out, avg_mode, avg_n, avg_state,
                               count_state_host, max_hold, min_hold, hold_valid_host, last_only, db_out);
}

// Group means in the FFT kernel's accumulating epilogue.  A group is the unit CTAs claim; when there are too few groups
// to fill the GPU (config 4 sharded over 8 ranks: 38 groups for 148 SMs) every group is split into `split` units whose
// float64 sums are combined by group_finish_kernel.  Rows go to db_rows or, with peers, into every rank's table.
static int run_group_mean(tdsa_plan* p, const void* iq, int64_t n_groups, int64_t frames, float* db_rows,
                          const uint64_t* peers, int n_peers, int64_t row0) {
  const int64_t n = p->n;
  // split = units per group.  Cost model (microseconds): the slowest CTA runs ceil(units / CTAs) units of frames/split
  // frames each (measured ~0.92 ns per point and frame in float64, ~0.52 ns in float32 on one CTA slot), and split > 1
  // adds a float64 partial row per unit that is written and read once (~4 TB/s through L2).
  static const int force_split = [] { const char* e = getenv("TDSA_GROUP_SPLIT"); return e ? atoi(e) : 0; }();
  const int64_t ctas = (int64_t)p->sm_count * (p->wl_nb == 1 ? 2 : 1);
  const double t_frame = (double)n * (p->precision == TDSA_PREC_F64 ? 0.92e-3 : 0.52e-3);
  int split = 1;
  double best = 0.0;
  for (int sp = 1; sp <= frames && sp <= 64; sp *= 2) {
    if (frames % sp) break;
    const int64_t rounds = (n_groups * sp + ctas - 1) / ctas;
    const double cost = (double)rounds * (double)(frames / sp) * t_frame +
                        (sp > 1 ? (double)(n_groups * sp) * (double)n * 16.0 / 4.0e6 : 0.0);
    if (sp == 1 || cost < best) { best = cost; split = sp; }
  }
  if (force_split > 0 && frames % force_split == 0) split = force_split;
  WlAcc acc;
  acc.group = (int)(frames / split);
  if (split == 1) {
    acc.group_db = db_rows; acc.n_peers = n_peers; acc.peer_row0 = row0;
    for (int i = 0; i < n_peers; ++i) acc.peer_rows[i] = (float*)(uintptr_t)peers[i];
    return run_wl(p, iq, n_groups * frames, n, nullptr, kEpiDb, nullptr, nullptr, kAccGroupMean, acc, nullptr, false);
  }
  int rc = ensure_scratch(&p->scratch, &p->scratch_bytes, (size_t)(n_groups * split * n) * sizeof(double));
  if (rc) return rc;
  acc.unit_sum = (double*)p->scratch;
  rc = run_wl(p, iq, n_groups * frames, n, nullptr, kEpiDb, nullptr, nullptr, kAccGroupMean, acc, nullptr, false);
  if (rc) return rc;
  GroupFinishArgs g;
  g.unit_sum = acc.unit_sum; g.n_groups = n_groups; g.width = n; g.split = split; g.frames = (int)frames;
  g.scale = (p->mode == TDSA_MODE_PSD) ? 1.0 / (p->fs * (double)n) : 1.0; g.floor = p->floor; g.mode = p->mode;
  g.group_db = db_rows; g.n_peers = n_peers; g.peer_row0 = row0;
  for (int i = 0; i < 8; ++i) g.peer_rows[i] = i < n_peers ? (float*)(uintptr_t)peers[i] : nullptr;
  const int64_t total = n_groups * n;
  group_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, p->stream>>>(g);
  count_launch();
  CK(cudaGetLastError());
  return TDSA_OK;
}

int tdsa_group_avg_db(tdsa_handle_t p, const void* iq, int64_t n_groups, int64_t frames_per_group, float* db_rows) {
  if (!p) return fail(TDSA_ERR_INVALID, "null plan");
  if (n_groups < 0 || frames_per_group < 1) return fail(TDSA_ERR_INVALID, "bad group geometry");
  if (p->mode == TDSA_MODE_MAG20) return fail(TDSA_ERR_INVALID, "group average works on power; use power or psd mode");
  if (n_groups == 0) return TDSA_OK;
  if (!iq || !db_rows) return fail(TDSA_ERR_INVALID, "null buffer");
  DeviceGuard guard(p->device);
  const int64_t n = p->n;
  // N = 4096 / 8192: the group mean is formed in the FFT kernel's accumulating epilogue (one launch, one dB row
  // per group out; no linear rows at all)
  if (frames_per_group < (1 << 20) && n_groups * frames_per_group < (1 << 30)) {
    if (p->win_dirty) { int rcw = upload_window(p); if (rcw) return rcw; }
    if (wl_prepare(p, iq, n_groups * frames_per_group, n, kEpiDb, kAccGroupMean, false))
      return run_group_mean(p, iq, n_groups, frames_per_group, db_rows, nullptr, 0, 0);
  }
  const int64_t per_group = frames_per_group * n * (int64_t)sizeof(double);
  // groups per launch: the float64 rows of a chunk are written by the FFT kernel and read straight back by the
  // group-mean kernel, so a chunk that fits the 126 MB L2 keeps that traffic off HBM (TDSA_GROUP_CHUNK_MB to tune)
  static const int64_t chunk_mb = [] { const char* e = getenv("TDSA_GROUP_CHUNK_MB"); return (int64_t)(e ? atoi(e) : 48); }();
  const int64_t chunk = std::max<int64_t>(1, (std::max<int64_t>(chunk_mb, 1) << 20) / per_group);
  for (int64_t g0 = 0; g0 < n_groups; g0 += chunk) {
    const int64_t ng = std::min(chunk, n_groups - g0);
    int rc = ensure_scratch(&p->scratch, &p->scratch_bytes, (size_t)(ng * per_group));
    if (rc) return rc;
    double* lin = (double*)p->scratch;
    rc = run_fused(p, (const float2*)iq + g0 * frames_per_group * n, ng * frames_per_group, n, nullptr, kEpiLinear, nullptr,
                   lin, nullptr, false);
    if (rc) return rc;
    const int64_t total = ng * n;
    group_mean_db_kernel<<<(unsigned)((total + 255) / 256), 256, 0, p->stream>>>(lin, ng, frames_per_group, n, p->floor,
                                                                              p->mode, db_rows + g0 * n);
    count_launch();
    CK(cudaGetLastError());
  }
  return TDSA_OK;
}

int tdsa_group_avg_db_peers(tdsa_handle_t p, const void* iq, int64_t n_groups, int64_t frames_per_group, int64_t row_offset,
                            const uint64_t* peer_rows_host, int n_peers) {
  if (!p) return fail(TDSA_ERR_INVALID, "null plan");
  if (n_groups < 0 || frames_per_group < 1 || row_offset < 0) return fail(TDSA_ERR_INVALID, "bad group geometry");
  if (!peer_rows_host || n_peers < 1 || n_peers > kMaxPeers) return fail(TDSA_ERR_INVALID, "1..%d peer buffers", kMaxPeers);
  if (p->mode == TDSA_MODE_MAG20) return fail(TDSA_ERR_INVALID, "group average works on power; use power or psd mode");
  if (n_groups == 0) return TDSA_OK;
  if (!iq) return fail(TDSA_ERR_INVALID, "null buffer");
  DeviceGuard guard(p->device);
  if (p->win_dirty) { int rcw = upload_window(p); if (rcw) return rcw; }
  if (frames_per_group >= (1 << 20) || n_groups * frames_per_group >= (1 << 30) ||
      !wl_prepare(p, iq, n_groups * frames_per_group, p->n, kEpiDb, kAccGroupMean, false))
    return fail(TDSA_ERR_UNSUPPORTED, "peer-store group mean needs N = 4096 or 8192 and 16-byte aligned frames");
  return run_group_mean(p, iq, n_groups, frames_per_group, nullptr, peer_rows_host, n_peers, row_offset);
}

int tdsa_welch(tdsa_handle_t p, const void* iq_stream, int64_t n_samples, int64_t hop, float* avg_db, float* peak_db) {
  if (!p || !iq_stream || !avg_db || !peak_db) return fail(TDSA_ERR_INVALID, "null argument");
  if (hop < 1 || n_samples < p->n) return fail(TDSA_ERR_INVALID, "need hop >= 1 and n_samples >= n_fft");
  if (p->mode == TDSA_MODE_MAG20) return fail(TDSA_ERR_INVALID, "welch averages power; use power or psd mode");
  const int64_t nseg = (n_samples - p->n) / hop + 1;
  const int64_t n = p->n;
  DeviceGuard guard(p->device);
  // segments that fit one CTA of the warp-local kernel: mean and peak accumulate in its epilogue (TMEM), one launch
  if (!is_big(p) && nseg < (1 << 30)) {
    if (p->win_dirty) { int rcw = upload_window(p); if (rcw) return rcw; }
    if (wl_prepare(p, iq_stream, nseg, hop, kEpiDb, kAccWelch, false)) {
      const int grid = wl_grid(p, nseg, kAccWelch);
      double* ps; float *pmx, *pmn;
      int rca = ensure_acc(p, grid, nseg, &ps, &pmx, &pmn, nullptr);
      if (rca) return rca;
      WlAcc acc;
      acc.part_sum = ps; acc.part_max = pmx;
      rca = run_wl(p, iq_stream, nseg, hop, nullptr, kEpiDb, nullptr, nullptr, kAccWelch, acc, nullptr, false);
      if (rca) return rca;
      const double scale = (p->mode == TDSA_MODE_PSD) ? 1.0 / (p->fs * (double)n) : 1.0;
      welch_acc_finish_kernel<<<partial_grid(n), 256, 0, p->stream>>>(ps, pmx, grid, n, nseg, scale, p->floor, p->mode,
                                                                           avg_db, peak_db);
      count_launch();
      CK(cudaGetLastError());
      return TDSA_OK;
    }
  }
  // 65536-point segments, opt-in (TDSA_WELCH_FUSED=1): ONE kernel (fft_wl_kernel, kAccFused).  Groups of sixteen CTAs own whole segments: each
  // CTA runs one column block of the radix-16 head pass into the group's ring of kFusedRing segments (complex T, stays in L2), then
  // the 4096-point sub-transform of its class with the Welch sum and peak in tensor memory; one finish kernel.
  if (p->log2n == 16 && welch_sub_enabled() && welch_fused_enabled() && nseg < (1 << 26)) {
    if (p->win_dirty) { int rcw = upload_window(p); if (rcw) return rcw; }
    const bool f32 = p->precision == TDSA_PREC_F32;
    const size_t csz = f32 ? sizeof(float2) : sizeof(double2);
    WlLaunch L;
    L.tmap = &p->tmap; L.sched = WlSched{p->d_sched, p->d_sched + 1}; L.nb = 1; L.acc_flags = kAccWelchFused;
    L.device = p->device; L.sm_count = p->sm_count; L.wperm = nullptr;
    LaunchInfo info;
    cudaError_t e;
    if (f32) { FftArgs<float> t{}; t.n_frames = nseg * 16; e = launch_wl_f32(kEpiDb, t, L, p->stream, &info, true); }
    else { FftArgs<double> t{}; t.n_frames = nseg * 16; e = launch_wl_f64(kEpiDb, t, L, p->stream, &info, true); }
    if (e != cudaSuccess) return fail(TDSA_ERR_CUDA, "welch fused query failed: %s", cudaGetErrorString(e));
    const int grid = info.grid, groups = grid / 16;
    if (info.ctas_per_sm * p->sm_count >= grid) {               // every CTA resident at once, or the groups could wait forever
      int rcs = ensure_scratch(&p->scratch2, &p->scratch2_bytes, (size_t)groups * kFusedRing * n * csz);
      if (rcs) return rcs;
      const size_t per = (size_t)4096 * (sizeof(double) + sizeof(float));
      rcs = ensure_scratch(&p->acc_parts, &p->acc_parts_bytes, per * (size_t)grid);
      if (rcs) return rcs;
      double* ps = (double*)p->acc_parts;
      float* pmx = (float*)(ps + (size_t)grid * 4096);
      WlAcc acc;
      acc.part_sum = ps; acc.part_max = pmx;
      acc.fused_iq = (const float2*)iq_stream; acc.fused_hop = hop; acc.fused_nseg = nseg;
      acc.fused_tw = f32 ? (const void*)p->d_twh32 : (const void*)p->d_twh64;
      acc.fused_y = p->scratch2; acc.fused_cnt = p->d_sched + 18;
      L.acc = acc;
      CK(cudaMemsetAsync(p->d_sched + 18 + 2 * groups, 0, sizeof(int), p->stream));   // the error word of an earlier call
      if (f32) {
        FftArgs<float> t{};
        t.n_frames = nseg * 16; t.window = p->d_win32; t.tw = p->d_twin32; t.ep = make_epi(p, nullptr, nullptr);
        e = launch_wl_f32(kEpiDb, t, L, p->stream, nullptr, false);
      } else {
        FftArgs<double> t{};
        t.n_frames = nseg * 16; t.window = p->d_win64; t.tw = p->d_twin64; t.ep = make_epi(p, nullptr, nullptr);
        e = launch_wl_f64(kEpiDb, t, L, p->stream, nullptr, false);
      }
      if (e != cudaSuccess) return fail(TDSA_ERR_CUDA, "welch fused launch failed: %s", cudaGetErrorString(e));
      const double scale = (p->mode == TDSA_MODE_PSD) ? 1.0 / (p->fs * (double)n) : 1.0;
      welch_sub_finish_kernel<<<65536 / 256, 256, 0, p->stream>>>(ps, pmx, grid, nseg, scale, p->floor, p->mode, avg_db, peak_db,
                                                                  p->d_sched + 18 + 2 * groups);
      count_launch();
      CK(cudaGetLastError());
      return TDSA_OK;
    }
  }
  // 65536-point segments, default: two launches: one radix-16 head pass over every segment (big_head_wl_kernel: windowed samples ->
  // sixteen 4096-point sub-transform inputs per segment, complex T, in the tail's thread order), then ONE launch of the
  // warp-local kernel over all 16 * nseg sub-transforms with the Welch sum and peak in tensor memory (kAccSub).  The
  // intermediate goes through HBM once each way (16 B/point in float64) under the tail's arithmetic; no linear rows.
  if (p->log2n == 16 && welch_sub_enabled() && nseg < (1 << 26)) {
    if (p->win_dirty) { int rcw = upload_window(p); if (rcw) return rcw; }
    const bool f32 = p->precision == TDSA_PREC_F32;
    const size_t csz = f32 ? sizeof(float2) : sizeof(double2);
    const int64_t chunk = std::min<int64_t>(nseg, 2048);                 // segments per head + tail pair (2 GiB of float64)
    const int64_t n_chunks = (nseg + chunk - 1) / chunk;
    int rcs = ensure_scratch(&p->scratch2, &p->scratch2_bytes, (size_t)chunk * n * csz);
    if (rcs) return rcs;
    WlLaunch L;
    L.tmap = &p->tmap; L.sched = WlSched{p->d_sched, p->d_sched + 1}; L.nb = 1; L.acc_flags = kAccWelchSub;
    L.device = p->device; L.sm_count = p->sm_count; L.wperm = nullptr;
    // grid of the tail = partial rows per chunk
    LaunchInfo info;
    {
      cudaError_t eq;
      if (f32) { FftArgs<float> t{}; t.n_frames = chunk * 16; eq = launch_wl_f32(kEpiDb, t, L, p->stream, &info, true); }
      else { FftArgs<double> t{}; t.n_frames = chunk * 16; eq = launch_wl_f64(kEpiDb, t, L, p->stream, &info, true); }
      if (eq != cudaSuccess) return fail(TDSA_ERR_CUDA, "welch tail query failed: %s", cudaGetErrorString(eq));
    }
    const int grid = info.grid;
    const size_t per = (size_t)4096 * (sizeof(double) + sizeof(float));
    rcs = ensure_scratch(&p->acc_parts, &p->acc_parts_bytes, per * (size_t)grid * (size_t)n_chunks);
    if (rcs) return rcs;
    double* ps = (double*)p->acc_parts;
    float* pmx = (float*)(ps + (size_t)grid * n_chunks * 4096);
    for (int64_t ci = 0; ci < n_chunks; ++ci) {
      const int64_t s0 = ci * chunk, ns = std::min(chunk, nseg - s0);
      cudaError_t e;
      WlAcc acc;
      acc.part_sum = ps + (size_t)ci * grid * 4096; acc.part_max = pmx + (size_t)ci * grid * 4096;
      L.acc = acc;
      if (ns < chunk) {      // a shorter last chunk may use a smaller grid: its unused partial rows must read as empty
        welch_sub_clear_kernel<<<(grid * 4096 + 255) / 256, 256, 0, p->stream>>>(acc.part_sum, acc.part_max, (int64_t)grid * 4096);
        count_launch();
      }
      if (f32) {
        BigArgs<float> a;
        a.iq = (const float2*)iq_stream + s0 * hop; a.n_frames = ns; a.frame_stride = hop; a.window = p->d_win32;
        a.tw = p->d_twh32; a.dc = nullptr; a.y = (float2*)p->scratch2; a.log2n = p->log2n;
        e = launch_big_head_f32(a, p->sm_count, p->stream, 0);
        if (e == cudaSuccess) {
          FftArgs<float> t;
          t.iq = nullptr; t.n_frames = ns * 16; t.frame_stride = 0; t.window = nullptr; t.tw = p->d_twin32; t.dc = nullptr;
          t.in_ct = (const float2*)p->scratch2; t.ep = make_epi(p, nullptr, nullptr);
          e = launch_wl_f32(kEpiDb, t, L, p->stream, nullptr, false);
        }
      } else {
        BigArgs<double> a;
        a.iq = (const float2*)iq_stream + s0 * hop; a.n_frames = ns; a.frame_stride = hop; a.window = p->d_win64;
        a.tw = p->d_twh64; a.dc = nullptr; a.y = (double2*)p->scratch2; a.log2n = p->log2n;
        e = launch_big_head_f64(a, p->sm_count, p->stream, 0);
        if (e == cudaSuccess) {
          FftArgs<double> t;
          t.iq = nullptr; t.n_frames = ns * 16; t.frame_stride = 0; t.window = nullptr; t.tw = p->d_twin64; t.dc = nullptr;
          t.in_ct = (const double2*)p->scratch2; t.ep = make_epi(p, nullptr, nullptr);
          e = launch_wl_f64(kEpiDb, t, L, p->stream, nullptr, false);
        }
      }
      if (e != cudaSuccess) return fail(TDSA_ERR_CUDA, "welch head/tail launch failed: %s", cudaGetErrorString(e));
    }
    const double scale = (p->mode == TDSA_MODE_PSD) ? 1.0 / (p->fs * (double)n) : 1.0;
    welch_sub_finish_kernel<<<65536 / 256, 256, 0, p->stream>>>(ps, pmx, grid * (int)n_chunks, nseg, scale, p->floor, p->mode,
                                                                avg_db, peak_db, nullptr);
    count_launch();
    CK(cudaGetLastError());
    return TDSA_OK;
  }
  // 65536-point segments: one cluster kernel keeps everything but the IQ samples on the SMs (tdsa_welch_cluster.cuh)
  if (p->log2n == kWcLog2N && welch_cluster_enabled()) {
    int max_clusters = 0;
    const bool f32 = p->precision == TDSA_PREC_F32;
    if (f32) launch_welch_cluster_f32(WelchClusterArgs<float>(), 0, p->stream, &max_clusters);
    else launch_welch_cluster_f64(WelchClusterArgs<double>(), 0, p->stream, &max_clusters);
    if (max_clusters > 0) {
      if (p->win_dirty) { int rcw = upload_window(p); if (rcw) return rcw; }
      const int clusters = (int)std::min<int64_t>(max_clusters, nseg);
      int rcs = ensure_scratch(&p->scratch, &p->scratch_bytes, (size_t)clusters * n * (sizeof(double) + sizeof(float)));
      if (rcs) return rcs;
      double* part_sum = (double*)p->scratch;
      float* part_peak = (float*)(part_sum + (size_t)clusters * n);
      cudaError_t e;
      if (f32) {
        WelchClusterArgs<float> a;
        a.iq = (const float2*)iq_stream; a.n_seg = nseg; a.hop = hop; a.window = p->d_win32; a.tw_head = p->d_twh32;
        a.tw_inner = p->d_twin32; a.part_sum = part_sum; a.part_peak = part_peak;
        e = launch_welch_cluster_f32(a, clusters, p->stream, nullptr);
      } else {
        WelchClusterArgs<double> a;
        a.iq = (const float2*)iq_stream; a.n_seg = nseg; a.hop = hop; a.window = p->d_win64; a.tw_head = p->d_twh64;
        a.tw_inner = p->d_twin64; a.part_sum = part_sum; a.part_peak = part_peak;
        e = launch_welch_cluster_f64(a, clusters, p->stream, nullptr);
      }
      if (e != cudaSuccess) return fail(TDSA_ERR_CUDA, "welch cluster launch failed: %s", cudaGetErrorString(e));
      const double scale = (p->mode == TDSA_MODE_PSD) ? 1.0 / (p->fs * (double)n) : 1.0;
      welch_cluster_finish_kernel<<<(int)(n / 256), 256, 0, p->stream>>>(part_sum, part_peak, clusters, nseg, scale, p->floor,
                                                                       avg_db, peak_db);
      count_launch();
      CK(cudaGetLastError());
      return TDSA_OK;
    }
  }
  // state: sum[N] double | peak[N] float, kept behind the linear scratch
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(nseg, ((int64_t)128 << 20) / (n * 8)));
  const size_t lin_bytes = (size_t)chunk * n * sizeof(double);
  int rc = ensure_scratch(&p->scratch, &p->scratch_bytes, lin_bytes + (size_t)n * (sizeof(double) + sizeof(float)));
  if (rc) return rc;
  double* lin = (double*)p->scratch;
  double* sum = (double*)((char*)p->scratch + lin_bytes);
  float* peak = (float*)(sum + n);
  const bool big = is_big(p);
  for (int64_t s0 = 0; s0 < nseg; s0 += chunk) {
    const int64_t ns = std::min(chunk, nseg - s0);
    // big plans leave the linear rows in the permuted [s][klow] order (coalesced); finish un-permutes
    rc = big ? run_big(p, (const float2*)iq_stream + s0 * hop, ns, hop, nullptr, kEpiLinearPermuted, nullptr, lin, nullptr, false)
             : run_fused(p, (const float2*)iq_stream + s0 * hop, ns, hop, nullptr, kEpiLinear, nullptr, lin, nullptr, false);
    if (rc) return rc;
    welch_reduce_kernel<<<(int)((n + 255) / 256), 256, 0, p->stream>>>(lin, ns, n, sum, peak, s0 == 0, p->floor, p->mode);
    count_launch();
    CK(cudaGetLastError());
  }
  welch_finish_kernel<<<(int)((n + 255) / 256), 256, 0, p->stream>>>(sum, peak, n, nseg,
                                                                   big ? p->log2n - 4 * big_head_passes(p) : 0,
                                                                   big ? big_head_passes(p) : 0, p->floor, p->mode, avg_db, peak_db);
  count_launch();
  CK(cudaGetLastError());
  return TDSA_OK;
}

// dB rows -> device-resident DataProcessor state (cal offset, tare, sweep-domain averaging, holds); flag block on
// the device, nothing read back
static int trace_update_dev_impl(const float* rows, int64_t n_rows, int64_t width, double cal_offset_db, int avg_mode, int avg_n,
                                 double* avg_state, float* max_hold, float* min_hold, int32_t* flags, float* rows_out,
                                 void* cuda_stream, int32_t* row_flags_scratch, int tare_target, double* tare_buf,
                                 double* tare_baseline) {
  if (!rows || n_rows < 0 || width < 1) return fail(TDSA_ERR_INVALID, "bad rows");
  if (!row_flags_scratch) return fail(TDSA_ERR_INVALID, "row_flags_scratch (int32[n_rows], device) is required");
  if (!flags) return fail(TDSA_ERR_INVALID, "flags block (device int32[8]) is required");
  const bool averaging = avg_mode != TDSA_AVG_OFF && avg_n > 1;
  if (averaging && !avg_state) return fail(TDSA_ERR_INVALID, "averaging needs avg_state");
  const bool has_tare = tare_buf != nullptr && tare_baseline != nullptr;
  if (has_tare && tare_target < 1) return fail(TDSA_ERR_INVALID, "tare needs a target >= 1");
  if (n_rows == 0) return TDSA_OK;
  cudaStream_t s = (cudaStream_t)cuda_stream;
  CK(cudaMemsetAsync(row_flags_scratch, 0, sizeof(int32_t) * n_rows, s));
  dim3 g((unsigned)((width + 255) / 256), (unsigned)n_rows);
  row_allnan_kernel<<<g, 256, 0, s>>>(rows, n_rows, width, row_flags_scratch);
  count_launch();
  TraceUpdateDevArgs a;
  a.rows = rows; a.n_rows = n_rows; a.width = width; a.cal = cal_offset_db;
  a.avg_mode = avg_mode; a.avg_n = avg_n; a.flags = flags;
  a.avg_state = avg_state; a.max_hold = max_hold; a.min_hold = min_hold;
  a.has_value = row_flags_scratch; a.rows_out = rows_out;
  a.tare_target = tare_target; a.tare_buf = tare_buf; a.tare_baseline = tare_baseline;
  trace_update_dev_kernel<<<(unsigned)((width + 255) / 256), 256, 0, s>>>(a);
  count_launch();
  trace_flags_after_update_kernel<<<1, 256, 0, s>>>(flags, row_flags_scratch, n_rows, avg_mode, avg_n, max_hold != nullptr,
                                                  min_hold != nullptr, has_tare ? 1 : 0, tare_target);
  count_launch();
  CK(cudaGetLastError());
  return TDSA_OK;
}

int tdsa_trace_update_dev(const float* rows, int64_t n_rows, int64_t width, double cal_offset_db, int avg_mode, int avg_n,
                          double* avg_state, float* max_hold, float* min_hold, int32_t* flags_dev, float* rows_out,
                          void* cuda_stream, int32_t* row_flags_scratch, int tare_target, double* tare_buf,
                          double* tare_baseline) {
  return trace_update_dev_impl(rows, n_rows, width, cal_offset_db, avg_mode, avg_n, avg_state, max_hold, min_hold, flags_dev,
                               rows_out, cuda_stream, row_flags_scratch, tare_target, tare_buf, tare_baseline);
}

// Host-scalar flavour of the above (count / valid / tare flags are host values, read and updated): a temporary flag
// block carries them to the device and back, which costs one stream synchronisation per call.
static int trace_update_impl(const float* rows, int64_t n_rows, int64_t width, double cal_offset_db, int avg_mode, int avg_n,
                             double* avg_state, int32_t* count_state_host, float* max_hold, float* min_hold,
                             int32_t* hold_valid_host, float* rows_out, void* cuda_stream, int32_t* row_flags_scratch,
                             int32_t* tare_flags_host, int32_t* tare_count_host, int tare_target, double* tare_buf,
                             double* tare_baseline) {
  const bool averaging = avg_mode != TDSA_AVG_OFF && avg_n > 1;
  if (averaging && (!avg_state || !count_state_host)) return fail(TDSA_ERR_INVALID, "averaging needs avg_state and count_state");
  if ((max_hold || min_hold) && !hold_valid_host) return fail(TDSA_ERR_INVALID, "holds need hold_valid");
  const bool tare = tare_flags_host && (tare_flags_host[0] || tare_flags_host[1]);
  if (tare && (!tare_count_host || !tare_buf || !tare_baseline || tare_target < 1))
    return fail(TDSA_ERR_INVALID, "tare needs count, buffer, baseline and a target >= 1");
  if (n_rows == 0) return TDSA_OK;
  cudaStream_t s = (cudaStream_t)cuda_stream;
  int32_t* d_flags = nullptr;
  CK(cudaMallocAsync((void**)&d_flags, kFlagWords * sizeof(int32_t), s));
  int32_t h[kFlagWords] = {};
  h[kFlagCount] = averaging ? *count_state_host : 0;
  h[kFlagMaxValid] = hold_valid_host ? hold_valid_host[0] : 0;
  h[kFlagMinValid] = hold_valid_host ? hold_valid_host[1] : 0;
  if (tare) { h[kFlagTareCollecting] = tare_flags_host[0]; h[kFlagTareActive] = tare_flags_host[1]; h[kFlagTareCount] = *tare_count_host; }
  cudaError_t e = cudaMemcpyAsync(d_flags, h, sizeof h, cudaMemcpyHostToDevice, s);
  int rc = e == cudaSuccess ? trace_update_dev_impl(rows, n_rows, width, cal_offset_db, avg_mode, avg_n, avg_state, max_hold, min_hold,
                                                    d_flags, rows_out, cuda_stream, row_flags_scratch, tare_target,
                                                    tare ? tare_buf : nullptr, tare ? tare_baseline : nullptr)
                            : fail(TDSA_ERR_CUDA, "flag upload failed: %s", cudaGetErrorString(e));
  if (rc == TDSA_OK) {
    e = cudaMemcpyAsync(h, d_flags, sizeof h, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = fail(TDSA_ERR_CUDA, "flag read-back failed: %s", cudaGetErrorString(e));
  }
  cudaFreeAsync(d_flags, s);
  if (rc) return rc;
  if (averaging) *count_state_host = h[kFlagCount];
  if (hold_valid_host) { hold_valid_host[0] = h[kFlagMaxValid]; hold_valid_host[1] = h[kFlagMinValid]; }
  if (tare) { tare_flags_host[0] = h[kFlagTareCollecting]; tare_flags_host[1] = h[kFlagTareActive]; *tare_count_host = h[kFlagTareCount]; }
  return TDSA_OK;
}

int tdsa_trace_update(const float* rows, int64_t n_rows, int64_t width, double cal_offset_db, int avg_mode, int avg_n,
                      double* avg_state, int32_t* count_state_host, float* max_hold, float* min_hold,
                      int32_t* hold_valid_host, float* rows_out, void* cuda_stream, int32_t* row_flags_scratch) {
  return trace_update_impl(rows, n_rows, width, cal_offset_db, avg_mode, avg_n, avg_state, count_state_host, max_hold,
                           min_hold, hold_valid_host, rows_out, cuda_stream, row_flags_scratch, nullptr, nullptr, 0, nullptr,
                           nullptr);
}

int tdsa_trace_update_tare(const float* rows, int64_t n_rows, int64_t width, double cal_offset_db, int avg_mode, int avg_n,
                           double* avg_state, int32_t* count_state_host, float* max_hold, float* min_hold,
                           int32_t* hold_valid_host, float* rows_out, void* cuda_stream, int32_t* row_flags_scratch,
                           int32_t* tare_flags_host, int32_t* tare_count_host, int tare_target, double* tare_buf,
                           double* tare_baseline) {
  return trace_update_impl(rows, n_rows, width, cal_offset_db, avg_mode, avg_n, avg_state, count_state_host, max_hold,
                           min_hold, hold_valid_host, rows_out, cuda_stream, row_flags_scratch, tare_flags_host,
                           tare_count_host, tare_target, tare_buf, tare_baseline);
}

int tdsa_colormap_rgba(const float* rows, int64_t n, float lo_db, float hi_db, const uint8_t* lut_rgba, uint8_t* rgba_out,
                       void* cuda_stream) {
  if (!rows || !lut_rgba || !rgba_out || n < 0) return fail(TDSA_ERR_INVALID, "bad argument");
  if (n == 0) return TDSA_OK;
  const float den = fmaxf(hi_db - lo_db, 1e-9f);
  colormap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(rows, n, lo_db, den, (const uchar4*)lut_rgba,
                                                                                   (uchar4*)rgba_out);
  count_launch();
  CK(cudaGetLastError());
  return TDSA_OK;
}

int tdsa_density_update(const float* live_db, int64_t width, double decay, float* hist, void* cuda_stream) {
  if (!live_db || !hist || width < 1) return fail(TDSA_ERR_INVALID, "bad argument");
  density_update_kernel<<<(unsigned)width, 256, 0, (cudaStream_t)cuda_stream>>>(live_db, width, (float)decay, decay < 1.0 ? 1 : 0, hist);
  count_launch();
  CK(cudaGetLastError());
  return TDSA_OK;
}

int tdsa_band_power(const double* bins, const float* levels, int64_t width, double f_lo, double f_hi, double* out,
                    void* cuda_stream) {
  if (!bins || !levels || !out || width < 1) return fail(TDSA_ERR_INVALID, "bad argument");
  band_power_kernel<<<1, 256, 0, (cudaStream_t)cuda_stream>>>(bins, levels, width, std::min(f_lo, f_hi), std::max(f_lo, f_hi), out);
  count_launch();
  CK(cudaGetLastError());
  return TDSA_OK;
}

int tdsa_top_peaks(const float* power, int64_t width, int n, int min_sep_bins, float min_excursion_db, int32_t* idx_out,
                   float* pwr_out, int32_t* count_out, void* cuda_stream) {
  if (!power || !idx_out || !pwr_out || !count_out) return fail(TDSA_ERR_INVALID, "null argument");
  if (n < 1 || n > 16) return fail(TDSA_ERR_INVALID, "n must be in [1, 16]");
  if (width > 16384) return fail(TDSA_ERR_UNSUPPORTED, "top_peaks supports rows of up to 16384 bins");
  constexpr int kSmem = 8192 * 8;
  static bool once[kMaxDevices] = {};
  const int dev = current_device();
  if (!once[dev]) { CK(cudaFuncSetAttribute(top_peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); once[dev] = true; }
  if (width < 3) { CK(cudaMemsetAsync(count_out, 0, sizeof(int32_t), (cudaStream_t)cuda_stream)); return TDSA_OK; }
  top_peaks_kernel<<<1, 1024, kSmem, (cudaStream_t)cuda_stream>>>(power, (int)width, n, min_sep_bins, min_excursion_db, idx_out,
                                                               pwr_out, count_out);
  count_launch();
  CK(cudaGetLastError());
  return TDSA_OK;
}

// ---- hackrf_sweep wire formats (host side; replaces the per-field Python float() loop of hackrf_sweep.py:138-146) ----
// CSV line: date, time, hz_low, hz_high, bin_width, num_samples, dB, dB, ...  Lines with fewer than 7 fields or a
// field that does not parse are skipped, like the reference's ValueError handler (:167-168).
int tdsa_parse_sweep_csv_host(const char* text, int64_t len, int64_t max_rows, int64_t max_bins, double* lo_hz, double* hi_hz,
                              float* values, int32_t* n_bins, int64_t* n_rows_out, int64_t* consumed_out) {
  if (!text || !lo_hz || !hi_hz || !values || !n_bins || !n_rows_out) return fail(TDSA_ERR_INVALID, "null argument");
  int64_t rows = 0, pos = 0, consumed = 0;
  std::string field;
  while (pos < len && rows < max_rows) {
    int64_t eol = pos;
    while (eol < len && text[eol] != '\n') ++eol;
    if (eol == len) break;                          // incomplete last line: leave it for the next call
    // split [pos, eol) on commas
    int nf = 0;
    bool ok = true;
    int64_t nb = 0;
    long long lo = 0, hi = 0;
    int64_t fs = pos;
    for (int64_t i = pos; i <= eol && ok; ++i) {
      if (i == eol || text[i] == ',') {
        int64_t a = fs, b = i;
        while (a < b && (text[a] == ' ' || text[a] == '\t' || text[a] == '\r')) ++a;
        while (b > a && (text[b - 1] == ' ' || text[b - 1] == '\t' || text[b - 1] == '\r')) --b;
        field.assign(text + a, (size_t)(b - a));
        char* end = nullptr;
        if (nf == 2 || nf == 3) {                   // int(fields[2]), int(fields[3])
          const long long v = strtoll(field.c_str(), &end, 10);
          if (field.empty() || *end != '\0') ok = false;
          (nf == 2 ? lo : hi) = v;
        } else if (nf >= 6) {                       // float(v) for v in fields[6:]
          const double v = strtod(field.c_str(), &end);
          if (field.empty() || *end != '\0') ok = false;
          else if (nb < max_bins) values[rows * max_bins + nb] = (float)v;
          ++nb;
        }
        ++nf;
        fs = i + 1;
      }
    }
    consumed = eol + 1;
    pos = eol + 1;
    if (!ok || nf < 7 || nb == 0 || nb > max_bins) continue;
    lo_hz[rows] = (double)lo; hi_hz[rows] = (double)hi; n_bins[rows] = (int32_t)nb;
    ++rows;
  }
  *n_rows_out = rows;
  if (consumed_out) *consumed_out = consumed;
  return TDSA_OK;
}

// Binary (-B) records, hackrf_sweep_binary_reference.py:29-43: uint32 record_length, then uint64 hz_low,
// uint64 hz_high, float32[] dB (little-endian). Incomplete trailing records are left for the next call.
int tdsa_parse_sweep_binary_host(const uint8_t* buf, int64_t len, int64_t max_rows, int64_t max_bins, double* lo_hz,
                                 double* hi_hz, float* values, int32_t* n_bins, int64_t* n_rows_out, int64_t* consumed_out) {
  if (!buf || !lo_hz || !hi_hz || !values || !n_bins || !n_rows_out) return fail(TDSA_ERR_INVALID, "null argument");
  int64_t rows = 0, pos = 0;
  while (rows < max_rows && pos + 4 <= len) {
    uint32_t rec;
    memcpy(&rec, buf + pos, 4);
    if (pos + 4 + (int64_t)rec > len) break;
    const uint8_t* r = buf + pos + 4;
    pos += 4 + (int64_t)rec;
    if (rec < 16) continue;
    const int64_t nb = ((int64_t)rec - 16) / 4;
    if (nb == 0 || nb > max_bins) continue;       // step_data.size == 0 -> return (:40-41)
    uint64_t lo, hi;
    memcpy(&lo, r, 8); memcpy(&hi, r + 8, 8);
    lo_hz[rows] = (double)lo; hi_hz[rows] = (double)hi; n_bins[rows] = (int32_t)nb;
    memcpy(values + rows * max_bins, r + 16, (size_t)nb * 4);
    ++rows;
  }
  *n_rows_out = rows;
  if (consumed_out) *consumed_out = pos;
  return TDSA_OK;
}

static void fill_stitch_invariants(StitchArgs& a) {
  volatile double num = a.stop - a.start, den = (double)(a.m - 1);      // volatile: no contraction / reassociation
  a.step = a.m > 1 ? num / den : 0.0;
  volatile double rh = a.row_hz, kk = (double)a.k;
  a.bw = rh / kk;
  volatile double b = a.bw;
  a.half_bw = b / 2.0;
  a.inv_bw = 1.0 / a.bw;
  a.inv_row = 1.0 / a.row_hz;
}

int tdsa_stitch(const float* rows, const double* row_lo_hz, double row_hz, int64_t n_rows, int64_t bins_per_row,
                double start_hz, double stop_hz, int64_t m, double* grid_out, void* cuda_stream, int32_t* order_scratch) {
  if (!rows || !row_lo_hz || !grid_out || !order_scratch) return fail(TDSA_ERR_INVALID, "null argument");
  if (n_rows < 1 || bins_per_row < 1 || m < 1) return fail(TDSA_ERR_INVALID, "bad sizes");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  stitch_rank_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, s>>>(row_lo_hz, n_rows, order_scratch);
  count_launch();
  StitchArgs a;
  a.rows = rows; a.lo = row_lo_hz; a.order = order_scratch; a.row_hz = row_hz; a.n_rows = n_rows; a.k = bins_per_row;
  a.start = start_hz; a.stop = stop_hz; a.m = m; a.out = grid_out;
  fill_stitch_invariants(a);
  stitch_interp_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(a);
  count_launch();
  CK(cudaGetLastError());
  return TDSA_OK;
}

int tdsa_stitch_range(const float* rows, const double* row_lo_hz, double row_hz, int64_t n_rows, int64_t bins_per_row,
                      double start_hz, double stop_hz, int64_t m, int64_t g0, int64_t count, double* grid_out,
                      void* cuda_stream, int32_t* order_scratch) {
  if (!rows || !row_lo_hz || !grid_out || !order_scratch) return fail(TDSA_ERR_INVALID, "null argument");
  if (n_rows < 1 || bins_per_row < 1 || m < 1 || g0 < 0 || count < 0 || g0 + count > m) return fail(TDSA_ERR_INVALID, "bad sizes");
  if (count == 0) return TDSA_OK;
  cudaStream_t s = (cudaStream_t)cuda_stream;
  stitch_rank_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, s>>>(row_lo_hz, n_rows, order_scratch);
  count_launch();
  StitchArgs a;
  a.rows = rows; a.lo = row_lo_hz; a.order = order_scratch; a.row_hz = row_hz; a.n_rows = n_rows; a.k = bins_per_row;
  a.start = start_hz; a.stop = stop_hz; a.m = m; a.g0 = g0; a.count = count; a.out = grid_out;
  fill_stitch_invariants(a);
  stitch_interp_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(a);
  count_launch();
  CK(cudaGetLastError());
  return TDSA_OK;
}

int tdsa_ring_push(const float* rows, int64_t n_rows, float* ring, int64_t H, int64_t W, int64_t* ptr_host,
                   void* cuda_stream) {
  if (!rows || !ring || !ptr_host || H < 1 || W < 1 || n_rows < 0) return fail(TDSA_ERR_INVALID, "bad argument");
  if (n_rows == 0) return TDSA_OK;
  const int64_t first = n_rows > H ? n_rows - H : 0;   // older rows would be overwritten anyway
  dim3 g((unsigned)((W + 255) / 256), (unsigned)(n_rows - first));
  ring_push_kernel<<<g, 256, 0, (cudaStream_t)cuda_stream>>>(rows, first, n_rows, ring, H, W, *ptr_host);
  count_launch();
  CK(cudaGetLastError());
  int64_t p = (*ptr_host - n_rows) % H;
  if (p < 0) p += H;
  *ptr_host = p;
  return TDSA_OK;
}

int tdsa_ring_push_dev(const float* rows, int64_t n_rows, float* ring, int64_t H, int64_t W, int64_t* state_dev,
                       float* last_row_dev, int dedupe, int64_t* slot_scratch, int32_t* differs_scratch, void* cuda_stream) {
  if (!rows || !ring || !state_dev || !slot_scratch || H < 1 || W < 1 || n_rows < 0) return fail(TDSA_ERR_INVALID, "bad argument");
  if (dedupe && (!last_row_dev || !differs_scratch)) return fail(TDSA_ERR_INVALID, "dedupe needs last_row_dev and differs_scratch");
  if (n_rows == 0) return TDSA_OK;
  cudaStream_t s = (cudaStream_t)cuda_stream;
  if (dedupe) {
    ring_row_differs_kernel<<<(unsigned)n_rows, 256, 0, s>>>(rows, n_rows, W, last_row_dev, state_dev, differs_scratch);
    count_launch();
  }
  ring_assign_kernel<<<1, 32, 0, s>>>(differs_scratch, n_rows, H, state_dev, slot_scratch, dedupe ? 1 : 0);
  count_launch();
  dim3 g((unsigned)((W + 255) / 256), (unsigned)n_rows);
  ring_scatter_kernel<<<g, 256, 0, s>>>(rows, n_rows, W, H, slot_scratch, state_dev, ring, last_row_dev);
  count_launch();
  CK(cudaGetLastError());
  return TDSA_OK;
}

int tdsa_ring_image_rgba(const float* ring, int64_t H, int64_t W, const int64_t* state_dev, float lo_db, float hi_db,
                         const uint8_t* lut_rgba, uint8_t* rgba_out, void* cuda_stream) {
  if (!ring || !state_dev || !lut_rgba || !rgba_out || H < 1 || W < 1) return fail(TDSA_ERR_INVALID, "bad argument");
  const float den = fmaxf(hi_db - lo_db, 1e-9f);
  const int64_t total = H * W;
  ring_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(ring, H, W, state_dev, lo_db, den,
                                                                                         (const uchar4*)lut_rgba, (uchar4*)rgba_out);
  count_launch();
  CK(cudaGetLastError());
  return TDSA_OK;
}

int tdsa_find_peaks_snap(const float* levels, int64_t width, float height, float prominence, int distance, int32_t* out,
                         void* cuda_stream) {
  if (!levels || !out) return fail(TDSA_ERR_INVALID, "null argument");
  if (width < 1) return fail(TDSA_ERR_INVALID, "empty trace");
  if (width > 16384) return fail(TDSA_ERR_UNSUPPORTED, "find_peaks_snap supports rows of up to 16384 bins");
  constexpr int kSmem = 8192 * (4 + 4 + 4 + 1);
  static bool once[kMaxDevices] = {};
  const int dev = current_device();
  if (!once[dev]) { CK(cudaFuncSetAttribute(find_peaks_snap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); once[dev] = true; }
  find_peaks_snap_kernel<<<1, 1024, kSmem, (cudaStream_t)cuda_stream>>>(levels, (int)width, height, prominence, distance < 1 ? 1 : distance, out);
  count_launch();
  CK(cudaGetLastError());
  return TDSA_OK;
}

int tdsa_h2d_async(const void* pinned_host, void* dev, size_t bytes, void* side_stream, void* done_event) {
  if (!pinned_host || !dev) return fail(TDSA_ERR_INVALID, "null pointer");
  CK(cudaMemcpyAsync(dev, pinned_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)side_stream));
  if (done_event) CK(cudaEventRecord((cudaEvent_t)done_event, (cudaStream_t)side_stream));
  return TDSA_OK;
}

int tdsa_psd_db_batch_host(tdsa_handle_t p, const void* iq_host, int64_t n_frames, int64_t stride, float* db_out_host,
                           int64_t chunk_frames) {
  if (!p) return fail(TDSA_ERR_INVALID, "null plan");
  if (n_frames < 0 || stride < 1) return fail(TDSA_ERR_INVALID, "bad frame geometry");
  if (n_frames == 0) return TDSA_OK;
  if (!iq_host || !db_out_host) return fail(TDSA_ERR_INVALID, "null buffer");
  DeviceGuard guard(p->device);
  const int64_t n = p->n;
  if (chunk_frames < 1) chunk_frames = std::max<int64_t>(1, ((int64_t)16 << 20) / (n * 8));   // ~16 MiB of IQ per chunk (measured best of 4..64 MiB)
  chunk_frames = std::min(chunk_frames, n_frames);
  const size_t in_bytes = (size_t)((chunk_frames - 1) * stride + n) * sizeof(float2);
  const size_t out_bytes = (size_t)chunk_frames * n * sizeof(float);
  if (!p->side) {
    CK(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&p->back, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CK(cudaEventCreateWithFlags(&p->ev_h2d[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&p->ev_k[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&p->ev_d2h[i], cudaEventDisableTiming));
    }
  }
  if (p->d_in_bytes < in_bytes) {
    for (int i = 0; i < 2; ++i) { cudaFree(p->d_in[i]); p->d_in[i] = nullptr; CK(cudaMalloc(&p->d_in[i], in_bytes)); }
    p->d_in_bytes = in_bytes;
  }
  if (p->d_out_bytes < out_bytes) {
    for (int i = 0; i < 2; ++i) { cudaFree(p->d_out[i]); p->d_out[i] = nullptr; CK(cudaMalloc(&p->d_out[i], out_bytes)); }
    p->d_out_bytes = out_bytes;
  }
  // The compute stream must be a real (non-legacy) stream for overlap; fall back to the side stream's sibling.
  cudaStream_t user = p->stream;
  int rc = TDSA_OK;
  int64_t c = 0;
  for (int64_t f0 = 0; f0 < n_frames; f0 += chunk_frames, ++c) {
    const int b = (int)(c & 1);
    const int64_t nf = std::min(chunk_frames, n_frames - f0);
    const size_t bytes = (size_t)((nf - 1) * stride + n) * sizeof(float2);
    // three streams, two buffers: the input buffer b is free once the kernel of chunk c-2 has run, the output buffer b
    // once the D2H of chunk c-2 has drained it; H2D, kernel and D2H of neighbouring chunks overlap (PCIe is full duplex)
    if (c >= 2) CK(cudaStreamWaitEvent(p->side, p->ev_k[b], 0));
    CK(cudaMemcpyAsync(p->d_in[b], (const float2*)iq_host + f0 * stride, bytes, cudaMemcpyHostToDevice, p->side));
    CK(cudaEventRecord(p->ev_h2d[b], p->side));
    CK(cudaStreamWaitEvent(user, p->ev_h2d[b], 0));
    if (c >= 2) CK(cudaStreamWaitEvent(user, p->ev_d2h[b], 0));
    rc = run_fused(p, p->d_in[b], nf, stride, nullptr, kEpiDb, (float*)p->d_out[b], nullptr, nullptr, false);
    if (rc) return rc;
    CK(cudaEventRecord(p->ev_k[b], user));
    CK(cudaStreamWaitEvent(p->back, p->ev_k[b], 0));
    CK(cudaMemcpyAsync(db_out_host + f0 * n, p->d_out[b], (size_t)nf * n * sizeof(float), cudaMemcpyDeviceToHost, p->back));
    CK(cudaEventRecord(p->ev_d2h[b], p->back));
  }
  CK(cudaStreamSynchronize(p->back));
  CK(cudaStreamSynchronize(user));
  return TDSA_OK;
}

int tdsa_plan_info(tdsa_handle_t p, int* n_fft, int* threads, int* ctas_per_sm, int* smem, int* grid) {
  if (!p) return fail(TDSA_ERR_INVALID, "null plan");
  LaunchInfo info;
  DeviceGuard guard(p->device);
  int rc = run_fused(p, nullptr, 1 << 20, p->n, nullptr, kEpiDb, nullptr, nullptr, &info, true);
  if (rc) return rc;
  if (n_fft) *n_fft = p->n;
  if (threads) *threads = info.threads;
  if (ctas_per_sm) *ctas_per_sm = info.ctas_per_sm;
  if (smem) *smem = info.smem;
  if (grid) *grid = info.grid;
  return TDSA_OK;
}

}  // extern "C"
