/*
 * tdsa.h — C ABI of libtdsa.so, the B200 (sm_100a) backend for the sample-mode
 * hot path of CWNE88/topdogspectrumanalyser.
 *
 * The reference is pure Python and has no FFI of its own: its boundary for
 * this path is the class datasources.base.SampleDataSource
 * (datasources/base.py:43-169).  These entry points are what a ctypes binding
 * behind that class needs; each one names the reference
 * code it replaces.  Plain pointers and sizes only; no torch / Python types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative TDSA_ERR_* code;
 *     tdsa_last_error() returns a thread-local message for the last failure.
 *   - pointers are DEVICE pointers unless the parameter name ends in _host.
 *   - a plan (handle) is bound to the CUDA device that was current at
 *     tdsa_create() and launches on the stream set with tdsa_set_stream()
 *     (default: the legacy default stream).  Calls are asynchronous with
 *     respect to the host unless documented otherwise.
 *   - a plan is thread-compatible (one thread at a time), like the reference's
 *     data sources, which are only driven from the Qt GUI thread.
 */
#ifndef TDSA_H_
#define TDSA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDSA_VERSION 100 /* 0.1.0 */

/* error codes */
#define TDSA_OK 0
#define TDSA_ERR_INVALID -1     /* bad argument */
#define TDSA_ERR_UNSUPPORTED -2 /* e.g. FFT size not a power of two in [64, 2^20] */
#define TDSA_ERR_CUDA -3        /* a CUDA runtime call failed; see tdsa_last_error() */
#define TDSA_ERR_NOMEM -4

/* window_id — datasources/rtl_samples.py:199-206 ("hanning"/"hamming"/"rectangle");
 * blackman: utils/constants.py:68-73 lists it, BASELINE.json asks for it. */
#define TDSA_WINDOW_HANN 0
#define TDSA_WINDOW_HAMMING 1
#define TDSA_WINDOW_RECT 2
#define TDSA_WINDOW_BLACKMAN 3
#define TDSA_WINDOW_CUSTOM 4 /* table supplied with tdsa_set_window_table_host */

/* window_norm — none: rtl_samples.py:22; rms: hackrf_samples.py:314-316 */
#define TDSA_NORM_NONE 0
#define TDSA_NORM_RMS_F32 1 /* w = float32(hann); w /= sqrt(mean(w^2)) in float32 */

/* mode — which dB branch is evaluated */
#define TDSA_MODE_POWER 0 /* 10*log10(|X|^2 + floor)          rtl_samples.py:180-184 */
#define TDSA_MODE_PSD 1   /* 10*log10(|X|^2/(fs*N) + floor)   rtl_samples.py:175-179 */
#define TDSA_MODE_MAG20 2 /* 20*log10(|X| + floor)            hackrf_samples.py:383  */

/* precision of the butterflies */
#define TDSA_PREC_F64 0 /* float64 window+FFT (the reference's arithmetic; strict 1e-4 dB) */
#define TDSA_PREC_F32 1 /* float32 window+FFT (fast path; deep-null tail documented)      */

/* averaging mode — utils/signal_processing.py:19-28 */
#define TDSA_AVG_OFF 0
#define TDSA_AVG_EXP 1
#define TDSA_AVG_LIN 2


typedef struct tdsa_plan* tdsa_handle_t;

int tdsa_version(void);
const char* tdsa_last_error(void);

/* Plan for one FFT size. Replaces RtlSamplesDataSource.__init__/set_fft_size
 * (rtl_samples.py:18-28,208-215) and HackrfSamplesDataSource._allocate_fft_resources
 * (hackrf_samples.py:311-324): builds the window, twiddle tables and scratch.
 * log_floor: utils/constants.py:152-155 (1e-10 power, 1e-12 psd/mag20).
 * fs: sample rate in Hz, used by TDSA_MODE_PSD only. */
int tdsa_create(int n_fft, int window_id, int window_norm, int mode, double log_floor, double fs,
                int precision, tdsa_handle_t* out);
int tdsa_destroy(tdsa_handle_t h);
int tdsa_set_stream(tdsa_handle_t h, void* cuda_stream);

/* set_window_type (rtl_samples.py:199-206). */
int tdsa_set_window(tdsa_handle_t h, int window_id, int window_norm);
/* Install an explicit float64 window of n_fft values (host pointer), e.g. np.hanning(N)
 * so the table is bit-identical to the reference's. Synchronous copy. */
int tdsa_set_window_table_host(tdsa_handle_t h, const double* window_host);
/* Copy the float64 window in use back to the host (n_fft doubles). Synchronous. */
int tdsa_get_window_table_host(tdsa_handle_t h, double* window_host);
/* set_psd_mode (rtl_samples.py:248-255) and the dB branch / floor / fs in use. */
int tdsa_set_mode(tdsa_handle_t h, int mode, double log_floor, double fs);
int tdsa_set_precision(tdsa_handle_t h, int precision);

/* Kernel 1 — the fused hot path, rtl_samples.py:169-184 applied to n_frames rows:
 *   db[f][k] = dB( | fftshift( FFT( iq[f] * window ) )[k] |^2 ), float32.
 * iq: complex64 (interleaved float re,im), frame f starts at iq + f*frame_stride
 * complex samples (frame_stride >= 1; frame_stride < n_fft gives overlapping frames;
 * frames must be 8-byte aligned, 16-byte for the bulk-copy staging path).
 * db_out: float32 [n_frames][n_fft], row-major, 16-byte aligned.
 * Algorithmic bytes: 12 per input sample (8 read + 4 written). */
int tdsa_psd_db_batch(tdsa_handle_t h, const void* iq, int64_t n_frames, int64_t frame_stride,
                      float* db_out);

/* Same transform, linear output (|X|^2 or PSD before the log), float64 [n_frames][n_fft].
 * Used where the reference averages in the linear domain (signal_processing.py:35-61). */
int tdsa_power_linear_batch(tdsa_handle_t h, const void* iq, int64_t n_frames, int64_t frame_stride,
                            double* lin_out);

/* HackRF-style per-frame front end, hackrf_samples.py:351-368: for each frame computes
 * mean |x|^2 and mean x; writes silent[f] = (mean|x|^2 < 1e-20); the DC estimate
 * dc = (1-alpha)*dc_prev + alpha*mean(x) is subtracted before the window inside the
 * fused kernel. dc_state: 2 doubles (re, im) carried across calls, updated frame by frame. */
int tdsa_psd_db_batch_dc(tdsa_handle_t h, const void* iq, int64_t n_frames, int64_t frame_stride,
                         double dc_alpha, double* dc_state, int32_t* silent_out, float* db_out);

/* Frames -> running trace state in one launch (TraceAverager + dB + holds):
 * for each of n_frames consecutive frames, linear power p is folded into avg_state
 * exactly as TraceAverager.process does (signal_processing.py:35-61; float64 state,
 * *count_state follows ._count), db = 10*log10(avg + floor) is written to db_out
 * (float32 [n_frames][n_fft], or only the last row if last_only != 0), and the
 * max/min hold rows are updated with fmax/fmin semantics of
 * display_data_processor.py:371-395 (NaN-ignoring; first frame initialises).
 * avg_mode TDSA_AVG_OFF skips the averager. max_hold/min_hold may be NULL.
 * count_state_host: HOST int32, TraceAverager._count (0 = empty buffer), read and updated.
 * hold_valid_host: HOST int32[2] (max, min): 0 = "hold not initialised yet", read and updated.
 * (Both are scalars of widget state, like the reference keeps them in Python attributes.) */
int tdsa_psd_db_avg_hold(tdsa_handle_t h, const void* iq, int64_t n_frames, int64_t frame_stride,
                         int avg_mode, int avg_n, double* avg_state, int32_t* count_state_host,
                         float* max_hold, float* min_hold, int32_t* hold_valid_host, int last_only,
                         float* db_out);

/* tdsa_psd_db_avg_hold with the HackRF front end of tdsa_psd_db_batch_dc in front of it
 * (hackrf_samples.py:351-381): silent frames update nothing and repeat the previous row;
 * silent_out: device int32[n_frames]. Synchronises the stream once (the host scalars depend on how many frames were
 * live); tdsa_psd_db_avg_hold_dev does not. */
int tdsa_psd_db_avg_hold_dc(tdsa_handle_t h, const void* iq, int64_t n_frames, int64_t frame_stride,
                            double dc_alpha, double* dc_state, int32_t* silent_out, int avg_mode, int avg_n,
                            double* avg_state, int32_t* count_state_host, float* max_hold,
                            float* min_hold, int32_t* hold_valid_host, int last_only, float* db_out);

/* Device-resident flavour of tdsa_psd_db_avg_hold[_dc]: nothing is read back, the call is fully asynchronous.
 * flags_dev: DEVICE int32[TDSA_FLAG_WORDS], the scalars the reference keeps in Python attributes:
 *   [TDSA_FLAG_COUNT] TraceAverager._count (0 = empty buffer), [TDSA_FLAG_MAX_VALID] / [TDSA_FLAG_MIN_VALID] "hold
 *   initialised", [TDSA_FLAG_LIVE] live (non-silent) frames folded by the last call; zero the block to reset.
 * last_row_dev: DEVICE float32[n_fft] or NULL: last good dB row, carried across calls; silent frames repeat it
 *   (hackrf_samples.py:351-355).  use_dc != 0 puts the HackRF front end of tdsa_psd_db_batch_dc in front.
 * Large batches take a fused path when the request is order-free (running average with last_only and no holds;
 * holds without averaging): the FFT kernel's accumulating epilogue keeps per-bin sums / maxima / minima in tensor
 * memory and a finish kernel folds them into the state, moving 8 B per input sample (+ 4 B per stored dB value).
 * Everything else scans float64 rows that are produced and consumed in L2-sized chunks. */
#define TDSA_FLAG_COUNT 0
#define TDSA_FLAG_MAX_VALID 1
#define TDSA_FLAG_MIN_VALID 2
#define TDSA_FLAG_LIVE 3
#define TDSA_FLAG_TARE_COLLECTING 4
#define TDSA_FLAG_TARE_ACTIVE 5
#define TDSA_FLAG_TARE_COUNT 6
#define TDSA_FLAG_WORDS 8
int tdsa_psd_db_avg_hold_dev(tdsa_handle_t h, const void* iq, int64_t n_frames, int64_t frame_stride, int use_dc,
                             double dc_alpha, double* dc_state, int32_t* silent_out, int avg_mode, int avg_n,
                             double* avg_state, float* max_hold, float* min_hold, int32_t* flags_dev,
                             float* last_row_dev, int last_only, float* db_out);

/* Config 4 (per sub-band rows): iq holds n_groups * frames_per_group contiguous frames; each
 * group is averaged in the linear domain as TraceAverager('lin', n >= frames_per_group) does
 * (signal_processing.py:56-59) and emitted as one dB row: db_rows float32 [n_groups][n_fft]. */
int tdsa_group_avg_db(tdsa_handle_t h, const void* iq, int64_t n_groups, int64_t frames_per_group,
                      float* db_rows);

/* tdsa_group_avg_db fused with the all-gather of config 4: the dB row of local group g is stored by the FFT kernel
 * itself into every rank's gathered row table, at row row_offset + g (peer_rows_host: HOST array of n_peers <= 8
 * device pointers to float32 [n_bands][n_fft] buffers, one per rank of the node, e.g. torch symmetric memory;
 * the caller's buffer is one of them).  The caller synchronises the ranks afterwards (a barrier, not a collective).
 * N = 4096 / 8192 only (TDSA_ERR_UNSUPPORTED otherwise: use tdsa_group_avg_db + an all-gather). */
int tdsa_group_avg_db_peers(tdsa_handle_t h, const void* iq, int64_t n_groups, int64_t frames_per_group,
                            int64_t row_offset, const uint64_t* peer_rows_host, int n_peers);

/* Config 3: Welch average + peak hold over a flat IQ stream:
 * segments s = 0 .. floor((n_samples - n_fft)/hop), each transformed as kernel 1;
 * avg_db = 10*log10(mean_s |X_s|^2 + floor) (TraceAverager 'lin' with n >= nseg,
 * signal_processing.py:56-59), peak_db = fmax_s dB_s (display_data_processor.py:382).
 * Outputs are float32 [n_fft]. Uses plan-owned scratch; synchronous only w.r.t. the stream.
 * n_fft <= 4096: one launch of the warp-local kernel, sum and maximum of |X|^2 per bin in tensor memory.
 * n_fft = 65536: a radix-16 head kernel writes the sixteen 4096-point sub-transform inputs of every segment (complex T,
 * 16 B per point in float64: plan scratch of n_segments * 2^20 bytes, at most 2048 segments per pass), one launch of the
 * warp-local kernel transforms them with the same tensor-memory state, a finish kernel un-permutes and takes the dB.
 * Environment switches, read at every call, select the measured alternatives: TDSA_WELCH_FUSED=1 (head and tails in one
 * kernel, intermediate in an L2-resident ring; NaN rows if its CTAs cannot all be resident), TDSA_WELCH_SUB=0 (round-1
 * paths: the 16-CTA cluster kernel, or with TDSA_WELCH_CLUSTER=0 the two-kernel path through float64 linear rows). */
int tdsa_welch(tdsa_handle_t h, const void* iq_stream, int64_t n_samples, int64_t hop, float* avg_db,
               float* peak_db);

/* Trace-state update on already-computed dB rows (device-resident DataProcessor state):
 * cal offset (display_data_processor.py:317-327), sweep-domain averaging (:214-218),
 * max/min hold (:371-395). rows: float32 [n_rows][width]. Any state pointer may be NULL.
 * row_flags_scratch: device int32[n_rows] work area. Rows that are entirely NaN are skipped
 * (display_data_processor.py:211). Synchronises the stream once (the host scalars depend on how many rows were live);
 * tdsa_trace_update_dev does not. */
int tdsa_trace_update(const float* rows, int64_t n_rows, int64_t width, double cal_offset_db,
                      int avg_mode, int avg_n, double* avg_state, int32_t* count_state_host,
                      float* max_hold, float* min_hold, int32_t* hold_valid_host, float* rows_out,
                      void* cuda_stream, int32_t* row_flags_scratch);

/* tdsa_trace_update with the tare ("normalisation") stage of display_data_processor.py:329-369 between the
 * cal offset and the holds: while collecting, 10^(dB/10) of each row is accumulated in tare_buf; after
 * tare_target rows (UIConstants.TARE_NUM_SAMPLES = 32) the baseline 10*log10(max(mean, 1e-30)) is captured
 * into tare_baseline and subtracted from that row on. tare_flags_host: HOST int32[2] {collecting, active};
 * tare_count_host: HOST int32 (TareState.count). tare_buf / tare_baseline: device float64[width]. */
int tdsa_trace_update_tare(const float* rows, int64_t n_rows, int64_t width, double cal_offset_db,
                           int avg_mode, int avg_n, double* avg_state, int32_t* count_state_host,
                           float* max_hold, float* min_hold, int32_t* hold_valid_host, float* rows_out,
                           void* cuda_stream, int32_t* row_flags_scratch, int32_t* tare_flags_host,
                           int32_t* tare_count_host, int tare_target, double* tare_buf,
                           double* tare_baseline);

/* Device-resident flavour of tdsa_trace_update_tare: the flag block (see tdsa_psd_db_avg_hold_dev; the tare words are
 * TareState.collecting / .active / .count) lives on the device, nothing is read back, no synchronisation.
 * tare_buf / tare_baseline may be NULL when tare is not used. */
int tdsa_trace_update_dev(const float* rows, int64_t n_rows, int64_t width, double cal_offset_db, int avg_mode,
                          int avg_n, double* avg_state, float* max_hold, float* min_hold, int32_t* flags_dev,
                          float* rows_out, void* cuda_stream, int32_t* row_flags_scratch, int tare_target,
                          double* tare_buf, double* tare_baseline);

/* Waterfall colour map, core/export_manager.py:72-79: index = uint8(clip((x - lo)/max(hi - lo, 1e-9), 0, 1) * 255)
 * in float32, rgba_out[i] = lut_rgba[index] (lut: device uint8[256][4]; rgba_out: device uint8[n][4]). */
int tdsa_colormap_rgba(const float* rows, int64_t n, float lo_db, float hi_db, const uint8_t* lut_rgba,
                       uint8_t* rgba_out, void* cuda_stream);

/* Density (persistence) histogram, displays/density_display.py:306-319: hist float32 [width][512] over
 * -200..+100 dB; hist *= float32(decay) when decay < 1, then +1 in the bin of every non-NaN live_db value. */
int tdsa_density_update(const float* live_db, int64_t width, double decay, float* hist, void* cuda_stream);

/* Band power between two markers, core/marker_manager.py:308-318; out: device float64[1] (NaN if no bin). */
int tdsa_band_power(const double* bins, const float* levels, int64_t width, double f_lo, double f_hi,
                    double* out, void* cuda_stream);

/* Top-n peak list, core/display_data_processor.py:432-471 (_find_top_peaks: strict local maxima by
 * decreasing power, min separation in bins and min valley excursion in dB). width <= 16384, n <= 16.
 * idx_out int32[n], pwr_out float32[n], count_out int32[1], all device. */
int tdsa_top_peaks(const float* power, int64_t width, int n, int min_sep_bins, float min_excursion_db,
                   int32_t* idx_out, float* pwr_out, int32_t* count_out, void* cuda_stream);

/* hackrf_sweep wire formats -> row arrays for tdsa_stitch (HOST functions, no GPU involved).
 * CSV (datasources/hackrf_sweep.py:135-146): "date, time, hz_low, hz_high, bin_width, num_samples, dB...";
 * binary -B (datasources/hackrf_sweep_binary_reference.py:29-43): uint32 length, uint64 lo, uint64 hi, float32[].
 * Outputs (host): lo_hz/hi_hz double[max_rows], values float32[max_rows][max_bins], n_bins int32[max_rows].
 * Malformed lines / records are skipped like the reference does; *consumed_out = bytes fully parsed. */
int tdsa_parse_sweep_csv_host(const char* text, int64_t len, int64_t max_rows, int64_t max_bins, double* lo_hz,
                              double* hi_hz, float* values, int32_t* n_bins, int64_t* n_rows_out,
                              int64_t* consumed_out);
int tdsa_parse_sweep_binary_host(const uint8_t* buf, int64_t len, int64_t max_rows, int64_t max_bins,
                                 double* lo_hz, double* hi_hz, float* values, int32_t* n_bins,
                                 int64_t* n_rows_out, int64_t* consumed_out);

/* hackrf_sweep stitch, datasources/hackrf_sweep.py:135-168: rows (float32 [n_rows][bins_per_row])
 * with per-row hz_low and a common row width row_hz are placed at bin centres
 * lo + bw/2 + i*bw, sorted by frequency and linearly interpolated (np.interp semantics,
 * edge-clamped) onto grid = linspace(start_hz, stop_hz, m). grid_out: float64 [m].
 * Rows must not overlap in frequency (hackrf_sweep's rows do not).
 * order_scratch: device int32[n_rows] work area (argsort of the rows). */
int tdsa_stitch(const float* rows, const double* row_lo_hz, double row_hz, int64_t n_rows,
                int64_t bins_per_row, double start_hz, double stop_hz, int64_t m, double* grid_out,
                void* cuda_stream, int32_t* order_scratch);

/* tdsa_stitch for a slice of the grid (sharded stitch): computes grid points g0 .. g0+count-1 of the m-point grid
 * into grid_out[0 .. count); bit-identical to the same elements of tdsa_stitch's output. */
int tdsa_stitch_range(const float* rows, const double* row_lo_hz, double row_hz, int64_t n_rows,
                      int64_t bins_per_row, double start_hz, double stop_hz, int64_t m, int64_t g0, int64_t count,
                      double* grid_out, void* cuda_stream, int32_t* order_scratch);

/* Waterfall history ring, displays/waterfall.py:163-180: ring is float32 [2*H][W];
 * for each pushed row: ptr = (ptr-1) mod H; ring[ptr] = ring[ptr+H] = row.
 * *ptr_host is read and updated on the host (it is a scalar of widget state). */
int tdsa_ring_push(const float* rows, int64_t n_rows, float* ring, int64_t H, int64_t W,
                   int64_t* ptr_host, void* cuda_stream);

/* tdsa_ring_push with the widget's duplicate filter (displays/waterfall.py:330-336: a row identical to the previous
 * one, np.array_equal semantics, is not added) and the write pointer on the device, because the number of new rows is
 * only known there.  state_dev: DEVICE int64[4] = {ptr, has_last, rows added by the last call, -}; zero it to reset.
 * last_row_dev: DEVICE float32[W], the last row added (may be NULL when dedupe == 0).
 * slot_scratch: DEVICE int64[n_rows]; differs_scratch: DEVICE int32[n_rows] (dedupe only). Asynchronous. */
int tdsa_ring_push_dev(const float* rows, int64_t n_rows, float* ring, int64_t H, int64_t W, int64_t* state_dev,
                       float* last_row_dev, int dedupe, int64_t* slot_scratch, int32_t* differs_scratch,
                       void* cuda_stream);

/* The ring's display view (displays/waterfall.py:179-180: rows ptr .. ptr+H, newest first) colour-mapped to the RGBA
 * image core/export_manager.py:67-84 builds: rgba_out uint8 [H][W][4]; lut_rgba: device uint8[256][4]. */
int tdsa_ring_image_rgba(const float* ring, int64_t H, int64_t W, const int64_t* state_dev, float lo_db, float hi_db,
                         const uint8_t* lut_rgba, uint8_t* rgba_out, void* cuda_stream);

/* Marker snap, core/marker_manager.py:74-99: scipy.signal.find_peaks(levels, height, prominence, distance) and the
 * position of the highest surviving peak; np.argmax(levels) when none survives. width <= 16384.
 * out: DEVICE int32[3] = {bin index, number of surviving peaks, 1 if the argmax fallback was used}. */
int tdsa_find_peaks_snap(const float* levels, int64_t width, float height, float prominence, int distance,
                         int32_t* out, void* cuda_stream);

/* Pinned-host -> device copy on a side stream with a completion event (config 5).
 * done_event may be NULL. */
int tdsa_h2d_async(const void* pinned_host, void* dev, size_t bytes, void* side_stream, void* done_event);

/* End-to-end convenience used by the Python class for host arrays: copies n_frames frames
 * from (preferably pinned) host memory in chunks, runs kernel 1 per chunk on the plan's
 * stream while the next chunk is copied on an internal side stream, and copies dB rows back.
 * Blocks until db_out_host is complete. */
int tdsa_psd_db_batch_host(tdsa_handle_t h, const void* iq_host, int64_t n_frames, int64_t frame_stride,
                           float* db_out_host, int64_t chunk_frames);

/* Introspection for bench.py / tests */
int tdsa_plan_info(tdsa_handle_t h, int* n_fft, int* threads_per_cta, int* ctas_per_sm, int* smem_bytes,
                   int* grid);
/* Number of kernels this library has launched since load (all plans). */
int64_t tdsa_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TDSA_H_ */
