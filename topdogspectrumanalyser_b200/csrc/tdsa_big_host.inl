// tdsa_big_host.inl — host driver of the two-kernel large-FFT path (included by tdsa_api.cu).
// N = 256*M: big_head_kernel (passes 0,1 + twiddles) -> scratch Y (kept <= 64 MiB so it stays in
// the 126 MB L2) -> fft_fused_kernel<TAIL> on 256 M-point sub-transforms per frame.
static int run_big(tdsa_plan* p, const void* iq, int64_t n_frames, int64_t stride, const double2* dc, int epi, float* db,
                   double* lin, LaunchInfo* info, bool dry) {
  if (p->win_dirty && !dry) { int rcw = upload_window(p); if (rcw) return rcw; }
  const int passes = big_head_passes(p);              // 1: N = 16*M (M <= 4096); 2: N = 256*M
  const int log2m = p->log2n - 4 * passes;
  const bool f32 = p->precision == TDSA_PREC_F32;
  const size_t csz = f32 ? sizeof(float2) : sizeof(double2);
  const int64_t n = p->n;
  const int64_t chunk = std::max<int64_t>(1, ((int64_t)64 << 20) / (n * (int64_t)csz));
  int tail_epi;
  switch (epi) {
    case kEpiDb: tail_epi = passes == 1 ? kEpiDbTail16 : kEpiDbTail; break;
    case kEpiLinear: tail_epi = passes == 1 ? kEpiLinearTail16 : kEpiLinearTail; break;
    case kEpiLinearPermuted: tail_epi = kEpiLinearPermuted; break;
    default: return fail(TDSA_ERR_INVALID, "bad epilogue for large FFT");
  }
  if (dry) {
    if (info) { info->threads = 256; info->smem = (int)(4096 * csz); info->ctas_per_sm = 2; info->grid = p->sm_count * 2; }
    return TDSA_OK;
  }
  int rc = ensure_scratch(&p->scratch2, &p->scratch2_bytes,
                          (size_t)std::min(chunk, std::max<int64_t>(n_frames, 1)) * n * csz);
  if (rc) return rc;
  for (int64_t f0 = 0; f0 < n_frames; f0 += chunk) {
    const int64_t nf = std::min(chunk, n_frames - f0);
    cudaError_t e;
    EpiParams ep = make_epi(p, db ? db + f0 * n : nullptr, lin ? lin + f0 * n : nullptr);
    if (f32) {
      BigArgs<float> a;
      a.iq = (const float2*)iq + f0 * stride; a.n_frames = nf; a.frame_stride = stride; a.window = p->d_win32;
      a.tw = p->d_twh32; a.dc = dc ? dc + f0 : nullptr; a.y = (float2*)p->scratch2; a.log2n = p->log2n;
      e = launch_big_head_f32(a, p->sm_count, p->stream, passes);
      if (e == cudaSuccess) {
        FftArgs<float> t;
        t.iq = nullptr; t.n_frames = nf * (passes == 1 ? 16 : 256); t.frame_stride = 0; t.window = nullptr; t.tw = p->d_twin32; t.dc = nullptr;
        t.in_ct = (const float2*)p->scratch2; t.ep = ep;
        e = launch_fft_f32(log2m, tail_epi, t, p->sm_count, p->stream, nullptr, false);
      }
    } else {
      BigArgs<double> a;
      a.iq = (const float2*)iq + f0 * stride; a.n_frames = nf; a.frame_stride = stride; a.window = p->d_win64;
      a.tw = p->d_twh64; a.dc = dc ? dc + f0 : nullptr; a.y = (double2*)p->scratch2; a.log2n = p->log2n;
      e = launch_big_head_f64(a, p->sm_count, p->stream, passes);
      if (e == cudaSuccess) {
        FftArgs<double> t;
        t.iq = nullptr; t.n_frames = nf * (passes == 1 ? 16 : 256); t.frame_stride = 0; t.window = nullptr; t.tw = p->d_twin64; t.dc = nullptr;
        t.in_ct = (const double2*)p->scratch2; t.ep = ep;
        e = launch_fft_f64(log2m, tail_epi, t, p->sm_count, p->stream, nullptr, false);
      }
    }
    if (e != cudaSuccess) return fail(TDSA_ERR_CUDA, "large FFT launch failed (N=%d): %s", p->n, cudaGetErrorString(e));
  }
  return TDSA_OK;
}
