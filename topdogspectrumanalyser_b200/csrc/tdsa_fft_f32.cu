// float32 instantiations of the fused kernel (fast path).
#include "tdsa_launch.cuh"
namespace tdsa {
cudaError_t launch_fft_f32(int log2n, int epi, const FftArgs<float>& a, int sm, cudaStream_t s, LaunchInfo* info, bool dry) {
  return launch_fft_impl<float>(log2n, epi, a, sm, s, info, dry);
}
int effective_logr_f32(int log2n) { return effective_logr<float>(log2n); }
cudaError_t launch_welch_cluster_f32(const WelchClusterArgs<float>& a, int clusters, cudaStream_t s, int* max_clusters) {
  return launch_welch_cluster_impl<float>(a, clusters, s, max_clusters);
}
}  // namespace tdsa

#include "tdsa_big.cuh"
namespace tdsa {
cudaError_t launch_big_head_f32(const BigArgs<float>& a, int sm, cudaStream_t s, int passes) {
  if (passes == 0) {                                        // head for fft_wl_kernel tails (N = 65536)
    // staging by 128-byte bulk copies (needs 16-byte aligned segment starts) or by 8-byte cp.async.  Measured (round 2, head
    // kernel alone, 2047 segments): float64 433 us against 575 us, float32 324 us against 290 us: bulk copies in float64 only;
    // TDSA_HEAD_BULK=0 / 1 forces one variant
    static const int bulk_env = [] { const char* e = getenv("TDSA_HEAD_BULK"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
    const bool bulk_ok = bulk_env < 0 ? sizeof(float) == 8 : bulk_env == 1;
    const bool bulk = bulk_ok && ((uintptr_t)a.iq & 15) == 0 && (a.frame_stride & 1) == 0;
    auto kern = bulk ? big_head_wl_kernel<float, true> : big_head_wl_kernel<float, false>;
    static int occ_of[kMaxDevices][2] = {};
    const int dev = current_device();
    if (occ_of[dev][bulk] == 0) {
      cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, head_wl_smem<float>());
      if (ea != cudaSuccess) return ea;
      int o = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, 256, head_wl_smem<float>()) != cudaSuccess) o = 2;
      occ_of[dev][bulk] = std::max(o, 1);
    }
    const int grid = 16 * (int)std::max<int64_t>(1, std::min<int64_t>(a.n_frames, (int64_t)sm * occ_of[dev][bulk] / 16));
    kern<<<grid, 256, head_wl_smem<float>(), s>>>(a);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
  }
  if (passes == 1) {
    const int64_t work = a.n_frames * (((int64_t)1 << a.log2n) >> 4);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)sm * 8));
    big_head1_kernel<float><<<grid, 256, 0, s>>>(a);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
  }
  constexpr int kSmem = 4096 * 2 * sizeof(float);
  static bool once[kMaxDevices] = {};
  const int dev = current_device();
  if (!once[dev]) {
    cudaError_t e = cudaFuncSetAttribute(big_head_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    once[dev] = true;
  }
  const int64_t work = a.n_frames * (((int64_t)1 << a.log2n) >> 12);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(work, (int64_t)sm * 4));
  big_head_kernel<float><<<grid, 256, kSmem, s>>>(a);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}
}  // namespace tdsa
