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
    static int occ_of[kMaxDevices] = {};
    const int dev = current_device();
    if (occ_of[dev] == 0) {
      cudaError_t ea = cudaFuncSetAttribute(big_head_wl_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, head_wl_smem<float>());
      if (ea != cudaSuccess) return ea;
      int o = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, big_head_wl_kernel<float>, 256, head_wl_smem<float>()) != cudaSuccess) o = 2;
      occ_of[dev] = std::max(o, 1);
    }
    const int grid = 16 * (int)std::max<int64_t>(1, std::min<int64_t>(a.n_frames, (int64_t)sm * occ_of[dev] / 16));
    big_head_wl_kernel<float><<<grid, 256, head_wl_smem<float>(), s>>>(a);
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
