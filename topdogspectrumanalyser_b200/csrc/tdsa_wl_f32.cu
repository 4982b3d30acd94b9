// float32 instantiations of the warp-local kernels (tdsa_fft_wl.cuh): N = 4096 / 8192, plain and accumulating epilogues.
#include "tdsa_launch.cuh"
namespace tdsa {
cudaError_t launch_wl_f32(int epi, const FftArgs<float>& a, const WlLaunch& L, cudaStream_t s, LaunchInfo* info, bool dry) {
  return launch_wl_impl<float>(epi, a, L, s, info, dry);
}
}  // namespace tdsa
