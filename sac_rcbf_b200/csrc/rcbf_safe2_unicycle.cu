// rcbf_safe2_unicycle.cu -- instantiations of the two-instances-per-lane Unicycle kernel (own translation unit: the
// library's .cu files compile in parallel).
#include <cuda_runtime.h>
#include <stdint.h>

#include "rcbf_safe2.cuh"

namespace rcbf {

int launch_safe2_unicycle(bool fused, const UniArgs& a, int64_t n, const UnicycleParams& p, const UnicycleEnvParams& e,
                          rcbf_counters_t* ws, cudaStream_t s, int64_t* handled) {
  return fused ? launch_safe2_tiles<true>(a, n, p, e, ws, s, handled)
               : launch_safe2_tiles<false>(a, n, p, e, ws, s, handled);
}

}  // namespace rcbf
