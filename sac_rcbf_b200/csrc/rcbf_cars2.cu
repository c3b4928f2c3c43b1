// rcbf_cars2.cu -- instantiation of the ring-compacted SimulatedCars fused-step kernel (own translation unit: the
// library's .cu files compile in parallel).
#include <cuda_runtime.h>
#include <stdint.h>

#include "rcbf_cars2.cuh"

namespace rcbf {

int launch_cars2(bool fused, const CarsArgs& a, int64_t n, const CarsParams& p, const CarsEnvParams& e,
                 rcbf_counters_t* ws, cudaStream_t s, int64_t* handled) {
  return fused ? launch_cars2_tiles<true>(a, n, p, e, ws, s, handled)
               : launch_cars2_tiles<false>(a, n, p, e, ws, s, handled);
}

}  // namespace rcbf
