// rcbf_safe_cars.cu -- C ABI entry points of the hot kernels for SimulatedCars (own translation unit so the three .cu files of the
// library compile in parallel; the kernels live in rcbf_safe_kernels.cuh).
#include <cuda_runtime.h>
#include <stdint.h>

#include "rcbf_safe_kernels.cuh"

using namespace rcbf;

namespace rcbf {
// rcbf_cars2.cu: the ring-compacted fused-step kernel on the leading full 32-instance tiles (if the call qualifies)
int launch_cars2(bool fused, const CarsArgs& a, int64_t n, const CarsParams& p, const CarsEnvParams& e,
                 rcbf_counters_t* ws, cudaStream_t s, int64_t* handled);
}

namespace {
// get_safe_action alone: k_cars2<false> on the full tiles, k_safe on the ragged rest (or on everything when the call
// does not qualify: dense saved tensors, interior-point mode, unaligned arrays, small n)
int launch_cars_layer(CarsArgs a, int64_t n, const rcbf_cars_params& p, rcbf_counters_t* ws, cudaStream_t s) {
  int64_t handled = 0;
  const int rc = launch_cars2(false, a, n, p, rcbf_cars_env_params{}, ws, s, &handled);
  if (rc != 0) return rc;
  if (handled == n) return 0;
  if (handled > 0) {
    a.st += handled * 10; a.ac += handled; a.sg += handled * 10; a.out += handled;
    if (a.status != nullptr) a.status += handled;
    if (a.meta != nullptr) a.meta += handled;
  }
  return launch_safe<CarsEnv<false>>(a, n - handled, p, rcbf_cars_env_params{}, ws, s);
}
}  // namespace

extern "C" {

int rcbf_cars_safe_action(const float* state, const float* action, const float* sigma, int64_t n,
                          const rcbf_cars_params* p, float* safe_action, float* x, float* lam, float* slack,
                          int32_t* status, int32_t* iters, rcbf_counters_t* workspace, void* stream) {
  CarsArgs a{};
  a.st = state; a.ac = action; a.sg = sigma;
  a.out = safe_action; a.x = x; a.lam = lam; a.slack = slack; a.status = status; a.iters = iters;
  return launch_cars_layer(a, n, *p, workspace, (cudaStream_t)stream);
}

int rcbf_cars_safe_action_saved(const float* state, const float* action, const float* sigma, int64_t n,
                                const rcbf_cars_params* p, float* safe_action, int32_t* meta, rcbf_counters_t* workspace,
                                void* stream) {
  CarsArgs a{};
  a.st = state; a.ac = action; a.sg = sigma;
  a.out = safe_action; a.meta = meta;
  return launch_cars_layer(a, n, *p, workspace, (cudaStream_t)stream);
}

int rcbf_cars_safe_step(float* state, float* t, int32_t* step, const float* action_rl, const float* sigma, int64_t n,
                        const rcbf_cars_params* p, const rcbf_cars_env_params* e, float* safe_action, float* obs,
                        float* reward, uint8_t* done, float* cost, int32_t* status, rcbf_counters_t* workspace,
                        void* stream) {
  CarsArgs a{};
  a.state = state; a.t = t; a.step = step; a.ac = action_rl; a.sg = sigma;
  a.out = safe_action; a.status = status;
  a.obs = obs; a.reward = reward; a.done = done; a.cost = cost;
  int64_t handled = 0;
  const int rc = launch_cars2(true, a, n, *p, *e, workspace, (cudaStream_t)stream, &handled);
  if (rc != 0) return rc;
  if (handled == n) return 0;
  if (handled > 0) {  // ragged rest (< 32 instances) through the one-tile-at-a-time kernel
    a.state += handled * 10; a.t += handled; a.step += handled; a.ac += handled; a.sg += handled * 10;
    a.out += handled; a.obs += handled * 10; a.reward += handled; a.done += handled; a.cost += handled;
    if (a.status != nullptr) a.status += handled;
  }
  return launch_safe<CarsEnv<true>>(a, n - handled, *p, *e, workspace, (cudaStream_t)stream);
}

}  // extern "C"
