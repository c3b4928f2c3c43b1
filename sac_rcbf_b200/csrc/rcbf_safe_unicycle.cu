// rcbf_safe_unicycle.cu -- C ABI entry points of the hot kernels for Unicycle (own translation unit so the three .cu files of the
// library compile in parallel; the kernels live in rcbf_safe_kernels.cuh).
#include <cuda_runtime.h>
#include <stdint.h>

#include "rcbf_safe_kernels.cuh"

using namespace rcbf;

extern "C" {

int rcbf_unicycle_safe_action(const float* state, const float* action, const float* mean, const float* sigma, int64_t n,
                              const rcbf_unicycle_params* p, float* safe_action, float* x, float* lam, float* slack,
                              int32_t* status, int32_t* iters, rcbf_counters_t* workspace, void* stream) {
  UniArgs a{};
  a.st = state; a.ac = action; a.mu = mean; a.sg = sigma;
  a.out = safe_action; a.x = x; a.lam = lam; a.slack = slack; a.status = status; a.iters = iters;
  return launch_safe<UniEnv<false>>(a, n, *p, rcbf_unicycle_env_params{}, workspace, (cudaStream_t)stream);
}

int rcbf_unicycle_safe_step(float* state4, int32_t* step, const float* action_rl, const float* mean, const float* sigma,
                            int64_t n, const rcbf_unicycle_params* p, const rcbf_unicycle_env_params* e,
                            float* safe_action, float* obs, float* reward, uint8_t* done, float* cost, uint8_t* goal_met,
                            int32_t* status, rcbf_counters_t* workspace, void* stream) {
  UniArgs a{};
  a.state4 = state4; a.step = step; a.ac = action_rl; a.mu = mean; a.sg = sigma;
  a.out = safe_action; a.status = status;
  a.obs = obs; a.reward = reward; a.done = done; a.cost = cost; a.goal_met = goal_met;
  return launch_safe<UniEnv<true>>(a, n, *p, *e, workspace, (cudaStream_t)stream);
}

}  // extern "C"
