// rcbf_safe_unicycle.cu -- C ABI entry points of the hot kernels for Unicycle (own translation unit so the three .cu files of the
// library compile in parallel; the kernels live in rcbf_safe_kernels.cuh).
#include <cuda_runtime.h>
#include <stdint.h>

#include "rcbf_safe_kernels.cuh"

using namespace rcbf;

namespace rcbf {
// rcbf_safe2_unicycle.cu: the two-instances-per-lane kernel on the leading full 64-instance tiles (if the call qualifies)
int launch_safe2_unicycle(bool fused, const UniArgs& a, int64_t n, const UnicycleParams& p, const UnicycleEnvParams& e,
                          rcbf_counters_t* ws, cudaStream_t s, int64_t* handled);
}

namespace {
template <bool kFused>
int launch_unicycle(UniArgs a, int64_t n, const rcbf_unicycle_params& p, const rcbf_unicycle_env_params& e,
                    rcbf_counters_t* ws, cudaStream_t s) {
  int64_t done = 0;
  const int rc = launch_safe2_unicycle(kFused, a, n, p, e, ws, s, &done);
  if (rc != 0) return rc;
  if (done == n) return 0;
  if (done > 0) {  // ragged rest (< 64 instances) through the one-per-lane kernel
    if (kFused) {
      a.state4 += done * 4; a.step += done; a.obs += done * 7; a.reward += done; a.done += done; a.cost += done;
      a.goal_met += done;
    } else {
      a.st += done * 3;
    }
    a.ac += done * 2; a.mu += done * 3; a.sg += done * 3; a.out += done * 2;
    if (a.status != nullptr) a.status += done;
    if (a.meta != nullptr) a.meta += done;
  }
  return launch_safe<UniEnv<kFused>>(a, n - done, p, e, ws, s);
}
}  // namespace

extern "C" {

int rcbf_unicycle_safe_action(const float* state, const float* action, const float* mean, const float* sigma, int64_t n,
                              const rcbf_unicycle_params* p, float* safe_action, float* x, float* lam, float* slack,
                              int32_t* status, int32_t* iters, rcbf_counters_t* workspace, void* stream) {
  UniArgs a{};
  a.st = state; a.ac = action; a.mu = mean; a.sg = sigma;
  a.out = safe_action; a.x = x; a.lam = lam; a.slack = slack; a.status = status; a.iters = iters;
  return launch_unicycle<false>(a, n, *p, rcbf_unicycle_env_params{}, workspace, (cudaStream_t)stream);
}

int rcbf_unicycle_safe_action_saved(const float* state, const float* action, const float* mean, const float* sigma,
                                    int64_t n, const rcbf_unicycle_params* p, float* safe_action, int32_t* meta,
                                    rcbf_counters_t* workspace, void* stream) {
  UniArgs a{};
  a.st = state; a.ac = action; a.mu = mean; a.sg = sigma;
  a.out = safe_action; a.meta = meta;
  return launch_unicycle<false>(a, n, *p, rcbf_unicycle_env_params{}, workspace, (cudaStream_t)stream);
}

int rcbf_unicycle_safe_step(float* state4, int32_t* step, const float* action_rl, const float* mean, const float* sigma,
                            int64_t n, const rcbf_unicycle_params* p, const rcbf_unicycle_env_params* e,
                            float* safe_action, float* obs, float* reward, uint8_t* done, float* cost, uint8_t* goal_met,
                            int32_t* status, rcbf_counters_t* workspace, void* stream) {
  UniArgs a{};
  a.state4 = state4; a.step = step; a.ac = action_rl; a.mu = mean; a.sg = sigma;
  a.out = safe_action; a.status = status;
  a.obs = obs; a.reward = reward; a.done = done; a.cost = cost; a.goal_met = goal_met;
  return launch_unicycle<true>(a, n, *p, *e, workspace, (cudaStream_t)stream);
}

}  // extern "C"
