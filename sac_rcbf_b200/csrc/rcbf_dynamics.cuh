// rcbf_dynamics.cuh -- per-instance environment steps and prior dynamics.
//
// Templated on the arithmetic/storage type T:
//   T = float   the throughput path (float4-packed Unicycle state, north_star item (1))
//   T = double  bit-faithful restatement of the reference's numpy float64 envs, used by the drop-in
//               num_envs=1 gym-style wrappers so 1000-step trajectories track the reference to ~1e-13.
//
// Reference: envs/unicycle_env.py:46-143,215-280 ; envs/simulated_cars_env.py:38-158 ;
//            rcbf_sac/dynamics.py:60-105,125-188 (prior model).
#pragma once

#include "rcbf_core.cuh"

namespace rcbf {

using UnicycleEnvParams = rcbf_unicycle_env_params;  // include/rcbf_b200.h
using CarsEnvParams = rcbf_cars_env_params;

RCBF_HD void sincos_t(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(x, s, c);
#else
  *s = sin(x);
  *c = cos(x);
#endif
}
RCBF_HD float t_exp(float x) { return expf(x); }
RCBF_HD double t_exp(double x) { return exp(x); }
RCBF_HD float t_sin(float x) { return sinf(x); }
RCBF_HD double t_sin(double x) { return sin(x); }

template <typename T>
struct UniEnvOut {
  T obs[7];
  T reward;
  T cost;     // 0.1 inside a hazard, else 0 (the reference omits info['cost'] when 0, unicycle_env.py:106-110)
  int done;
  int goal_met;
};

RCBF_HD float t_min(float a, float b) { return fminf(a, b); }
RCBF_HD double t_min(double a, double b) { return fmin(a, b); }
// (a / n, b / n): float32 shares one reciprocal and corrects each quotient once (div_by: correctly rounded for the
// n in [1e-3, 1e2] that occur here, i.e. the same bits as two IEEE divisions at a third of the instructions)
RCBF_HD void div2(float a, float b, float n, float& qa, float& qb) {
  const float r = rcp_refined(n);
  qa = div_by(a, n, r);
  qb = div_by(b, n, r);
}
RCBF_HD void div2(double a, double b, double n, double& qa, double& qb) {
  qa = a / n;
  qb = b / n;
}

// obs = [x, y, cos th, sin th, compass_x, compass_y, exp(-dist)]     unicycle_env.py:215-231,260-277
template <typename T>
RCBF_HD void unicycle_obs(const UnicycleEnvParams& p, const T st[3], T c, T s, T dist, T obs[7]) {
  const T vx = T(p.goal_x) - st[0], vy = T(p.goal_y) - st[1];
  const T cx = vx * c + vy * s;    // row-vector times R(theta)    :272-274
  const T cy = vx * (-s) + vy * c;
  const T nrm = t_sqrt(cx * cx + cy * cy) + T(0.001);  // :276
  obs[0] = st[0];
  obs[1] = st[1];
  obs[2] = c;
  obs[3] = s;
  div2(cx, cy, nrm, obs[4], obs[5]);
  obs[6] = t_exp(-dist);
}

template <typename T>
RCBF_HD T unicycle_goal_dist(const UnicycleEnvParams& p, const T st[3]) {
  const T vx = T(p.goal_x) - st[0], vy = T(p.goal_y) - st[1];
  return t_sqrt(vx * vx + vy * vy);
}

// sin/cos of theta + delta from sin/cos of theta for the small per-step heading increment |delta| <= dt * 1 rad/s:
// Taylor to delta^5 (relative error < 1e-12 at |delta| = 0.02) + one plane rotation.  Saves a full sincos per step.
RCBF_HD void rotate_small(float s0, float c0, float delta, float* s1, float* c1) {
  const float d2 = delta * delta;
  const float sd = delta * fmaf(d2, fmaf(d2, 8.333333333e-3f, -1.666666667e-1f), 1.0f);
  const float cd = fmaf(d2, fmaf(d2, 4.166666667e-2f, -0.5f), 1.0f);
  *s1 = fmaf(s0, cd, c0 * sd);
  *c1 = fmaf(c0, cd, -(s0 * sd));
}

template <typename T>
RCBF_HD void unicycle_env_finish(const UnicycleEnvParams& p, T st[3], T& last_dist, int& step, T s, T c,
                                 UniEnvOut<T>& o) {
  const T k = T(p.dt) * T(0.1);
  st[0] -= (k * c) * c;  // :87 uses g() and cos() of the UPDATED theta
  st[1] -= (k * s) * c;
  step += 1;  // :89
  const T dist = unicycle_goal_dist(p, st);
  T reward = last_dist - dist;  // :93-95
  last_dist = dist;
  const bool goal = dist <= T(p.goal_size);  // :97,113-123
  if (goal) reward += T(p.reward_goal);
  o.done = goal || (step >= p.max_episode_steps);  // :100-102
  o.goal_met = goal;
  const T r2 = T(p.hazards_radius) * T(p.hazards_radius);
  T d2min = T(3.0e38);  // any(d2_i < r2) == (min_i d2_i < r2); a NaN position compares false either way   :106
  RCBF_UNROLL
  for (int i = 0; i < kUniHaz; ++i) {
    const T dx = st[0] - T(p.hazards[i][0]), dy = st[1] - T(p.hazards[i][1]);
    d2min = t_min(d2min, dx * dx + dy * dy);
  }
  o.cost = (d2min < r2) ? T(0.1) : T(0);
  o.reward = reward;
  unicycle_obs(p, st, c, s, dist, o.obs);
}

// UnicycleEnv.step (:46-111).  st, last_dist, step are updated in place.
// float64: sin/cos evaluated exactly where numpy evaluates them (bit-faithful).  float32 (throughput layout): sin/cos
// of the current heading may be passed in (the fused kernel already has them from the constraint assembly) and the
// updated heading's pair comes from rotate_small.
template <typename T>
RCBF_HD void unicycle_env_step(const UnicycleEnvParams& p, T st[3], T& last_dist, int& step, const T a_in[2],
                               UniEnvOut<T>& o) {
  const T dt = T(p.dt);
  const T a0 = t_min(t_max(a_in[0], T(-1)), T(1));  // :62
  const T a1 = t_min(t_max(a_in[1], T(-1)), T(1));
  T s, c;
  sincos_t(st[2], &s, &c);
  st[0] += dt * (c * a0);  // :86   state += dt * (f + g(state) @ action), f = 0
  st[1] += dt * (s * a0);
  st[2] += dt * a1;
  sincos_t(st[2], &s, &c);
  unicycle_env_finish<T>(p, st, last_dist, step, s, c, o);
}

RCBF_HD void unicycle_env_step_sc(const UnicycleEnvParams& p, float st[3], float& last_dist, int& step,
                                  const float a_in[2], float s, float c, UniEnvOut<float>& o) {
  const float dt = (float)p.dt;
  const float a0 = fminf(fmaxf(a_in[0], -1.f), 1.f);  // :62
  const float a1 = fminf(fmaxf(a_in[1], -1.f), 1.f);
  st[0] += dt * (c * a0);  // :86
  st[1] += dt * (s * a0);
  const float delta = dt * a1;
  st[2] += delta;
  float s1, c1;
  rotate_small(s, c, delta, &s1, &c1);
  unicycle_env_finish<float>(p, st, last_dist, step, s1, c1, o);
}

template <typename T>
RCBF_HD void unicycle_reset(const UnicycleEnvParams& p, T st[3], T& last_dist, int& step) {  // :125-143
  st[0] = T(p.init_x);
  st[1] = T(p.init_y);
  st[2] = T(p.init_theta);
  step = 0;
  last_dist = unicycle_goal_dist(p, st);
}

template <typename T>
struct CarsEnvOut {
  T obs[10];
  T reward;
  T cost;
  int done;
};

// x / n for a compile-time n: float32 multiplies by the (correctly rounded) reciprocal and corrects the quotient once
// with the exact residual -- the fast path of an IEEE division without its range checks (|x| <= 1e6 here)
RCBF_HD float div_const(float x, float n, float r) { return div_by(x, n, r); }
RCBF_HD double div_const(double x, double n, double) { return x / n; }

template <typename T>
RCBF_HD void cars_obs(const T st[10], T obs[10]) {  // simulated_cars_env.py:143-158
  RCBF_UNROLL
  for (int i = 0; i < 5; ++i) {
    obs[2 * i] = div_const(st[2 * i], T(100), T(0.01));
    obs[2 * i + 1] = div_const(st[2 * i + 1], T(30), T(1.0 / 30.0));
  }
}

// SimulatedCarsEnv.step (:38-106).  st, t, step updated in place.
template <typename T>
RCBF_HD void cars_env_step(const CarsEnvParams& p, T st[10], T& t, int& step, T a, CarsEnvOut<T>& o) {
  T pos[5], vel[5], acc[5];
  RCBF_UNROLL
  for (int i = 0; i < 5; ++i) {
    pos[i] = st[2 * i];
    vel[i] = st[2 * i + 1];
  }
  const T v0 = T(30) - T(10) * t_sin(T(0.2) * t);            // :59-60
  cars_accels<T>(T(p.kp), T(p.k_brake), pos, vel, v0, acc);  // :61-64 (car 4 keeps its own P-term)
  const T dt = T(p.dt);
  RCBF_UNROLL
  for (int i = 0; i < 5; ++i) {
    const T ai = acc[i] * T(1.1);  // :67
    const T gi = (i == 3) ? T(50) * a : T(0);
    st[2 * i] = pos[i] + dt * vel[i];          // :77  state += dt * (f + g * action)
    st[2 * i + 1] = vel[i] + dt * (ai + gi);
  }
  t = t + dt;  // :79
  step += 1;   // :81
  o.done = step >= p.max_episode_steps;  // :83
  T cost = T(0);
  if (st[4] - st[6] < T(2.99)) cost -= T(0.1);  // :100-101
  if (st[6] - st[8] < T(2.99)) cost -= T(0.1);  // :103-104
  o.cost = cost;
  const T a2 = a * a;
  o.reward = T(-5) * (a2 < T(0) ? -a2 : a2) / T(p.max_episode_steps);  // :93
  cars_obs(st, o.obs);
}

template <typename T>
RCBF_HD void cars_reset(T st[10], T& t, int& step, T v_noise) {  // :108-125
  const T p0[5] = {T(34), T(28), T(22), T(16), T(10)};
  RCBF_UNROLL
  for (int i = 0; i < 5; ++i) {
    st[2 * i] = p0[i];
    st[2 * i + 1] = T(30) + v_noise;
  }
  st[7] = T(35);
  t = T(0);
  step = 0;
}

// ---------------------------------------------------------------------------------------------------
// prior model  next = s + dt (f(s,t) + g(s) u) + dt * mean        rcbf_sac/dynamics.py:86-92
// ---------------------------------------------------------------------------------------------------
template <typename T>
RCBF_HD void unicycle_prior_next(T dt, const T st[3], const T u[2], const T mean[3], T nxt[3]) {
  T s, c;
  sincos_t(st[2], &s, &c);
  nxt[0] = st[0] + dt * (c * u[0]);  // dynamics.py:145-151 g = [[c,0],[s,0],[0,1]], f = 0
  nxt[1] = st[1] + dt * (s * u[0]);
  nxt[2] = st[2] + dt * u[1];
  RCBF_UNROLL
  for (int j = 0; j < 3; ++j) nxt[j] += dt * mean[j];  // :92
}

template <typename T>
RCBF_HD void cars_prior_next(T dt, T kp, T kb, const T st[10], T u, T t, const T mean[10], T nxt[10]) {
  T pos[5], vel[5], acc[5];
  RCBF_UNROLL
  for (int i = 0; i < 5; ++i) {
    pos[i] = st[2 * i];
    vel[i] = st[2 * i + 1];
  }
  const T v0 = T(30) - T(10) * t_sin(T(0.2) * t);  // dynamics.py:172
  cars_accels<T>(kp, kb, pos, vel, v0, acc);       // :173-177
  acc[3] = T(0);                                   // :176 (and no x1.1: that gap is what the GP learns)
  RCBF_UNROLL
  for (int i = 0; i < 5; ++i) {
    const T gi = (i == 3) ? T(50) * u : T(0);  // :158-162
    nxt[2 * i] = pos[i] + dt * vel[i];
    nxt[2 * i + 1] = vel[i] + dt * (acc[i] + gi);
  }
  RCBF_UNROLL
  for (int j = 0; j < 10; ++j) nxt[j] += dt * mean[j];
}

// ---------------------------------------------------------------------------------------------------
// model rollouts: one transition of generate_model_rollouts        rcbf_sac/generate_rollouts.py:29-66
//   state = get_state(obs) ; next ~ N(prior_next(state, a) + dt*mean, (dt*std)^2) ; next_obs, reward, done
// `eps` is the standard-normal draw (np.random.normal(mu, std) == mu + std * eps), passed in so that the step is
// deterministic given its inputs.
// ---------------------------------------------------------------------------------------------------
RCBF_HD float t_atan2(float y, float x) { return atan2f(y, x); }
RCBF_HD double t_atan2(double y, double x) { return atan2(y, x); }
RCBF_HD float t_log(float x) { return logf(x); }
RCBF_HD double t_log(double x) { return log(x); }

template <typename T>
RCBF_HD void unicycle_rollout_step(T dt, T goal_x, T goal_y, const T obs[7], const T a[2], const T mean[3],
                                   const T std[3], const T eps[3], T next_obs[7], T& reward, int& done) {
  T st[3] = {obs[0], obs[1], t_atan2(obs[3], obs[2])};  // dynamics.py:216-221
  T nx[3];
  unicycle_prior_next<T>(dt, st, a, mean, nx);          // generate_rollouts.py:30 (dynamics.py:86-92)
  RCBF_UNROLL
  for (int j = 0; j < 3; ++j) nx[j] = nx[j] + (dt * std[j]) * eps[j];  // :31
  T s, c;
  sincos_t(nx[2], &s, &c);
  const T dist_prev = -t_log(obs[6]);                   // :37  (obs[-1] = exp(-dist))
  const T gx = goal_x - nx[0], gy = goal_y - nx[1];     // :38
  const T dist = t_sqrt(gx * gx + gy * gy);             // :39
  const T cx = gx * c + gy * s, cy = gx * (-s) + gy * c;  // :42 (row vector times R(theta'))
  const T nrm = t_sqrt(cx * cx + cy * cy) + T(0.001);   // :43
  next_obs[0] = nx[0];
  next_obs[1] = nx[1];
  next_obs[2] = c;
  next_obs[3] = s;
  next_obs[4] = cx / nrm;
  next_obs[5] = cy / nrm;
  next_obs[6] = t_exp(-dist);                           // :44
  const bool reached = dist <= T(0.3);                  // :47,52
  // :50 + :53 -- the reference adds reward_goal TWICE when the goal is reached
  reward = (dist_prev - dist) * T(1) + (reached ? T(1) : T(0)) + (reached ? T(1) : T(0));
  done = reached;                                       // :54
}

template <typename T>
RCBF_HD void cars_rollout_step(T dt, T kp, T kb, int max_steps, const T obs[10], T a, T t, const T mean[10],
                               const T std[10], const T eps[10], T next_obs[10], T& reward, int& done, T& next_t) {
  T st[10], nx[10];
  RCBF_UNROLL
  for (int i = 0; i < 5; ++i) {                         // dynamics.py:222-225
    st[2 * i] = obs[2 * i] * T(100);
    st[2 * i + 1] = obs[2 * i + 1] * T(30);
  }
  cars_prior_next<T>(dt, kp, kb, st, a, t, mean, nx);   // generate_rollouts.py:30
  RCBF_UNROLL
  for (int j = 0; j < 10; ++j) nx[j] = nx[j] + (dt * std[j]) * eps[j];  // :31
  cars_obs<T>(nx, next_obs);                            // :32
  const T a2 = a * a;
  reward = T(-5) * (a2 < T(0) ? -a2 : a2) / T(max_steps);  // :61
  next_t = t + dt;                                      // dynamics.py:102
  done = next_t >= T(max_steps) * dt;                   // :64
}

}  // namespace rcbf
