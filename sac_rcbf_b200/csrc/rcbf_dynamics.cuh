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

// ---- float32 throughput path, written once for T = float (one instance per lane) and T = f2 (two instances per lane
// through FMUL2 / FADD2 / FFMA2).  Every operation is spelled out (rcbf_f2.cuh): the compiler contracts nothing, so
// both instantiations -- and therefore the scalar env.step kernel, the one-per-lane fused kernel and the two-per-lane
// fused kernel -- produce the same bits.  unicycle_env.py:46-111,215-280.
template <typename T>
struct UniEnvOutV {
  T obs[7];
  T reward;
  T cost;
  typename VecOf<T>::mask done, goal_met;
};

// the env parameters as the float32 path uses them (converted once, on the host or at kernel entry: the double -> float
// conversions and the two derived products are the same IEEE operations either way)
struct UniEnvF {
  float hz[kUniHaz][2];
  float r2;            // hazards_radius^2
  float dt, ndt, nk;   // dt, -dt, -(dt * 0.1)
  float gx, gy, goal_size, reward_goal;
  float init_x, init_y, init_th;
  int max_steps, auto_reset;
};
RCBF_HD UniEnvF make_env_f(const UnicycleEnvParams& p) {
  UniEnvF f;
  RCBF_UNROLL
  for (int i = 0; i < kUniHaz; ++i) {
    f.hz[i][0] = (float)p.hazards[i][0];
    f.hz[i][1] = (float)p.hazards[i][1];
  }
  const float hr = (float)p.hazards_radius;
  f.r2 = mul_rn(hr, hr);
  f.dt = (float)p.dt;
  f.ndt = -f.dt;
  f.nk = -mul_rn(f.dt, 0.1f);
  f.gx = (float)p.goal_x;
  f.gy = (float)p.goal_y;
  f.goal_size = (float)p.goal_size;
  f.reward_goal = (float)p.reward_goal;
  f.init_x = (float)p.init_x;
  f.init_y = (float)p.init_y;
  f.init_th = (float)p.init_theta;
  f.max_steps = p.max_episode_steps;
  f.auto_reset = p.auto_reset;
  return f;
}

template <typename T>
RCBF_HD T unicycle_goal_dist_v(const UniEnvF& p, T x, T y, T* vx_out, T* vy_out) {
  const T vx = sub_rn(T(p.gx), x), vy = sub_rn(T(p.gy), y);
  *vx_out = vx;
  *vy_out = vy;
  return sqrt_pos(t_fma(vx, vx, mul_rn(vy, vy)));
}

// s, c = sin / cos of the CURRENT heading (the constraint assembly already has them); the updated heading's pair comes
// from one plane rotation by the small per-step increment |delta| <= dt * 1 rad/s (Taylor to delta^5, relative error
// < 1e-12 at |delta| = 0.02) instead of a second full sincos.
template <typename T>
RCBF_HD void unicycle_env_step_v(const UniEnvF& p, T st[3], T& last_dist, typename VecOf<T>::ivec& step,
                                 const T a_in[2], T s, T c, UniEnvOutV<T>& o) {
  const T dt = T(p.dt);
  const T a0 = t_fmin(t_fmax(a_in[0], T(-1.f)), T(1.f));  // :62
  const T a1 = t_fmin(t_fmax(a_in[1], T(-1.f)), T(1.f));
  T x = t_fma(dt, mul_rn(c, a0), st[0]);                  // :86   state += dt * (f + g(state) @ action), f = 0
  T y = t_fma(dt, mul_rn(s, a0), st[1]);
  const T delta = mul_rn(dt, a1);
  st[2] = add_rn(st[2], delta);
  const T d2 = mul_rn(delta, delta);
  const T sp = t_fma(d2, t_fma(d2, T(8.333333333e-3f), T(-1.666666667e-1f)), T(1.0f));
  const T sd = mul_rn(delta, sp);
  const T nsd = mul_rn(mul_rn(T(p.ndt), a1), sp);         // -sd (a packed multiply is cheaper than two sign flips)
  const T cd = t_fma(d2, t_fma(d2, T(4.166666667e-2f), T(-0.5f)), T(1.0f));
  const T s1 = t_fma(s, cd, mul_rn(c, sd));               // sin / cos of the UPDATED heading
  const T c1 = t_fma(c, cd, mul_rn(s, nsd));
  const T nk = T(p.nk);
  x = t_fma(mul_rn(nk, c1), c1, x);                       // :87 uses g() and cos() of the UPDATED theta
  y = t_fma(mul_rn(nk, s1), c1, y);
  st[0] = x;
  st[1] = y;
  step = t_iadd(step, 1);                                 // :89
  T vx, vy;
  const T dist = unicycle_goal_dist_v<T>(p, x, y, &vx, &vy);
  const T rw = sub_rn(last_dist, dist);                   // :93-95
  last_dist = dist;
  const typename VecOf<T>::mask goal = t_le(dist, T(p.goal_size));  // :97,113-123
  o.reward = t_sel(goal, add_rn(rw, T(p.reward_goal)), rw);
  o.done = t_or(goal, t_ige(step, p.max_steps));          // :100-102
  o.goal_met = goal;
  T d2min = T(3.0e38f);  // any(d2_i < r2) == (min_i d2_i < r2); a NaN position compares false either way   :106
  RCBF_UNROLL
  for (int i = 0; i < kUniHaz; ++i) {
    const T dx = sub_rn(x, T(p.hz[i][0])), dy = sub_rn(y, T(p.hz[i][1]));
    d2min = t_fmin(d2min, t_fma(dx, dx, mul_rn(dy, dy)));
  }
  o.cost = t_sel(t_lt(d2min, T(p.r2)), T(0.1f), T(0.f));
  // obs = [x, y, cos th, sin th, compass_x, compass_y, exp(-dist)]     :215-231,260-277
  const T cx = t_fma(vx, c1, mul_rn(vy, s1));             // row-vector times R(theta)    :272-274
  const T cy = t_fma(vy, c1, mul_rn(sub_rn(x, T(p.gx)), s1));
  const T nrm = add_rn(sqrt_pos(t_fma(cx, cx, mul_rn(cy, cy))), T(0.001f));  // :276
  const T rn = rcp_refined(nrm);
  o.obs[0] = x;
  o.obs[1] = y;
  o.obs[2] = c1;
  o.obs[3] = s1;
  o.obs[4] = div_by_v(cx, nrm, rn);
  o.obs[5] = div_by_v(cy, nrm, rn);
  // exp(-dist) = 2^(-dist log2 e): the product's rounding error enters as a first-order correction (<= ~2 ulp)
#if defined(__CUDA_ARCH__)
  const T nhi = mul_rn(dist, T(-1.44269502e+00f));
  const T tl = t_fma(dist, T(1.92596299e-08f), t_fma(dist, T(1.44269502e+00f), nhi));
  const T ev = ex2_approx(nhi);
  o.obs[6] = t_fma(ev, mul_rn(tl, T(-6.93147182e-01f)), ev);
#else
  o.obs[6] = exp_neg(t_neg(dist));
#endif
}

// reset values (:125-143); the goal distance of the initial pose is the same for every instance
RCBF_HD float unicycle_reset_dist(const UniEnvF& p) {
  float vx, vy;
  return unicycle_goal_dist_v<float>(p, p.init_x, p.init_y, &vx, &vy);
}

RCBF_HD void unicycle_env_step_sc(const UnicycleEnvParams& p, float st[3], float& last_dist, int& step,
                                  const float a_in[2], float s, float c, UniEnvOut<float>& o) {
  const UniEnvF pf = make_env_f(p);
  UniEnvOutV<float> v;
  unicycle_env_step_v<float>(pf, st, last_dist, step, a_in, s, c, v);
  RCBF_UNROLL
  for (int j = 0; j < 7; ++j) o.obs[j] = v.obs[j];
  o.reward = v.reward;
  o.cost = v.cost;
  o.done = v.done;
  o.goal_met = v.goal_met;
}

template <typename T>
RCBF_HD void unicycle_reset(const UnicycleEnvParams& p, T st[3], T& last_dist, int& step) {  // :125-143
  st[0] = T(p.init_x);
  st[1] = T(p.init_y);
  st[2] = T(p.init_theta);
  step = 0;
  last_dist = unicycle_goal_dist(p, st);
}
template <>
RCBF_HD void unicycle_reset<float>(const UnicycleEnvParams& p, float st[3], float& last_dist, int& step) {  // float32: same
  st[0] = (float)p.init_x;                                                               // bits as the packed kernel
  st[1] = (float)p.init_y;
  st[2] = (float)p.init_theta;
  step = 0;
  last_dist = unicycle_reset_dist(make_env_f(p));
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
