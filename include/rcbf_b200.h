/* rcbf_b200.h -- C ABI of librcbf_b200.so: the sm_100a implementation of SAC-RCBF's per-step safety hot path.
 *
 * Drop-in boundary.  The reference (yemam3/SAC-RCBF) is pure Python; the functions below are what a ctypes / torch
 * binding replaces in it (see INTEGRATION.md for the stubs):
 *
 *   rcbf_*_assemble          CBFQPLayer.get_cbf_qp_constraints          rcbf_sac/diff_cbf_qp.py:146-379
 *   rcbf_qp_solve            CBFQPLayer.cbf_layer / solve_qp (qpth)     rcbf_sac/diff_cbf_qp.py:81-144
 *   rcbf_*_safe_action       CBFQPLayer.get_safe_action (forward)       rcbf_sac/diff_cbf_qp.py:44-79
 *   rcbf_*_safe_action_bwd   autograd of the above (qpth backward)      rcbf_sac/diff_cbf_qp.py:139 + :103-106,:77
 *   rcbf_*_env_step_*        UnicycleEnv.step / SimulatedCarsEnv.step   envs/unicycle_env.py:46-111 ; envs/simulated_cars_env.py:38-87
 *   rcbf_*_env_reset_*       .reset()                                   envs/unicycle_env.py:125-143 ; envs/simulated_cars_env.py:108-125
 *   rcbf_*_predict_next_*    DynamicsModel.predict_next_state (prior)   rcbf_sac/dynamics.py:60-105
 *   rcbf_*_safe_step         get_safe_action + env.step fused           rcbf_sac/sac_cbf.py:218-238 + main.py:93-95
 *   rcbf_*_rollout_step_*    one transition of generate_model_rollouts  rcbf_sac/generate_rollouts.py:29-66
 *   rcbf_replay_push/sample  ReplayMemory.batch_push / .sample           rcbf_sac/replay_memory.py:12-32
 *   (CascadeCBFLayer.get_u_safe, rcbf_sac/cbf_qp.py:29-53, is rcbf_unicycle_safe_action with sigma_scale = k_d,
 *    abs_sigma_map = 0, p_diag = (10, 1e-4, 1e7) and the unclamped correction read from `x`.)
 *
 * Conventions: every pointer is a DEVICE pointer unless its name ends in _host; arrays are row-major (n, feat) like
 * the reference's tensors; `n` instances; `stream` is a cudaStream_t passed as void*; nullable outputs are marked.
 * Every call is asynchronous on `stream`, allocates nothing and returns 0 or the cudaError_t of the launch.
 */
#ifndef RCBF_B200_H
#define RCBF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCBF_UNI_HAZ 5 /* envs/unicycle_env.py:26 */
#define RCBF_UNI_M 9   /* 5 CBF rows + 4 actuator rows, diff_cbf_qp.py:42 */
#define RCBF_UNI_NZ 3  /* (u_v, u_w, eps) */
#define RCBF_CARS_M 4
#define RCBF_CARS_NZ 2

/* per-instance status written by the solvers */
#define RCBF_OK_TRIVIAL 0   /* x = 0 feasible */
#define RCBF_OK_CERTIFIED 1 /* exact KKT certificate */
#define RCBF_OK_IPM 2       /* interior-point residual < tol (float64 pass) */
#define RCBF_MAXITER 3      /* iteration limit: best iterate returned (what qpth does) */
#define RCBF_NAN 4          /* -> Python raises Exception('QP Failed to solve'), diff_cbf_qp.py:141-143 */

typedef struct {
  float hazards[RCBF_UNI_HAZ][2];
  float collision_radius_sq; /* float((1.2*hazards_radius)^2) diff_cbf_qp.py:207,246 */
  float gamma_b;
  float l_p;
  float sigma_scale;  /* 1 (CBFQPLayer: k_d is dead, :261) or k_d (CascadeCBFLayer, cbf_qp.py:141) */
  int abs_sigma_map;  /* 1: |g_p| (diff_cbf_qp.py:241); 0: signed (cbf_qp.py:119) */
  float u_min[2], u_max[2];
  float p_diag[3];
  int solver_mode; /* 0: greedy active-set presolve + KKT certificate, interior point as fallback (default);
                      1: "pdipm": every non-trivial QP goes through the primal-dual interior point */
} rcbf_unicycle_params;

typedef struct {
  float gamma_2, gamma_sq; /* float(gamma_b+gamma_b), float(gamma_b*gamma_b) diff_cbf_qp.py:348 */
  float kp, k_brake;
  float collision_radius_sq;
  float sigma_scale;
  float u_min, u_max;
  float p_diag[2];
  float slack_coeff;
  int solver_mode; /* as in rcbf_unicycle_params */
} rcbf_cars_params;

typedef struct {
  double hazards[RCBF_UNI_HAZ][2];
  double hazards_radius;
  double dt, goal_x, goal_y, goal_size, reward_goal;
  double init_x, init_y, init_theta;
  int max_episode_steps;
  int auto_reset;
} rcbf_unicycle_env_params;

typedef struct {
  double dt, kp, k_brake;
  int max_episode_steps;
  int auto_reset;
} rcbf_cars_env_params;

/* Solver workspace: RCBF_WS_WORDS 64-bit words owned by the caller, zero-initialised once, one per concurrent stream.
 *   [0]=#NaN  [1]=#uncertified(max-iter)  [2]=#float64 interior-point passes  [3]=#trivial
 *   [4]=sum over instances of pass-1 iterations (presolve rounds in mode 0, interior-point iterations in mode 1)
 *   [5]=#instances left pending by the fast path  [6]=sum of their interior-point iterations  [7] reserved
 *   [8..15] queue bookkeeping, [16..) queue of pending instances (index + 1, 0 = empty; reset by the library).
 * The counters [0..7] accumulate with atomics across calls; zero them when you want per-call numbers.  With a workspace
 * the default solver mode finishes pending instances inside the same kernel (one launch per call).  Nullable: a second
 * kernel then scans safe_action for the pending sentinel instead (slower, two launches). */
#define RCBF_WS_WORDS 32768
typedef unsigned long long rcbf_counters_t;

/* ---- constraint assembly (raw P,q are constants: P = diag(p_diag), q = 0) ------------------------------------- */
int rcbf_unicycle_assemble(const float* state, const float* action, const float* mean, const float* sigma, int64_t n,
                           const rcbf_unicycle_params* p_host, float* G /* n*9*3 */, float* h /* n*9 */, void* stream);
int rcbf_cars_assemble(const float* state, const float* action, const float* sigma, int64_t n,
                       const rcbf_cars_params* p_host, float* G /* n*4*2 */, float* h /* n*4 */, void* stream);

/* float64 variants for the single-instance numpy layer (rcbf_sac/cbf_qp.py:84-240 works in float64): double in, double
 * out; normalise != 0 also applies that file's [G|h] row normalisation (:262-265), so the result feeds rcbf_qp_solve. */
int rcbf_unicycle_assemble_f64(const double* state, const double* action, const double* mean, const double* sigma,
                               int64_t n, const rcbf_unicycle_params* p_host, int normalise, double* G /* n*9*3 */,
                               double* h /* n*9 */, void* stream);
int rcbf_cars_assemble_f64(const double* state, const double* action, const double* sigma, int64_t n,
                           const rcbf_cars_params* p_host, int normalise, double* G /* n*4*2 */, double* h /* n*4 */,
                           void* stream);

/* ---- get_safe_action forward.  x/lam/slack (float32, saved for the backward), status, iters: nullable ---------- */
int rcbf_unicycle_safe_action(const float* state, const float* action, const float* mean, const float* sigma,
                              int64_t n, const rcbf_unicycle_params* p_host, float* safe_action /* n*2 */,
                              float* x /* n*3 */, float* lam /* n*9 */, float* slack /* n*9 */, int32_t* status,
                              int32_t* iters, rcbf_counters_t* workspace, void* stream);
int rcbf_cars_safe_action(const float* state, const float* action, const float* sigma, int64_t n,
                          const rcbf_cars_params* p_host, float* safe_action /* n*1 */, float* x /* n*2 */,
                          float* lam /* n*4 */, float* slack /* n*4 */, int32_t* status, int32_t* iters,
                          rcbf_counters_t* workspace, void* stream);

/* ---- get_safe_action backward: d loss / d action given d loss / d safe_action -------------------------------- */
int rcbf_unicycle_safe_action_bwd(const float* state, const float* action, const float* mean, const float* sigma,
                                  const float* x, const float* lam, const float* slack, const float* grad_out,
                                  int64_t n, const rcbf_unicycle_params* p_host, float* grad_action, void* stream);
int rcbf_cars_safe_action_bwd(const float* state, const float* action, const float* sigma, const float* x,
                              const float* lam, const float* slack, const float* grad_out, int64_t n,
                              const rcbf_cars_params* p_host, float* grad_action, void* stream);

/* ---- the differentiable path in compact form (what CBFQPLayer's autograd function uses) ---------------------------
 * The forward saves ONE int32 per instance, meta = (status << 16) | active-set mask (bit i = row i of G~x <= h~ is
 * active at the optimum; 0xffff: no vertex, interior-point result), instead of x / lam / slack (84 bytes).  The backward
 * finishes a trivial instance (status RCBF_OK_TRIVIAL: x = 0, gradient = the clamp mask of diff_cbf_qp.py:77) from 28
 * bytes; for the others it re-assembles the constraints and differentiates the closed-form solution on the <= nz active
 * rows (the exact active-set form of the implicit-KKT backward: rcbf_*_safe_action_bwd with qpth's 1e-8 clamps taken to
 * zero; the two agree to ~1e-5 of the batch gradient norm).  One launch, no scratch memory. */
int rcbf_unicycle_safe_action_saved(const float* state, const float* action, const float* mean, const float* sigma,
                                    int64_t n, const rcbf_unicycle_params* p_host, float* safe_action /* n*2 */,
                                    int32_t* meta /* n */, rcbf_counters_t* workspace, void* stream);
int rcbf_cars_safe_action_saved(const float* state, const float* action, const float* sigma, int64_t n,
                                const rcbf_cars_params* p_host, float* safe_action /* n */, int32_t* meta /* n */,
                                rcbf_counters_t* workspace, void* stream);
int rcbf_unicycle_safe_action_bwd_meta(const float* state, const float* action, const float* mean, const float* sigma,
                                       const int32_t* meta, const float* grad_out, int64_t n,
                                       const rcbf_unicycle_params* p_host, float* grad_action, void* stream);
int rcbf_cars_safe_action_bwd_meta(const float* state, const float* action, const float* sigma, const int32_t* meta,
                                   const float* grad_out, int64_t n, const rcbf_cars_params* p_host, float* grad_action,
                                   void* stream);

/* ---- Unicycle layer on an arbitrary hazard set (the reference sizes its layer from len(env.hazards_locations),
 * rcbf_sac/diff_cbf_qp.py:35,243-261).  hazards_xy_host: n_hazards x 2 floats on the HOST, 1 <= n_hazards <=
 * RCBF_MAX_HAZARDS (p_host->hazards is ignored; every other field of p_host applies).  One instance per thread, same
 * per-instance source as the 5-hazard hot kernels (rows: n_hazards CBF rows, then the 4 actuator rows; meta's active-set
 * mask indexes rows of the PADDED problem: 8 or 12 CBF rows, then the actuator rows).  Returns -1 for an unsupported
 * hazard count. */
#define RCBF_MAX_HAZARDS 12
int rcbf_unicycle_safe_action_general(const float* state, const float* action, const float* mean, const float* sigma,
                                      int64_t n, const rcbf_unicycle_params* p_host, const float* hazards_xy_host,
                                      int n_hazards, float* safe_action /* n*2 */, int32_t* meta /* n, nullable */,
                                      int32_t* status /* n, nullable */, rcbf_counters_t* counters /* nullable */,
                                      void* stream);
int rcbf_unicycle_safe_action_bwd_general(const float* state, const float* action, const float* mean, const float* sigma,
                                          const int32_t* meta, const float* grad_out, int64_t n,
                                          const rcbf_unicycle_params* p_host, const float* hazards_xy_host, int n_hazards,
                                          float* grad_action, void* stream);
int rcbf_unicycle_assemble_general(const float* state, const float* action, const float* mean, const float* sigma,
                                   int64_t n, const rcbf_unicycle_params* p_host, const float* hazards_xy_host,
                                   int n_hazards, float* G /* n*(n_hazards+4)*3 */, float* h /* n*(n_hazards+4) */,
                                   void* stream);

/* ---- generic small QP  min 1/2 x'Qx + p'x  s.t. Gx <= h  (cbf_layer / solve_qp), float64 tensors like qpth sees ---
 * (nz, m) in {(3,9), (3,12), (3,16), (2,4)}; a system with fewer rows is padded by the caller with copies of one of its
 * rows (a duplicate constraint changes neither the feasible set nor the optimum).  (solve_qp's [G|h] row normalisation is applied by the caller, diff_cbf_qp.py:103-106.) */
int rcbf_qp_solve(const double* Q /* n*nz*nz */, const double* p /* n*nz */, const double* G /* n*m*nz */,
                  const double* h /* n*m */, int64_t n, int nz, int m, double* x /* n*nz */, double* lam /* n*m */,
                  double* slack /* n*m */, int32_t* status, int32_t* iters, rcbf_counters_t* counters, void* stream);
int rcbf_qp_solve_bwd(const double* Q, const double* G, const double* x, const double* lam, const double* slack,
                      const double* grad_x, int64_t n, int nz, int m, double* dQ, double* dp, double* dG, double* dh,
                      void* stream);

/* ---- environments.  *_f32: throughput layout; *_f64: bit-faithful to the numpy reference ----------------------
 * Unicycle state: 4 values per instance (x, y, theta, last_goal_dist) -> one 16-byte load in the f32 layout.
 * Cars state: 10 values per instance + t + episode_step. */
int rcbf_unicycle_env_reset_f32(float* state4, int32_t* step, const uint8_t* mask /* nullable: all */, int64_t n,
                                const rcbf_unicycle_env_params* e_host, float* obs /* n*7 nullable */, void* stream);
int rcbf_unicycle_env_reset_f64(double* state4, int32_t* step, const uint8_t* mask, int64_t n,
                                const rcbf_unicycle_env_params* e_host, double* obs, void* stream);
int rcbf_unicycle_env_step_f32(float* state4, int32_t* step, const float* action, int64_t n,
                               const rcbf_unicycle_env_params* e_host, float* obs /* n*7 */, float* reward,
                               uint8_t* done, float* cost, uint8_t* goal_met, void* stream);
int rcbf_unicycle_env_step_f64(double* state4, int32_t* step, const double* action, int64_t n,
                               const rcbf_unicycle_env_params* e_host, double* obs, double* reward, uint8_t* done,
                               double* cost, uint8_t* goal_met, void* stream);
int rcbf_cars_env_reset_f32(float* state, float* t, int32_t* step, const float* v_noise, const uint8_t* mask,
                            int64_t n, float* obs /* nullable */, void* stream);
int rcbf_cars_env_reset_f64(double* state, double* t, int32_t* step, const double* v_noise, const uint8_t* mask,
                            int64_t n, double* obs, void* stream);
int rcbf_cars_env_step_f32(float* state, float* t, int32_t* step, const float* action, int64_t n,
                           const rcbf_cars_env_params* e_host, float* obs /* n*10 */, float* reward, uint8_t* done,
                           float* cost, void* stream);
int rcbf_cars_env_step_f64(double* state, double* t, int32_t* step, const double* action, int64_t n,
                           const rcbf_cars_env_params* e_host, double* obs, double* reward, uint8_t* done,
                           double* cost, void* stream);

/* ---- prior model (DynamicsModel.predict_next_state): next = s + dt (f + g u) + dt*mean ; mean nullable (= 0) ---- */
int rcbf_unicycle_predict_next_f32(const float* state, const float* action, const float* mean, int64_t n, double dt,
                                   float* next, void* stream);
int rcbf_unicycle_predict_next_f64(const double* state, const double* action, const double* mean, int64_t n,
                                   double dt, double* next, void* stream);
int rcbf_cars_predict_next_f32(const float* state, const float* action, const float* t, const float* mean, int64_t n,
                               double dt, double kp, double k_brake, float* next, void* stream);
int rcbf_cars_predict_next_f64(const double* state, const double* action, const double* t, const double* mean,
                               int64_t n, double dt, double kp, double k_brake, double* next, void* stream);

/* ---- model-rollout transition (generate_model_rollouts, rcbf_sac/generate_rollouts.py:29-66): state = get_state(obs),
 * next ~ N(prior_next(state, a) + dt*mean, (dt*std)^2) with the standard-normal draw `eps` supplied by the caller,
 * next_obs / reward / done rebuilt like the reference (incl. its double reward_goal).  mean, std, eps nullable. ------ */
int rcbf_unicycle_rollout_step_f32(const float* obs /* n*7 */, const float* action, const float* mean,
                                   const float* std, const float* eps, int64_t n, double dt, double goal_x,
                                   double goal_y, float* next_obs /* n*7 */, float* reward, uint8_t* done, void* stream);
int rcbf_unicycle_rollout_step_f64(const double* obs, const double* action, const double* mean, const double* std,
                                   const double* eps, int64_t n, double dt, double goal_x, double goal_y,
                                   double* next_obs, double* reward, uint8_t* done, void* stream);
int rcbf_cars_rollout_step_f32(const float* obs /* n*10 */, const float* action, const float* t, const float* mean,
                               const float* std, const float* eps, int64_t n, double dt, double kp, double k_brake,
                               int max_steps, float* next_obs, float* reward, uint8_t* done, float* next_t,
                               void* stream);
int rcbf_cars_rollout_step_f64(const double* obs, const double* action, const double* t, const double* mean,
                               const double* std, const double* eps, int64_t n, double dt, double kp, double k_brake,
                               int max_steps, double* next_obs, double* reward, uint8_t* done, double* next_t,
                               void* stream);

/* ---- fused safe step: assemble + QP + clamp + env.step in ONE launch (float32 env layout) --------------------- */
int rcbf_unicycle_safe_step(float* state4, int32_t* step, const float* action_rl, const float* mean,
                            const float* sigma, int64_t n, const rcbf_unicycle_params* p_host,
                            const rcbf_unicycle_env_params* e_host, float* safe_action /* n*2 */, float* obs /* n*7 */,
                            float* reward, uint8_t* done, float* cost, uint8_t* goal_met, int32_t* status /* nullable */,
                            rcbf_counters_t* workspace, void* stream);
int rcbf_cars_safe_step(float* state, float* t, int32_t* step, const float* action_rl, const float* sigma, int64_t n,
                        const rcbf_cars_params* p_host, const rcbf_cars_env_params* e_host, float* safe_action,
                        float* obs /* n*10 */, float* reward, uint8_t* done, float* cost, int32_t* status,
                        rcbf_counters_t* workspace, void* stream);

/* ---- host-buffer entry points (the e2e path): pinned or pageable HOST arrays in, HOST arrays out; the library
 * stages through its own device scratch and pipelines H2D / compute / D2H over `chunks` slices on internal streams.
 * Synchronous (returns when the outputs are valid). */
int rcbf_unicycle_safe_action_host(const float* state_host, const float* action_host, const float* mean_host,
                                   const float* sigma_host, int64_t n, const rcbf_unicycle_params* p_host,
                                   float* safe_action_host, int32_t* n_failed_host, int device, int chunks);
int rcbf_cars_safe_action_host(const float* state_host, const float* action_host, const float* sigma_host, int64_t n,
                               const rcbf_cars_params* p_host, float* safe_action_host, int32_t* n_failed_host,
                               int device, int chunks);

/* fused step with HOST inputs/outputs: the env state (state4/state, t, step) stays on the DEVICE, the per-step inputs
 * (u_rl, GP mean/std) come from host arrays and the per-step outputs go back to host arrays.  Same pipeline.
 * Every *_host OUTPUT pointer is nullable: an output the caller does not need on the host is not copied back (a caller
 * that only wants safe_action / reward / done / cost moves 17 instead of 46 bytes per instance over PCIe).
 * The library keeps its streams, staging buffers and workspaces per device; the entry points are thread safe
 * (serialised by one lock). */
int rcbf_unicycle_safe_step_host(float* state4, int32_t* step, const float* action_host, const float* mean_host,
                                 const float* sigma_host, int64_t n, const rcbf_unicycle_params* p_host,
                                 const rcbf_unicycle_env_params* e_host, float* safe_action_host, float* obs_host,
                                 float* reward_host, uint8_t* done_host, float* cost_host, uint8_t* goal_met_host,
                                 int32_t* n_failed_host, int device, int chunks);
int rcbf_cars_safe_step_host(float* state, float* t, int32_t* step, const float* action_host, const float* sigma_host,
                             int64_t n, const rcbf_cars_params* p_host, const rcbf_cars_env_params* e_host,
                             float* safe_action_host, float* obs_host, float* reward_host, uint8_t* done_host,
                             float* cost_host, int32_t* n_failed_host, int device, int chunks);

/* ---- disturbance-GP posterior ----------------------------------------------------------
 * Replaces GPyDisturbanceEstimator.predict (rcbf_sac/gp_model.py:86-114) + the fitted branch of
 * DynamicsModel.predict_disturbance (rcbf_sac/dynamics.py:371-379) for all n_gp output dimensions in ONE launch:
 *   z*      = test_x * inv_x_scale                              (dynamics.py:375)
 *   k*_j    = outputscale_g * exp(-|z* - z_j|^2 * inv_2l2_g)    (ScaleKernel(RBFKernel), gp_model.py:17-20)
 *   w       = F_g k*            F_g (rank_g x n) with  F_g^T F_g = (K_g + noise_g I)^-1  (exactly, or truncated to
 *                               the numerical rank by the host side, which validates the truncation when it builds F)
 *   mean    = y_scale_g * (w . proj_y_g)                        proj_y_g = F_g y_g          (gp_model.py:99, dynamics.py:378)
 *   std     = y_scale_g * sqrt(max(outputscale_g - |w|^2 + include_noise * noise_g, min_variance))   (:100,:379)
 * All arithmetic is float64 (outputscale - |w|^2 cancels ~ n outputscale / noise digits).
 * Device arrays, caller-owned, read-only:
 *   train_z     [n_pad][dim_pad]   normalised training inputs, zero padded; n_pad % 64 == 0, dim_pad in {4, 8, 12, 16}
 *   inv_x_scale [dim_pad]          zero beyond n_in
 *   hyp         [n_gp][4]          {inv_2l2, outputscale, noise, y_scale}
 *   r_tiles     [n_gp]             number of tile_rows-wide row tiles of F_g actually used (<= max_tiles)
 *   factor      [n_gp][max_tiles][n_pad][tile_rows]   F_g, tile-major and transposed (zero padded)
 *   proj_y      [n_gp][max_tiles * tile_rows] */
typedef struct {
  const double* train_z;
  const double* inv_x_scale;
  const double* hyp;
  const int32_t* r_tiles;
  const double* factor;
  const double* proj_y;
  int32_t n_pad, n_in, dim_pad, n_gp, max_tiles, tile_rows; /* tile_rows in {4, 8, 16, 64} */
  int32_t include_noise;                                    /* 1: 'f_var' as the reference returns it (gp_model.py:100) */
  double min_variance;                                      /* gpytorch float32 clamp: 1e-6 */
  /* far-field fast path (optional, ff_coef == NULL disables it; needs max_tiles == 1): when every
   * a_j = |z* - z_j|^2 * inv_2l2 is tiny -- the reference pins the lengthscale at 1e5 (gp_model.py:18-21), a ~ 1e-8 --
   * exp(-a) = 1 - a + a^2/2 to float64 accuracy and w = F k* collapses to a quadratic-in-(z*, |z*|^2) polynomial whose
   * coefficients the host sums once:  O(rank * n_in^2) per test point instead of O(n * rank).
   *   ff_coef [n_gp][tile_rows][3 + 2 dim_pad + dim_pad (dim_pad + 1) / 2]:  w_r = c0 + s (c1 + s c2)
   *            + sum_k z*_k (lin_k + s slin_k) + sum_{k<=l} quad_kl z*_k z*_l,   s = |z*|^2
   *   ff_amax [n_gp]: a test point may use the polynomial for GP g iff (|z*| + ff_zmax)^2 * inv_2l2_g <= ff_amax[g]
   *            (the host derives it from the a^3/6 remainder so that mean / variance move by < 1e-9 relative);
   *            other points are evaluated exactly by the same kernel (their warp sums over the training set). */
  const double* ff_coef;
  const double* ff_amax;
  double ff_zmax; /* max_j |z_j| */
  int64_t test_stride; /* elements between consecutive test_x rows; 0 = n_in (dense).  Lets the float4 env state
                        * (x, y, theta, last_goal_dist) of rcbf_unicycle_env_step_f32 be read in place with stride 4 */
} rcbf_gp_posterior;

/* test_x (n_test, n_in) row-major -> mean, std (n_test, n_gp) row-major, same scalar type as test_x. */
int rcbf_gp_predict_f32(const float* test_x, int64_t n_test, const rcbf_gp_posterior* post_host, float* mean,
                        float* std, void* stream);
int rcbf_gp_predict_f64(const double* test_x, int64_t n_test, const rcbf_gp_posterior* post_host, double* mean,
                        double* std, void* stream);

/* fused step with the disturbance GP evaluated ON THE DEVICE from the resident state (what RCBF_SAC.get_safe_action does
 * logically, rcbf_sac/sac_cbf.py:230-236: predict_disturbance(state) -> get_safe_action): per slice the posterior
 * kernel reads the float4 state in place (test_stride = 4) and its mean / std feed the fused step without ever leaving
 * the GPU, so the only host input is the nominal action (8 bytes per instance).  post_host->n_in and n_gp must be 3.
 * Output pointers are nullable as above. */
int rcbf_unicycle_safe_step_host_gp(float* state4, int32_t* step, const float* action_host,
                                    const rcbf_gp_posterior* post_host, int64_t n, const rcbf_unicycle_params* p_host,
                                    const rcbf_unicycle_env_params* e_host, float* safe_action_host, float* obs_host,
                                    float* reward_host, uint8_t* done_host, float* cost_host, uint8_t* goal_met_host,
                                    int32_t* n_failed_host, int device, int chunks);

/* ---- device-resident replay ring ------------------------------------------------------------------
 * Replaces ReplayMemory (rcbf_sac/replay_memory.py:4-35): the list of (state, action, reward, next_state, mask, t,
 * next_t) tuples (:17) becomes seven caller-owned device arrays of `capacity` rows, field order as in the tuple:
 *   field[0] state (capacity, obs_dim)   field[1] action (capacity, action_dim)   field[2] reward (capacity)
 *   field[3] next_state (capacity, obs_dim)   field[4] mask   field[5] t   field[6] next_t   (capacity each)
 * all of one scalar type (elem_bytes = 4: float32, 8: float64).  row_stride[f] = elements between consecutive rows of
 * field f (0: the field's own width, i.e. a separate dense array).  The layout this library is built for is ROW-MAJOR:
 * the seven fields of a transition adjacent in one (capacity, stride) matrix -- field[f] = base + (elements of the
 * fields before f), every row_stride[f] = stride, base and stride * elem_bytes multiples of 16 (pad the stride to 32
 * bytes), stride * elem_bytes <= 256 -- so that a drawn transition is three 32-byte sectors instead of nine; any other
 * layout works through word-granular kernels.  The ring cursor (position, size) stays with the caller, as
 * `self.position` / `len(self.buffer)` do in the reference. */
#define RCBF_REPLAY_FIELDS 7
typedef struct {
  void* field[RCBF_REPLAY_FIELDS];
  int64_t row_stride[RCBF_REPLAY_FIELDS];
  int64_t capacity;
  int32_t obs_dim, action_dim, elem_bytes;
} rcbf_replay_ring;

/* batch_push (replay_memory.py:20-26 = n x push, :12-18): row i of every src[f] (n rows, same layout and scalar type as
 * the ring field) goes to ring row (position + i) % capacity.  0 <= position < capacity, 0 <= n <= capacity (of a longer
 * batch only the newest `capacity` rows survive n pushes; the caller passes those).  src[5] / src[6] may be NULL
 * (push without t / next_t, :26): those ring fields are left as they are.  One launch. */
int rcbf_replay_push(const rcbf_replay_ring* ring_host, int64_t position, const void* const src[RCBF_REPLAY_FIELDS],
                     int64_t n, void* stream);

/* sample (replay_memory.py:28-32 = random.sample + seven np.stack): out[f] (batch rows, contiguous) receives ring rows
 * perm(0), ..., perm(batch - 1), where perm is a bijection of [0, size) selected by `key` (Feistel network on
 * ceil(log2 size) bits, cycle-walked into the range): `batch` DISTINCT rows, i.e. sampling without replacement, drawn
 * and gathered in one launch.  1 <= size <= capacity, 0 <= batch <= size (random.sample raises ValueError beyond; the
 * Python class does the same before calling).  out[f] may be NULL (field not wanted); idx_out (batch int64, nullable)
 * receives the drawn row indices. */
int rcbf_replay_sample(const rcbf_replay_ring* ring_host, int64_t size, int64_t batch, uint64_t key,
                       void* const out[RCBF_REPLAY_FIELDS], int64_t* idx_out, void* stream);

/* ---- low-latency completion / counter read-back (small batches: the reference's B = 1 / 25 / 256 call shapes) ----------
 * rcbf_counters_publish enqueues a one-warp kernel that copies the 8 counters of `workspace` into host_mirror[1..8] and
 * then stores `token` into host_mirror[0].  host_mirror is 9 words of page-locked host memory (cudaHostAlloc / a pinned
 * torch tensor: device-accessible under unified addressing).  A host thread that polls host_mirror[0] for its token knows
 * that every launch enqueued on `stream` before this one has finished and holds the counters (counter 0 = NaN results:
 * the check of rcbf_sac/diff_cbf_qp.py:141-143) -- about 10 us less than cudaMemcpyAsync + cudaStreamSynchronize. */
int rcbf_counters_publish(const rcbf_counters_t* workspace, uint64_t* host_mirror /* 9 words, pinned */, uint64_t token,
                          void* stream);
int rcbf_stream_synchronize(void* stream); /* cudaStreamSynchronize (for hosts that hold no CUDA runtime binding) */
/* In-kernel publication (saves the second launch): bind a 9-word page-locked host mirror to a workspace once
 * (rcbf_counters_bind_mirror stores its address in workspace word 11), then OR
 *     RCBF_SOLVER_PUBLISH | (token << RCBF_SOLVER_TOKEN_SHIFT)          (token < 2^22)
 * into the `solver_mode` of the params passed to rcbf_*_safe_action / _saved / _safe_step (solver mode 0, workspace
 * given): the last block of the call's LAST kernel copies the 8 counters into host_mirror[1..8] and then stores the
 * token into host_mirror[0].  Bits 0..7 of solver_mode stay the solver mode.  Without a bound mirror the bits are
 * ignored. */
#define RCBF_SOLVER_PUBLISH 0x100
#define RCBF_SOLVER_TOKEN_SHIFT 9
int rcbf_counters_bind_mirror(rcbf_counters_t* workspace, uint64_t* host_mirror /* 9 words, pinned; NULL unbinds */,
                              void* stream);

/* ---- measurement helpers --------------------------------------------------------------------------------------
 * FP32 FMA throughput probe: `iters` dependent-chain FMAs x 8 chains per thread; returns nothing, time it outside.
 * flops per launch = 2 * 8 * iters * blocks * threads. */
int rcbf_fp32_fma_probe(float* sink, int blocks, int threads, int iters, void* stream);
int rcbf_fp64_fma_probe(double* sink, int blocks, int threads, int iters, void* stream); /* same, float64 FMAs */
const char* rcbf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RCBF_B200_H */
