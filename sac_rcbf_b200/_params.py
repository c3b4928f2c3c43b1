"""ctypes mirrors of the plain-C parameter structs in include/rcbf_b200.h (== csrc/rcbf_core.cuh)."""
import ctypes as C

import numpy as np

UNI_HAZ = 5          # hazards of the specialised hot kernels (envs/unicycle_env.py:26)
MAX_HAZARDS = 12     # RCBF_MAX_HAZARDS: general per-instance kernels (rcbf_general.cu)
WS_WORDS = 32768  # RCBF_WS_WORDS: solver workspace (counters + fallback queue), 64-bit words


class UnicycleParams(C.Structure):
    _fields_ = [
        ("hazards", (C.c_float * 2) * UNI_HAZ),
        ("collision_radius_sq", C.c_float),
        ("gamma_b", C.c_float),
        ("l_p", C.c_float),
        ("sigma_scale", C.c_float),
        ("abs_sigma_map", C.c_int),
        ("u_min", C.c_float * 2),
        ("u_max", C.c_float * 2),
        ("p_diag", C.c_float * 3),
        ("solver_mode", C.c_int),
    ]


class GpPosterior(C.Structure):  # rcbf_gp_posterior
    _fields_ = [
        ("train_z", C.c_void_p),
        ("inv_x_scale", C.c_void_p),
        ("hyp", C.c_void_p),
        ("r_tiles", C.c_void_p),
        ("factor", C.c_void_p),
        ("proj_y", C.c_void_p),
        ("n_pad", C.c_int32),
        ("n_in", C.c_int32),
        ("dim_pad", C.c_int32),
        ("n_gp", C.c_int32),
        ("max_tiles", C.c_int32),
        ("tile_rows", C.c_int32),
        ("include_noise", C.c_int32),
        ("min_variance", C.c_double),
        ("ff_coef", C.c_void_p),
        ("ff_amax", C.c_void_p),
        ("ff_zmax", C.c_double),
        ("test_stride", C.c_int64),
    ]


class ReplayRing(C.Structure):  # rcbf_replay_ring
    _fields_ = [
        ("field", C.c_void_p * 7),
        ("row_stride", C.c_int64 * 7),
        ("capacity", C.c_int64),
        ("obs_dim", C.c_int32),
        ("action_dim", C.c_int32),
        ("elem_bytes", C.c_int32),
    ]


class CarsParams(C.Structure):
    _fields_ = [
        ("gamma_2", C.c_float),
        ("gamma_sq", C.c_float),
        ("kp", C.c_float),
        ("k_brake", C.c_float),
        ("collision_radius_sq", C.c_float),
        ("sigma_scale", C.c_float),
        ("u_min", C.c_float),
        ("u_max", C.c_float),
        ("p_diag", C.c_float * 2),
        ("slack_coeff", C.c_float),
        ("solver_mode", C.c_int),
    ]


class UnicycleEnvParams(C.Structure):
    _fields_ = [
        ("hazards", (C.c_double * 2) * UNI_HAZ),
        ("hazards_radius", C.c_double),
        ("dt", C.c_double),
        ("goal_x", C.c_double),
        ("goal_y", C.c_double),
        ("goal_size", C.c_double),
        ("reward_goal", C.c_double),
        ("init_x", C.c_double),
        ("init_y", C.c_double),
        ("init_theta", C.c_double),
        ("max_episode_steps", C.c_int),
        ("auto_reset", C.c_int),
    ]


class CarsEnvParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double),
        ("kp", C.c_double),
        ("k_brake", C.c_double),
        ("max_episode_steps", C.c_int),
        ("auto_reset", C.c_int),
    ]


DEFAULT_HAZARDS = np.array([[0., 0.], [-1., 1.], [-1., -1.], [1., -1.], [1., 1.]]) * 1.5  # envs/unicycle_env.py:26


FAR_HAZARD = 1.0e4  # padding hazard: h = gamma*(1/2 |d|^2)^3 ~ 1e24 > 0 always, so its row can never be active


def _check_hazards(hazards_locations):
    """(K,2) with 1 <= K <= 5 hazards -> (5,2): the kernels are specialised for the reference's 5 hazards
    (unicycle_env.py:26); fewer are padded with hazards 1e4 m away whose CBF rows are inert."""
    hz = np.asarray(DEFAULT_HAZARDS if hazards_locations is None else hazards_locations, np.float64)
    if hz.ndim != 2 or hz.shape[1] != 2 or not (1 <= hz.shape[0] <= UNI_HAZ):
        raise ValueError("the sm_100a kernels support 1..%d hazards of shape (K, 2) (reference: unicycle_env.py:26), "
                         "got %r" % (UNI_HAZ, hz.shape))
    if hz.shape[0] < UNI_HAZ:
        pad = np.full((UNI_HAZ - hz.shape[0], 2), FAR_HAZARD)
        pad[:, 1] += np.arange(pad.shape[0]) * 10.0
        hz = np.concatenate([hz, pad], 0)
    return hz


def unicycle_params(hazards_locations=None, hazards_radius=0.6, gamma_b=100.0, l_p=0.03, u_min=(-2.5, -2.5),
                    u_max=(2.5, 2.5), p_diag=(1.0, 1e-2, 1e5), sigma_scale=1.0, abs_sigma_map=True,
                    solver_mode=0):
    """Defaults = envs/unicycle_env.py:21-26 and rcbf_sac/diff_cbf_qp.py:12,207,265."""
    hz = _check_hazards(hazards_locations)
    p = UnicycleParams()
    for i in range(UNI_HAZ):
        p.hazards[i][0], p.hazards[i][1] = float(hz[i, 0]), float(hz[i, 1])
    p.collision_radius_sq = (1.2 * hazards_radius) ** 2  # squared in double, rounded once (diff_cbf_qp.py:207,246)
    p.gamma_b, p.l_p, p.sigma_scale, p.abs_sigma_map = gamma_b, l_p, sigma_scale, int(abs_sigma_map)
    for c in range(2):
        p.u_min[c], p.u_max[c] = float(u_min[c]), float(u_max[c])
    for j in range(3):
        p.p_diag[j] = float(p_diag[j])
    p.solver_mode = int(solver_mode)
    return p


def cars_params(gamma_b=100.0, kp=4.0, k_brake=20.0, u_min=-10.0, u_max=10.0, p_diag=(0.1, 10.0), sigma_scale=1.0,
                slack_coeff=200.0, solver_mode=0):
    """Defaults = envs/simulated_cars_env.py:18-26 and rcbf_sac/diff_cbf_qp.py:272,352,356."""
    p = CarsParams()
    p.gamma_2, p.gamma_sq = gamma_b + gamma_b, gamma_b * gamma_b  # python-double scalars of diff_cbf_qp.py:348
    p.kp, p.k_brake, p.collision_radius_sq, p.sigma_scale = kp, k_brake, 3.5 ** 2, sigma_scale
    p.u_min, p.u_max = float(u_min), float(u_max)
    p.p_diag[0], p.p_diag[1] = float(p_diag[0]), float(p_diag[1])
    p.slack_coeff = slack_coeff
    p.solver_mode = int(solver_mode)
    return p


def unicycle_env_params(hazards_locations=None, hazards_radius=0.6, dt=0.02, goal_pos=(2.5, 2.5), goal_size=0.3,
                        reward_goal=1.0, init_state=(-2.5, -2.5, 0.0), max_episode_steps=1000, auto_reset=False):
    """Defaults = envs/unicycle_env.py:24-33,137."""
    hz = _check_hazards(hazards_locations)
    e = UnicycleEnvParams()
    for i in range(UNI_HAZ):
        e.hazards[i][0], e.hazards[i][1] = float(hz[i, 0]), float(hz[i, 1])
    e.hazards_radius, e.dt, e.goal_x, e.goal_y = hazards_radius, dt, float(goal_pos[0]), float(goal_pos[1])
    e.goal_size, e.reward_goal = goal_size, reward_goal
    e.init_x, e.init_y, e.init_theta = (float(v) for v in init_state)
    e.max_episode_steps, e.auto_reset = int(max_episode_steps), int(auto_reset)
    return e


def cars_env_params(dt=0.02, kp=4.0, k_brake=20.0, max_episode_steps=300, auto_reset=False):
    """Defaults = envs/simulated_cars_env.py:21-26."""
    e = CarsEnvParams()
    e.dt, e.kp, e.k_brake = dt, kp, k_brake
    e.max_episode_steps, e.auto_reset = int(max_episode_steps), int(auto_reset)
    return e
