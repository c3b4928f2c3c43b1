"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck), one tool per gpurun call."""
import sys, types
sys.path.insert(0, '.')
import numpy as np, torch
import sac_rcbf_b200 as S
from oracle import rcbf_oracle as O
args = types.SimpleNamespace(cuda=True, gp_model_size=2000, l_p=0.03)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
B = 4099
st, ac, mu, sg = O.synth_unicycle(B, seed=1, hazard_frac=0.7)
env = S.UnicycleEnv(num_envs=B, auto_reset=True); layer = S.CBFQPLayer(env, args, gamma_b=20, k_d=3.0, l_p=0.03)
for solver in ("presolve", "pdipm"):
    layer.solver = solver
    a = dev(ac).requires_grad_(True)
    out = layer.get_safe_action(dev(st), a, dev(mu), dev(sg)); out.sum().backward()
    out2 = layer.get_safe_action(dev(st)[1:778], dev(ac)[1:778], dev(mu)[1:778], dev(sg)[1:778])   # unaligned -> plain path
    env.state = dev(st)
    for _ in range(3):
        env.safe_step(layer, dev(ac), dev(mu), dev(sg))
env.step(dev(ac)); env.reset()
stc, acc, muc, sgc, t = O.synth_cars(B, seed=1)
envc = S.SimulatedCarsEnv(num_envs=B); layc = S.CBFQPLayer(envc, args, gamma_b=20, k_d=3.0, l_p=0.03)
for solver in ("presolve", "pdipm"):
    layc.solver = solver
    a = dev(acc).requires_grad_(True)
    layc.get_safe_action(dev(stc), a, dev(muc), dev(sgc)).sum().backward()
    envc.state = dev(stc); envc._t.copy_(dev(t))
    for _ in range(3):                    # presolve: k_cars2 on the 128 full tiles + k_safe on the ragged 3 instances
        envc.safe_step(layc, dev(acc), dev(sgc))
envc.step(dev(acc))
P, q, G, h = layer.get_cbf_qp_constraints(dev(st), dev(ac), dev(mu), dev(sg))
layer.solve_qp(P, q, G.clone(), h)
dm = S.DynamicsModel(env, args); dm.predict_next_state(st.astype(np.float64), ac.astype(np.float64))
S.rollout_transition(env, dm, O.unicycle_obs(st.astype(np.float64)), ac.astype(np.float64), np.zeros(B), np.zeros((B, 3)))
pin = lambda x: torch.from_numpy(x).pin_memory()
env.safe_step_host(layer, pin(ac), pin(mu), pin(sg), chunks=3)
mem = S.DeviceReplayMemory(1000, seed=0, obs_dim=7, action_dim=2)            # row-major ring: tiled kernels
for nb in (25, 700, 1300):
    mem.batch_push(*[dev(x[:nb]) for x in (O.unicycle_obs(st.astype(np.float64)).astype(np.float32), ac,
                                             st[:, 0], O.unicycle_obs(st.astype(np.float64)).astype(np.float32),
                                             st[:, 1], st[:, 2], st[:, 2])])
mem.sample(256); mem.sample(1000)
mem.batch_push(dev(np.zeros((5, 7), np.float32)), dev(np.zeros((5, 2), np.float32)), dev(np.zeros(5, np.float32)),
               dev(np.zeros((5, 7), np.float32)), dev(np.ones(5, np.float32)))    # no t / next_t: word-granular kernel
torch.cuda.synchronize()
print("sanitize workload done")
