"""Exercise k_cars2's pending path (never taken by real SimulatedCars inputs: the 2 x 4 QP is enumerated in place): run
with a library built with -DRCBF_C2_FORCE_PENDING (scripts/mk_variant.sh) and with the default one; the outputs of two
fused steps and of the layer must be identical (the tail's exhaustive enumeration finds the same KKT point)."""
import os, sys, types
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import sac_rcbf_b200 as S
from sac_rcbf_b200 import workloads as O
B = 50000 + 7
stc, acc, muc, sgc, t = (torch.from_numpy(a).cuda() for a in O.synth_cars(B, seed=77))
env = S.SimulatedCarsEnv(num_envs=B)
lay = S.CBFQPLayer(env, types.SimpleNamespace(cuda=True), gamma_b=20, k_d=3.0, l_p=0.03)
env.state = stc; env._t.copy_(t)
res = {}
for k in range(2):
    us, obs, rew, done, info = env.safe_step(lay, acc, sgc, want_status=True)
    res.update({"us%d" % k: us, "obs%d" % k: obs, "rew%d" % k: rew, "done%d" % k: done, "cost%d" % k: info["cost"],
                "state%d" % k: env.state, "t%d" % k: env._t, "step%d" % k: env._step})
    res = {a: (b.clone() if torch.is_tensor(b) else b) for a, b in res.items()}
out, meta = lay._forward_meta(stc, acc, muc, sgc)
res["layer"], res["meta_active"] = out, meta & 0xffff
stats = lay.solver_stats()
np.savez(sys.argv[1], **{a: b.cpu().numpy() for a, b in res.items()})
print(sys.argv[1], "fallback (pending drained in-kernel):", stats["fallback"], "nan", stats["nan"], "uncertified", stats["uncertified"])
