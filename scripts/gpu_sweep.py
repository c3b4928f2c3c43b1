"""BASELINE config 5: SimulatedCars (and Unicycle) RCBF-QP sweep over the batch size, N = 2^10 .. 2^24 TOTAL instances
sharded over the ranks of this launch (strong scaling; run plain for 1 GPU or under torchrun for N GPUs).
Per size: microseconds per `get_safe_action` launch (raw C-ABI call on device tensors, max over ranks) and QP/s;
with --cpu also the oracle port of the reference (reference-order assembly + restated qpth f64) up to 2^14, which is
what the larger sizes are extrapolated from linearly (SURVEY 8d config 5 says to state this).

    python scripts/gpu_sweep.py [--cpu] > profiles/r01_config5_sweep.txt
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/gpu_sweep.py
"""
import os
import sys
import time
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sac_rcbf_b200 as S  # noqa: E402
from sac_rcbf_b200.sharding import shard_range  # noqa: E402
from oracle import rcbf_oracle as O  # noqa: E402  (synthetic inputs + the CPU baseline leg only)


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    ns = types.SimpleNamespace(cuda=True, device_num=dev.index)
    rows = []
    for mode, synth in (("SimulatedCars", O.synth_cars), ("Unicycle", O.synth_unicycle)):
        env = (S.SimulatedCarsEnv if mode == "SimulatedCars" else S.UnicycleEnv)(num_envs=1, device=dev)
        layer = S.CBFQPLayer(env, ns, gamma_b=20, k_d=3.0, l_p=0.03)
        layer.check_nan = False
        base = synth(1 << 20, seed=12345)[:4]
        for lg in range(10, 25):
            n_total = 1 << lg
            lo, hi = shard_range(n_total, rank, world)
            n = hi - lo
            idx = (torch.arange(lo, hi) % (1 << 20)).numpy()
            st, ac, mu, sg = (torch.from_numpy(a[idx]).to(dev) for a in base)
            for _ in range(5):
                layer._forward_raw(st, ac, mu, sg)
            reps = 200 if lg <= 16 else 30 if lg <= 20 else 8
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                layer._forward_raw(st, ac, mu, sg)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
            # Python-layer latency of the drop-in call (autograd Function, NaN check with its host sync)
            layer.check_nan = True
            t0 = time.perf_counter()
            for _ in range(20):
                layer.get_safe_action(st, ac, mu, sg)
            torch.cuda.synchronize(dev)
            ms_api = torch.tensor([(time.perf_counter() - t0) / 20 * 1e3], device=dev)
            layer.check_nan = False
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                dist.all_reduce(ms_api, op=dist.ReduceOp.MAX)
            rows.append((mode, lg, n_total, float(ms), float(ms_api)))
    if rank == 0:
        print("# config 5 sweep, %d GPU(s), total instances sharded contiguously by rank, max over ranks" % world)
        print("# %-14s %4s %10s %14s %12s %16s" % ("env", "lg2", "N", "us/launch", "QP/s", "us/get_safe_action"))
        for mode, lg, n_total, ms, ms_api in rows:
            print("  %-14s %4d %10d %14.1f %12.3e %16.1f" % (mode, lg, n_total, ms * 1e3, n_total / ms * 1e3, ms_api * 1e3))
        if "--cpu" in sys.argv:
            torch.set_num_threads(os.cpu_count() or 1)
            print("# CPU oracle port of the reference path (f32 assembly + restated qpth f64, B = 512 per solve), %d threads"
                  % torch.get_num_threads())
            tt = torch.from_numpy
            for mode, synth in (("SimulatedCars", O.synth_cars), ("Unicycle", O.synth_unicycle)):
                for lg in (10, 12, 14):
                    n = 1 << lg
                    st, ac, mu, sg = synth(n, seed=12345)[:4]
                    t0 = time.perf_counter()
                    for b in range(0, n, 512):
                        O.safe_action(mode, tt(st[b:b + 512]), tt(ac[b:b + 512]), tt(mu[b:b + 512]), tt(sg[b:b + 512]),
                                      solver="qpth", gamma_b=20.0)
                    dt = time.perf_counter() - t0
                    print("  %-14s %4d %10d %14.1f %12.3e   (linear beyond: %.1f us per instance)" % (
                        mode, lg, n, dt * 1e6, n / dt, dt * 1e6 / n))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
