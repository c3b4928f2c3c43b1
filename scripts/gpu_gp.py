"""GP posterior kernel timing (run on the GPU box): reference-scale bank (3000 training points, main.py:247) and a
full-rank stress case; prints points/s and the float64-FMA share."""
import sys, time, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sac_rcbf_b200 import _lib
from sac_rcbf_b200.gp_model import DisturbanceGPBank


def time_fn(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def fp64_peak():
    lib = _lib.load()
    sink = torch.zeros(1, dtype=torch.float64, device="cuda")
    blocks, threads, iters = 148 * 8, 256, 4096
    s = _lib.stream_ptr(torch.device("cuda"))
    ms = time_fn(lambda: lib.rcbf_fp64_fma_probe(_lib.ptr(sink), blocks, threads, iters, s))
    return 2 * 8 * iters * blocks * threads / ms / 1e9


def main():
    rng = np.random.default_rng(0)
    peak = fp64_peak()
    print("fp64 FMA probe: %.2f TFLOP/s" % peak)
    for name, n, d, n_gp, hyp, B in (("unicycle_ref_scale", 3000, 3, 3, None, 1 << 22),
                                      ("cars_ref_scale", 3000, 10, 10, None, 1 << 20),
                                      ("unicycle_b512", 3000, 3, 3, None, 512),
                                      ("full_rank_n1024", 1024, 3, 1, (0.5, 1.0, 0.01), 1 << 14)):
        x = rng.uniform(-3, 3, (n, d))
        y = 0.1 * np.sin(x[:, :1]) + 0.05 * rng.standard_normal((n, n_gp)) - 0.1
        xs, ys = x.std(0), y.std(0)
        bank = DisturbanceGPBank(x / (xs + 1e-8), y / (ys + 1e-8), [0.2] * n_gp, x_scale=xs, y_scale=ys + 1e-8)
        t0 = time.time()
        if hyp is None:
            bank.train(70)
        else:
            bank.set_hyperparameters(lengthscale=[hyp[0]] * n_gp, outputscale=[hyp[1]] * n_gp, noise=[hyp[2]] * n_gp)
        torch.cuda.synchronize(); t1 = time.time()
        bank.build_posterior(); torch.cuda.synchronize(); t2 = time.time()
        test = torch.as_tensor(rng.uniform(-3, 3, (B, d)), dtype=torch.float32).cuda()
        post = bank._post[0]
        rows = sum(int(t) for t in bank._post[1][3].cpu()) * post.tile_rows
        print("%-20s n=%d d=%d gps=%d ranks=%s tile_rows=%d  fit %.2fs  factor %.2fs" % (
            name, n, d, n_gp, bank.ranks[:4], post.tile_rows, t1 - t0, t2 - t1))
        if bank.far_field_active:
            ms = time_fn(lambda: bank.predict(test))
            nc = 3 + 2 * post.dim_pad + post.dim_pad * (post.dim_pad + 1) // 2
            flop = 2.0 * B * n_gp * post.tile_rows * (nc + 4)
            print("    far-field polynomial   B=%d: %.3f ms = %.3e points/s  (%.2f TFLOP/s f64 = %.2f of probe)" % (
                B, ms, B / ms * 1e3, flop / ms / 1e9, flop / ms / 1e9 / peak))
            bank.far_field = False
            bank.build_posterior()
            post = bank._post[0]
        Bx = min(B, 1 << 18)
        sub = test[:Bx]
        ms = time_fn(lambda: bank.predict(sub), reps=3)
        # float64 flop model of the exact kernels: per (test, train) pair and GP: distance 3*dim_pad, exp ~25
        # FMA-equivalents, scale 2; plus 2 per factor row
        flop = Bx * post.n_pad * (n_gp * (3 * post.dim_pad + 2 + 25) + 2 * rows)
        print("    exact %-16s B=%d: %.3f ms = %.3e points/s  (%.2f TFLOP/s f64-equivalent = %.2f of probe)" % (
            "lowrank" if post.tile_rows <= 16 else "tiled", Bx, ms, Bx / ms * 1e3, flop / ms / 1e9, flop / ms / 1e9 / peak))


if __name__ == "__main__":
    main()
