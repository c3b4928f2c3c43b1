"""TEST HELPER: numpy restatement of what csrc/rcbf_gp.cu computes from a packed `rcbf_gp_posterior` (checks the host-side
packing / rank truncation on a box without a GPU)."""
import numpy as np


def eval_packed(bank, test_x):
    post, keep = bank._post
    train_z, inv_x, hyp, r_tiles, factor, proj_y = [k.cpu().numpy() for k in keep[:6]]
    test_x = np.asarray(test_x, np.float64)
    z = np.zeros((len(test_x), post.dim_pad))
    z[:, :post.n_in] = test_x
    z *= inv_x
    mean = np.zeros((len(test_x), post.n_gp))
    std = np.zeros_like(mean)
    for g in range(post.n_gp):
        d2 = ((train_z[:, None, :] - z[None]) ** 2).sum(-1)
        k = hyp[g, 1] * np.exp(-d2 * hyp[g, 0])
        q = 0.0
        m = 0.0
        for t in range(r_tiles[g]):
            w = factor[g, t].T @ k
            q = q + (w * w).sum(0)
            m = m + (w * proj_y[g, t * post.tile_rows:(t + 1) * post.tile_rows, None]).sum(0)
        var = np.maximum(hyp[g, 1] - q + (hyp[g, 2] if post.include_noise else 0.0), post.min_variance)
        mean[:, g] = m * hyp[g, 3]
        std[:, g] = np.sqrt(var) * hyp[g, 3]
    return mean, std


def eval_far_field(bank, test_x):
    """numpy restatement of k_gp_farfield's polynomial branch; also returns the per-point validity mask."""
    post, keep = bank._post
    train_z, inv_x, hyp, r_tiles, factor, proj_y, coef, amax = [k.cpu().numpy() for k in keep]
    dp, rt = post.dim_pad, post.tile_rows
    test_x = np.asarray(test_x, np.float64)
    z = np.zeros((len(test_x), dp))
    z[:, :post.n_in] = test_x
    z *= inv_x
    s = (z * z).sum(1)
    iu = np.triu_indices(dp)
    mean = np.zeros((len(test_x), post.n_gp))
    std = np.zeros_like(mean)
    ok = np.zeros(mean.shape, bool)
    for g in range(post.n_gp):
        c = coef[g]
        w = c[:, 0:1] + s[None] * (c[:, 1:2] + s[None] * c[:, 2:3]) + c[:, 3:3 + dp] @ z.T \
            + (c[:, 3 + dp:3 + 2 * dp] @ z.T) * s[None] + c[:, 3 + 2 * dp:] @ (z[:, iu[0]] * z[:, iu[1]]).T
        q = (w * w).sum(0)
        m = (w * proj_y[g, :rt, None]).sum(0)
        var = np.maximum(hyp[g, 1] - q + (hyp[g, 2] if post.include_noise else 0.0), post.min_variance)
        mean[:, g] = m * hyp[g, 3]
        std[:, g] = np.sqrt(var) * hyp[g, 3]
        ok[:, g] = (np.sqrt(s) + post.ff_zmax) ** 2 * hyp[g, 0] <= amax[g]
    return mean, std, ok
