"""DeviceReplayMemory vs the reference's ReplayMemory (rcbf_sac/replay_memory.py:4-35).

The ring content after every batch of a push schedule (wraps, exact fill, a batch larger than the capacity, an empty
batch, a single push) comes from the REAL class: tests/golden/replay_memory.npz is written by oracle/make_golden.py from
the unmodified reference source, and where /root/reference is present the real class is also driven live, side by side.

CPU (`-m "not gpu"`): the ring cursor arithmetic (`ring_plan`, the only host-side logic of the class) against the golden
positions / lengths and against the live class.  GPU: the two kernels (rcbf_replay_push / rcbf_replay_sample) -- ring
content bit-exact against the golden rows, draws without replacement, every drawn row a stored transition."""
import numpy as np
import pytest
import torch

from oracle import make_golden, ref_loader
from sac_rcbf_b200.replay_memory import ring_plan, _splitmix64
from tests import replay_perm


def _apply_plan(pos, size, cap, n):
    skip, count, wpos, pos, size = ring_plan(pos, size, cap, n)
    assert 0 <= skip <= n and skip + count == max(n, 0) and count <= cap and 0 <= wpos < cap
    return (skip, count, wpos), pos, size


def test_ring_cursor_matches_reference_golden(golden):
    g = golden("replay_memory.npz")
    cap = int(g["capacity"])
    pos = size = 0
    for k, b in enumerate(make_golden.replay_inputs(cap)):
        _, pos, size = _apply_plan(pos, size, cap, b[0].shape[0])
        assert size == g["after_%d" % k].shape[0] and pos == int(g["position_%d" % k])
    _, pos, size = _apply_plan(pos, size, cap, 1)
    assert pos == int(g["position_push"]) and size == g["after_push"].shape[0]


def test_ring_plan_places_rows_like_successive_pushes():
    """Emulate n single pushes (:12-18) on a list of row ids and compare with where the plan puts the batch rows."""
    rng = np.random.default_rng(3)
    for cap in (1, 2, 7, 23):
        ring, pos, size = [None] * cap, 0, 0
        ref_ring, ref_pos = [], 0
        uid = 0
        for n in [0, 1, cap - 1, cap, cap + 1, 3 * cap + 2] + list(rng.integers(0, 2 * cap + 2, 12)):
            n = int(n)
            ids = list(range(uid, uid + n))
            uid += n
            for i in ids:                                   # the reference's push
                if len(ref_ring) < cap:
                    ref_ring.append(None)
                ref_ring[ref_pos] = i
                ref_pos = (ref_pos + 1) % cap
            (skip, count, wpos), pos, size = _apply_plan(pos, size, cap, n)
            for j in range(count):
                ring[(wpos + j) % cap] = ids[skip + j]
            assert pos == ref_pos and size == len(ref_ring) and ring[:size] == ref_ring


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference source not mounted")
def test_ring_cursor_matches_live_reference():
    ref = ref_loader.load_reference()
    cap = 23
    real = ref.ReplayMemory(cap, 0)
    pos = size = 0
    for b in make_golden.replay_inputs(cap, seed=4):
        real.batch_push(*b)
        _, pos, size = _apply_plan(pos, size, cap, b[0].shape[0])
        assert size == len(real) and pos == real.position
    with pytest.raises(ValueError):
        real.sample(cap + 1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_device_replay_memory_fails_loudly_without_cuda():
    import sac_rcbf_b200 as S
    with pytest.raises(S.RcbfLibraryError):
        S.DeviceReplayMemory(10, 0, 7, 2)


def test_keyed_permutation_statistics_host_restatement():
    """The construction the kernel uses (Feistel network + cycle walking + key-derived rotation), restated in numpy:
    a bijection for every key; rows and (output slot, row) pairings uniform (chi-square, 6-sigma bands)."""
    for size in (1, 2, 3, 5, 16, 17, 255, 1000, 4097):
        for key in (0, 1, 0xDEADBEEFCAFEF00D):
            assert sorted(replay_perm.perm(np.arange(size), size, key).tolist()) == list(range(size))
    size, batch, draws = 1000, 100, 2000
    cnt = np.zeros(size)
    s0 = _splitmix64(11)
    for d in range(draws):
        cnt += np.bincount(replay_perm.perm(np.arange(batch), size, _splitmix64(s0 ^ d)), minlength=size)
    exp = draws * batch / size
    assert abs(((cnt - exp) ** 2 / exp).sum() - (size - 1)) < 6 * np.sqrt(2 * (size - 1))
    size = 7
    pair = np.zeros((size, size))
    s0 = _splitmix64(5)
    for d in range(7000):
        pair[np.arange(size), replay_perm.perm(np.arange(size), size, _splitmix64(s0 ^ d))] += 1
    assert ((pair - 1000.0) ** 2 / 1000.0).sum() < 36 + 6 * np.sqrt(72)


# ------------------------------------------------------------------------------------------------------------ GPU

def _rows(mem):
    n = len(mem)
    cols = [mem.state[:n], mem.action[:n], mem.reward[:n, None], mem.next_state[:n], mem.mask[:n, None], mem.t[:n, None],
            mem.next_t[:n, None]]
    return torch.cat([c.double().cpu() for c in cols], 1).numpy()


def _cat(smp):
    return torch.cat([smp[0], smp[1], smp[2][:, None], smp[3], smp[4][:, None], smp[5][:, None], smp[6][:, None]],
                     1).double().cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_ring_content_matches_reference_golden_on_device(golden, dtype):
    from sac_rcbf_b200.replay_memory import DeviceReplayMemory
    g = golden("replay_memory.npz")
    cap = int(g["capacity"])
    cast = (lambda a: a) if dtype == torch.float64 else (lambda a: a.astype(np.float32).astype(np.float64))
    mem = DeviceReplayMemory(cap, seed=0, obs_dim=7, action_dim=2, dtype=dtype)
    assert mem.state.is_cuda
    for k, b in enumerate(make_golden.replay_inputs(cap)):
        mem.batch_push(*b)
        assert len(mem) == g["after_%d" % k].shape[0] and mem.position == int(g["position_%d" % k])
        np.testing.assert_array_equal(_rows(mem), cast(g["after_%d" % k]))
    mem.push(np.ones(7), np.ones(2), 1.0, np.ones(7), 1.0, t=0.5, next_t=0.52)
    assert mem.position == int(g["position_push"])
    np.testing.assert_array_equal(_rows(mem), cast(g["after_push"]))
    smp = mem.sample(16)
    assert [x.dim() for x in smp] == list(g["sample_shapes"]) and all(x.is_cuda for x in smp)
    # every drawn row is a stored transition, no row twice, and the index output names the rows that were gathered
    out = mem.sample(len(mem), return_indices=True)
    idx = out[7].cpu().numpy()
    assert sorted(idx.tolist()) == list(range(len(mem)))
    np.testing.assert_array_equal(_cat(out), _rows(mem)[idx])
    with pytest.raises(ValueError):
        mem.sample(cap + 1)
    # push without t / next_t (:26) leaves those ring fields alone
    t_before = mem.t.clone()
    mem.batch_push(np.zeros((3, 7)), np.zeros((3, 2)), np.zeros(3), np.zeros((3, 7)), np.ones(3))
    assert torch.equal(mem.t, t_before)


@pytest.mark.gpu
@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference source not mounted")
def test_device_ring_matches_live_reference():
    from sac_rcbf_b200.replay_memory import DeviceReplayMemory
    ref = ref_loader.load_reference()
    cap = 23
    real = ref.ReplayMemory(cap, 0)
    mem = DeviceReplayMemory(cap, seed=0, obs_dim=7, action_dim=2, dtype=torch.float64)
    for b in make_golden.replay_inputs(cap, seed=4):
        real.batch_push(*b)
        mem.batch_push(*b)
        want = np.stack([np.concatenate([np.ravel(x) for x in it]) for it in real.buffer]) if len(real) else np.zeros((0, 20))
        np.testing.assert_array_equal(_rows(mem), want)
    out_real, out_mine = real.sample(8), mem.sample(8)
    assert [np.asarray(x).shape for x in out_real] == [tuple(x.shape) for x in out_mine]


@pytest.mark.gpu
@pytest.mark.parametrize("size,batch", [(1, 1), (2, 2), (5, 3), (64, 64), (1000, 256), (4097, 4097), (1 << 20, 100000)])
def test_sample_is_without_replacement(size, batch):
    from sac_rcbf_b200.replay_memory import DeviceReplayMemory
    mem = DeviceReplayMemory(size + 3, seed=7, obs_dim=3, action_dim=1)
    x = torch.arange(size, dtype=torch.float32, device="cuda")
    mem.batch_push(x[:, None].expand(size, 3), x[:, None], x, x[:, None].expand(size, 3) + 0.5, torch.ones(size), x, x + 1)
    for _ in range(3):
        s, a, r, s2, m, t, nt, idx = mem.sample(batch, return_indices=True)
        assert s.shape == (batch, 3) and a.shape == (batch, 1) and r.shape == (batch,)
        ids = idx.cpu().numpy()
        key = _splitmix64(_splitmix64(mem._seed) ^ (mem._draws - 1))
        np.testing.assert_array_equal(ids, replay_perm.perm(np.arange(batch), size, key))   # bit-exact vs the host port
        assert ids.min() >= 0 and ids.max() < size and len(np.unique(ids)) == batch
        assert torch.equal(r, idx.float()) and torch.equal(s[:, 2], r) and torch.equal(s2[:, 0], r + 0.5)
        assert torch.equal(a[:, 0], r) and torch.equal(nt, t + 1) and bool((m == 1).all())
    a1 = mem.sample(batch, return_indices=True)[7]
    a2 = mem.sample(batch, return_indices=True)[7]
    if size > 64:
        assert not torch.equal(a1, a2)                      # a new key per draw


@pytest.mark.gpu
def test_sample_is_uniform_over_the_ring():
    """Marginal uniformity of the keyed permutation: over many draws every stored row is picked equally often
    (chi-square against the uniform expectation, 6-sigma band), and so is every (output slot, row) pairing in a small
    ring (the draw is a random ORDER too, like random.sample)."""
    from sac_rcbf_b200.replay_memory import DeviceReplayMemory
    size, batch, draws = 1000, 100, 4000
    mem = DeviceReplayMemory(size, seed=11, obs_dim=1, action_dim=1)
    x = torch.arange(size, dtype=torch.float32, device="cuda")
    mem.batch_push(x[:, None], x[:, None], x, x[:, None], x, x, x)
    counts = torch.zeros(size, dtype=torch.int64, device="cuda")
    for _ in range(draws):
        counts += torch.bincount(mem.sample(batch, return_indices=True)[7], minlength=size)
    c = counts.cpu().numpy().astype(np.float64)
    exp = draws * batch / size
    chi2 = ((c - exp) ** 2 / exp).sum()
    assert abs(chi2 - (size - 1)) < 6 * np.sqrt(2 * (size - 1)), chi2
    size = 7
    mem = DeviceReplayMemory(size, seed=5, obs_dim=1, action_dim=1)
    x = torch.arange(size, dtype=torch.float32, device="cuda")
    mem.batch_push(x[:, None], x[:, None], x, x[:, None], x, x, x)
    pair = np.zeros((size, size))
    for _ in range(7000):
        ids = mem.sample(size, return_indices=True)[7].cpu().numpy()
        pair[np.arange(size), ids] += 1
    chi2 = ((pair - 1000.0) ** 2 / 1000.0).sum()
    assert chi2 < 36 + 6 * np.sqrt(72), chi2               # (size-1)^2 degrees of freedom


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,od,ad", [(torch.float32, 7, 2), (torch.float64, 10, 1), (torch.float32, 70, 3)])
def test_separate_array_layout_through_the_c_abi(dtype, od, ad):
    """The C ABI also takes a ring of seven separate dense arrays (row_stride 0: word-granular kernels; the (70, 3)
    row-major ring is wider than the tiled kernels' 256 bytes and takes them too): same pushes, same key -> same draw."""
    import ctypes as C
    from sac_rcbf_b200 import _lib, _params as P
    from sac_rcbf_b200.replay_memory import DeviceReplayMemory
    lib = _lib.load()
    cap, n, eb = 301, 450, (4 if dtype == torch.float32 else 8)
    mem = DeviceReplayMemory(cap, seed=3, obs_dim=od, action_dim=ad, dtype=dtype)
    g = torch.Generator(device="cuda").manual_seed(1)
    r = lambda *s: torch.randn(s, generator=g, device="cuda", dtype=dtype)  # noqa: E731
    widths = (od, ad, 1, od, 1, 1, 1)
    ring = [torch.zeros(cap, w, device="cuda", dtype=dtype) for w in widths]
    desc = P.ReplayRing((C.c_void_p * 7)(*[x.data_ptr() for x in ring]), (C.c_int64 * 7)(), cap, od, ad, eb)
    pos = size = 0
    for nb in (n, 17, 1):
        src = [r(nb, w) for w in widths]
        mem.batch_push(src[0], src[1], src[2][:, 0], src[3], src[4][:, 0], src[5][:, 0], src[6][:, 0])
        skip, count, wpos, pos, size = ring_plan(pos, size, cap, nb)
        ptrs = (C.c_void_p * 7)(*[x[skip:].contiguous().data_ptr() for x in src])
        keep = [x[skip:].contiguous() for x in src]
        ptrs = (C.c_void_p * 7)(*[x.data_ptr() for x in keep])
        assert lib.rcbf_replay_push(C.byref(desc), wpos, C.byref(ptrs), count, None) == 0
        torch.cuda.synchronize()
    mine = [mem.state, mem.action, mem.reward[:, None], mem.next_state, mem.mask[:, None], mem.t[:, None], mem.next_t[:, None]]
    for a, b in zip(mine, ring):
        assert torch.equal(a, b)
    batch = 128
    out = mem.sample(batch, return_indices=True)
    key = _splitmix64(_splitmix64(mem._seed) ^ (mem._draws - 1))
    out2 = [torch.empty(batch, w, device="cuda", dtype=dtype) for w in widths]
    idx2 = torch.empty(batch, dtype=torch.int64, device="cuda")
    ptrs = (C.c_void_p * 7)(*[x.data_ptr() for x in out2])
    assert lib.rcbf_replay_sample(C.byref(desc), size, batch, key, C.byref(ptrs), idx2.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert torch.equal(out[7], idx2)
    for a, b in zip(out[:7], out2):
        assert torch.equal(a.reshape(batch, -1), b)
    # argument checks of the ABI
    assert lib.rcbf_replay_sample(C.byref(desc), size, size + 1, key, C.byref(ptrs), None, None) != 0
    assert lib.rcbf_replay_push(C.byref(desc), cap, C.byref(ptrs), 1, None) != 0
