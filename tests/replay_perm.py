"""Host restatement (numpy) of the replay sampler's keyed permutation (csrc/rcbf_replay.cu: rp_perm) and of the key
schedule of DeviceReplayMemory.sample -- test infrastructure: the GPU test compares the kernel's indices with it bit for
bit, the CPU test checks the statistics of the construction."""
import numpy as np

_M = 0xFFFFFFFF
_M64 = (1 << 64) - 1


def _mix(x):
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & _M
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & _M
    return x ^ (x >> 16)


def _mixv(x):
    u = np.uint64
    x = x.astype(u)
    x ^= x >> u(16)
    x = (x * u(0x85EBCA6B)) & u(_M)
    x ^= x >> u(13)
    x = (x * u(0xC2B2AE35)) & u(_M)
    return x ^ (x >> u(16))


def half_bits(size):
    bits = 2
    while (1 << bits) < size:
        bits += 1
    return (bits + (bits & 1)) // 2


def perm(i, size, key):
    """rows drawn for output slots i (array) of a ring holding `size` rows under `key`."""
    u = np.uint64
    hb = half_bits(size)
    hm = (1 << hb) - 1
    k0, k1 = key & _M, key >> 32
    v = np.asarray(i).astype(u).copy()
    todo = np.ones(len(v), bool)
    while todo.any():
        vv = v[todo]
        l, r = (vv >> u(hb)) & u(hm), vv & u(hm)
        for rd in range(6):
            c = (k0 * (2 * rd + 1) + _mix((k1 + 0x9E3779B9 * (rd + 1)) & _M)) & _M
            f = _mixv((r + u(c)) & u(_M)) & u(hm)
            l, r = r, l ^ f
        v[todo] = (l << u(hb)) | r
        todo = v >= size
    o = (key * 0x9E3779B97F4A7C15) & _M64
    o ^= o >> 29
    o = (o * 0xBF58476D1CE4E5B9) & _M64
    o ^= o >> 32
    return (v.astype(np.int64) + (o % size)) % size
