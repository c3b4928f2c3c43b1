"""DeviceReplayMemory -- the replay ring of rcbf_sac/replay_memory.py:4-35 resident in HBM.

Same tuple layout (state, action, reward, next_state, mask, t, next_t) and method names as the reference's ReplayMemory;
the storage is ONE preallocated row-major device matrix (a transition = one padded row; `state`, `action`, ... are
strided views of it) and the two operations are one kernel launch each (csrc/rcbf_replay.cu through the C ABI):

  batch_push  rcbf_replay_push: one scatter of the whole batch instead of the reference's per-item Python loop
              (replay_memory.py:20-26); only the ring cursor (`position`, `len`) is host arithmetic (`ring_plan`).
  sample      rcbf_replay_sample: index draw WITHOUT replacement (what random.sample does, :30) fused with the gather of
              all seven fields (:31); the draw is a keyed bijection of [0, len), so it costs O(batch), not O(len).

Device-only: there is no host fallback (use the reference's own ReplayMemory for a host-side buffer)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import _params as P

_M64 = (1 << 64) - 1


def ring_plan(position, size, capacity, n):
    """Cursor arithmetic of n successive push() calls (replay_memory.py:12-18) on a ring at (position, size):
    -> (skip, count, write_position, new_position, new_size): rows [skip, skip + count) of the batch are written to ring
    rows (write_position + i) % capacity.  Of a batch longer than the capacity only the newest `capacity` rows survive."""
    position, size, capacity, n = int(position), int(size), int(capacity), int(n)
    if n <= 0:
        return 0, 0, position, position, size
    skip = max(0, n - capacity)
    count = n - skip
    write_position = (position + skip) % capacity
    return skip, count, write_position, (position + n) % capacity, min(capacity, size + n)


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


class DeviceReplayMemory:

    def __init__(self, capacity, seed, obs_dim, action_dim, device=None, dtype=torch.float32):
        _lib.require_cuda()
        self._lib = _lib.load()
        self.capacity = int(capacity)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.RcbfLibraryError("DeviceReplayMemory lives on a CUDA device (no host fallback); "
                                        "got device=%r" % (device,))
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be float32 or float64")
        self.dtype = dtype
        self.obs_dim, self.action_dim = int(obs_dim), int(action_dim)
        # ROW-MAJOR ring: the seven fields of a transition adjacent in one (capacity, stride) matrix, stride padded to a
        # 32-byte multiple -- a drawn transition is three sectors of HBM instead of nine (include/rcbf_b200.h).  The
        # reference-named attributes are strided views of that matrix.
        eb = 4 if dtype == torch.float32 else 8
        widths = (self.obs_dim, self.action_dim, 1, self.obs_dim, 1, 1, 1)
        self._widths = widths
        row = sum(widths)
        # (a 128-byte stride -- one L2 line per row instead of 1.5 on average -- measured slower: 237 against 214 us per
        #  4 Mi-row draw, 225 against 179 us per push; fewer rows fit a tile and the push writes a third more)
        stride = (row * eb + 31) // 32 * 32 // eb
        self._rows = torch.zeros((self.capacity, stride), dtype=dtype, device=self.device)
        offs = [sum(widths[:f]) for f in range(7)]
        cols = [self._rows[:, o:o + w] for o, w in zip(offs, widths)]
        self.state, self.action, self.next_state = cols[0], cols[1], cols[3]
        self.reward, self.mask, self.t, self.next_t = cols[2][:, 0], cols[4][:, 0], cols[5][:, 0], cols[6][:, 0]
        self.position = 0
        self.size = 0
        self._seed = int(seed) & _M64
        self._draws = 0
        base = self._rows.data_ptr()
        self._ring = P.ReplayRing((C.c_void_p * 7)(*[base + o * eb for o in offs]), (C.c_int64 * 7)(*([stride] * 7)),
                                  self.capacity, self.obs_dim, self.action_dim, eb)
        self._eb = eb

    def _t(self, x, shape):
        x = x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))
        return x.to(self.device, self.dtype).reshape(shape).contiguous()

    def push(self, state, action, reward, next_state, mask, t=None, next_t=None):
        self.batch_push(self._t(state, (1, -1)), self._t(action, (1, -1)), self._t(reward, (1,)),
                        self._t(next_state, (1, -1)), self._t(mask, (1,)),
                        None if t is None else self._t(t, (1,)), None if next_t is None else self._t(next_t, (1,)))

    def batch_push(self, state_batch, action_batch, reward_batch, next_state_batch, mask_batch, t_batch=None,
                   next_t_batch=None):
        n = int(state_batch.shape[0])
        skip, count, wpos, new_pos, new_size = ring_plan(self.position, self.size, self.capacity, n)
        if count == 0:
            return
        with_t = t_batch is not None and next_t_batch is not None        # replay_memory.py:23
        src = [self._t(state_batch, (n, self.obs_dim)), self._t(action_batch, (n, self.action_dim)),
               self._t(reward_batch, (n,)), self._t(next_state_batch, (n, self.obs_dim)), self._t(mask_batch, (n,)),
               self._t(t_batch, (n,)) if with_t else None, self._t(next_t_batch, (n,)) if with_t else None]
        src = [None if s is None else s[skip:] for s in src]             # (row slices of contiguous arrays stay contiguous)
        ptrs = (C.c_void_p * 7)(*[None if s is None else s.data_ptr() for s in src])
        _lib.check(self._lib.rcbf_replay_push(C.byref(self._ring), wpos, C.byref(ptrs), count,
                                              _lib.stream_ptr(self.device)), "rcbf_replay_push")
        self.position, self.size = new_pos, new_size

    def sample(self, batch_size, return_indices=False):
        """(state, action, reward, next_state, mask, t, next_t) device tensors, drawn without replacement."""
        batch_size = int(batch_size)
        if batch_size > self.size or batch_size < 0:
            raise ValueError("Sample larger than population or is negative")    # what random.sample raises
        # one allocation for the seven outputs (each a contiguous (batch, width) block of it)
        b, ws = batch_size, self._widths
        flat = torch.empty(b * sum(ws), dtype=self.dtype, device=self.device)
        out = list(flat.split_with_sizes([b * w for w in ws]))
        for f in (0, 1, 3):
            out[f] = out[f].view(b, ws[f])
        out = tuple(out)
        idx = torch.empty(batch_size, dtype=torch.int64, device=self.device) if return_indices else None
        if batch_size:
            key = _splitmix64(_splitmix64(self._seed) ^ self._draws)
            self._draws += 1
            p0, eb, b = flat.data_ptr(), self._eb, batch_size
            offs = [0]
            for w in self._widths[:-1]:
                offs.append(offs[-1] + b * w * eb)
            ptrs = (C.c_void_p * 7)(*[p0 + x for x in offs])
            _lib.check(self._lib.rcbf_replay_sample(C.byref(self._ring), self.size, batch_size, key, C.byref(ptrs),
                                                    _lib.ptr(idx), _lib.stream_ptr(self.device)), "rcbf_replay_sample")
        return out + (idx,) if return_indices else out

    def __len__(self):
        return self.size
