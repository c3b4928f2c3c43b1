"""DeviceReplayMemory -- ring buffer with the tuple layout and method names of rcbf_sac/replay_memory.py, stored as
preallocated device tensors.  `batch_push` is one vectorised scatter instead of the reference's
per-item Python loop (replay_memory.py:20-26); `sample` draws without replacement on the device like
`random.sample` does on the host (replay_memory.py:30)."""
import numpy as np
import torch


class DeviceReplayMemory:

    def __init__(self, capacity, seed, obs_dim, action_dim, device=None, dtype=torch.float32):
        self.capacity = int(capacity)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = dtype
        z = lambda *s: torch.zeros(s, dtype=dtype, device=self.device)  # noqa: E731
        self.state, self.next_state = z(self.capacity, obs_dim), z(self.capacity, obs_dim)
        self.action = z(self.capacity, action_dim)
        self.reward, self.mask, self.t, self.next_t = z(self.capacity), z(self.capacity), z(self.capacity), z(self.capacity)
        self.position = 0
        self.size = 0
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(int(seed))

    def _t(self, x, shape):
        return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(self.device, self.dtype).reshape(shape)

    def push(self, state, action, reward, next_state, mask, t=None, next_t=None):
        self.batch_push(self._t(state, (1, -1)), self._t(action, (1, -1)), self._t(reward, (1,)),
                        self._t(next_state, (1, -1)), self._t(mask, (1,)),
                        None if t is None else self._t(t, (1,)), None if next_t is None else self._t(next_t, (1,)))

    def batch_push(self, state_batch, action_batch, reward_batch, next_state_batch, mask_batch, t_batch=None,
                   next_t_batch=None):
        n = int(state_batch.shape[0])
        if n == 0:
            return
        if n > self.capacity:  # only the newest `capacity` items survive, exactly like pushing one by one
            sl = slice(n - self.capacity, n)
            self.position = (self.position + n - self.capacity) % self.capacity
            self.size = self.capacity
            return self.batch_push(state_batch[sl], action_batch[sl], reward_batch[sl], next_state_batch[sl],
                                   mask_batch[sl], None if t_batch is None else t_batch[sl],
                                   None if next_t_batch is None else next_t_batch[sl])
        idx = (self.position + torch.arange(n, device=self.device)) % self.capacity
        self.state[idx] = self._t(state_batch, (n, -1))
        self.action[idx] = self._t(action_batch, (n, -1))
        self.reward[idx] = self._t(reward_batch, (n,))
        self.next_state[idx] = self._t(next_state_batch, (n, -1))
        self.mask[idx] = self._t(mask_batch, (n,))
        if t_batch is not None and next_t_batch is not None:
            self.t[idx] = self._t(t_batch, (n,))
            self.next_t[idx] = self._t(next_t_batch, (n,))
        self.position = (self.position + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def sample(self, batch_size):
        """(state, action, reward, next_state, mask, t, next_t) device tensors, drawn without replacement."""
        if batch_size > self.size:
            raise ValueError("Sample larger than population or is negative")    # what random.sample raises
        idx = torch.randperm(self.size, generator=self._gen, device=self.device)[:batch_size]
        return (self.state[idx], self.action[idx], self.reward[idx], self.next_state[idx], self.mask[idx],
                self.t[idx], self.next_t[idx])

    def __len__(self):
        return self.size
