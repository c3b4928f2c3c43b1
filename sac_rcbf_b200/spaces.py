"""Minimal gym.spaces.Box stand-in (gym is not installed in this image).  If gym / gymnasium is importable its Box is
used instead so the envs plug into real gym tooling.  Fields used by the reference: .shape .low .high .sample() .seed()
.contains()  (unicycle_env.py:21-23, sac_cbf.py:70, main.py:304,310, diff_cbf_qp.py:464)."""
import numpy as np

try:  # pragma: no cover - not present in the build image
    from gym.spaces import Box  # type: ignore
except Exception:  # noqa: BLE001
    try:  # pragma: no cover
        from gymnasium.spaces import Box  # type: ignore
    except Exception:  # noqa: BLE001

        class Box:
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.shape = tuple(shape) if shape is not None else tuple(np.shape(low))
                self.dtype = np.dtype(dtype)
                self.low = np.full(self.shape, low, dtype=self.dtype)
                self.high = np.full(self.shape, high, dtype=self.dtype)
                self._rng = np.random.RandomState()

            def seed(self, seed=None):
                self._rng = np.random.RandomState(seed)
                return [seed]

            def sample(self):
                return self._rng.uniform(self.low, self.high).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

            def __repr__(self):
                return "Box(%s, %s, %s)" % (self.low.min(), self.high.max(), self.shape)
