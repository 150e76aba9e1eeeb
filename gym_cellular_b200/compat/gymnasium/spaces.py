"""Minimal `gymnasium.spaces` look-alikes: Discrete, MultiDiscrete, MultiBinary, Box, Tuple, Dict."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._np_random = None
        self._seed = seed

    @property
    def shape(self):
        return self._shape

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.default_rng(self._seed)
        return self._np_random

    def seed(self, seed=None):
        self._seed = seed
        self._np_random = np.random.default_rng(seed)
        return seed

    def sample(self, mask=None):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        assert int(n) > 0
        self.n = int(n)
        self.start = int(start)
        super().__init__((), np.int64, seed)

    def sample(self, mask=None):
        return self.start + int(self.np_random.integers(self.n))

    def contains(self, x):
        try:
            xi = int(x)
        except (TypeError, ValueError):
            return False
        return xi == x and self.start <= xi < self.start + self.n

    def __repr__(self):
        return f"Discrete({self.n})" if self.start == 0 else f"Discrete({self.n}, start={self.start})"

    def __eq__(self, other):
        return isinstance(other, Discrete) and (self.n, self.start) == (other.n, other.start)


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None, start=None):
        self.nvec = np.array(nvec, dtype=dtype, copy=True)
        self.start = np.zeros_like(self.nvec) if start is None else np.array(start, dtype=dtype)
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self, mask=None):
        return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype) + self.start

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.nvec.shape and bool(np.all(x >= self.start) and np.all(x - self.start < self.nvec))

    def __repr__(self):
        return f"MultiDiscrete({self.nvec})"


class MultiBinary(Space):
    def __init__(self, n, seed=None):
        if isinstance(n, (tuple, list, np.ndarray)):
            self.n = tuple(int(i) for i in n)
            shape = self.n
        else:
            self.n = int(n)
            shape = (self.n,)
        super().__init__(shape, np.int8, seed)

    def sample(self, mask=None):
        return self.np_random.integers(0, 2, size=self.shape, dtype=self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all((x == 0) | (x == 1)))

    def __repr__(self):
        return f"MultiBinary({self.n})"


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        self.low = np.broadcast_to(np.asarray(low), shape).astype(dtype)
        self.high = np.broadcast_to(np.asarray(high), shape).astype(dtype)
        super().__init__(shape, dtype, seed)

    def sample(self, mask=None):
        if np.issubdtype(self.dtype, np.integer):
            return self.np_random.integers(self.low, self.high + 1, size=self.shape).astype(self.dtype)
        return self.np_random.uniform(self.low, self.high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.tolist()}, {self.high.tolist()}, {self.shape}, {self.dtype})"


class Tuple(Space):
    def __init__(self, spaces, seed=None):
        self.spaces = tuple(spaces)
        super().__init__(None, None, seed)

    def sample(self, mask=None):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        if isinstance(x, (list, np.ndarray)):
            x = tuple(x)
        return (isinstance(x, tuple) and len(x) == len(self.spaces)
                and all(s.contains(p) for s, p in zip(self.spaces, x)))

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)

    def __iter__(self):
        return iter(self.spaces)

    def __repr__(self):
        return "Tuple(" + ", ".join(repr(s) for s in self.spaces) + ")"

    def __eq__(self, other):
        return isinstance(other, Tuple) and self.spaces == other.spaces


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **kw):
        d = dict(spaces or {})
        d.update(kw)
        self.spaces = d
        super().__init__(None, None, seed)

    def sample(self, mask=None):
        return {k: s.sample() for k, s in self.spaces.items()}

    def contains(self, x):
        return (isinstance(x, dict) and x.keys() == self.spaces.keys()
                and all(self.spaces[k].contains(v) for k, v in x.items()))

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k!r}: {s!r}" for k, s in self.spaces.items()) + ")"


__all__ = ["Space", "Discrete", "MultiDiscrete", "MultiBinary", "Box", "Tuple", "Dict"]
