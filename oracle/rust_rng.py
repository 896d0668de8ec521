"""The seeded random streams of the reference's path -- TEST INFRASTRUCTURE ONLY.

Pure-Python restatement of the third-party crates Cargo.lock pins for crates/gp, crates/moe and crates/doe
(absent from /root/reference): rand_xoshiro 0.6.0 (`Xoshiro256Plus`, `seed_from_u64` through SplitMix64),
rand 0.8.5 (`RngCore::next_u32` = upper half of `next_u64`; `Uniform::new(0., 1.)` = `UniformFloat::sample`,
52 mantissa bits; `SliceRandom::shuffle` / `gen_index` = `UniformInt::<u32>::sample_single`, widening
multiply with the rejection zone `(range << lz) - 1`; `Standard` for bool / Option<u64>).
Call sites: crates/doe/src/lhs.rs:235-258, 283-304 (`_classic_lhs`, `_maximin_lhs`), crates/gp/src/optimization.rs:59-62
(multistart seeds), crates/gp/src/sparse_algorithm.rs:457-462, 833-847 (`make_inducings`),
crates/moe/src/algorithm.rs:330 (`let seed = self.rng().gen()` -- an Option<u64>!).
Pinned on the reference's own fixture crates/doe/src/lhs.rs:332-347 (`test_classic_lhs`, seed 42) in
tests/test_host_rng.py: every printed digit is reproduced.
"""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1


class Xoshiro256Plus:
    """rand_xoshiro 0.6.0 src/xoshiro256plus.rs; seed_from_u64 = SplitMix64 stream (src/common.rs from_splitmix!)."""

    def __init__(self, seed: int):
        z, self.s = seed & _M64, []
        for _ in range(4):
            z = (z + 0x9E3779B97F4A7C15) & _M64
            x = z
            x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
            x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
            self.s.append(x ^ (x >> 31))

    def next_u64(self) -> int:
        s = self.s
        r = (s[0] + s[3]) & _M64
        t = (s[1] << 17) & _M64
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = ((s[3] << 45) | (s[3] >> 19)) & _M64
        return r

    def next_u32(self) -> int:
        return self.next_u64() >> 32

    def uniform01(self) -> float:
        """rand 0.8.5 distributions/uniform.rs UniformFloat::sample for Uniform::new(0., 1.) (scale 1, low 0)."""
        return (self.next_u64() >> 12) * (1.0 / 4503599627370496.0)

    def gen_bool_standard(self) -> bool:
        """rand 0.8.5 distributions/other.rs: Standard for bool = sign bit of next_u32."""
        return (self.next_u32() >> 31) == 1

    def gen_option_u64(self):
        """Standard for Option<T>: `if rng.gen::<bool>() { Some(rng.gen()) } else { None }`."""
        return self.next_u64() if self.gen_bool_standard() else None

    def gen_index(self, ubound: int) -> int:
        """rand 0.8.5 seq/mod.rs gen_index -> gen_range(0..ubound as u32) -> UniformInt::sample_single_inclusive."""
        assert 0 < ubound <= 0xFFFFFFFF
        zone = ((ubound << (32 - ubound.bit_length())) & 0xFFFFFFFF) - 1
        while True:
            m = self.next_u32() * ubound
            if (m & 0xFFFFFFFF) <= zone:
                return m >> 32

    def shuffle(self, seq: list) -> list:
        """SliceRandom::shuffle, in place."""
        for i in range(len(seq) - 1, 0, -1):
            j = self.gen_index(i + 1)
            seq[i], seq[j] = seq[j], seq[i]
        return seq


def lhs_classic_normalized(ns: int, nx: int, rng: Xoshiro256Plus) -> np.ndarray:
    """doe/src/lhs.rs:235-258: all ns * nx uniforms first (column-major fill), then one shuffle per column."""
    cut = np.array([i * (1.0 / ns) for i in range(ns + 1)])          # Array::linspace(0., 1., ns + 1)
    a, b = cut[:ns], cut[1:]
    rnd = np.array([[rng.uniform01() for _ in range(ns)] for _ in range(nx)]).T
    out = np.empty((ns, nx))
    for j in range(nx):
        col = list(rnd[:, j] * (b - a) + a)
        out[:, j] = rng.shuffle(col)
    return out


def lhs_maximin_normalized(ns: int, nx: int, rng: Xoshiro256Plus, max_iters: int = 5) -> np.ndarray:
    """doe/src/lhs.rs:283-304."""
    from scipy.spatial.distance import pdist
    best = lhs_classic_normalized(ns, nx, rng)
    dbest = pdist(best).min() if ns > 1 else np.inf
    for _ in range(max_iters - 1):
        cand = lhs_classic_normalized(ns, nx, rng)
        dm = pdist(cand).min() if ns > 1 else np.inf
        if dbest < dm:
            dbest, best = dm, cand
    return best


def lhs_sample(xlimits, ns: int, seed: int, kind: str = "classic") -> np.ndarray:
    """doe/src/lhs.rs:67-88 + traits.rs sample(): normalized * (upper - lower) + lower."""
    xl = np.asarray(xlimits, dtype=np.float64)
    rng = Xoshiro256Plus(seed)
    f = lhs_classic_normalized if kind == "classic" else lhs_maximin_normalized
    return f(ns, xl.shape[0], rng) * (xl[:, 1] - xl[:, 0]) + xl[:, 0]


def inducing_indices(n: int, n_inducing: int, seed: int) -> list:
    """gp/src/sparse_algorithm.rs:457-462, 833-847: shuffle 0..n with Xoshiro256Plus::seed_from_u64(seed), keep the head."""
    return Xoshiro256Plus(seed).shuffle(list(range(n)))[: min(n_inducing, n)]
