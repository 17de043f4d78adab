"""Oracle (TEST INFRASTRUCTURE): the uniform-at-random sparse-depth sampler.

Reference lines followed: dataloaders/dense_to_sparse.py:29-48 (UniformSampling).
The random numbers come from numpy's legacy global MT19937 stream
(``np.random.uniform`` -> ``random_sample``: 53-bit doubles ``(a>>5, b>>6)``);
numpy (third-party, pinned here at 2.3) is that dependency.  ``mt19937_uniform``
restates that published generator in pure numpy integer arithmetic so the
bit-exactness claim does not rest on calling numpy's own RNG.
"""
import numpy as np


def uniform_mask(depth, num_samples, max_depth=np.inf, rng=None):
    """dense_to_sparse.py:29-48 using numpy's global legacy RNG (or ``rng``, a RandomState)."""
    rs = np.random if rng is None else rng
    keep = depth > 0
    if max_depth is not np.inf:
        keep = np.bitwise_and(keep, depth <= max_depth)
    n_keep = np.count_nonzero(keep)
    if n_keep == 0 or num_samples == 0:
        return keep
    if num_samples == -1:
        prob = rs.randint(5, 100, 1) / 100.0
    elif num_samples == -2:
        prob = rs.randint(5, 100, 1) / 1000.0
    else:
        prob = float(num_samples) / n_keep
    return np.bitwise_and(keep, rs.uniform(0, 1, depth.shape) < prob)


def mt19937_init(seed):
    """init_genrand (Matsumoto & Nishimura 2002), as numpy's legacy seeding for an int seed."""
    mt = np.zeros(624, dtype=np.uint64)
    mt[0] = seed & 0xFFFFFFFF
    for i in range(1, 624):
        mt[i] = (1812433253 * (int(mt[i - 1]) ^ (int(mt[i - 1]) >> 30)) + i) & 0xFFFFFFFF
    return mt.astype(np.uint32)


def _mt19937_twist(mt):
    mt = mt.astype(np.uint64)
    out = mt.copy()
    for k in range(624):
        y = (int(out[k]) & 0x80000000) | (int(out[(k + 1) % 624]) & 0x7FFFFFFF)
        v = int(out[(k + 397) % 624]) ^ (y >> 1)
        if y & 1:
            v ^= 0x9908B0DF
        out[k] = v
    return out.astype(np.uint32)


def _temper(y):
    y = y.astype(np.uint64)
    y ^= y >> np.uint64(11)
    y ^= (y << np.uint64(7)) & np.uint64(0x9D2C5680)
    y ^= (y << np.uint64(15)) & np.uint64(0xEFC60000)
    y ^= y >> np.uint64(18)
    return (y & np.uint64(0xFFFFFFFF)).astype(np.uint64)


def mt19937_uniform(seed, n):
    """First ``n`` doubles of ``np.random.seed(seed); np.random.uniform(0,1,n)``:
    each double = ((a>>5)*2^26 + (b>>6)) / 2^53 from two consecutive 32-bit outputs."""
    need = 2 * n
    mt = mt19937_init(seed)
    words = []
    while sum(len(w) for w in words) < need:
        mt = _mt19937_twist(mt)
        words.append(_temper(mt))
    w = np.concatenate(words)[:need]
    a, b = w[0::2] >> np.uint64(5), w[1::2] >> np.uint64(6)
    return (a.astype(np.float64) * 67108864.0 + b.astype(np.float64)) / 9007199254740992.0


def uniform_mask_mt(depth, num_samples, seed):
    """uniform_mask for num_samples > 0 and max_depth = inf, driven by mt19937_uniform."""
    keep = depth > 0
    n_keep = np.count_nonzero(keep)
    if n_keep == 0 or num_samples == 0:
        return keep
    prob = float(num_samples) / n_keep
    u = mt19937_uniform(seed, depth.size).reshape(depth.shape)
    return np.bitwise_and(keep, u < prob)
