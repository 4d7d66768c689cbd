"""ORACLE (test infrastructure, NOT product code) -- NumPy restatement of the JAX PRNG.

The reference (lasgroup/Model-based-policy-optimizers) draws every random number
through ``jax.random`` (call sites: mbpo/optimizers/trajectory_optimizers/
icem_optimizer.py:123,155,174-180,246 and mbpo/utils/general_utils.py:189-191).
JAX/jaxlib are a third-party dependency that is NOT under /root/reference and is
NOT installable in this environment (setup.py:7,9 pin only ``jax>=0.4.13``), so
this file restates JAX's published algorithm:

  * threefry2x32 (Random123, 20 rounds)            jax/_src/prng.py  threefry2x32_p
  * PRNGKey(seed)                                   jax/_src/prng.py  threefry_seed
  * split / random_bits, legacy and partitionable   jax/_src/prng.py  threefry_split,
                                                    threefry_random_bits
  * uniform / normal                                jax/_src/random.py _uniform, _normal_real
  * erf_inv (f32)                                   xla/client/lib/math.cc ErfInv32 (Giles)

PARITY PINNING: pinned against external known-answer vectors only (Random123
threefry2x32-20 KATs, the values printed in the JAX documentation for
PRNGKey(0)/PRNGKey(42)); see tests/test_oracle_prng.py and tests/golden/prng_kats.json.
The reference itself holds no golden vectors (tests/test_icemopt.py:37-38 is a
threshold), so float parity against the real JAX/XLA run is UNPINNED.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may
import this module.
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
_ROT_A = (13, 15, 26, 6)
_ROT_B = (17, 29, 16, 24)


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds.  All args uint32 scalars/arrays (broadcast)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=U32)
        k1 = np.asarray(k1, dtype=U32)
        x0 = np.asarray(x0, dtype=U32).copy()
        x1 = np.asarray(x1, dtype=U32).copy()
        ks = (k0, k1, k0 ^ k1 ^ U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for g in range(5):
            for r in (_ROT_A if g % 2 == 0 else _ROT_B):
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + U32(g + 1)
    return x0, x1


def PRNGKey(seed: int, enable_x64: bool = False) -> np.ndarray:
    """jax.random.PRNGKey (jax/_src/prng.py random_seed -> threefry_seed).  A Python int goes through
    ``jnp.asarray(np.int64(seed))``: with ``jax_enable_x64`` off (the reference never turns it on) that is an
    int32 truncation, and ``shift_right_logical(seed, 32)`` of a 32-bit value is 0 -- so the key is
    ``[0, seed & 0xffffffff]`` whatever the seed (PRNGKey(-1) = [0, 0xffffffff], PRNGKey((7 << 32) | 9) = [0, 9]).
    With x64 on, the high word is ``(seed >> 32) & 0xffffffff`` of the int64."""
    seed = int(seed)
    hi = (seed >> 32) & 0xFFFFFFFF if enable_x64 else 0
    return np.array([hi, seed & 0xFFFFFFFF], dtype=U32)


def _threefry_2x32_counts(key, counts):
    """jax/_src/prng.py threefry_2x32(keypair, count): odd sizes are zero padded,
    the flat counter array is cut in halves (x0 = first half, x1 = second half)."""
    counts = np.asarray(counts, dtype=U32).ravel()
    odd = counts.size % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, dtype=U32)])
    half = counts.size // 2
    y0, y1 = threefry2x32(key[0], key[1], counts[:half], counts[half:])
    out = np.concatenate([y0, y1])
    return out[:-1] if odd else out


def split(key, num: int = 2, partitionable: bool = False) -> np.ndarray:
    """jax.random.split(key, num) -> uint32[num, 2]."""
    key = np.asarray(key, dtype=U32)
    if partitionable:
        y0, y1 = threefry2x32(key[0], key[1], np.zeros(num, dtype=U32), np.arange(num, dtype=U32))
        return np.stack([y0, y1], axis=-1)
    return _threefry_2x32_counts(key, np.arange(2 * num, dtype=U32)).reshape(num, 2)


def random_bits(key, n: int, partitionable: bool = False) -> np.ndarray:
    """32-bit random_bits(key, shape=(n,)) -> uint32[n]."""
    key = np.asarray(key, dtype=U32)
    if partitionable:
        y0, y1 = threefry2x32(key[0], key[1], np.zeros(n, dtype=U32), np.arange(n, dtype=U32))
        return y0 ^ y1
    return _threefry_2x32_counts(key, np.arange(n, dtype=U32))


def bits_to_uniform(bits, lo, hi) -> np.ndarray:
    """jax/_src/random.py _uniform for float32: mantissa trick, then affine + max."""
    bits = np.asarray(bits, dtype=U32)
    lo = np.float32(lo)
    hi = np.float32(hi)
    f = ((bits >> U32(9)) | U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    return np.maximum(lo, f * np.float32(hi - lo) + lo)


def uniform(key, n: int, lo=0.0, hi=1.0, partitionable: bool = False) -> np.ndarray:
    return bits_to_uniform(random_bits(key, n, partitionable), lo, hi)


_ERFINV_SMALL = np.array(
    [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
     -0.00125372503, -0.00417768164, 0.246640727, 1.50140941], dtype=np.float32)
_ERFINV_LARGE = np.array(
    [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
     -0.0076224613, 0.00943887047, 1.00167406, 2.83297682], dtype=np.float32)


def erf_inv_f32(x) -> np.ndarray:
    """XLA ErfInv32: w = -log1p(-x*x); two 9-term Horner branches; result p*x."""
    x = np.asarray(x, dtype=np.float32)
    w = -np.log1p(-(x * x))
    small = w < np.float32(5.0)
    ws = w - np.float32(2.5)
    with np.errstate(invalid="ignore"):
        wl = np.sqrt(w) - np.float32(3.0)
    ww = np.where(small, ws, wl).astype(np.float32)
    p = np.where(small, _ERFINV_SMALL[0], _ERFINV_LARGE[0]).astype(np.float32)
    for i in range(1, 9):
        c = np.where(small, _ERFINV_SMALL[i], _ERFINV_LARGE[i]).astype(np.float32)
        p = c + p * ww
    out = p * x
    return np.where(np.abs(x) == np.float32(1.0), np.float32(np.inf) * x, out).astype(np.float32)


_NORMAL_LO = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)


def bits_to_normal(bits) -> np.ndarray:
    """jax/_src/random.py _normal_real: sqrt(2) * erf_inv(uniform(nextafter(-1,0), 1))."""
    u = bits_to_uniform(bits, _NORMAL_LO, np.float32(1.0))
    return (np.float32(np.sqrt(2)) * erf_inv_f32(u)).astype(np.float32)


def normal(key, n: int, partitionable: bool = False) -> np.ndarray:
    return bits_to_normal(random_bits(key, n, partitionable))


def randint(key, n: int, minval: int, maxval: int, partitionable: bool = False) -> np.ndarray:
    """jax.random.randint(key, (n,), minval, maxval) for the default int32 dtype.

    Restates jax/_src/random.py ``_randint`` (JAX 0.4.x): ``k1, k2 = split(key)``; 32 bits from each;
    ``span = uint32(maxval - minval)`` (1 when ``maxval <= minval`` so that minval is returned);
    ``multiplier = ((2**16 % span) ** 2) % span``; ``offset = ((hi % span) * multiplier + lo % span) % span``
    with uint32 wrap-around; result ``minval + int32(offset)``.  (The ``maxval`` out-of-range branch cannot
    trigger for int32 bounds.)  UNPINNED against a JAX run (no JAX in this image); used by
    brax's UniformSamplingQueue.sample, reference call site mbpo/systems/brax_wrapper.py:29.
    """
    k = split(key, 2, partitionable)
    hi = random_bits(k[0], n, partitionable).astype(np.uint64)
    lo = random_bits(k[1], n, partitionable).astype(np.uint64)
    minval = int(np.int32(minval))
    maxval = int(np.int32(maxval))
    span = np.uint64((maxval - minval) & 0xFFFFFFFF)
    if maxval <= minval:
        span = np.uint64(1)
    m32 = np.uint64(0xFFFFFFFF)
    mult = np.uint64(2 ** 16) % span
    mult = ((mult * mult) & m32) % span
    off = (((hi % span) * mult) & m32) + (lo % span)
    off = (off & m32) % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int64).astype(np.int32)
