"""jax.random work-alikes on CUDA tensors (threefry2x32, bit-identical to JAX).

Replaces the reference's calls at mbpo/optimizers/trajectory_optimizers/icem_optimizer.py:
123,155,174-180,246 and mbpo/utils/general_utils.py:189-191.  Keys are uint32[..., 2] CUDA
tensors; every function accepts a batch of keys (leading dims) = ``jax.vmap`` of the scalar call.
"""
from __future__ import annotations

import torch

from . import _lib
from .config import config


def PRNGKey(seed: int, device=None) -> torch.Tensor:
    """jax.random.PRNGKey: uint32[2].  With ``config.enable_x64`` False (JAX's default, which the reference never
    changes) the seed is truncated to 32 bits and the high word is 0: PRNGKey(-1) = [0, 0xffffffff],
    PRNGKey((7 << 32) | 9) = [0, 9].  With it True the high word is (seed >> 32) & 0xffffffff."""
    dev = _lib.require_cuda(device)
    seed = int(seed)
    hi = (seed >> 32) & 0xFFFFFFFF if config.enable_x64 else 0
    lo = seed & 0xFFFFFFFF
    # torch has no uint32 constructor from python ints > 2^31 on all versions: go through int64
    return torch.tensor([hi, lo], dtype=torch.int64).to(torch.uint32).to(dev)


def _as_keys(key: torch.Tensor) -> torch.Tensor:
    if key.dtype not in (torch.uint32, torch.int32):
        raise TypeError("PRNG keys must be uint32 (or int32 bit patterns), got %s" % key.dtype)
    if key.shape[-1] != 2:
        raise ValueError("PRNG keys must have shape [..., 2], got %s" % (tuple(key.shape),))
    return key.contiguous()


def split(key: torch.Tensor, num: int = 2) -> torch.Tensor:
    """jax.random.split: uint32[..., 2] -> uint32[..., num, 2]."""
    key = _as_keys(key)
    m = key.numel() // 2
    out = torch.empty(key.shape[:-1] + (num, 2), dtype=torch.uint32, device=key.device)
    with _lib.cuda_guard(key):
        _lib.check(_lib.lib.mbpo_prng_split(_lib.ptr(key), m, num, config.prng_mode, _lib.ptr(out),
                                            _lib.stream_ptr(key.device)))
    return out


def random_bits(key: torch.Tensor, n: int) -> torch.Tensor:
    """jax.random.bits(key, (n,)) : uint32[..., 2] -> uint32[..., n]."""
    key = _as_keys(key)
    out = torch.empty(key.shape[:-1] + (n,), dtype=torch.uint32, device=key.device)
    with _lib.cuda_guard(key):
        _lib.check(_lib.lib.mbpo_prng_random_bits(_lib.ptr(key), key.numel() // 2, n, config.prng_mode,
                                                  _lib.ptr(out), _lib.stream_ptr(key.device)))
    return out


def uniform(key: torch.Tensor, n: int, minval: float = 0.0, maxval: float = 1.0) -> torch.Tensor:
    """jax.random.uniform(key, (n,), minval, maxval) in float32."""
    key = _as_keys(key)
    out = torch.empty(key.shape[:-1] + (n,), dtype=torch.float32, device=key.device)
    with _lib.cuda_guard(key):
        _lib.check(_lib.lib.mbpo_prng_uniform(_lib.ptr(key), key.numel() // 2, n, config.prng_mode, minval, maxval,
                                              _lib.ptr(out), _lib.stream_ptr(key.device)))
    return out


def normal(key: torch.Tensor, n: int) -> torch.Tensor:
    """jax.random.normal(key, (n,)) in float32."""
    key = _as_keys(key)
    out = torch.empty(key.shape[:-1] + (n,), dtype=torch.float32, device=key.device)
    with _lib.cuda_guard(key):
        _lib.check(_lib.lib.mbpo_prng_normal(_lib.ptr(key), key.numel() // 2, n, config.prng_mode, _lib.ptr(out),
                                             _lib.stream_ptr(key.device)))
    return out


def randint(key: torch.Tensor, n, minval: int, maxval: int) -> torch.Tensor:
    """jax.random.randint(key, shape, minval, maxval) (int32).  ``n``: an int (shape ``(n,)``) or a shape tuple -- JAX
    draws the words of a multi-dimensional shape in row-major order, so ``shape=(a, b)`` is the ``a * b`` draw reshaped
    (bptt_optimizer.py:389-391 ``transition_indices``)."""
    key = _as_keys(key)
    shape = (int(n),) if isinstance(n, int) else tuple(int(d) for d in n)
    count = 1
    for d in shape:
        count *= d
    out = torch.empty(key.shape[:-1] + (count,), dtype=torch.int32, device=key.device)
    with _lib.cuda_guard(key):
        _lib.check(_lib.lib.mbpo_prng_randint(_lib.ptr(key), key.numel() // 2, count, config.prng_mode, int(minval),
                                              int(maxval), _lib.ptr(out), _lib.stream_ptr(key.device)))
    return out.reshape(key.shape[:-1] + shape)
