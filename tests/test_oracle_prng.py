"""Pins the oracle's JAX-PRNG restatement (NumPy and the plain-C twin) to external known answers."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import jax_prng as jr
from oracle import mbpo_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "prng_kats.json")))


def _h(x):
    return np.uint32(int(x, 16))


@pytest.mark.parametrize("kat", KATS["threefry2x32"])
def test_threefry_random123_kats(kat, c_oracle):
    y0, y1 = jr.threefry2x32(_h(kat["key"][0]), _h(kat["key"][1]), _h(kat["ctr"][0]), _h(kat["ctr"][1]))
    assert (int(y0), int(y1)) == (int(kat["out"][0], 16), int(kat["out"][1], 16))
    out = (C.c_uint32 * 2)()
    c_oracle.orc_threefry2x32(int(kat["key"][0], 16), int(kat["key"][1], 16), int(kat["ctr"][0], 16),
                              int(kat["ctr"][1], 16), out)
    assert (out[0], out[1]) == (int(kat["out"][0], 16), int(kat["out"][1], 16))


def test_legacy_known_answers():
    k = KATS["legacy"]
    key0 = jr.PRNGKey(0)
    assert key0.tolist() == [0, 0] and jr.PRNGKey(42).tolist() == [0, 42]
    # x64 disabled (the reference's setting): the seed is truncated to 32 bits, the high word is always 0
    assert jr.PRNGKey((7 << 32) | 9).tolist() == [0, 9] and jr.PRNGKey(-1).tolist() == [0, 0xFFFFFFFF]
    assert jr.PRNGKey((7 << 32) | 9, enable_x64=True).tolist() == [7, 9]
    assert jr.PRNGKey(-1, enable_x64=True).tolist() == [0xFFFFFFFF, 0xFFFFFFFF]
    sp = jr.split(key0)
    assert sp.tolist() == k["split_prngkey0"]
    assert jr.normal(key0, 1)[0] == pytest.approx(k["normal_prngkey0"], rel=1e-6)
    # the JAX quick-start: key, subkey = split(key); normal(subkey); then normal(new key)
    assert jr.normal(sp[1], 1)[0] == pytest.approx(k["normal_split_subkey"], rel=1e-6)
    assert jr.normal(sp[0], 1)[0] == pytest.approx(k["normal_split_newkey"], rel=1e-6)
    assert jr.uniform(key0, 1)[0] == pytest.approx(k["uniform_prngkey0"], rel=1e-7)
    assert jr.normal(jr.PRNGKey(42), 1)[0] == pytest.approx(k["normal_prngkey42"], rel=1e-6)


def test_partitionable_restatement():
    k = KATS["partitionable_restatement_only"]
    assert jr.split(jr.PRNGKey(0), 2, partitionable=True).tolist() == k["split_key0"]
    assert jr.normal(jr.PRNGKey(42), 1, partitionable=True)[0] == pytest.approx(k["normal_key42"], rel=1e-6)


@pytest.mark.parametrize("partitionable", [False, True])
@pytest.mark.parametrize("n", [1, 2, 3, 11, 16, 501])
def test_c_twin_prng_bit_exact(c_oracle, partitionable, n):
    rng = np.random.default_rng(n)
    key = rng.integers(0, 2 ** 32, 2, dtype=np.uint64).astype(np.uint32)
    out = np.zeros((n, 2), np.uint32)
    c_oracle.orc_split(key.ctypes.data_as(C.c_void_p), n, int(partitionable), out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, jr.split(key, n, partitionable))
    bits = np.zeros(n, np.uint32)
    c_oracle.orc_random_bits(key.ctypes.data_as(C.c_void_p), n, int(partitionable), bits.ctypes.data_as(C.c_void_p))
    assert np.array_equal(bits, jr.random_bits(key, n, partitionable))
    z = np.zeros(n, np.float32)
    c_oracle.orc_normal(key.ctypes.data_as(C.c_void_p), n, int(partitionable), z.ctypes.data_as(C.c_void_p))
    np.testing.assert_allclose(z, jr.normal(key, n, partitionable), rtol=2e-6, atol=1e-7)


def test_vmapped_helpers_equal_scalar_calls():
    keys = np.random.default_rng(0).integers(0, 2 ** 32, (5, 2), dtype=np.uint64).astype(np.uint32)
    for part in (False, True):
        for num in (1, 3, 6):
            assert np.array_equal(orc.split_keys(keys, num, part), np.stack([jr.split(k, num, part) for k in keys]))
        for n in (1, 11, 16):
            assert np.array_equal(orc.random_bits_keys(keys, n, part),
                                  np.stack([jr.random_bits(k, n, part) for k in keys]))


def test_normal_distribution_sanity():
    z = jr.normal(jr.PRNGKey(1), 200000)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert np.isfinite(z).all()
    # extreme words: all-zero bits give the clamped lower end, all-one bits stay finite
    assert np.isfinite(jr.bits_to_normal(np.array([0, 0xFFFFFFFF], np.uint32))).all()
