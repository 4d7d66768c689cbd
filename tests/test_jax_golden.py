"""Checks against outputs of the REAL reference (tests/golden/jax_golden.npz, written by
tests/golden/make_golden_jax.py where JAX is installed).  Skipped while that file is absent -- which is the case in
this repository today (no JAX wheel in the image): see DESIGN.md section 2 for what pins the float path meanwhile."""
import os

import numpy as np
import pytest

from oracle import jax_prng as ojr
from oracle import mbpo_oracle as orc

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jax_golden.npz")
needs_fixture = pytest.mark.skipif(not os.path.exists(PATH), reason="tests/golden/jax_golden.npz not generated "
                                   "(needs a JAX install: python tests/golden/make_golden_jax.py)")
RTOL = 1e-5


@pytest.fixture(scope="module")
def g():
    return np.load(PATH)


def test_generator_script_is_committed_and_parses():
    import ast
    src = os.path.join(os.path.dirname(PATH), "make_golden_jax.py")
    ast.parse(open(src).read())



@needs_fixture
def test_oracle_prng_against_jax(g):
    part = bool(g["threefry_partitionable"])
    key = ojr.PRNGKey(1234)
    assert np.array_equal(g["prngkey_1234"], key) and np.array_equal(g["prngkey_minus1"], ojr.PRNGKey(-1))
    assert np.array_equal(g["split5"], ojr.split(key, 5, part))
    assert np.array_equal(g["bits11"], ojr.random_bits(key, 11, part))
    np.testing.assert_allclose(g["normal16"], ojr.normal(key, 16, part), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(g["uniform9"], ojr.uniform(key, 9, -2.0, 3.0, part), rtol=0, atol=0)
    assert np.array_equal(g["randint33_0_10"], ojr.randint(key, 33, 0, 10, part))
    assert np.array_equal(g["randint7_5_1000000"], ojr.randint(key, 7, 5, 1000000, part))


@needs_fixture
def test_oracle_floats_against_jax(g):
    part = bool(g["threefry_partitionable"])
    for name in g.files:
        if name.startswith("noise_h"):
            h, ex = name[len("noise_h"):].split("_e")
            want = orc.powerlaw_psd_gaussian_keys(float(ex), int(h), g["noise_keys"], part)
            np.testing.assert_allclose(want, g[name], rtol=RTOL, atol=5e-6)
    xn, r = orc.pendulum_step(g["step_x"], g["step_u"][:, 0])
    np.testing.assert_allclose(xn, g["step_xn"], rtol=RTOL, atol=2e-6)
    np.testing.assert_allclose(r, g["step_r"], rtol=RTOL, atol=2e-6)
    for h in (20, 30, 50):
        obs = g["roll_h%d_observation" % h]
        xn, r = orc.pendulum_step(obs.reshape(-1, 3), g["roll_h%d_actions" % h].reshape(-1))     # teacher-forced
        np.testing.assert_allclose(xn, g["roll_h%d_next_observation" % h].reshape(-1, 3), rtol=RTOL, atol=3e-6)
        np.testing.assert_allclose(r, g["roll_h%d_reward" % h].reshape(-1), rtol=RTOL, atol=3e-6)


@needs_fixture
def test_oracle_icem_keys_against_jax(g):
    part = bool(g["threefry_partitionable"])
    for tag, h in (("defaults_h20", 20), ("config2_h30", 30), ("colored_h30", 30)):
        st = orc.icem_init(ojr.PRNGKey(11), h, 1, part)
        assert np.array_equal(g["icem_%s_state_key" % tag], st.key)
        assert np.array_equal(g["icem_%s_key" % tag][0], ojr.split(st.key, 2, part)[1])
    assert float(g["mpc_rewards"].sum()) >= -400                                  # tests/test_icemopt.py:37-38


@needs_fixture
@pytest.mark.gpu
def test_cuda_path_against_jax(g, cuda_device):
    import torch
    import mbpo_b200
    from mbpo_b200.optimizers import iCemTO, iCemParams
    from mbpo_b200.systems import PendulumSystem
    mbpo_b200.config.threefry_partitionable = bool(g["threefry_partitionable"])
    try:
        dev = cuda_device
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        system = PendulumSystem()
        sp = system.reset(device=dev).system_params
        out = system.step(d(g["step_x"]), d(g["step_u"]), sp)
        np.testing.assert_allclose(out.x_next.cpu().numpy(), g["step_xn"], rtol=RTOL, atol=2e-6)
        np.testing.assert_allclose(out.reward.cpu().numpy(), g["step_r"], rtol=RTOL, atol=2e-6)
        for tag, h, params in (("defaults_h20", 20, {}), ("config2_h30", 30, dict(num_samples=512, num_particles=1)),
                               ("colored_h30", 30, dict(num_samples=512, num_particles=1, exponent=2.0, alpha=0.1))):
            opt = iCemTO(horizon=h, action_dim=1, opt_params=iCemParams(**params))
            opt.set_system(system)
            st = opt.init(mbpo_b200.random.PRNGKey(11, dev))
            assert np.array_equal(st.key.cpu().numpy(), g["icem_%s_state_key" % tag])
            agree = 0
            for i in range(4):
                new = opt.optimize(d(g["icem_%s_x0" % tag][i]), st)
                assert np.array_equal(new.key.cpu().numpy(), g["icem_%s_key" % tag][i])
                # free-running: equal unless a sub-tolerance elite flip separated the two runs (counted, printed)
                if np.allclose(new.best_sequence.cpu().numpy(), g["icem_%s_best_sequence" % tag][i], rtol=RTOL, atol=1e-5):
                    agree += 1
                    np.testing.assert_allclose(float(new.best_reward), float(g["icem_%s_best_reward" % tag][i]),
                                               rtol=2e-5, atol=2e-6)
            print("%s: %d of 4 plans agree end to end with the reference" % (tag, agree))
            assert agree >= 1
    finally:
        mbpo_b200.config.threefry_partitionable = False
