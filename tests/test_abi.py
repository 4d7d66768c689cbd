"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/mbpo_b200.h declares, agrees with the ctypes struct layouts, and fails loudly (error
codes, no fallback) when asked to compute without a CUDA device.  No kernels run here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import mbpo_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mbpo_b200.h")


@pytest.fixture(scope="module")
def mb():
    import mbpo_b200
    return mbpo_b200


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mbpo_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(mb):
    names = _declared_symbols()
    assert len(names) >= 20
    lib = C.CDLL(mb._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libmbpo_b200.so does not export %s" % n
    assert set(names) == set(mb._lib.SIGNATURES), "ctypes binding and header disagree"


def test_abi_version_and_struct_layout(mb):
    L = mb._lib
    assert L.lib.mbpo_abi_version() == L.MBPO_ABI_VERSION == 1
    for which, st in enumerate((L.IcemCfgC, L.PendulumParamsC, L.MlpEnsembleParamsC, L.IcemTraceC, L.PolicyParamsC,
                                L.ReplayStateC, L.ReplayFieldsC)):
        assert L.lib.mbpo_struct_size(which) == C.sizeof(st)
    assert L.lib.mbpo_struct_size(99) == 0
    assert C.sizeof(L.PendulumParamsC) == 36


@pytest.mark.parametrize("horizon,exponent", [(20, 0.0), (30, 2.0), (15, 1.0), (50, 0.5), (5, 3.0)])
def test_cfg_init_matches_reference_trace_time_constants(mb, horizon, exponent):
    """mbpo_icem_cfg_init runs on the host: s_scale / sigma (general_utils.py:143-178) and
    num_prev_elites (icem_optimizer.py:170) against the oracle."""
    L = mb._lib
    cfg = L.IcemCfgC()
    L.check(L.lib.mbpo_icem_cfg_init(C.byref(cfg), horizon, 1, 3, 10, 500, 50, 0.5, 0.0, 5, exponent, 0.3, -1.0, 1.0,
                                     1, 1e4))
    s_scale, sigma = orc.powerlaw_tables(exponent, horizon)
    F = horizon // 2 + 1
    np.testing.assert_allclose(np.array(cfg.s_scale[:F]), s_scale, rtol=3e-7)
    assert cfg.sigma == pytest.approx(float(sigma), rel=5e-7)
    assert cfg.num_prev_elites == orc.ICemParams().num_prev_elites == 15
    assert (cfg.horizon, cfg.num_samples, cfg.num_elites, cfg.num_steps, cfg.warm_start) == (horizon, 500, 50, 5, 1)
    assert cfg.prng_mode == L.PRNG_LEGACY and cfg.math_mode == L.MATH_REFERENCE
    for frac, k, want in ((0.3, 50, 15), (0.01, 50, 1), (0.5, 7, 3), (0.3, 10, 3)):
        L.check(L.lib.mbpo_icem_cfg_init(C.byref(cfg), horizon, 1, 3, 1, 64, k, 0.5, 0.0, 5, 0.0, frac, -1.0, 1.0, 1, 1e4))
        assert cfg.num_prev_elites == want == max(int(frac * k), 1)


def test_error_codes_without_gpu(mb):
    L = mb._lib
    cfg = L.IcemCfgC()
    assert L.lib.mbpo_icem_cfg_init(C.byref(cfg), 1000, 1, 3, 1, 64, 8, 0.5, 0.0, 5, 0.0, 0.3, -1.0, 1.0, 1, 1e4) == L.MBPO_EINVAL
    assert b"horizon" in L.lib.mbpo_last_error()
    L.check(L.lib.mbpo_icem_cfg_init(C.byref(cfg), 30, 1, 3, 1, 512, 50, 0.5, 0.0, 5, 0.0, 0.3, -1.0, 1.0, 1, 1e4))
    assert L.lib.mbpo_icem_plan_is_fused(C.byref(cfg)) == 1
    assert L.lib.mbpo_icem_workspace_bytes(C.byref(cfg), 16) > 16 * 527 * 30 * 4
    cfg.horizon = 21                                   # no unrolled instance for H=21: the any-horizon fused kernel
    assert L.lib.mbpo_icem_plan_is_fused(C.byref(cfg)) == 1 and L.lib.mbpo_icem_plan_cluster_size(C.byref(cfg), 1) == 0
    cfg.horizon = 128                                  # 513 x 129 floats + 256 staging rows do not fit: staged
    assert L.lib.mbpo_icem_plan_is_fused(C.byref(cfg)) == 0
    cfg.horizon = 30
    cfg.num_samples = 4000                             # 4000 x 31 floats do not fit 227 KB of shared memory
    assert L.lib.mbpo_icem_plan_is_fused(C.byref(cfg)) == 0
    cfg.num_samples, cfg.num_elites = 8, 100           # K > N + Np
    assert L.lib.mbpo_icem_plan_is_fused(C.byref(cfg)) == 0
    # null pointers are EINVAL, never a crash
    assert L.lib.mbpo_prng_split(None, 4, 2, 0, None, None) == L.MBPO_EINVAL
    assert L.lib.mbpo_system_step(0, None, 0, None, None, 4, None, None, None) == L.MBPO_EINVAL
    assert L.lib.mbpo_system_step(9, None, 0, None, None, 4, None, None, None) == L.MBPO_EINVAL
    # replay queue: argument errors are reported before anything is launched
    st = L.ReplayStateC(data=None, capacity=10, row_width=9)
    assert L.lib.mbpo_replay_insert(C.byref(st), None, 1, None) == L.MBPO_EINVAL
    st = L.ReplayStateC(data=0x1000, capacity=10, row_width=9)
    fields = L.ReplayFieldsC(num_fields=1)
    fields.width[0], fields.ptr[0] = 9, 0x2000
    assert L.lib.mbpo_replay_insert(C.byref(st), C.byref(fields), 11, None) == L.MBPO_EINVAL   # brax raises ValueError
    assert b"larger than the maximum replay size" in L.lib.mbpo_last_error()
    fields.width[0] = 8
    assert L.lib.mbpo_replay_insert(C.byref(st), C.byref(fields), 1, None) == L.MBPO_EINVAL
    assert b"row_width" in L.lib.mbpo_last_error()
    assert L.lib.mbpo_replay_sample(C.byref(st), None, 0, 4, None, None, None, None) == L.MBPO_EINVAL
    assert L.lib.mbpo_env_reset_from_buffer(C.byref(st), None, 4, 0, 1, 3, 9, None, None, None, None, None) == L.MBPO_EINVAL
    assert L.lib.mbpo_prng_randint(None, 4, 2, 0, 0, 10, None, None) == L.MBPO_EINVAL
    assert L.lib.mbpo_running_statistics_workspace_bytes(3) == 148 * 4 * 2 * 3 * 8
    assert L.lib.mbpo_running_statistics_workspace_bytes(1000) == 0
    assert L.lib.mbpo_running_statistics_accumulate(None, 4, 3, None, None, 0, None, None) == L.MBPO_EINVAL
    assert L.lib.mbpo_running_statistics_accumulate(0x1000, 4, 3, 0x1000, 0x1000, 8, 0x1000, None) == L.MBPO_EWORKSPACE
    assert L.lib.mbpo_running_statistics_finalize(None, 3, None, None, None, 1e-6, 1e6, None, None, None, None, None) == L.MBPO_EINVAL
    assert L.lib.mbpo_compute_gae(None, None, None, None, None, 4, 4, 4, 1, 0.99, 0.95, None, None, None) == L.MBPO_EINVAL
    assert L.lib.mbpo_eval_metrics(None, None, None, None, 1, 4, 4, 4, 1, None, None, None, None) == L.MBPO_EINVAL
    with pytest.raises(mb.MbpoError):
        L.check(L.MBPO_EINVAL)
    with pytest.raises(NotImplementedError):
        L.check(L.MBPO_EUNSUPPORTED)


def test_no_cpu_fallback(mb):
    """Without a GPU the product path refuses to run (the oracle is never consulted)."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mbpo_b200.optimizers import iCemTO, iCemParams
    from mbpo_b200.systems import PendulumSystem, SystemParams
    with pytest.raises(mb.MbpoError):
        mb.random.PRNGKey(0)
    with pytest.raises(mb.MbpoError):
        PendulumSystem().step(torch.zeros(3), torch.zeros(1), SystemParams())
    with pytest.raises(mb.MbpoError):
        mb.random.split(torch.zeros(2, dtype=torch.uint32))
    # the replay queue, the normalisers, GAE and the evaluation metrics have no CPU path either
    from mbpo_b200 import running_statistics as rs
    from mbpo_b200.replay_buffers import UniformSamplingQueue
    from mbpo_b200.utils import compute_gae
    from mbpo_b200.utils.optimizer_utils import Transition
    z = torch.zeros
    q = UniformSamplingQueue(8, Transition(z(3), z(1), z(()), z(()), z(3)), 2)
    assert q.row_width == 9
    st = q.init(torch.zeros(2, dtype=torch.int32).view(torch.uint32))      # host tensors: every launch refuses them
    with pytest.raises(mb.MbpoError):
        q.insert(st, Transition(z(4, 3), z(4, 1), z(4), z(4), z(4, 3)))
    with pytest.raises(mb.MbpoError):
        q.sample(st)
    with pytest.raises(mb.MbpoError):
        rs.init_state(3)
    with pytest.raises(mb.MbpoError):
        rs.update(rs.RunningStatisticsState(z(()), z(3), z(3), torch.ones(3)), z(5, 3))
    with pytest.raises(mb.MbpoError):
        compute_gae(z(4, 2), z(4, 2), z(4, 2), z(4, 2), z(2))
    opt = iCemTO(horizon=20, action_dim=1, opt_params=iCemParams())
    opt.set_system(PendulumSystem())
    cfg = opt._cfg()                                   # host-only: allowed
    assert cfg.num_prev_elites == 15 and cfg.system_kind == mb._lib.SYSTEM_PENDULUM
    import sys
    assert not any(m.startswith("oracle") for m in sys.modules if "mbpo_b200" in m)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "model-based-policy-optimizers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in text and "import oracle" not in text, os.path.join(dirpath, f)


def test_api_surface_matches_reference(mb):
    """Names and argument order of the reference interface."""
    import inspect
    from mbpo_b200.optimizers import BaseOptimizer, iCEMOptimizer, iCemParams, iCemTO
    from mbpo_b200.systems import PendulumSystem, System, SystemParams, SystemState
    from mbpo_b200.utils import rollout_actions
    assert iCemParams._fields == ("num_particles", "num_samples", "num_elites", "init_std", "alpha", "num_steps",
                                  "exponent", "elite_set_fraction", "u_min", "u_max", "warm_start", "lambda_constraint")
    assert iCemParams() == (10, 500, 50, 0.5, 0.0, 5, 0.0, 0.3, -1.0, 1.0, True, 1e4)
    assert list(inspect.signature(iCemTO.__init__).parameters)[:8] == [
        "self", "horizon", "action_dim", "key", "opt_params", "cost_fn", "use_optimism", "use_pessimism"]
    assert list(inspect.signature(iCemTO.optimize).parameters) == ["self", "initial_state", "opt_state"]
    assert list(inspect.signature(iCemTO.act).parameters) == ["self", "obs", "opt_state", "evaluate"]
    assert list(inspect.signature(iCemTO.init).parameters) == ["self", "key", "true_buffer_state"]
    assert list(inspect.signature(iCEMOptimizer.__init__).parameters)[:5] == ["self", "horizon", "opt_params", "system", "key"]
    assert list(inspect.signature(System.step).parameters) == ["self", "x", "u", "system_params"]
    assert list(inspect.signature(rollout_actions).parameters) == ["system", "system_params", "init_state", "actions", "horizon"]
    from mbpo_b200 import acting
    from mbpo_b200.utils import lambda_return, rollout_policy
    assert list(inspect.signature(rollout_policy).parameters) == [                       # optimizer_utils.py:63-71
        "system", "system_params", "init_state", "policy", "policy_state", "horizon", "stop_grads"]
    assert inspect.signature(rollout_policy).parameters["stop_grads"].default is True
    assert list(inspect.signature(lambda_return).parameters) == ["reward", "next_values", "discount", "lambda_"]   # :120-123
    assert list(inspect.signature(acting.actor_step).parameters) == [                    # sac/acting.py:35-40
        "env", "env_state", "policy", "key", "extra_fields"]
    assert list(inspect.signature(acting.generate_unroll).parameters)[:6] == [           # sac/acting.py:58-64
        "env", "env_state", "policy", "key", "unroll_length", "extra_fields"]                # (+ additive sharding kwargs)
    from mbpo_b200.replay_buffers import UniformSamplingQueue
    from mbpo_b200.systems import BraxWrapper
    assert list(inspect.signature(UniformSamplingQueue.__init__).parameters)[:4] == [    # brax replay_buffers.QueueBase
        "self", "max_replay_size", "dummy_data_sample", "sample_batch_size"]
    assert list(inspect.signature(UniformSamplingQueue.insert).parameters) == ["self", "buffer_state", "samples"]
    assert list(inspect.signature(UniformSamplingQueue.sample).parameters) == ["self", "buffer_state"]
    assert list(inspect.signature(BraxWrapper.__init__).parameters) == [                 # brax_wrapper.py:15-19
        "self", "system", "system_params", "sample_buffer_state", "sample_buffer"]
    assert list(inspect.signature(BraxWrapper.reset).parameters) == ["self", "rng"]
    assert list(inspect.signature(BraxWrapper.step).parameters) == ["self", "state", "action"]
    from mbpo_b200.utils import compute_gae
    assert list(inspect.signature(compute_gae).parameters) == [                          # ppo/losses.py:128-134 (+ brax's keywords)
        "truncation", "termination", "rewards", "values", "bootstrap_value", "lambda_", "discount"]
    assert list(inspect.signature(acting.Evaluator.__init__).parameters) == [            # sac/acting.py:85-88
        "self", "eval_env", "eval_policy_fn", "num_eval_envs", "episode_length", "action_repeat", "key"]
    from mbpo_b200 import running_statistics as rs
    assert list(inspect.signature(rs.update).parameters)[:7] == [                        # brax acme/running_statistics.update
        "state", "batch", "weights", "std_min_value", "std_max_value", "pmap_axis_name", "validate_shapes"]
    assert list(inspect.signature(rs.normalize).parameters) == ["batch", "mean_std", "max_abs_value"]
    assert list(inspect.signature(rs.Normalizer.update).parameters)[:2] == ["x", "state"]   # bptt_optimizer.py:51
    assert list(inspect.signature(acting.ExperienceCollector.get_experience).parameters) == [   # sac/sac.py:283-285
        "self", "normalizer_params", "policy_params", "env_state", "buffer_state", "key"]
    assert list(inspect.signature(acting.Evaluator.run_evaluation).parameters) == [      # :118-122
        "self", "policy_params", "training_metrics", "unroll_key", "aggregate_episodes"]
    assert issubclass(iCemTO, BaseOptimizer) and issubclass(PendulumSystem, System)
    assert iCEMOptimizer(horizon=20).can_act_in_batches is False
    s = PendulumSystem()
    assert (s.x_dim, s.u_dim, s.min_action, s.max_action) == (3, 1, -1.0, 1.0)
    sp = System.system_params_vmap_axes(0)
    assert sp.key == 0 and sp.dynamics_params is None
    st = SystemState(x_next=1, reward=2, system_params=SystemParams())
    assert st.done == 0.0 and st.replace(reward=3).reward == 3
