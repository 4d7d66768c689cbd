"""GPU parity tests of the replay queue and BraxWrapper.reset (mbpo_replay_*, mbpo_env_reset_from_buffer,
mbpo_prng_randint) against oracle/brax_replay.py and oracle/jax_prng.randint.  Everything here is integer / byte
work: the bar is bit-exact."""
import numpy as np
import pytest
import torch

from oracle import brax_replay as obr
from oracle import jax_prng as ojr
from oracle import mbpo_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb(cuda_device):
    import mbpo_b200
    return mbpo_b200


@pytest.fixture(params=[False, True], ids=["legacy", "partitionable"])
def prng_mode(request, mb):
    mb.config.threefry_partitionable = request.param
    yield request.param
    mb.config.threefry_partitionable = False


def _dev(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _sac_dummy(mb, dev):
    """sac.py:194-200: scalar reward / discount, truncation in extras -> a row of 2X + A + 3 = 10 floats."""
    from mbpo_b200.utils.optimizer_utils import Transition
    z = torch.zeros
    return Transition(observation=z(3, device=dev), action=z(1, device=dev), reward=z((), device=dev),
                      discount=z((), device=dev), next_observation=z(3, device=dev),
                      extras={"state_extras": {"truncation": z((), device=dev)}, "policy_extras": {}})


def _true_dummy(mb, dev):
    """tests/test_sac.py:15-19: no extras -> a row of 9 floats."""
    from mbpo_b200.utils.optimizer_utils import Transition
    z = torch.zeros
    return Transition(observation=z(3, device=dev), action=z(1, device=dev), reward=z((), device=dev),
                      discount=z((), device=dev), next_observation=z(3, device=dev))


def _rows_of(tr, n):
    """ravel_pytree order of a Transition batch: the fields side by side."""
    cols = [tr.observation.reshape(n, 3), tr.action.reshape(n, 1), tr.reward.reshape(n, 1),
            tr.discount.reshape(n, 1), tr.next_observation.reshape(n, 3)]
    if isinstance(tr.extras, dict):
        cols.append(tr.extras["state_extras"]["truncation"].reshape(n, 1))
    return torch.cat(cols, dim=1).cpu().numpy()


@pytest.mark.parametrize("minval,maxval", [(0, 10), (3, 3), (7, 2), (-5, 5), (0, 2 ** 31 - 1), (-2 ** 31, 2 ** 31 - 1),
                                            (0, 1), (0, 65537), (100, 1_000_000)])
def test_randint_bit_exact(mb, cuda_device, prng_mode, minval, maxval):
    rng = np.random.default_rng(maxval & 0xFFFF)
    keys = rng.integers(0, 2 ** 32, size=(5, 2), dtype=np.uint64).astype(np.uint32)
    for n in (1, 2, 7, 64):
        got = mb.random.randint(_dev(keys, cuda_device), n, minval, maxval).cpu().numpy()
        want = np.stack([ojr.randint(k, n, minval, maxval, prng_mode) for k in keys])
        assert got.dtype == np.int32 and np.array_equal(got, want)
    # a shape tuple is the flat draw reshaped (bptt_optimizer.py:389-391: shape=(updates, batch_size))
    got = mb.random.randint(_dev(keys[0], cuda_device), (4, 16), minval, maxval).cpu().numpy()
    assert got.shape == (4, 16) and np.array_equal(got.reshape(-1), ojr.randint(keys[0], 64, minval, maxval, prng_mode))


@pytest.mark.parametrize("layout", ["sac", "true"])
def test_queue_insert_sample_match_oracle(mb, cuda_device, prng_mode, layout):
    from mbpo_b200.replay_buffers import UniformSamplingQueue
    from mbpo_b200.utils.optimizer_utils import Transition
    dev = cuda_device
    dummy = _sac_dummy(mb, dev) if layout == "sac" else _true_dummy(mb, dev)
    D = 10 if layout == "sac" else 9
    R, batch = 53, 17
    q = UniformSamplingQueue(R, dummy, batch)
    assert q.row_width == D
    oq = obr.UniformSamplingQueue(R, D, batch, prng_mode)
    key = ojr.PRNGKey(11)
    st, ost = q.init(_dev(key, dev)), oq.init(key)
    rng = np.random.default_rng(3)
    for n in [1, 6, 40, 0, 7, 53, 2, 52, 9, 9, 9]:        # fills, wraps, replaces the whole queue, empty insert
        g = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32)).to(dev)
        extras = {"state_extras": {"truncation": g(n)}, "policy_extras": {}} if layout == "sac" else ()
        tr = Transition(observation=g(n, 3), action=g(n, 1), reward=g(n), discount=g(n), next_observation=g(n, 3),
                        extras=extras)
        st = q.insert(st, tr)
        ost = oq.insert(ost, _rows_of(tr, n))
        assert (st.insert_position, st.sample_position) == (ost.insert_position, ost.sample_position)
        assert q.size(st) == ost.insert_position - ost.sample_position
        live = ost.insert_position
        assert np.array_equal(st.data.cpu().numpy()[:live], ost.data[:live])
        st, got, idx = q.sample_with_indices(st)
        ost, want_rows, want_idx = oq.sample(ost)
        assert np.array_equal(idx.cpu().numpy(), want_idx)
        assert np.array_equal(st.key.cpu().numpy(), ost.key)
        assert np.array_equal(_rows_of(got, batch), want_rows)
        assert got.observation.shape == (batch, 3) and got.reward.shape == (batch,)
    with pytest.raises(ValueError):
        g = lambda *s: torch.zeros(s, device=dev)
        q.insert(st, Transition(g(R + 1, 3), g(R + 1, 1), g(R + 1), g(R + 1), g(R + 1, 3),
                                {"state_extras": {"truncation": g(R + 1)}, "policy_extras": {}} if layout == "sac" else ()))


def test_sample_from_empty_queue_returns_row_zero(mb, cuda_device):
    from mbpo_b200.replay_buffers import UniformSamplingQueue
    q = UniformSamplingQueue(10, _true_dummy(mb, cuda_device), 8)
    st = q.init(_dev(ojr.PRNGKey(0), cuda_device))
    st2, batch, idx = q.sample_with_indices(st)
    assert torch.all(idx == 0) and torch.all(batch.observation == 0)
    assert np.array_equal(st2.key.cpu().numpy(), ojr.split(ojr.PRNGKey(0), 2)[0])


@pytest.mark.parametrize("sample_batch_size", [1, 5])
def test_brax_wrapper_reset_matches_oracle(mb, cuda_device, prng_mode, sample_batch_size):
    """tests/test_sac.py:15-28 builds the true buffer; BraxWrapper.reset under VmapWrapper draws every env's first
    observation from it (brax_wrapper.py:25-38)."""
    from mbpo_b200.replay_buffers import UniformSamplingQueue
    from mbpo_b200.systems import BraxWrapper, PendulumSystem
    from mbpo_b200.utils.optimizer_utils import Transition
    dev = cuda_device
    system = PendulumSystem()
    q = UniformSamplingQueue(10, _true_dummy(mb, dev), sample_batch_size)
    oq = obr.UniformSamplingQueue(10, 9, sample_batch_size, prng_mode)
    st, ost = q.init(_dev(ojr.PRNGKey(0), dev)), oq.init(ojr.PRNGKey(0))
    rng = np.random.default_rng(5)
    for n in (4, 4, 4):                                    # the third insert wraps the ring (head != 0)
        g = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32)).to(dev)
        tr = Transition(g(n, 3), g(n, 1), g(n), g(n), g(n, 3))
        st, ost = q.insert(st, tr), oq.insert(ost, _rows_of(tr, n))
    assert st.head != 0
    params = system.init_params(_dev(ojr.PRNGKey(1), dev))
    env = BraxWrapper(system, params, st, q)
    E = 300
    rngs = ojr.split(ojr.PRNGKey(9), E, prng_mode)
    state = env.reset(_dev(rngs, dev))
    obs, reward, sys_keys, _ = obr.brax_wrapper_reset(rngs, oq, ost, 3, 1)
    assert np.array_equal(state.obs.cpu().numpy(), obs)
    assert np.array_equal(state.reward.cpu().numpy(), reward)
    assert np.array_equal(state.system_params.key.cpu().numpy(), sys_keys)
    assert torch.all(state.done == 0) and state.pipeline_state is None
    assert len(np.unique(obs, axis=0)) > 1                # several different rows were drawn
    one = env.reset(_dev(rngs[7], dev))                  # un-vmapped call
    assert np.array_equal(one.obs.cpu().numpy(), obs[7]) and one.reward.shape == ()
    nxt = env.step(state, torch.zeros((E, 1), device=dev))
    want = orc.pendulum_step(obs, np.zeros(E, np.float32))
    np.testing.assert_allclose(nxt.obs.cpu().numpy(), want[0], rtol=1e-5, atol=1e-6)


def test_collect_insert_sample_cycle(mb, cuda_device):
    """SAC's get_experience (sac.py:283-304): wrap(BraxWrapper) -> reset(keys) -> unroll -> insert the time-major
    Transition -> sample; every row of the queue is a transition of the unroll in (t, e) order."""
    from mbpo_b200 import envs
    from mbpo_b200.replay_buffers import UniformSamplingQueue
    from mbpo_b200.systems import BraxWrapper, PendulumSystem
    from mbpo_b200.utils.optimizer_utils import Transition
    dev = cuda_device
    system = PendulumSystem()
    true_q = UniformSamplingQueue(10, _true_dummy(mb, dev), 1)
    first = system.reset(device=dev)
    true_st = true_q.insert(true_q.init(_dev(ojr.PRNGKey(0), dev)),
                            Transition(first.x_next[None], torch.zeros((1, 1), device=dev), first.reward[None],
                                       torch.full((1,), 0.99, device=dev), first.x_next[None]))
    env = envs.wrap(BraxWrapper(system, system.init_params(_dev(ojr.PRNGKey(1), dev)), true_st, true_q),
                    episode_length=7, action_repeat=1)
    E, T = 96, 20
    state = env.reset(mb.random.split(_dev(ojr.PRNGKey(2), dev), E))
    assert torch.equal(state.obs, first.x_next.expand(E, 3))          # the true buffer holds one row
    actions = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, (T, E, 1)).astype(np.float32)).to(dev)
    state, tr = env.unroll(state, actions)
    R = 2000                                                          # T * E = 1920: the second insert wraps the ring
    q = UniformSamplingQueue(R, _sac_dummy(mb, dev), 256)
    st = q.init(_dev(ojr.PRNGKey(3), dev))
    st = q.insert(st, tr)
    state, tr2 = env.unroll(state, actions[:5])
    st = q.insert(st, tr2)
    want = np.concatenate([_rows_of(tr, T * E), _rows_of(tr2, 5 * E)])[-R:]
    assert st.insert_position == R
    assert np.array_equal(st.data.cpu().numpy(), want)
    st, batch, idx = q.sample_with_indices(st)
    assert np.array_equal(_rows_of(batch, 256), want[idx.cpu().numpy()])
    assert batch.extras["state_extras"]["truncation"].shape == (256,)


def test_insert_at_full_size_is_a_permutation_free_copy(mb, cuda_device):
    """Config-3 sized rows (65,536 envs x 16 steps = 1 M rows of 10 floats), size-independent check: column sums of
    the queue equal the fields' sums exactly (integers below 2**24 in float64 accumulation), before and after the
    ring wraps, and the logical order is the insert order."""
    from mbpo_b200.replay_buffers import UniformSamplingQueue
    from mbpo_b200.utils.optimizer_utils import Transition
    dev = cuda_device
    E, T = 65536, 16
    n = E * T
    q = UniformSamplingQueue(n + 1000, _sac_dummy(mb, dev), 1)
    st = q.init(_dev(ojr.PRNGKey(0), dev))
    for rep in range(2):                                              # the second insert wraps the ring
        gen = torch.Generator(device=dev).manual_seed(rep)
        g = lambda *s: torch.randint(0, 1000, s, generator=gen, device=dev).to(torch.float32)
        tr = Transition(g(T, E, 3), g(T, E, 1), g(T, E), g(T, E), g(T, E, 3),
                        {"state_extras": {"truncation": g(T, E)}, "policy_extras": {}})
        st = q.insert(st, tr)
        data = st.data
        live = data[st.insert_position - n:st.insert_position]
        want = torch.cat([tr.observation.reshape(n, 3), tr.action.reshape(n, 1), tr.reward.reshape(n, 1),
                          tr.discount.reshape(n, 1), tr.next_observation.reshape(n, 3),
                          tr.extras["state_extras"]["truncation"].reshape(n, 1)], dim=1)
        assert torch.equal(live, want)
        assert torch.equal(live.double().sum(0), want.double().sum(0))
    assert st.head != 0 and st.insert_position == n + 1000


def test_eval_metrics_match_oracle_and_fold_in_chunks(mb, cuda_device):
    from mbpo_b200 import acting
    from mbpo_b200.envs import EnvState
    from mbpo_b200.utils.optimizer_utils import Transition
    dev = cuda_device
    rng = np.random.default_rng(1)
    T, E, rep = 57, 1000, 2
    reward = rng.standard_normal((T, E)).astype(np.float32)
    discount = (rng.random((T, E)) > 0.02).astype(np.float32)         # dones at arbitrary steps
    steps0 = (rng.integers(0, 5, E) * rep).astype(np.float32)
    done0 = (rng.random(E) < 0.1).astype(np.float32)
    want = obr.eval_metrics(reward, discount, steps0, done0, rep)
    st = EnvState(obs=None, reward=None, done=_dev(done0, dev), system_params=None, info={"steps": _dev(steps0, dev)})
    tr = Transition(None, None, _dev(reward, dev), _dev(discount, dev), None)
    got = acting.eval_metrics(st, tr, rep)
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)
    # two chunks: the second starts from the env state the first ended in
    k = 20
    a = acting.eval_metrics(st, Transition(None, None, tr.reward[:k], tr.discount[:k], None), rep)
    steps, done = steps0.copy(), done0.copy()
    for t in range(k):
        steps = np.where(done != 0, 0, steps) + rep
        done = 1 - discount[t]
    st2 = EnvState(obs=None, reward=None, done=_dev(done.astype(np.float32), dev), system_params=None,
                   info={"steps": _dev(steps.astype(np.float32), dev)})
    b = acting.eval_metrics(st2, Transition(None, None, tr.reward[k:], tr.discount[k:], None), rep, carry=a)
    for g, w in zip(b, want):
        assert np.array_equal(g.cpu().numpy(), w)


def test_evaluator_runs_one_episode_per_env(mb, cuda_device):
    """sac/acting.py:82-151: reset from the true buffer, one deterministic episode per eval env, EvalWrapper's metrics.
    Checked against the oracle's actor rollout teacher-forced on the GPU's observations and its fold of the rewards."""
    from mbpo_b200 import acting, envs
    from mbpo_b200.replay_buffers import UniformSamplingQueue
    from mbpo_b200.systems import BraxWrapper, PendulumSystem
    from mbpo_b200.utils.optimizer_utils import Transition
    dev = cuda_device
    system = PendulumSystem()
    q = UniformSamplingQueue(10, _true_dummy(mb, dev), 1)
    first = system.reset(device=dev)
    st = q.insert(q.init(_dev(ojr.PRNGKey(0), dev)),
                  Transition(first.x_next[None], torch.zeros((1, 1), device=dev), first.reward[None],
                             torch.full((1,), 0.99, device=dev), first.x_next[None]))
    L, rep, E = 50, 1, 64
    env = envs.wrap(BraxWrapper(system, system.init_params(_dev(ojr.PRNGKey(1), dev)), st, q), L, rep)
    pol = orc.make_policy_params(seed=7)
    make_policy = acting.make_inference_fn()
    import functools
    params = acting.PolicyParams([_dev(w, dev) for w in pol.weights], [_dev(b, dev) for b in pol.biases])
    ev = acting.Evaluator(env, functools.partial(make_policy, deterministic=True), num_eval_envs=E, episode_length=L,
                          action_repeat=rep, key=_dev(ojr.PRNGKey(4), dev))
    key = _dev(ojr.PRNGKey(6), dev)
    state = ev._generate_eval_unroll(params, key)
    em = state.info["eval_metrics"]
    assert torch.all(em["active_episodes"] == 0) and torch.all(em["episode_steps"] == L)
    # the same unroll through the public pieces
    first_state = env.reset(mb.random.split(key, E))
    nst, tr = acting.generate_unroll(env, first_state, make_policy(params, deterministic=True), key, L // rep)
    want = obr.eval_metrics(tr.reward.cpu().numpy(), tr.discount.cpu().numpy(), np.zeros(E, np.float32),
                            np.zeros(E, np.float32), rep)
    assert np.array_equal(em["episode_metrics"]["reward"].cpu().numpy(), want[0])
    x0 = np.tile(np.array([-1, 0, 0], np.float32), (E, 1))
    o, _ = orc.actor_rollout(pol, x0, ojr.PRNGKey(6), L, L, key_convention="unroll", deterministic=True,
                             teacher_obs=tr.observation.cpu().numpy())
    np.testing.assert_allclose(tr.reward.cpu().numpy(), o["reward"], rtol=1e-5, atol=3e-6)
    metrics = ev.run_evaluation(params, {"training/sps": 1.0}, unroll_key=key)
    assert metrics["eval/episode_reward"] == np.mean(want[0]) and metrics["eval/avg_episode_length"] == L
    assert metrics["training/sps"] == 1.0 and metrics["eval/sps"] > 0 and metrics["eval/walltime"] > 0
    m2 = ev.run_evaluation(params, {}, aggregate_episodes=False)       # draws its own key
    assert m2["eval/episode_reward"].shape == (E,)


@pytest.mark.parametrize("X", [3, 1, 17])
def test_running_statistics_update_and_normalize(mb, cuda_device, X):
    """running_statistics.update / normalize (sac.py:298-301) against the oracle's restatement of the update as
    written (float64 sums); tolerance rel 1e-5 (XLA's reduction order is unspecified).  The kernel is deterministic:
    the same batch gives the same bits."""
    from mbpo_b200 import running_statistics as rs
    dev = cuda_device
    rng = np.random.default_rng(X)
    st, ost = rs.init_state(X, dev), obr.running_statistics_init(X)
    for shape in ((1,), (7,), (20, 32), (3, 1000, 5)):
        batch = (rng.standard_normal(shape + (X,)) * rng.uniform(0.1, 8, X) + rng.uniform(-2, 2, X)).astype(np.float32)
        new = rs.update(st, _dev(batch, dev))
        again = rs.update(st, _dev(batch, dev))
        ost = obr.running_statistics_update(ost, batch, accumulate=np.float64)
        for k in ("count", "mean", "summed_variance", "std"):
            np.testing.assert_allclose(getattr(new, k).cpu().numpy(), ost[k], rtol=1e-5, atol=1e-6)
            assert torch.equal(getattr(new, k), getattr(again, k))
        st = new
    batch = rng.standard_normal((100, X)).astype(np.float32)
    got = rs.normalize(_dev(batch, dev), st).cpu().numpy()
    np.testing.assert_allclose(got, (batch - ost["mean"]) / ost["std"], rtol=1e-5, atol=1e-6)
    got = rs.normalize(_dev(batch, dev), st, max_abs_value=0.5).cpu().numpy()
    assert np.abs(got).max() <= 0.5


def test_running_statistics_at_full_size_and_sharded(mb, cuda_device):
    """65,536 envs x 200 observations: the update equals float64 mean / std of the rows (size-independent property),
    and shard sums added together (the all-reduce of the multi-GPU path) give the unsharded statistics."""
    from mbpo_b200 import _lib as L, running_statistics as rs
    dev = cuda_device
    gen = torch.Generator(device=dev).manual_seed(0)
    obs = torch.randn((200, 65536, 3), generator=gen, device=dev) * torch.tensor([1.0, 0.5, 8.0], device=dev) + 0.25
    st = rs.update(rs.init_state(3, dev), obs)
    flat = obs.reshape(-1, 3).double()
    np.testing.assert_allclose(st.mean.cpu().numpy(), flat.mean(0).cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st.std.cpu().numpy(), flat.std(0, unbiased=False).cpu().numpy(), rtol=1e-5)
    assert float(st.count) == 200 * 65536
    # two "ranks": accumulate each half, add the sums (what all_reduce_sums does), finalize once
    st0 = rs.init_state(3, dev)
    ws_bytes = L.lib.mbpo_running_statistics_workspace_bytes(3)
    ws = torch.empty(ws_bytes // 8, dtype=torch.float64, device=dev)
    total = torch.zeros(7, dtype=torch.float64, device=dev)
    for part in (obs[:, :30000].contiguous(), obs[:, 30000:].contiguous()):
        sums = torch.empty(7, dtype=torch.float64, device=dev)
        L.check(L.lib.mbpo_running_statistics_accumulate(L.ptr(part), part.numel() // 3, 3, L.ptr(st0.mean), L.ptr(ws),
                                                         ws_bytes, L.ptr(sums), L.stream_ptr(dev)))
        total += sums
    out = [torch.empty(1, device=dev)] + [torch.empty(3, device=dev) for _ in range(3)]
    L.check(L.lib.mbpo_running_statistics_finalize(L.ptr(total), 3, L.ptr(st0.count.reshape(1)), L.ptr(st0.mean),
                                                   L.ptr(st0.summed_variance), 1e-6, 1e6, *[L.ptr(o) for o in out],
                                                   L.stream_ptr(dev)))
    assert float(out[0]) == float(st.count)
    np.testing.assert_allclose(out[1].cpu().numpy(), st.mean.cpu().numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(out[3].cpu().numpy(), st.std.cpu().numpy(), rtol=1e-6)


def test_get_experience_in_full(mb, cuda_device):
    """sac.py:283-304: actor steps -> running_statistics.update(transitions.observation) -> replay_buffer.insert, twice
    (the second collection runs the policy on observations normalised with the first one's statistics)."""
    from mbpo_b200 import acting, envs, running_statistics as rs
    from mbpo_b200.replay_buffers import UniformSamplingQueue
    from mbpo_b200.systems import BraxWrapper, PendulumSystem
    from mbpo_b200.utils.optimizer_utils import Transition
    dev = cuda_device
    system = PendulumSystem()
    tq = UniformSamplingQueue(10, _true_dummy(mb, dev), 1)
    first = system.reset(device=dev)
    tst = tq.insert(tq.init(_dev(ojr.PRNGKey(0), dev)),
                    Transition(first.x_next[None], torch.zeros((1, 1), device=dev), first.reward[None],
                               torch.full((1,), 0.99, device=dev), first.x_next[None]))
    E, T, L = 256, 20, 200
    env = envs.wrap(BraxWrapper(system, system.init_params(_dev(ojr.PRNGKey(1), dev)), tst, tq), L, 1)
    pol = orc.make_policy_params(seed=7)
    params = acting.PolicyParams([_dev(w, dev) for w in pol.weights], [_dev(b, dev) for b in pol.biases])
    q = UniformSamplingQueue(2 ** 14, _sac_dummy(mb, dev), 64)
    col = acting.ExperienceCollector(env, acting.make_normalized_inference_fn(), q, T)
    norm, buf = rs.init_state(3, dev), q.init(_dev(ojr.PRNGKey(2), dev))
    state = env.reset(mb.random.split(_dev(ojr.PRNGKey(3), dev), E))
    key = _dev(ojr.PRNGKey(4), dev)
    onorm = obr.running_statistics_init(3)
    okey = ojr.PRNGKey(4)
    rows = []
    for it in range(2):
        prev_norm, prev_state = norm, state
        norm, state, buf = col.get_experience(norm, params, state, buf, key)      # sac.py:283-304: a 3-tuple
        key = col.last_key
        live = buf.data[buf.insert_position - T * E:buf.insert_position].cpu().numpy()
        obs = live[:, :3].reshape(T, E, 3)
        # the policy saw (obs - mean) / std of the statistics before this collection (the GPU's own: after 20 steps
        # from the hanging state std(cos) is ~1e-3, so a 1e-7 difference in the mean is 1e-4 after normalisation;
        # the statistics themselves are compared below)
        m, sd = prev_norm.mean.cpu().numpy(), prev_norm.std.cpu().numpy()
        want, okey = orc.actor_rollout(pol, prev_state.obs.cpu().numpy(), okey, T, L, key_convention="sac",
                                       teacher_obs=((obs - m) / sd).astype(np.float32))
        np.testing.assert_allclose(live[:, 3].reshape(T, E), want["action"][..., 0], rtol=5e-5, atol=5e-6)
        assert np.array_equal(key.cpu().numpy(), okey)
        onorm = obr.running_statistics_update(onorm, obs, accumulate=np.float64)
        for k in ("count", "mean", "std"):
            np.testing.assert_allclose(getattr(norm, k).cpu().numpy(), onorm[k], rtol=1e-5, atol=1e-6)
        rows.append(live)
    assert buf.insert_position == 2 * T * E and float(norm.count) == 2 * T * E
    assert np.array_equal(buf.data[:T * E].cpu().numpy(), rows[0])
    # observation[t+1] = next_observation[t] inside a collection, and the second starts where the first ended
    r0, r1 = rows[0].reshape(T, E, 10), rows[1].reshape(T, E, 10)
    assert np.array_equal(r0[1:, :, :3], r0[:-1, :, 6:9]) and np.array_equal(r1[0, :, :3], r0[-1, :, 6:9])


def test_guard_zones_around_every_output(mb, cuda_device):
    """No out-of-bounds writes (compute-sanitizer is not available on the pool): every output buffer of the replay /
    reset / eval / statistics entry points sits between sentinel-filled guard zones, for ragged sizes and for ring
    positions that make a tile cross the ring's end and start on every alignment modulo 16 bytes."""
    L = mb._lib
    dev = cuda_device
    G, SENT = 64, -12345.0

    def guarded(n, dtype=torch.float32):
        full = torch.full((n + 2 * G,), SENT, dtype=dtype, device=dev)
        return full, full[G:G + n]

    def intact(full, n):
        ref = torch.tensor(SENT, device=dev).to(full.dtype)
        return bool(torch.all(full[:G] == ref)) and bool(torch.all(full[G + n:] == ref))

    rng = np.random.default_rng(0)
    for R, D, widths in ((101, 10, (3, 1, 1, 1, 3, 1)), (64, 9, (3, 1, 1, 1, 3)), (37, 5, (5,)), (1000, 23, (20, 3))):
        full, ring = guarded(R * D)
        ring.zero_()
        st = L.ReplayStateC(data=ring.data_ptr(), capacity=R, row_width=D)
        logical = np.zeros((R, D), np.float32)
        oq = obr.UniformSamplingQueue(R, D, 1)
        ost = oq.init(ojr.PRNGKey(0))
        for n in (1, 3, R // 2, R - 1, 7, R, 2, 33 % R + 1):
            rows = rng.standard_normal((n, D)).astype(np.float32)
            fields = L.ReplayFieldsC(num_fields=len(widths))
            keep, col = [], 0
            for f, w in enumerate(widths):
                t = _dev(rows[:, col:col + w].copy(), dev)
                keep.append(t)
                fields.width[f], fields.ptr[f] = w, t.data_ptr()
                col += w
            L.check(L.lib.mbpo_replay_insert(L.C.byref(st), L.C.byref(fields), n, L.stream_ptr(dev)))
            ost = oq.insert(ost, rows)
            assert intact(full, R * D)
            out_full, out = guarded(R * D)
            L.check(L.lib.mbpo_replay_read(L.C.byref(st), 0, R, out.data_ptr(), L.stream_ptr(dev)))
            assert intact(out_full, R * D)
            live = ost.insert_position
            assert np.array_equal(out.cpu().numpy().reshape(R, D)[:live], ost.data[:live])
            assert (st.insert_position, st.sample_position) == (ost.insert_position, ost.sample_position)
        # sample / reset outputs
        key = _dev(ojr.PRNGKey(1), dev)
        for batch in (1, 5, 129):
            kf, k = guarded(2)
            idf, idx = guarded(batch)
            bf, b = guarded(batch * D)
            L.check(L.lib.mbpo_replay_sample(L.C.byref(st), key.data_ptr(), 0, batch, k.view(torch.uint32).data_ptr(),
                                             idx.view(torch.int32).data_ptr(), b.data_ptr(), L.stream_ptr(dev)))
            assert intact(kf, 2) and intact(idf, batch) and intact(bf, batch * D)
        for E in (1, 31, 130):
            rngs = _dev(ojr.split(ojr.PRNGKey(2), E), dev)
            of, o = guarded(E * 3)
            rf, r = guarded(E)
            kf, k = guarded(E * 2)
            L.check(L.lib.mbpo_env_reset_from_buffer(L.C.byref(st), rngs.data_ptr(), E, 0, 1, min(3, D), D - 1,
                                                     o.data_ptr(), r.data_ptr(), k.view(torch.uint32).data_ptr(), None,
                                                     L.stream_ptr(dev)))
            assert intact(of, E * 3) and intact(rf, E) and intact(kf, E * 2)
    # eval metrics and running statistics
    for E, T in ((1, 1), (33, 7), (257, 3)):
        r, d = torch.randn((T, E), device=dev), torch.ones((T, E), device=dev)
        z = torch.zeros(E, device=dev)
        outs = [guarded(E) for _ in range(3)]
        for _, v in outs:
            v.fill_(1.0)
        L.check(L.lib.mbpo_eval_metrics(r.data_ptr(), d.data_ptr(), z.data_ptr(), z.data_ptr(), 1, E, T, E, 1,
                                        outs[0][1].data_ptr(), outs[1][1].data_ptr(), outs[2][1].data_ptr(),
                                        L.stream_ptr(dev)))
        assert all(intact(f, E) for f, _ in outs)
    for X, n in ((1, 1), (3, 1000), (17, 33), (64, 5)):
        batch = torch.randn((n, X), device=dev)
        ws_bytes = L.lib.mbpo_running_statistics_workspace_bytes(X)
        wf, ws = guarded(ws_bytes // 8, torch.float64)
        sf, sums = guarded(2 * X + 1, torch.float64)
        mean = torch.zeros(X, device=dev)
        L.check(L.lib.mbpo_running_statistics_accumulate(batch.data_ptr(), n, X, mean.data_ptr(), ws.data_ptr(), ws_bytes,
                                                         sums.data_ptr(), L.stream_ptr(dev)))
        assert intact(wf, ws_bytes // 8) and intact(sf, 2 * X + 1)
        assert float(sums[2 * X]) == n
        np.testing.assert_allclose(sums[:X].cpu().numpy(), batch.double().sum(0).cpu().numpy(), rtol=1e-12, atol=1e-9)
        outs = [guarded(1)] + [guarded(X) for _ in range(3)]
        cnt, sv = torch.zeros(1, device=dev), torch.zeros(X, device=dev)
        L.check(L.lib.mbpo_running_statistics_finalize(sums.data_ptr(), X, cnt.data_ptr(), mean.data_ptr(), sv.data_ptr(),
                                                       1e-6, 1e6, *[v.data_ptr() for _, v in outs], L.stream_ptr(dev)))
        assert intact(outs[0][0], 1) and all(intact(f, X) for f, _ in outs[1:])
        nf, nout = guarded(n * X)
        L.check(L.lib.mbpo_running_statistics_normalize(batch.data_ptr(), n, X, outs[1][1].data_ptr(), outs[3][1].data_ptr(),
                                                        0.0, nout.data_ptr(), L.stream_ptr(dev)))
        assert intact(nf, n * X)


def test_bptt_normalizer_and_take(mb, cuda_device):
    """bptt_optimizer.py:31-75 Normalizer (update / normalize / inverse), :297-303 update_normalizers on state and
    reward, and :447-450 randint + take(mode='wrap') of evaluation rows from the true buffer."""
    from mbpo_b200.replay_buffers import UniformSamplingQueue
    from mbpo_b200.running_statistics import Normalizer
    from mbpo_b200.utils.optimizer_utils import Transition
    dev = cuda_device
    rng = np.random.default_rng(0)
    for X in (3, 1):
        nz = Normalizer((X,), dev)
        st, ost = nz.initialize_normalizer_state(), obr.normalizer_init(X)
        for n in (1, 10, 2000, 65536):
            x = (rng.standard_normal((n, X)) * rng.uniform(0.1, 8, X) + rng.uniform(-2, 2, X)).astype(np.float32)
            st, ost = nz.update(_dev(x, dev), st), obr.normalizer_update(x, ost)
            assert float(st.size) == ost["size"]
            np.testing.assert_allclose(st.mean.cpu().numpy(), ost["mean"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(st.std.cpu().numpy(), ost["std"], rtol=1e-5, atol=1e-6)
        x = rng.standard_normal((50, X)).astype(np.float32)
        z = nz.normalize(_dev(x, dev), st)
        np.testing.assert_allclose(z.cpu().numpy(), (x - ost["mean"]) / ost["std"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(nz.inverse(z, st).cpu().numpy(), x, rtol=1e-5, atol=1e-5)
    const = nz.update(torch.full((8, 1), 2.5, device=dev), nz.initialize_normalizer_state())
    assert float(const.std) == np.float32(1e-8) and float(const.mean) == 2.5            # std floor EPS
    # evaluation rows: randint over the live range, take with wrap
    q = UniformSamplingQueue(50, _true_dummy(mb, dev), 10)
    g = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32)).to(dev)
    st = q.init(_dev(ojr.PRNGKey(0), dev))
    for n in (30, 30):                                                                   # head != 0
        st = q.insert(st, Transition(g(n, 3), g(n, 1), g(n), g(n), g(n, 3)))
    eval_rng = ojr.PRNGKey(5)
    idx = mb.random.randint(_dev(eval_rng, dev), 100, st.sample_position, st.insert_position)
    assert np.array_equal(idx.cpu().numpy(), ojr.randint(eval_rng, 100, 0, 50))
    rows = q.take(st, idx)
    data = st.data.cpu().numpy()
    assert np.array_equal(rows.observation.cpu().numpy(), data[idx.cpu().numpy(), :3])
    wrapped = q.take(st, idx + 50 * 3)                                                    # mode='wrap'
    assert torch.equal(wrapped.observation, rows.observation)
    neg = q.take(st, idx - 50)
    assert torch.equal(neg.reward, rows.reward)


def test_graphed_rollout_replays_the_same_bits(mb, cuda_device):
    """The collection call captured in a CUDA graph (launch-bound regime: 32 envs x 20 steps, tests/test_sac.py's
    shape): three replays equal three get_experience calls chained by hand, bit for bit."""
    from mbpo_b200 import acting
    from mbpo_b200.envs import wrap
    from mbpo_b200.systems import PendulumSystem
    dev = cuda_device
    E, T, L = 32, 20, 50
    pol = orc.make_policy_params(seed=3)
    policy = acting.Policy(acting.PolicyParams([_dev(w, dev) for w in pol.weights], [_dev(b, dev) for b in pol.biases]))
    system = PendulumSystem()
    env = wrap(system, system.reset(device=dev).system_params, episode_length=L)
    rng = np.random.default_rng(0)
    th, w = rng.uniform(-np.pi, np.pi, E), rng.uniform(-8, 8, E)
    x0 = _dev(np.stack([np.cos(th), np.sin(th), w], -1).astype(np.float32), dev)
    key = _dev(ojr.PRNGKey(9), dev)
    st = env.reset(x0)
    collect = acting.GraphedRollout(env, st, policy, key, T)
    k, s = key, st
    for it in range(3):
        k, s, tr = acting.get_experience(env, s, policy, k, T)
        gk, gs, gtr = collect()
        assert torch.equal(gk, k) and torch.equal(gs.obs, s.obs) and torch.equal(gs.done, s.done)
        assert torch.equal(gs.info["steps"], s.info["steps"])
        for a, b in ((gtr.observation, tr.observation), (gtr.action, tr.action), (gtr.reward, tr.reward),
                     (gtr.discount, tr.discount), (gtr.next_observation, tr.next_observation),
                     (gtr.extras["state_extras"]["truncation"], tr.extras["state_extras"]["truncation"])):
            assert torch.equal(a, b), it


@pytest.mark.parametrize("T,B", [(1, 1), (5, 33), (20, 2048), (200, 300)])
def test_compute_gae_bit_exact(mb, cuda_device, T, B):
    """ppo/losses.py:128-184 on the Transition of an unroll: elementwise float32 operations in the reference's order,
    so the scan is bit-exact against the oracle."""
    from mbpo_b200.utils import compute_gae
    rng = np.random.default_rng(T * 1000 + B)
    r = rng.standard_normal((T, B)).astype(np.float32)
    v = rng.standard_normal((T, B)).astype(np.float32) * 3
    boot = rng.standard_normal(B).astype(np.float32)
    trunc = (rng.random((T, B)) < 0.05).astype(np.float32)
    done = np.maximum(trunc, (rng.random((T, B)) < 0.05).astype(np.float32))
    discount_field = 1 - done                                      # Transition.discount = 1 - done
    term = ((1 - discount_field) * (1 - trunc)).astype(np.float32)  # losses.py:89
    for lam, d in ((0.95, 0.99), (1.0, 0.9), (0.0, 0.99)):
        vs, adv = compute_gae(*[_dev(x, cuda_device) for x in (trunc, term, r, v, boot)], lambda_=lam, discount=d)
        want_vs, want_adv = obr.compute_gae(trunc, term, r, v, boot, lam, d)
        assert np.array_equal(vs.cpu().numpy(), want_vs) and np.array_equal(adv.cpu().numpy(), want_adv)


def test_kernels_reproduce_the_replay_golden_vectors(mb, cuda_device, prng_mode):
    """The committed fixture tests/golden/replay_golden.npz through the CUDA path: randint words, the queue's insert /
    sample history with the ring wrapping, BraxWrapper.reset draws (bit for bit), the two normalisers (rel 1e-5), the
    evaluation metrics and PPO's GAE (bit for bit)."""
    import os
    from mbpo_b200 import acting, running_statistics as rs
    from mbpo_b200.envs import EnvState
    from mbpo_b200.replay_buffers import UniformSamplingQueue
    from mbpo_b200.systems import BraxWrapper, PendulumSystem
    from mbpo_b200.utils import compute_gae
    from mbpo_b200.utils.optimizer_utils import Transition
    dev = cuda_device
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "replay_golden.npz"))
    tag = "part" if prng_mode else "legacy"
    for s, (lo, hi) in enumerate([(0, 10), (-5, 5), (0, 65537), (3, 3)]):
        got = mb.random.randint(_dev(ojr.PRNGKey(s), dev), 16, lo, hi)
        assert np.array_equal(got.cpu().numpy(), g["randint_%s" % tag][s])
    q = UniformSamplingQueue(32, _sac_dummy(mb, dev), 8)
    st = q.init(_dev(ojr.PRNGKey(7), dev))
    for k in range(5):
        r = _dev(g["rows_%d" % k], dev)
        tr = Transition(r[:, 0:3], r[:, 3:4], r[:, 4], r[:, 5], r[:, 6:9],
                        {"state_extras": {"truncation": r[:, 9]}, "policy_extras": {}})
        st = q.insert(st, tr)
        st, batch, idx = q.sample_with_indices(st)
        assert np.array_equal(idx.cpu().numpy(), g["q_%s_%d_idx" % (tag, k)])
        assert np.array_equal(_rows_of(batch, 8), g["q_%s_%d_batch" % (tag, k)])
        assert np.array_equal(st.key.cpu().numpy(), g["q_%s_%d_key" % (tag, k)])
        assert [st.insert_position, st.sample_position] == g["q_%s_%d_positions" % (tag, k)].tolist()
    assert np.array_equal(st.data.cpu().numpy(), g["q_%s_data" % tag])
    system = PendulumSystem()
    one = UniformSamplingQueue(32, _sac_dummy(mb, dev), 1)
    env = BraxWrapper(system, system.init_params(_dev(ojr.PRNGKey(1), dev)), st, one)
    state = env.reset(_dev(g["reset_%s_rngs" % tag], dev))
    assert np.array_equal(state.obs.cpu().numpy(), g["reset_%s_obs" % tag])
    assert np.array_equal(state.reward.cpu().numpy(), g["reset_%s_reward" % tag])
    assert np.array_equal(state.system_params.key.cpu().numpy(), g["reset_%s_keys" % tag])
    stats, nz = rs.init_state(3, dev), rs.Normalizer((3,), dev)
    norm = nz.initialize_normalizer_state()
    for k in range(6):
        o = _dev(g["obs"][k], dev)
        stats, norm = rs.update(stats, o), nz.update(o, norm)
        got = torch.cat([stats.count.reshape(1), stats.mean, stats.summed_variance, stats.std]).cpu().numpy()
        np.testing.assert_allclose(got, g["stats_%d" % k], rtol=1e-5, atol=1e-6)
        got = torch.cat([norm.size.reshape(1).float(), norm.mean, norm.std]).cpu().numpy()
        np.testing.assert_allclose(got, g["norm_%d" % k], rtol=1e-5, atol=1e-6)
    vs, adv = compute_gae(*[_dev(g["gae_" + k], dev) for k in ("truncation", "termination", "reward", "values",
                                                                "bootstrap")], lambda_=0.95, discount=0.99)
    assert np.array_equal(vs.cpu().numpy(), g["gae_vs"]) and np.array_equal(adv.cpu().numpy(), g["gae_advantages"])
    z = torch.zeros(9, device=dev)
    st0 = EnvState(obs=None, reward=None, done=z, system_params=None, info={"steps": z})
    er, es, ea = acting.eval_metrics(st0, Transition(None, None, _dev(g["gae_reward"], dev), _dev(g["gae_discount"], dev),
                                                     None), 1)
    assert np.array_equal(er.cpu().numpy(), g["eval_reward"]) and np.array_equal(es.cpu().numpy(), g["eval_steps"])
    assert np.array_equal(ea.cpu().numpy(), g["eval_active"])
