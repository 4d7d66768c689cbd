"""ORACLE (test infrastructure, NOT product code) -- NumPy restatement of brax's UniformSamplingQueue
and of the reference's BraxWrapper.reset that draws initial states from it.

brax is a third-party dependency absent from /root/reference (setup.py:19, unpinned) and not installable
here; this file restates the published algorithm of ``brax/training/replay_buffers.py`` (brax 0.9-0.10,
``QueueBase.init / insert_internal``, ``UniformSamplingQueue.sample_internal``):

  * data: float32 [max_replay_size, D], a row = ``ravel_pytree`` of one Transition (fields in NamedTuple
    order, dict keys sorted), ``insert_position`` / ``sample_position`` int32, ``key``;
  * insert: ``roll = min(0, len(data) - position - len(update))``; if roll: ``data = roll(data, roll, axis=0)``;
    ``position += roll``; write the update at ``position``; ``position = (position + len(update)) %
    (len(data) + 1)``; ``sample_position = max(0, sample_position + roll)``;
  * sample: ``key, sample_key = split(key)``; ``idx = randint(sample_key, (batch,), sample_position,
    insert_position)``; ``batch = take(data, idx, axis=0, mode='wrap')``.

Reference call sites: mbpo/systems/brax_wrapper.py:25-38 (reset), mbpo/optimizers/base_optimizer.py:44-57
(dummy buffer), mbpo/optimizers/policy_optimizers/sac/sac.py:202-205,303 (SAC's replay buffer and insert),
tests/test_sac.py:15-28.  PARITY UNPINNED against a brax / JAX run; the randint arithmetic is restated in
oracle/jax_prng.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import numpy as np

from . import jax_prng as jp


@dataclass
class ReplayBufferState:
    data: np.ndarray            # float32 [R, D]
    insert_position: int
    sample_position: int
    key: np.ndarray             # uint32 [2]


class UniformSamplingQueue:
    def __init__(self, max_replay_size: int, row_width: int, sample_batch_size: int, partitionable: bool = False):
        self.R, self.D, self.batch = int(max_replay_size), int(row_width), int(sample_batch_size)
        self.partitionable = partitionable

    def init(self, key) -> ReplayBufferState:
        return ReplayBufferState(np.zeros((self.R, self.D), np.float32), 0, 0, np.asarray(key, np.uint32).copy())

    def insert(self, st: ReplayBufferState, rows: np.ndarray) -> ReplayBufferState:
        rows = np.asarray(rows, np.float32).reshape(-1, self.D)
        n = rows.shape[0]
        if n > self.R:
            raise ValueError("Trying to insert a batch of samples larger than the maximum replay size")
        data = st.data
        position = st.insert_position
        roll = min(0, self.R - position - n)
        if roll:
            data = np.roll(data, roll, axis=0)
        else:
            data = data.copy()
        position = position + roll
        data[position:position + n] = rows
        position = (position + n) % (self.R + 1)
        sample_position = max(0, st.sample_position + roll)
        return ReplayBufferState(data, position, sample_position, st.key)

    def sample(self, st: ReplayBufferState):
        k = jp.split(st.key, 2, self.partitionable)
        idx = jp.randint(k[1], self.batch, st.sample_position, st.insert_position, self.partitionable)
        batch = st.data[np.mod(idx.astype(np.int64), self.R)]
        return replace(st, key=k[0]), batch, idx


def brax_wrapper_reset(rngs: np.ndarray, queue: UniformSamplingQueue, st: ReplayBufferState, x_dim: int,
                       action_dim: int):
    """vmap(BraxWrapper.reset)(rngs) (brax_wrapper.py:25-38; VmapWrapper.reset, brax_utils/training.py:66-69):
    per env ``keys = split(rng, 2)``, the buffer is sampled under ``keys[0]``, element 0 of the batch gives
    ``obs`` and ``reward``, ``system_params.key = keys[1]``, ``done = 0``."""
    rngs = np.asarray(rngs, np.uint32).reshape(-1, 2)
    E = rngs.shape[0]
    obs = np.zeros((E, x_dim), np.float32)
    reward = np.zeros(E, np.float32)
    sys_keys = np.zeros((E, 2), np.uint32)
    idx0 = np.zeros(E, np.int32)
    for e in range(E):
        keys = jp.split(rngs[e], 2, queue.partitionable)
        _, batch, idx = queue.sample(replace(st, key=keys[0]))
        obs[e] = batch[0, :x_dim]
        reward[e] = batch[0, x_dim + action_dim]
        sys_keys[e] = keys[1]
        idx0[e] = idx[0]
    return obs, reward, sys_keys, idx0


def eval_metrics(reward, discount, steps_in, done_in, action_repeat: int):
    """EvalWrapper.step (brax_utils/training.py:172-199) folded over the [T, E] reward / discount streams of an unroll
    that started from (steps_in, done_in), from EvalWrapper.reset's zeros (:159-170).  float32, step order.
    -> (episode_reward [E], episode_steps [E], active_episodes [E])."""
    F = np.float32
    reward, discount = np.asarray(reward, F), np.asarray(discount, F)
    steps, done = np.asarray(steps_in, F).copy(), np.asarray(done_in, F).copy()
    E = steps.shape[0]
    ep_reward, ep_steps, active = np.zeros(E, F), np.zeros(E, F), np.ones(E, F)
    for t in range(reward.shape[0]):
        steps = (np.where(done != 0, F(0), steps) + F(action_repeat)).astype(F)     # :98,120-124
        ep_steps = np.where(active != 0, steps, ep_steps)                           # :181-185
        ep_reward = (ep_reward + (reward[t] * active).astype(F)).astype(F)          # :186-190
        active = (active * discount[t]).astype(F)                                   # :191, discount = 1 - done
        done = (F(1) - discount[t]).astype(F)
    return ep_reward, ep_steps, active


# ---- brax.training.acme.running_statistics (third party, absent; restated from the published algorithm) --------------
def running_statistics_init(size: int):
    F = np.float32
    return dict(count=F(0), mean=np.zeros(size, F), summed_variance=np.zeros(size, F), std=np.ones(size, F))


def running_statistics_update(state, batch, std_min_value=1e-6, std_max_value=1e6, accumulate=np.float32):
    """running_statistics.update as written (call site sac/sac.py:298-301): count += n; diff_to_old_mean = batch - mean;
    mean += sum(diff_to_old_mean) / count; summed_variance += sum(diff_to_old_mean * (batch - new_mean));
    std = clip(sqrt(max(summed_variance, 0) / count)).  ``accumulate`` is the dtype of the two sums (XLA's reduction
    order is unspecified: float32 pairwise here, float64 for a tight check)."""
    F = np.float32
    X = state["mean"].shape[0]
    batch = np.asarray(batch, F).reshape(-1, X)
    count = F(state["count"] + F(batch.shape[0]))
    d_old = (batch - state["mean"]).astype(F)
    mean_update = (d_old.sum(0, dtype=accumulate) / accumulate(count)).astype(F)
    mean = (state["mean"] + mean_update).astype(F)
    d_new = (batch - mean).astype(F)
    var_update = (d_old * d_new).astype(F).sum(0, dtype=accumulate).astype(F)
    sv = (state["summed_variance"] + var_update).astype(F)
    std = np.clip(np.sqrt(np.maximum(sv, F(0)) / count).astype(F), F(std_min_value), F(std_max_value))
    return dict(count=count, mean=mean, summed_variance=sv, std=std)


def running_statistics_sums(batch, mean):
    """What mbpo_running_statistics_accumulate returns for one rank: float64 [2X] = (sum d, sum d*d), d in float32."""
    X = mean.shape[0]
    d = (np.asarray(batch, np.float32).reshape(-1, X) - mean).astype(np.float32).astype(np.float64)
    return np.concatenate([d.sum(0), (d * d).sum(0)])


def running_statistics_finalize(state, sums, step_increment, std_min_value=1e-6, std_max_value=1e6):
    """mbpo_running_statistics_finalize: sum(d_old * d_new) = sum d*d - mean_update * sum d."""
    F = np.float32
    X = state["mean"].shape[0]
    count = F(state["count"] + F(step_increment))
    mean_update = sums[:X] / np.float64(count)
    mean = (state["mean"] + mean_update.astype(F)).astype(F)
    var_update = sums[X:] - mean_update * sums[:X]
    sv = (state["summed_variance"] + var_update.astype(F)).astype(F)
    std = np.clip(np.sqrt(np.maximum(sv, F(0)) / count).astype(F), F(std_min_value), F(std_max_value))
    return dict(count=count, mean=mean, summed_variance=sv, std=std)


# ---- BPTT's own Normalizer (reference code: mbpo/optimizers/policy_optimizers/bptt_optimizer.py:31-75) ---------------
def normalizer_init(size: int):
    return dict(mean=np.zeros(size, np.float32), std=np.ones(size, np.float32), size=0)


def normalizer_update(x, state, accumulate=np.float64):
    """Normalizer.update as written (:51-66).  ``accumulate``: dtype of the two sums."""
    F = np.float32
    X = state["mean"].shape[0]
    x = np.asarray(x, F).reshape(-1, X)
    new_size = x.shape[0]
    total = new_size + state["size"]
    new_mean = ((state["mean"] * F(state["size"]) + x.sum(0, dtype=accumulate).astype(F)) / F(total)).astype(F)
    new_s_n = (np.square(state["std"]) * F(state["size"])
               + np.square((x - new_mean).astype(F)).sum(0, dtype=accumulate).astype(F)
               + F(state["size"]) * np.square(state["mean"] - new_mean)).astype(F)
    new_std = np.sqrt(new_s_n / F(total)).astype(F)
    return dict(mean=new_mean, std=np.maximum(new_std, F(1e-8)), size=total)


def compute_gae(truncation, termination, rewards, values, bootstrap_value, lambda_=1.0, discount=0.99):
    """ppo/losses.py:128-184 compute_gae as written, float32 ([T, B] time-major; python floats weakly typed)."""
    F = np.float32
    truncation, termination, rewards, values = (np.asarray(x, F) for x in (truncation, termination, rewards, values))
    bootstrap_value = np.asarray(bootstrap_value, F)
    d, lam = F(discount), F(lambda_)
    truncation_mask = (F(1) - truncation).astype(F)
    values_t_plus_1 = np.concatenate([values[1:], bootstrap_value[None]], axis=0)
    dn = (d * (F(1) - termination).astype(F)).astype(F)
    deltas = (((rewards + (dn * values_t_plus_1).astype(F)).astype(F) - values).astype(F) * truncation_mask).astype(F)
    acc = np.zeros_like(bootstrap_value)
    out = np.empty_like(values)
    for t in range(values.shape[0] - 1, -1, -1):
        acc = (deltas[t] + ((((dn[t] * truncation_mask[t]).astype(F)) * lam).astype(F) * acc).astype(F)).astype(F)
        out[t] = acc
    vs = (out + values).astype(F)
    vs_t_plus_1 = np.concatenate([vs[1:], bootstrap_value[None]], axis=0)
    advantages = (((rewards + (dn * vs_t_plus_1).astype(F)).astype(F) - values).astype(F) * truncation_mask).astype(F)
    return vs, advantages
