"""CPU tests of the replay queue's host logic (no kernels run): the row layout follows jax's pytree flattening order
(ravel_pytree of the dummy Transition: NamedTuple fields in order, dict keys sorted, empty containers dropped) and the
column views handed back by ``unflatten`` have the dummy sample's shapes."""
import torch

from mbpo_b200.replay_buffers import UniformSamplingQueue, _leaves, _unflatten
from mbpo_b200.utils.optimizer_utils import Transition


def test_leaf_order_is_jax_tree_order():
    z = torch.zeros
    tree = Transition(observation=z(3), action=z(2), reward=z(()), discount=z(()), next_observation=z(3),
                      extras={"state_extras": {"truncation": z(()), "a_first": z(4)}, "policy_extras": {}})
    shapes = [tuple(l.shape) for l in _leaves(tree)]
    # observation, action, reward, discount, next_observation, then extras: 'policy_extras' (empty) < 'state_extras',
    # inside it 'a_first' < 'truncation'
    assert shapes == [(3,), (2,), (), (), (3,), (4,), ()]
    assert _leaves(Transition(z(3), z(1), z(()), z(()), z(3))) and len(_leaves(Transition(z(3), z(1), z(()), z(()), z(3)))) == 5
    assert _leaves({"b": 1, "a": (2, None, [3])}) == [2, 3, 1]


def test_row_layout_and_unflatten_views():
    z = torch.zeros
    sac = Transition(z(3), z(1), z(()), z(()), z(3), {"state_extras": {"truncation": z(())}, "policy_extras": {}})
    q = UniformSamplingQueue(16, sac, 4)
    assert q.row_width == 10 and [q.column_of(i) for i in range(6)] == [0, 3, 4, 5, 6, 9]
    rows = torch.arange(40, dtype=torch.float32).reshape(4, 10)
    tr = q.unflatten(rows)
    assert isinstance(tr, Transition) and tr.observation.shape == (4, 3) and tr.reward.shape == (4,)
    assert torch.equal(tr.next_observation, rows[:, 6:9]) and torch.equal(tr.action, rows[:, 3:4])
    assert torch.equal(tr.extras["state_extras"]["truncation"], rows[:, 9]) and tr.extras["policy_extras"] == {}
    assert list(tr.extras) == ["state_extras", "policy_extras"]            # the caller's key order is kept
    true = Transition(z(3), z(1), z(1), z(1), z(3))                         # base_optimizer.py:44-50: reward / discount [1]
    q2 = UniformSamplingQueue(10, true, 1)
    assert q2.row_width == 9 and q2.unflatten(torch.zeros(2, 9)).reward.shape == (2, 1)
    leaves = [1, 2, 3]
    assert _unflatten({"b": 0, "a": (0, 0)}, leaves) == {"b": 3, "a": (1, 2)}


def test_queue_size_and_too_many_leaves():
    import pytest
    from mbpo_b200 import MbpoUnsupported
    from mbpo_b200.replay_buffers import ReplayBufferState
    z = torch.zeros
    q = UniformSamplingQueue(16, Transition(z(3), z(1), z(()), z(()), z(3)), 4)
    assert q.size(ReplayBufferState(ring=None, head=3, insert_position=12, sample_position=2)) == 10
    with pytest.raises(MbpoUnsupported):
        UniformSamplingQueue(16, tuple(z(1) for _ in range(9)), 4)
