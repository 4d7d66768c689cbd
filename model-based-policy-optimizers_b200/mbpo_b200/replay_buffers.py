"""brax.training.replay_buffers.UniformSamplingQueue work-alike on CUDA tensors.

The data format either side of the env rollouts: SAC appends every collected Transition to this queue
(mbpo/optimizers/policy_optimizers/sac/sac.py:202-205,303), the true buffer handed to the optimizers is one
(mbpo/optimizers/base_optimizer.py:44-57, tests/test_sac.py:15-28), and BraxWrapper.reset draws first observations
from it (mbpo/systems/brax_wrapper.py:25-38).  Same constructor, ``init`` / ``insert`` / ``sample`` / ``size`` and state
fields (``data``, ``insert_position``, ``sample_position``, ``key``) as brax.

Storage is a ring in HBM behind ``mbpo_replay_*`` (include/mbpo_b200.h): brax's ``jnp.roll`` of the whole buffer on
every insert into a full queue becomes a head offset.  Like a donated jit argument, the state passed to ``insert`` must
not be used afterwards: the returned state shares (and has overwritten part of) its storage.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Tuple

import torch

from . import _lib
from .config import config
from .systems.base_systems import _Replaceable


def _leaves(tree: Any) -> List[Any]:
    """jax.tree_util leaf order: NamedTuple / tuple / list fields in order, dict keys sorted, None dropped."""
    if tree is None:
        return []
    if isinstance(tree, dict):
        return [l for k in sorted(tree) for l in _leaves(tree[k])]
    if isinstance(tree, (tuple, list)):
        return [l for t in tree for l in _leaves(t)]
    return [tree]


def _unflatten(tree: Any, leaves: List[Any]) -> Any:
    if tree is None:
        return None
    if isinstance(tree, dict):
        out = {k: _unflatten(tree[k], leaves) for k in sorted(tree)}
        return {k: out[k] for k in tree}
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return type(tree)(*[_unflatten(t, leaves) for t in tree])
    if isinstance(tree, (tuple, list)):
        return type(tree)(_unflatten(t, leaves) for t in tree)
    return leaves.pop(0)


def _leaf_shape(leaf: Any) -> Tuple[int, ...]:
    return tuple(leaf.shape) if hasattr(leaf, "shape") else ()


@dataclass
class ReplayBufferState(_Replaceable):
    """brax ReplayBufferState.  ``ring`` holds the physical rows, ``head`` the physical row of ``data[0]``."""
    ring: torch.Tensor = None           # float32 [max_replay_size, D]
    head: int = 0
    insert_position: int = 0
    sample_position: int = 0
    key: torch.Tensor = None            # uint32 [2]

    def _c(self) -> _lib.ReplayStateC:
        return _lib.ReplayStateC(data=_lib.ptr(self.ring), capacity=self.ring.shape[0], row_width=self.ring.shape[1],
                                 reserved=0, head=self.head, insert_position=self.insert_position,
                                 sample_position=self.sample_position)

    @property
    def data(self) -> torch.Tensor:
        """The queue in brax's logical row order (a copy)."""
        out = torch.empty_like(self.ring)
        st = self._c()
        with _lib.cuda_guard(self.ring):
            _lib.check(_lib.lib.mbpo_replay_read(_lib.C.byref(st), 0, self.ring.shape[0], _lib.ptr(out),
                                                 _lib.stream_ptr(self.ring.device)))
        return out


class UniformSamplingQueue:
    """brax UniformSamplingQueue(max_replay_size, dummy_data_sample, sample_batch_size)."""

    def __init__(self, max_replay_size: int, dummy_data_sample: Any, sample_batch_size: int, device=None):
        self._dummy = dummy_data_sample
        self._shapes = [_leaf_shape(l) for l in _leaves(dummy_data_sample)]
        self._widths = [int(torch.Size(s).numel()) for s in self._shapes]
        if not 1 <= len(self._widths) <= _lib.MBPO_REPLAY_MAX_FIELDS:
            raise _lib.MbpoUnsupported(_lib.MBPO_EUNSUPPORTED, "a row holds 1..%d leaves, the sample has %d" % (
                _lib.MBPO_REPLAY_MAX_FIELDS, len(self._widths)))
        self._max_replay_size = int(max_replay_size)
        self._sample_batch_size = int(sample_batch_size)
        self._row_width = sum(self._widths)
        self._device = device

    @property
    def row_width(self) -> int:
        return self._row_width

    def column_of(self, leaf_index: int) -> int:
        return sum(self._widths[:leaf_index])

    def init(self, key: torch.Tensor) -> ReplayBufferState:
        dev = key.device if self._device is None else torch.device(self._device)
        ring = torch.zeros((self._max_replay_size, self._row_width), dtype=torch.float32, device=dev)
        return ReplayBufferState(ring=ring, head=0, insert_position=0, sample_position=0, key=key)

    def insert(self, buffer_state: ReplayBufferState, samples: Any) -> ReplayBufferState:
        """samples: the dummy sample's structure with leading batch dimensions (any number: time-major rollout
        buffers [T, E, ...] insert in the (t, e) order ``jnp.concatenate`` gives at sac.py:296)."""
        leaves = _leaves(samples)
        if len(leaves) != len(self._widths):
            raise ValueError("insert: %d leaves, the queue's rows hold %d" % (len(leaves), len(self._widths)))
        fields = _lib.ReplayFieldsC(num_fields=len(leaves))
        keep, n_rows = [], None
        for f, (leaf, shape, width) in enumerate(zip(leaves, self._shapes, self._widths)):
            t = leaf.to(torch.float32).contiguous()
            lead = t.shape[:t.dim() - len(shape)]
            if tuple(t.shape[t.dim() - len(shape):]) != shape:
                raise ValueError("insert: leaf %d has shape %s, the dummy sample's is %s" % (f, tuple(t.shape), shape))
            rows = int(torch.Size(lead).numel())
            if n_rows is None:
                n_rows = rows
            elif rows != n_rows:
                raise ValueError("insert: leaves disagree on the batch size (%d vs %d)" % (rows, n_rows))
            keep.append(t)
            fields.width[f] = width
            fields.ptr[f] = _lib.ptr(t)
        if n_rows > self._max_replay_size:
            raise ValueError("Trying to insert a batch of samples larger than the maximum replay size. "
                             "num_samples: %d, max replay size %d" % (n_rows, self._max_replay_size))
        st = buffer_state._c()
        with _lib.cuda_guard(buffer_state.ring):
            _lib.check(_lib.lib.mbpo_replay_insert(_lib.C.byref(st), _lib.C.byref(fields), n_rows,
                                                   _lib.stream_ptr(buffer_state.ring.device)))
        return buffer_state.replace(head=st.head, insert_position=st.insert_position,
                                    sample_position=st.sample_position)

    def sample(self, buffer_state: ReplayBufferState):
        """-> (new buffer_state, batch with the dummy sample's structure and a leading [sample_batch_size])."""
        new_state, batch, _ = self.sample_with_indices(buffer_state)
        return new_state, batch

    def sample_with_indices(self, buffer_state: ReplayBufferState):
        dev = buffer_state.ring.device
        n = self._sample_batch_size
        key = buffer_state.key.contiguous()
        key_out = torch.empty_like(key)
        idx = torch.empty((n,), dtype=torch.int32, device=dev)
        rows = torch.empty((n, self._row_width), dtype=torch.float32, device=dev)
        st = buffer_state._c()
        with _lib.cuda_guard(buffer_state.ring):
            _lib.check(_lib.lib.mbpo_replay_sample(_lib.C.byref(st), _lib.ptr(key), config.prng_mode, n,
                                                   _lib.ptr(key_out), _lib.ptr(idx), _lib.ptr(rows),
                                                   _lib.stream_ptr(dev)))
        return buffer_state.replace(key=key_out), self.unflatten(rows), idx

    def take(self, buffer_state: ReplayBufferState, idx: torch.Tensor) -> Any:
        """unflatten(jnp.take(buffer_state.data, idx, axis=0, mode='wrap')) (bptt_optimizer.py:447-450)."""
        idx = idx.to(torch.int32).contiguous().reshape(-1)
        rows = torch.empty((idx.numel(), self._row_width), dtype=torch.float32, device=buffer_state.ring.device)
        st = buffer_state._c()
        with _lib.cuda_guard(buffer_state.ring):
            _lib.check(_lib.lib.mbpo_replay_take(_lib.C.byref(st), _lib.ptr(idx), idx.numel(), _lib.ptr(rows),
                                                 _lib.stream_ptr(rows.device)))
        return self.unflatten(rows)

    def unflatten(self, rows: torch.Tensor) -> Any:
        """[n, D] rows -> the dummy sample's structure (column views of ``rows``)."""
        leaves, col = [], 0
        for shape, width in zip(self._shapes, self._widths):
            leaves.append(rows[:, col:col + width].reshape(rows.shape[0], *shape))
            col += width
        return _unflatten(self._dummy, leaves)

    def size(self, buffer_state: ReplayBufferState) -> int:
        return buffer_state.insert_position - buffer_state.sample_position
