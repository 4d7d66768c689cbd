"""brax.training.acme.running_statistics work-alike: the observation normaliser SAC / PPO update after every
collection (mbpo/optimizers/policy_optimizers/sac/sac.py:298-301, ppo/ppo.py) and the policies read
(sac_networks.py:58-73; the actor kernels take ``mean`` / ``std`` through acting.Policy).

``update`` is one pass over the rows (``mbpo_running_statistics_accumulate``), then -- the one real exchange step of the
data-collection path, the reference's ``pmap_axis_name`` psum -- an all-reduce of 2X + 1 float64 sums over the ranks
when a process group is initialised (NCCL over NVLink; gloo in the CPU tests), then ``mbpo_running_statistics_finalize``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from .systems.base_systems import _Replaceable


@dataclass
class RunningStatisticsState(_Replaceable):
    """brax RunningStatisticsState: count [] , mean [X], summed_variance [X], std [X] (float32)."""
    count: torch.Tensor = None
    mean: torch.Tensor = None
    summed_variance: torch.Tensor = None
    std: torch.Tensor = None


def init_state(size: int, device=None) -> RunningStatisticsState:
    """running_statistics.init_state(specs.Array((size,), float32)): zeros, std = ones."""
    dev = _lib.require_cuda(device)
    z = lambda: torch.zeros(size, dtype=torch.float32, device=dev)
    return RunningStatisticsState(count=torch.zeros((), dtype=torch.float32, device=dev), mean=z(), summed_variance=z(),
                                  std=torch.ones(size, dtype=torch.float32, device=dev))


def all_reduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """The psum of the reference's pmap_axis_name: float64 [2X + 1] (sum d, sum d*d, row count), summed over ranks."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def update(state: RunningStatisticsState, batch: torch.Tensor, *, weights=None, std_min_value: float = 1e-6,
           std_max_value: float = 1e6, pmap_axis_name: Optional[str] = None, validate_shapes: bool = True,
           group=None) -> RunningStatisticsState:
    """running_statistics.update(state, batch, pmap_axis_name=...): batch [..., X] (any leading batch dims).
    With ``pmap_axis_name`` (or an explicit ``group``) the sums are all-reduced over the process group, so every rank
    ends with the statistics of the union of the ranks' batches."""
    X = state.mean.shape[-1]
    if weights is not None:      # brax's keyword; no caller in the reference passes it (sac.py:298-301, ppo.py)
        raise _lib.MbpoUnsupported(_lib.MBPO_EUNSUPPORTED, "running_statistics.update: weighted updates have no kernel")
    if validate_shapes and batch.shape[-1] != X:
        raise ValueError("running_statistics.update: batch [..., %d] does not match the statistics [%d]" % (
            batch.shape[-1], X))
    b = batch.to(torch.float32).contiguous()
    n_rows = b.numel() // X
    dev = b.device
    ws_bytes = _lib.lib.mbpo_running_statistics_workspace_bytes(X)
    if ws_bytes == 0:
        raise _lib.MbpoUnsupported(_lib.MBPO_EUNSUPPORTED, "running_statistics: observation size %d has no kernel" % X)
    workspace = torch.empty(ws_bytes // 8, dtype=torch.float64, device=dev)
    sums = torch.empty(2 * X + 1, dtype=torch.float64, device=dev)
    mean_in = state.mean.contiguous()
    with _lib.cuda_guard(b):
        _lib.check(_lib.lib.mbpo_running_statistics_accumulate(_lib.ptr(b), n_rows, X, _lib.ptr(mean_in),
                                                               _lib.ptr(workspace), ws_bytes, _lib.ptr(sums),
                                                               _lib.stream_ptr(dev)))
    if pmap_axis_name is not None or group is not None:
        all_reduce_sums(sums, group)            # sums[2X] carries the row count: step_increment is psum-ed with it
    count = torch.empty_like(state.count)
    mean, sv, std = torch.empty_like(mean_in), torch.empty_like(mean_in), torch.empty_like(mean_in)
    with _lib.cuda_guard(b):
        _lib.check(_lib.lib.mbpo_running_statistics_finalize(
            _lib.ptr(sums), X, _lib.ptr(state.count.reshape(1).contiguous()), _lib.ptr(mean_in),
            _lib.ptr(state.summed_variance.contiguous()), std_min_value, std_max_value, _lib.ptr(count.reshape(1)),
            _lib.ptr(mean), _lib.ptr(sv), _lib.ptr(std), _lib.stream_ptr(dev)))
    return RunningStatisticsState(count=count, mean=mean, summed_variance=sv, std=std)


def normalize(batch: torch.Tensor, mean_std: RunningStatisticsState, max_abs_value: Optional[float] = None):
    """running_statistics.normalize: (batch - mean) / std, optionally clipped."""
    X = mean_std.mean.shape[-1]
    b = batch.to(torch.float32).contiguous()
    out = torch.empty_like(b)
    with _lib.cuda_guard(b):
        _lib.check(_lib.lib.mbpo_running_statistics_normalize(
            _lib.ptr(b), b.numel() // X, X, _lib.ptr(mean_std.mean.contiguous()), _lib.ptr(mean_std.std.contiguous()),
            float(max_abs_value) if max_abs_value is not None else 0.0, _lib.ptr(out), _lib.stream_ptr(b.device)))
    return out


# ---- BPTT's own Normalizer (mbpo/optimizers/policy_optimizers/bptt_optimizer.py:31-75) -------------------------------
EPS = 1e-8


@dataclass
class NormalizerState(_Replaceable):
    """bptt_optimizer.py:31-35.  ``size`` is a float64 scalar on the device (the reference's traced int)."""
    mean: torch.Tensor = None
    std: torch.Tensor = None
    size: torch.Tensor = None


class Normalizer:
    """bptt_optimizer.py:37-75: ``update(x, state)`` / ``normalize(x, state)`` / ``inverse(x, state)``; the state
    normaliser's mean / std are what acting.BpttActorPolicy reads.  ``group`` (additive): all-reduce the sums when the
    batch is sharded over ranks."""

    def __init__(self, input_shape, device=None):
        self.input_shape = tuple(input_shape)
        self._device = device

    def initialize_normalizer_state(self) -> NormalizerState:
        dev = _lib.require_cuda(self._device)
        n = self.input_shape[0]
        return NormalizerState(mean=torch.zeros(n, dtype=torch.float32, device=dev),
                               std=torch.ones(n, dtype=torch.float32, device=dev),
                               size=torch.zeros((), dtype=torch.float64, device=dev))

    @staticmethod
    def update(x: torch.Tensor, state: NormalizerState, group=None) -> NormalizerState:
        X = state.mean.shape[-1]
        b = x.to(torch.float32).contiguous()
        n_rows = b.numel() // X
        dev = b.device
        ws_bytes = _lib.lib.mbpo_running_statistics_workspace_bytes(X)
        if ws_bytes == 0:
            raise _lib.MbpoUnsupported(_lib.MBPO_EUNSUPPORTED, "Normalizer: input size %d has no kernel" % X)
        workspace = torch.empty(ws_bytes // 8, dtype=torch.float64, device=dev)
        sums = torch.empty(2 * X + 1, dtype=torch.float64, device=dev)
        mean_in, std_in = state.mean.contiguous(), state.std.contiguous()
        size_in = state.size.to(torch.float64).reshape(1).contiguous()
        size, mean, std = torch.empty_like(size_in), torch.empty_like(mean_in), torch.empty_like(std_in)
        with _lib.cuda_guard(b):
            _lib.check(_lib.lib.mbpo_running_statistics_accumulate(_lib.ptr(b), n_rows, X, _lib.ptr(mean_in),
                                                                   _lib.ptr(workspace), ws_bytes, _lib.ptr(sums),
                                                                   _lib.stream_ptr(dev)))
            if group is not None:
                all_reduce_sums(sums, group)
            _lib.check(_lib.lib.mbpo_normalizer_finalize(_lib.ptr(sums), X, _lib.ptr(size_in), _lib.ptr(mean_in),
                                                         _lib.ptr(std_in), EPS, _lib.ptr(size), _lib.ptr(mean),
                                                         _lib.ptr(std), _lib.stream_ptr(dev)))
        return NormalizerState(mean=mean, std=std, size=size.reshape(()))

    @staticmethod
    def normalize(x: torch.Tensor, state: NormalizerState) -> torch.Tensor:
        return normalize(x, RunningStatisticsState(mean=state.mean, std=state.std))

    @staticmethod
    def inverse(x: torch.Tensor, state: NormalizerState) -> torch.Tensor:
        X = state.mean.shape[-1]
        b = x.to(torch.float32).contiguous()
        out = torch.empty_like(b)
        with _lib.cuda_guard(b):
            _lib.check(_lib.lib.mbpo_normalizer_inverse(_lib.ptr(b), b.numel() // X, X, _lib.ptr(state.mean.contiguous()),
                                                        _lib.ptr(state.std.contiguous()), _lib.ptr(out),
                                                        _lib.stream_ptr(b.device)))
        return out
