"""Multi-GPU sharding of independent planning problems / envs (one process per GPU).

Problems never exchange data (every problem owns its key, mean/std and initial state:
icem_optimizer.py:62-69), so ranks take contiguous blocks and the only collective is the
final gather of first actions (or of rollout buffers).  Results are bit-identical for every
world size because keys are per problem.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `total` units owned by `rank`; sizes differ by at most 1."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside [0, %d)" % (rank, world_size))
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_gather_blocks(local: torch.Tensor, total: int) -> torch.Tensor:
    """Concatenates every rank's block (leading axis sized by shard_bounds) into [total, ...]
    on all ranks.  NCCL over NVLink on GPUs; gloo in the CPU tests."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_bounds(total, r, ws) for r in range(ws)]
    max_rows = max(hi - lo for lo, hi in sizes)
    if total % ws == 0:                       # equal blocks: gather straight into the result
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((ws * max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    out = out.reshape((ws, max_rows) + tuple(local.shape[1:]))
    return torch.cat([out[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def plan_sharded(optimizer, initial_states: torch.Tensor, opt_state, gather: bool = True):
    """Each rank plans its block of problems; first actions [B, A] are gathered on all ranks.

    `initial_states` [B, X] and the batched `opt_state` are the GLOBAL problem set (host or
    device); every rank slices its block, so a G-rank run equals the 1-rank run bit for bit.
    Returns (actions [B, A] if gather else local block, local new opt_state, (lo, hi))."""
    rank, ws = world()
    total = initial_states.shape[0]
    lo, hi = shard_bounds(total, rank, ws)
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else initial_states.device
    local_state = opt_state.replace(key=opt_state.key[lo:hi].to(dev),
                                    best_sequence=opt_state.best_sequence[lo:hi].to(dev),
                                    best_reward=opt_state.best_reward[lo:hi].to(dev))
    actions, new_state = optimizer.act(initial_states[lo:hi].to(dev), local_state)
    if gather:
        actions = all_gather_blocks(actions.contiguous(), total)
    return actions, new_state, (lo, hi)
