"""Process-wide switches mirroring the JAX flags the reference's results depend on."""
from __future__ import annotations

from . import _lib


class _Config:
    """threefry_partitionable mirrors ``jax_threefry_partitionable``.  The reference's era
    (JAX 0.4.x: it uses ``jax.tree_map``, removed in 0.6) defaults to False = legacy layout.

    math_mode selects how the pendulum angle is carried between steps:
    ``"reference"`` re-derives it with atan2 from [cos, sin] every step exactly as
    pendulum_dynamics.py:35,43 does; ``"theta_carry"`` keeps it in a register (same
    mathematics, different rounding)."""

    threefry_partitionable: bool = False
    math_mode: str = "reference"
    enable_x64: bool = False      # mirrors jax_enable_x64; only random.PRNGKey's seed handling depends on it

    @property
    def prng_mode(self) -> int:
        return _lib.PRNG_PARTITIONABLE if self.threefry_partitionable else _lib.PRNG_LEGACY

    @property
    def math_mode_id(self) -> int:
        if self.math_mode not in ("reference", "theta_carry"):
            raise ValueError("config.math_mode must be 'reference' or 'theta_carry'")
        return _lib.MATH_REFERENCE if self.math_mode == "reference" else _lib.MATH_THETA_CARRY


config = _Config()
