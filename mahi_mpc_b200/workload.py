"""Synthetic inputs of the benchmark / demo drivers (harness utility, pure numpy; no arithmetic of the hot path).

Mirrors the reference drivers' reference-trajectory construction (examples/model_control_example.cpp:60-68:
positions +-a sin(2 pi f t), velocities +-2 pi f a cos(2 pi f t)), randomised per instance as SURVEY.md 8d
specifies: Philox counter RNG, seed 0x5EED, x0 = [U(-.5,.5)^(nx/2), U(-1,1)^(nx/2)], u_prev = 0, a ~ U(.25,1),
f ~ U(.25,1) Hz, phi ~ U(0,2 pi).  Each instance owns a fixed slice of the counter space, so a shard
[offset, offset+B) reproduces exactly the same instances as the un-sharded batch.
"""
from __future__ import annotations

import numpy as np

DEFAULT_Q = {4: [10.0, 1.0, 5.0, 5.0]}  # examples/thread_model_control_example.cpp:24
DEFAULT_R = {2: [5.0, 5.0]}  # :25


def default_weights(nx, nu):
    Q = np.array(DEFAULT_Q.get(nx, ([10.0, 1.0] * nx)[: nx // 2] + [5.0] * (nx - nx // 2)), float)
    R = np.array(DEFAULT_R.get(nu, [5.0] * nu), float)
    Rm = np.ones(nu)  # torque cost on (the reference ships no default)
    return Q, R, Rm


def synthetic_batch(nx, nu, N, h, B, seed=0x5EED, offset=0):
    """x0 (B,nx), u_prev (B,nu), traj (B,N,nx)."""
    half = nx // 2
    assert half <= 6
    bg = np.random.Philox(key=seed)
    bg.advance(offset * 4)  # 4 counter blocks = 16 doubles reserved per instance
    draws = np.random.Generator(bg).random((B, 16))
    x0 = np.empty((B, nx))
    x0[:, :half] = draws[:, :half] - 0.5
    x0[:, half:] = 2.0 * draws[:, 6:6 + nx - half] - 1.0
    amp = 0.25 + 0.75 * draws[:, 12]
    frq = 0.25 + 0.75 * draws[:, 13]
    phs = 2 * np.pi * draws[:, 14]
    arg = 2 * np.pi * frq[:, None] * (np.arange(N) * h)[None, :] + phs[:, None]
    s, c = np.sin(arg), np.cos(arg)
    traj = np.empty((B, N, nx))
    for j in range(nx):
        sign = 1.0 if ((j - half) if j >= half else j) % 2 == 0 else -1.0
        traj[:, :, j] = sign * amp[:, None] * (s if j < half else 2 * np.pi * frq[:, None] * c)
    return x0, np.zeros((B, nu)), traj
