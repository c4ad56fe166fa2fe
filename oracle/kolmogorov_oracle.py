r"""NumPy restatement of the Kolmogorov-flow stepper -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  `sda.mcs.KolmogorovFlow` (sda/mcs.py:244-338) delegates all
arithmetic to jax-cfd (`jax_cfd.base`), installed by the reference's README from
git HEAD (README.md:22-26; no version pin) on jax==0.4.4 (environment.yml:16-17).
Neither is vendored in /root/reference nor installable here, and the reference
holds no test or golden vector for this path.  This file restates the PUBLISHED
algorithm of the functions the reference calls, anchored on the reference's call
sites:

* grid            cfd.grids.Grid(shape=(N, N), domain=[0, 2 pi]^2)              mcs.py:259-262
* boundary        periodic                                                      mcs.py:264
* forcing         simple_turbulence_forcing(constant_magnitude=1, constant_wavenumber=4,
                  linear_coefficient=-0.1, forcing_type='kolmogorov')            mcs.py:266-272
* inner steps     dt_min = stable_time_step(max_velocity=5, max_courant_number=0.5)
                  steps = ceil(dt / dt_min) (1 if dt_min > dt)                   mcs.py:274-284
* step            semi_implicit_navier_stokes(density=1, viscosity=1/Re, dt/steps)
                  repeated `steps` times                                         mcs.py:286-295
* prior           filtered_velocity_field(maximum_velocity=3, peak_wavenumber=4) mcs.py:297-305

jax-cfd scheme (staggered MAC grid, arrays indexed [x, y]; u lives at offset (1, 1/2),
v at (1/2, 1), pressure at cell centres):
  explicit terms  F(v) = -div(c_face * u_face) + nu * laplacian(v) + f
      u_face  : linear interpolation of the advecting component to the face
      c_face  : van-Leer-limited Lax-Wendroff (upwind + phi(r) * (LW - upwind))
  forward Euler   v* = v + dt F(v)
  projection      q = pinv(laplacian) div(v*)  (FFT-diagonalised, zero mode -> 0)
                  v = v* - grad q
State layout is the reference's: x[..., 0, :, :] = u, x[..., 1, :, :] = v, axis -2 = x, axis -1 = y.
"""

from __future__ import annotations

import math

import numpy as np


def inner_steps(size: int, dt: float) -> int:
    r"""mcs.py:274-284 with jax-cfd's advection stable_time_step = courant * dx / max_velocity."""

    dx = 2 * math.pi / size
    dt_min = 0.5 * dx / 5.0

    return 1 if dt_min > dt else math.ceil(dt / dt_min)


def _shift(a, s, axis):
    r"""jax-cfd `GridArray.shift(s, axis)`: result[i] = a[i + s] (periodic)."""

    return np.roll(a, -s, axis=axis)


def _safe_div(x, y):
    return x / np.where(y != 0, y, 1)


def _van_leer(r):
    return np.where(r > 0, _safe_div(2 * r, 1 + r), 0.0).astype(r.dtype)


def _face_value(c, u_face, axis, dt, h):
    r"""Van-Leer limited Lax-Wendroff value of c on the + face along `axis` (apply_tvd_limiter)."""

    c_left, c_right, c_next = _shift(c, -1, axis), _shift(c, 1, axis), _shift(c, 2, axis)
    pos_r = _safe_div(c - c_left, c_right - c)
    neg_r = _safe_div(c_next - c_right, c_right - c)
    phi = np.where(u_face > 0, _van_leer(pos_r), _van_leer(neg_r))
    upwind = np.where(u_face > 0, c, c_right)
    courant = (dt / h) * u_face
    lw_pos = c + 0.5 * (1 - courant) * (c_right - c)
    lw_neg = c_right - 0.5 * (1 + courant) * (c_right - c)
    high = np.where(u_face > 0, lw_pos, lw_neg)

    return upwind - (upwind - high) * phi


def explicit_terms(u, v, dt, h, nu, forcing_u, drag=0.1):
    r"""convection + diffusion + forcing of navier_stokes_explicit_terms, for both components.
    `forcing_u` = 0 and `drag` = 0 switch the Kolmogorov forcing off (unforced Navier-Stokes: the analytic
    anchors of tests/test_oracle.py); the reference always runs with both on (mcs.py:266-272)."""

    ax, ay = -2, -1
    out = []

    for comp, c in enumerate((u, v)):
        if comp == 0:  # c = u at (1, 1/2): faces at (3/2, 1/2) and (1, 1)
            ux = 0.5 * (u + _shift(u, 1, ax))
            vy = 0.5 * (v + _shift(v, 1, ax))
        else:  # c = v at (1/2, 1): faces at (1, 1) and (1/2, 3/2)
            ux = 0.5 * (u + _shift(u, 1, ay))
            vy = 0.5 * (v + _shift(v, 1, ay))

        fx = _face_value(c, ux, ax, dt, h) * ux
        fy = _face_value(c, vy, ay, dt, h) * vy
        conv = -((fx - _shift(fx, -1, ax)) + (fy - _shift(fy, -1, ay))) / h
        lap = (_shift(c, 1, ax) + _shift(c, -1, ax) + _shift(c, 1, ay) + _shift(c, -1, ay) - 4 * c) / h ** 2
        force = (forcing_u if comp == 0 else 0) - drag * c
        out.append(conv + nu * lap + force)

    return out


def laplacian_eigenvalues(size: int, h: float, dtype):
    k = np.arange(size)
    lam = (2 * np.cos(2 * np.pi * k / size) - 2) / h ** 2

    return lam.astype(dtype)


def project(u, v, h):
    r"""pressure.projection with solve_fast_diag (circulant, pseudo-inverse)."""

    ax, ay = -2, -1
    size = u.shape[-1]
    div = ((u - _shift(u, -1, ax)) + (v - _shift(v, -1, ay))) / h
    lam = laplacian_eigenvalues(size, h, np.float64)
    denom = lam[:, None] + lam[None, : size // 2 + 1]
    inv = np.where(np.abs(denom) > 1e-9, 1 / np.where(denom == 0, 1, denom), 0.0)
    q = np.fft.irfft2(np.fft.rfft2(div) * inv, s=div.shape[-2:]).astype(u.dtype)

    return u - (_shift(q, 1, ax) - q) / h, v - (_shift(q, 1, ay) - q) / h


def forcing_profile(size: int, dtype):
    r"""kolmogorov_forcing(k=4) evaluated at u's offset: sin(4 y_{j+1/2}), constant along x."""

    h = 2 * math.pi / size
    y = (np.arange(size) + 0.5) * h

    return np.sin(4 * y).astype(dtype)[None, :]


def transition(x: np.ndarray, dt: float = 0.2, reynolds: float = 1e3, n_inner: int | None = None, forced: bool = True) -> np.ndarray:
    r"""KolmogorovFlow.transition (mcs.py:333-338, inner function :307-316). x: (..., 2, N, N).
    forced=False: no Kolmogorov forcing and no linear drag (test anchors only)."""

    size = x.shape[-1]
    h = x.dtype.type(2 * math.pi / size)
    steps = inner_steps(size, dt)
    sub = x.dtype.type(dt / steps)
    nu = x.dtype.type(1 / reynolds)
    f = forcing_profile(size, x.dtype) if forced else x.dtype.type(0)
    drag = x.dtype.type(0.1 if forced else 0.0)
    u, v = x[..., 0, :, :], x[..., 1, :, :]

    for _ in range(steps if n_inner is None else n_inner):
        du, dv = explicit_terms(u, v, sub, h, nu, f, drag)
        u, v = project(u + sub * du, v + sub * dv, h)

    return np.stack((u, v), axis=-3).astype(x.dtype)


def divergence(x: np.ndarray) -> np.ndarray:
    size = x.shape[-1]
    h = 2 * math.pi / size
    u, v = x[..., 0, :, :], x[..., 1, :, :]

    return ((u - _shift(u, -1, -2)) + (v - _shift(v, -1, -1))) / h


def prior(shape, size: int, rng: np.random.Generator, dtype=np.float32) -> np.ndarray:
    r"""filtered_velocity_field(maximum_velocity=3, peak_wavenumber=4): white noise filtered by the
    square root of a log-normal spectral density peaked at wavenumber 4 (divided by k for the
    circle circumference), then 3 x (project, rescale to max speed 3).  jax's threefry keys cannot
    be reproduced here: distribution-level agreement only."""

    h = 2 * math.pi / size
    k1 = 2 * np.pi * np.fft.fftfreq(size, h)
    kk = np.sqrt(k1[:, None] ** 2 + k1[None, :] ** 2)
    variance, mode = 0.25, 4.0
    mean = math.log(mode) + variance

    with np.errstate(divide='ignore', invalid='ignore'):
        logk = np.log(kk)
        dens = np.exp(-(mean - logk) ** 2 / 2 / variance - logk) / math.sqrt(2 * math.pi * variance) / kk

    filt = np.where(kk > 0, dens, 0.0)
    noise = rng.standard_normal(tuple(shape) + (2, size, size))
    field = np.fft.ifft2(np.fft.fft2(noise) * np.sqrt(filt)).real.astype(dtype)
    u, v = field[..., 0, :, :], field[..., 1, :, :]

    for _ in range(3):
        u, v = project(u, v, dtype(h))
        speed = np.sqrt(u ** 2 + v ** 2).max(axis=(-2, -1), keepdims=True)
        u, v = 3.0 * u / speed, 3.0 * v / speed

    return np.stack((u, v), axis=-3).astype(dtype)
