r"""The oracle restatement (oracle/score_oracle.py) against the golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""

import numpy as np
import pytest
import torch

from oracle import score_oracle as so
from oracle.testing import rel_l2

from helpers import build_state


@pytest.mark.parametrize('L,k', [(5, 2), (9, 2), (7, 1), (12, 3)])
def test_window_maps_bit_exact(golden, L, k):
    g = golden('maps')
    B, C, H, W = 2, 2, 2, 3
    x = torch.arange(B * L * C * H * W, dtype=torch.float64).reshape(B, L, C, H, W)
    u = so.unfold(x, k)
    assert torch.equal(u, torch.from_numpy(g[f'unfold_L{L}_k{k}']))
    assert torch.equal(so.fold(u, k), torch.from_numpy(g[f'fold_L{L}_k{k}']))
    gg = torch.from_numpy(g[f'adjoint_g_L{L}_k{k}'])
    assert torch.equal(so.unfold_transpose(gg, k), torch.from_numpy(g[f'adjoint_L{L}_k{k}']))
    assert [i * 100 + s for i, s in so.fold_map(L, k)] == [int(v) for v in g[f'foldtag_L{L}_k{k}']]


def test_unfold_short_trajectory_raises():
    with pytest.raises(RuntimeError):
        so.unfold(torch.zeros(1, 4, 2, 2, 2), 2)


def test_schedule(golden):
    g = golden('schedule')
    t = torch.from_numpy(g['t'])

    for kind in ('cos', 'lin', 'exp'):
        assert torch.equal(so.mu(t, kind), torch.from_numpy(g[f'mu_{kind}']))
        assert torch.equal(so.sigma(t, kind), torch.from_numpy(g[f'sigma_{kind}']))

    assert torch.equal(so.sigma(t, sde='subvp'), torch.from_numpy(g['sigma_subvp']))
    assert torch.equal(so.sigma(t, sde='subsubvp'), torch.from_numpy(g['sigma_subsubvp']))
    # mu(1) = 1e-3, sigma(1) = 1, mu(0) = 1, sigma(0) = 1e-3
    assert abs(float(so.mu(torch.tensor(1.0))) - 1e-3) < 1e-6
    assert abs(float(so.sigma(torch.tensor(0.0))) - 1e-3) < 1e-6


def test_helpers(golden):
    g = golden('helpers')
    x = torch.from_numpy(g['x'])
    assert rel_l2(so.coarsen(x, 2), torch.from_numpy(g['coarsen2'])) < 1e-7
    assert rel_l2(so.coarsen(x, 4), torch.from_numpy(g['coarsen4'])) < 1e-7
    assert rel_l2(so.vorticity(x), torch.from_numpy(g['vorticity'])) < 1e-7


@pytest.mark.parametrize('name', ['net_small', 'net_config'])
def test_network(golden, name):
    g = golden(name)
    state, k = build_state(name, 16)
    x, t = torch.from_numpy(g['x']), torch.tensor(float(g['t']))
    assert rel_l2(so.time_embedding(state, torch.from_numpy(g['emb_t'])), torch.from_numpy(g['emb'])) < 1e-6
    wins = so.unfold(x, k)
    assert rel_l2(so.score_unet(state, wins[:, :1], t, state['forcing']), torch.from_numpy(g['kernel_out'])) < 2e-6
    assert rel_l2(so.mc_score(state, x, t, k), torch.from_numpy(g['mc_score'])) < 2e-6
    state64 = {kk: v.double() for kk, v in state.items()}
    assert rel_l2(so.mc_score(state64, x.double(), t.double(), k), torch.from_numpy(g['mc_score_fp64'])) < 1e-12

    A = lambda v: v[..., ::2, ::2]  # noqa: E731
    gs = so.gaussian_score(lambda a, b: so.mc_score(state, a, b, k), torch.from_numpy(g['y']), A, 0.1, x, t, gamma=1e-2)
    assert rel_l2(gs, torch.from_numpy(g['gaussian_score'])) < 2e-5
    A2 = lambda v: so.coarsen(v[:, ::2], 4)  # noqa: E731
    gs2 = so.gaussian_score(lambda a, b: so.mc_score(state, a, b, k), torch.from_numpy(g['y2']), A2, 0.1, x, torch.tensor(0.8), gamma=1e-2)
    assert rel_l2(gs2, torch.from_numpy(g['gaussian_score_coarsen'])) < 2e-5


def test_sampler(golden):
    g = golden('net_small')
    state, k = build_state('net_small', 16)
    A = lambda v: v[..., ::2, ::2]  # noqa: E731
    y = torch.from_numpy(g['y'])
    full, n, corr = (int(v) for v in g['sample_meta'])
    out = so.pc_sample(
        lambda a, b: so.gaussian_score(lambda c, d: so.mc_score(state, c, d, k), y, A, 0.1, a, b, gamma=1e-2),
        torch.from_numpy(g['sample_x1']), steps=full, corrections=corr, tau=0.5,
        noise=list(torch.from_numpy(g['sample_noise'])), n_steps=n,
    )
    assert rel_l2(out, torch.from_numpy(g['sample_after'])) < 1e-4


def test_kolmogorov_oracle_invariants():
    r"""The stepper restatement is unpinned by the reference; check what the scheme guarantees."""

    from oracle import kolmogorov_oracle as ko

    assert ko.inner_steps(256, 0.2) == 82
    rng = np.random.default_rng(0)
    x = ko.prior((2,), 32, rng, np.float64)
    assert np.allclose(np.sqrt((x ** 2).sum(-3)).max(axis=(-2, -1)), 3.0)
    assert np.abs(ko.divergence(x)).max() < 1e-10
    y = ko.transition(x, dt=0.2)
    assert np.abs(ko.divergence(y)).max() < 1e-10
    assert np.sqrt((y ** 2).sum(-3)).max() < 5.0
    y32 = ko.transition(x.astype(np.float32), dt=0.2)
    assert np.linalg.norm(y32 - y) / np.linalg.norm(y) < 1e-5


# --------------------------------------------------------------------------- stepper oracle: analytic anchors
# jax-cfd is absent (parity unpinned by the reference, oracle/kolmogorov_oracle.py header), so the restated
# scheme is anchored on exact solutions of the incompressible Navier-Stokes equations instead of on memory.
def _taylor_green(N, amp=1.0):
    r"""u = sin x cos y, v = -cos x sin y on the staggered MAC positions (u at ((i+1)h, (j+1/2)h), v at
    ((i+1/2)h, (j+1)h)): an exact solution of unforced Navier-Stokes decaying as exp(-2 nu t)."""

    import math

    h = 2 * math.pi / N
    xu, yu = (np.arange(N) + 1) * h, (np.arange(N) + 0.5) * h
    xv, yv = (np.arange(N) + 0.5) * h, (np.arange(N) + 1) * h

    return amp * np.stack((np.sin(xu)[:, None] * np.cos(yu)[None, :], -np.cos(xv)[:, None] * np.sin(yv)[None, :]))


def test_stepper_oracle_reproduces_taylor_green_decay_at_second_order():
    import math

    from oracle import kolmogorov_oracle as ko

    nu, T, errs = 1e-2, 1.0, {}

    for N in (16, 32, 64):
        h = 2 * math.pi / N
        n = int(round(T / (0.2 * h * h)))  # dt ~ h^2: the forward-Euler error stays below the spatial one
        x = _taylor_green(N)

        for _ in range(n):
            x = ko.transition(x, dt=T / n, reynolds=1 / nu, forced=False)

        exact = _taylor_green(N, math.exp(-2 * nu * T))
        errs[N] = np.linalg.norm(x - exact) / np.linalg.norm(exact)
        assert np.abs(ko.divergence(x)).max() < 1e-12

    assert errs[64] < 1e-3
    # second-order convergence in h (van Leer limited Lax-Wendroff advection, centred diffusion and projection)
    assert errs[16] / errs[32] > 3.5 and errs[32] / errs[64] > 3.5


def test_stepper_oracle_reaches_the_laminar_kolmogorov_flow():
    r"""Forced, at Re = 10 (linearly stable): the steady state is u = A sin(4 y), v = 0 with
    A = 1 / (nu |lambda_4| + 0.1), lambda_4 the discrete Laplacian eigenvalue of sin(4 y).  Pins the forcing
    amplitude (1), wavenumber (4), offset (u sits at y_{j+1/2}), the drag (-0.1 u) and the viscosity."""

    import math

    from oracle import kolmogorov_oracle as ko

    N, nu = 32, 0.1
    h = 2 * math.pi / N
    x = np.zeros((2, N, N))

    for _ in range(60):
        x = ko.transition(x, dt=0.2, reynolds=1 / nu)

    amp = 1 / (nu * (2 - 2 * math.cos(4 * h)) / h ** 2 + 0.1)
    exact = amp * np.sin(4 * (np.arange(N) + 0.5) * h)[None, :] * np.ones((N, 1))
    assert np.abs(x[0] - exact).max() < 1e-7 * amp
    assert np.abs(x[1]).max() < 1e-12
    assert abs(amp - 1 / (16 * nu + 0.1)) < 0.05  # the continuum value, for orientation
