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
