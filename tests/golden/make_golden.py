#!/usr/bin/env python
r"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Every vector is produced by the reference's own classes (sda.score / sda.nn /
sda.mcs static helpers, experiments/kolmogorov/utils.py:make_score) on CPU, with
weights filled by oracle.testing.fill_state_ (keyed by state_dict name) and
seeded inputs.  The script also asserts that the oracle restatement
(oracle/score_oracle.py) reproduces each vector, i.e. it PINS the oracle.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_import, score_oracle as so  # noqa: E402
from oracle.testing import fill_state_, randn, rel_l2  # noqa: E402

OUT = Path(__file__).resolve().parent

ref = ref_import.import_reference()
exp = ref_import.import_experiments(ref)


def save(name, **arrays):
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(OUT / f'{name}.npz', **arrays)
    size = (OUT / f'{name}.npz').stat().st_size
    print(f'{name}.npz: {size / 1024:.1f} KiB', {k: v.shape for k, v in arrays.items()})


def check(what, ours, theirs, tol):
    err = rel_l2(ours, theirs)
    print(f'  oracle vs reference [{what}]: rel-L2 = {err:.3e}')
    assert err <= tol, (what, err)


# ---------------------------------------------------------------- window maps
def golden_maps():
    out = {}

    for L, k in [(5, 2), (9, 2), (7, 1), (12, 3)]:
        B, C, H, W = 2, 2, 2, 3
        x = torch.arange(B * L * C * H * W, dtype=torch.float64).reshape(B, L, C, H, W)
        u = ref.score.MCScoreNet.unfold(x, k)
        f = ref.score.MCScoreNet.fold(u * 1.0, k)
        # adjoint of unfold through autograd (UnfoldBackward0)
        xr = x.clone().requires_grad_(True)
        g = torch.arange(u.numel(), dtype=torch.float64).reshape(u.shape) % 7 + 1
        (gx,) = torch.autograd.grad((ref.score.MCScoreNet.unfold(xr, k) * g).sum(), xr)
        # fold applied to a window-tagged tensor: value = window * 100 + slot
        nw = L - 2 * k
        tag = torch.zeros(1, nw, (2 * k + 1) * C, 1, 1, dtype=torch.float64)
        for i in range(nw):
            for s in range(2 * k + 1):
                tag[0, i, s * C:(s + 1) * C] = i * 100 + s
        ft = ref.score.MCScoreNet.fold(tag, k)[0, :, 0, 0, 0]

        assert torch.equal(so.unfold(x, k), u)
        assert torch.equal(so.fold(u, k), f)
        assert torch.equal(so.unfold_transpose(g, k), gx)
        assert [int(v) for v in ft] == [i * 100 + s for i, s in so.fold_map(L, k)]

        out[f'unfold_L{L}_k{k}'] = u
        out[f'fold_L{L}_k{k}'] = f
        out[f'adjoint_g_L{L}_k{k}'] = g
        out[f'adjoint_L{L}_k{k}'] = gx
        out[f'foldtag_L{L}_k{k}'] = ft

    save('maps', **out)


# ---------------------------------------------------------------- schedule
def golden_schedule():
    t = torch.linspace(0, 1, 33)
    out = {'t': t}

    for kind in ('cos', 'lin', 'exp'):
        sde = ref.score.VPSDE(None, shape=(), alpha=kind)
        out[f'mu_{kind}'] = sde.mu(t)
        out[f'sigma_{kind}'] = sde.sigma(t)
        assert torch.equal(so.mu(t, kind), out[f'mu_{kind}'])
        assert torch.equal(so.sigma(t, kind), out[f'sigma_{kind}'])

    out['sigma_subvp'] = ref.score.SubVPSDE(None, shape=()).sigma(t)
    out['sigma_subsubvp'] = ref.score.SubSubVPSDE(None, shape=()).sigma(t)
    assert torch.equal(so.sigma(t, sde='subvp'), out['sigma_subvp'])
    assert torch.equal(so.sigma(t, sde='subsubvp'), out['sigma_subsubvp'])

    save('schedule', **out)


# ---------------------------------------------------------------- helpers
def golden_helpers():
    x = randn((3, 2, 16, 16), seed=11)
    c2 = ref.mcs.KolmogorovFlow.coarsen(x, 2)
    c4 = ref.mcs.KolmogorovFlow.coarsen(x, 4)
    w = ref.mcs.KolmogorovFlow.vorticity(x)
    check('coarsen2', so.coarsen(x, 2), c2, 1e-7)
    check('coarsen4', so.coarsen(x, 4), c4, 1e-7)
    check('vorticity', so.vorticity(x), w, 1e-7)
    save('helpers', x=x, coarsen2=c2, coarsen4=c4, vorticity=w)


# ---------------------------------------------------------------- networks
def build(window, hidden_channels, hidden_blocks, size, seed, dtype=torch.float32):
    score = exp.make_score(
        window=window,
        embedding=64,
        hidden_channels=hidden_channels,
        hidden_blocks=hidden_blocks,
        kernel_size=3,
        activation='SiLU',
    )
    # LocalScoreUNet bakes size=64 (experiments/kolmogorov/utils.py:35-43): rebuild the buffer for `size`
    domain = 2 * torch.pi / size * (torch.arange(size) + 1 / 2)
    score.kernel.forcing = torch.sin(4 * domain).expand(1, size, size).clone()
    fill_state_(score.state_dict(), seed=seed)
    return score.to(dtype).eval()


def kernel_state(score):
    return {k[len('kernel.'):]: v for k, v in score.state_dict().items()}


def golden_network(name, window, hidden_channels, hidden_blocks, size, B, L, seed, steps):
    k = window // 2
    score = build(window, hidden_channels, hidden_blocks, size, seed)
    state = kernel_state(score)
    x = randn((B, L, 2, size, size), seed=seed + 1)
    t = torch.tensor(0.37)
    out = {'x': x, 't': t}

    with torch.no_grad():
        # time embedding + the bare kernel on the first windows
        emb = score.kernel.embedding(torch.tensor([0.37, 0.9]))
        check('embedding', so.time_embedding(state, torch.tensor([0.37, 0.9])), emb, 1e-6)
        out['emb_t'] = torch.tensor([0.37, 0.9])
        out['emb'] = emb

        wins = ref.score.MCScoreNet.unfold(x, k)
        kern = score.kernel(wins[:, :1], t)
        check('kernel', so.score_unet(state, wins[:, :1], t, state['forcing']), kern, 2e-6)
        out['kernel_out'] = kern

        eps = score(x, t)
        check('mc_score', so.mc_score(state, x, t, k), eps, 2e-6)
        out['mc_score'] = eps

        # fp64 truth of the same call
        score64 = build(window, hidden_channels, hidden_blocks, size, seed, torch.float64)
        # identical weights: copy the fp32 values exactly
        score64.load_state_dict({kk: v.double() for kk, v in score.state_dict().items()})
        eps64 = score64(x.double(), t.double())
        out['mc_score_fp64'] = eps64
        print(f'  fp32 vs fp64 reference: rel-L2 = {rel_l2(eps, eps64):.3e}')

    # guided score: subsampled observation (experiments/kolmogorov/figures.ipynb:777)
    A = lambda v: v[..., ::2, ::2]  # noqa: E731
    y = randn((B, L, 2, size // 2, size // 2), seed=seed + 2)
    guided = ref.score.GaussianScore(y, A=A, std=0.1, sde=ref.score.VPSDE(score, shape=()), gamma=1e-2)
    gs = guided(x, t)
    gs = gs.detach()
    ours = so.gaussian_score(lambda a, b: so.mc_score(state, a, b, k), y, A, 0.1, x, t, gamma=1e-2)
    check('gaussian_score', ours, gs, 2e-5)
    out['y'] = y
    out['gaussian_score'] = gs

    # coarsened observation through the reference helper (figures.ipynb:204)
    A2 = lambda v: ref.mcs.KolmogorovFlow.coarsen(v[:, ::2], 4)  # noqa: E731
    y2 = randn((B, (L + 1) // 2, 2, size // 4, size // 4), seed=seed + 3)
    guided2 = ref.score.GaussianScore(y2, A=A2, std=0.1, sde=ref.score.VPSDE(score, shape=()), gamma=1e-2)
    out['y2'] = y2
    out['gaussian_score_coarsen'] = guided2(x, torch.tensor(0.8)).detach()

    # sampler with recorded noise: VPSDE.sample draws randn(shape) then randn_like per correction
    if steps:
        drawn = []
        real_randn_like = torch.randn_like

        def recording(v, *a, **kw):
            z = real_randn_like(v, *a, **kw)
            drawn.append(z.clone())
            return z

        sde = ref.score.VPSDE(guided, shape=(L, 2, size, size))
        full_steps = 16
        torch.manual_seed(seed + 4)
        x1 = torch.randn((B, L, 2, size, size))
        torch.manual_seed(seed + 4)
        torch.randn_like = recording

        # run only the first `steps` iterations of the reference loop: patch linspace length via a subclass-free trick:
        # sample() iterates time[:-1]; we stop by raising from eps after the wanted number of evaluations.
        class Stop(Exception):
            pass

        calls = {'n': 0}
        snapshots = []
        inner = sde.eps

        class Counting(torch.nn.Module):
            def forward(self, v, tt, c=None):
                if calls['n'] == 2 * steps:
                    snapshots.append(v.clone())
                    raise Stop()
                calls['n'] += 1
                return inner(v, tt, c)

        sde.eps = Counting()

        try:
            sde.sample((B,), steps=full_steps, corrections=1, tau=0.5)
        except Stop:
            pass
        finally:
            torch.randn_like = real_randn_like

        x_after = snapshots[0]
        noise = drawn[:steps]
        ours = so.pc_sample(
            lambda a, b: so.gaussian_score(lambda c, d: so.mc_score(state, c, d, k), y, A, 0.1, a, b, gamma=1e-2),
            x1, steps=full_steps, corrections=1, tau=0.5, noise=noise, n_steps=steps,
        )
        check('pc_sample', ours, x_after, 1e-4)
        out['sample_x1'] = x1
        out['sample_noise'] = torch.stack(noise)
        out['sample_after'] = x_after
        out['sample_meta'] = np.array([full_steps, steps, 1])

    save(name, **out)


if __name__ == '__main__':
    torch.set_num_threads(8)
    golden_maps()
    golden_schedule()
    golden_helpers()
    golden_network('net_small', window=3, hidden_channels=(32, 64), hidden_blocks=(1, 2), size=16, B=2, L=5, seed=100, steps=2)
    golden_network('net_config', window=5, hidden_channels=(96, 192, 384), hidden_blocks=(3, 3, 3), size=16, B=1, L=6, seed=200, steps=0)
