r"""The score path on the GPU through the public classes (-> ctypes -> C ABI): window maps bit-exact,
U-Net / MCScoreNet / GaussianScore / sampler against the golden vectors of the unmodified reference
and against the CPU oracle on larger seeded inputs.  Bar (north_star): per-step score relative L2
<= 1e-4 in the default bf16x3 mode."""

import os

import numpy as np
import pytest
import torch

from helpers import build_score, build_state
from oracle import score_oracle as so
from oracle.testing import randn, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.mark.parametrize('L,k', [(5, 2), (9, 2), (7, 1), (12, 3)])
def test_window_maps_bit_exact(golden, L, k):
    from sda_b200.score import MCScoreNet

    g = golden('maps')
    B, C, H, W = 2, 2, 2, 3
    x = torch.arange(B * L * C * H * W, dtype=torch.float32).reshape(B, L, C, H, W).cuda().requires_grad_(True)
    u = MCScoreNet.unfold(x, k)
    assert torch.equal(u.cpu().double(), torch.from_numpy(g[f'unfold_L{L}_k{k}']))
    assert torch.equal(MCScoreNet.fold(u, k).cpu().double(), torch.from_numpy(g[f'fold_L{L}_k{k}']))
    gg = torch.from_numpy(g[f'adjoint_g_L{L}_k{k}']).float().cuda()
    (gx,) = torch.autograd.grad((u * gg).sum(), x)
    assert torch.equal(gx.cpu().double(), torch.from_numpy(g[f'adjoint_L{L}_k{k}']))
    # adjoint of fold against autograd of the reference formula on the CPU
    s = torch.randn(B, L - 2 * k, (2 * k + 1) * C, H, W, device='cuda', requires_grad=True)
    go = torch.randn(B, L, C, H, W, device='cuda')
    (a,) = torch.autograd.grad(MCScoreNet.fold(s, k), s, go)
    sr = s.detach().cpu().requires_grad_(True)
    (b,) = torch.autograd.grad(MCScoreNet.fold(sr, k), sr, go.cpu())
    assert torch.equal(a.cpu(), b)


def test_short_trajectory_raises():
    from sda_b200.score import MCScoreNet

    with pytest.raises(RuntimeError):
        MCScoreNet.unfold(torch.zeros(1, 4, 2, 8, 8, device='cuda'), 2)


@pytest.mark.parametrize('name', ['net_small', 'net_config'])
@pytest.mark.parametrize('engine', ['umma', 'simt'])
def test_golden_vectors(golden, name, engine, monkeypatch):
    import sda_b200.score as sc

    monkeypatch.setenv('SDAB_ENGINE', engine)
    g = golden(name)
    score, k = build_score(name, 16, 'cuda')
    x, t = torch.from_numpy(g['x']).cuda(), torch.tensor(float(g['t'])).cuda()

    with torch.no_grad():
        wins = sc.MCScoreNet.unfold(x, k)
        assert rel_l2(score.kernel(wins[:, :1], t), torch.from_numpy(g['kernel_out'])) < TOL
        eps = score(x, t)

    assert rel_l2(eps, torch.from_numpy(g['mc_score'])) < TOL
    assert rel_l2(eps, torch.from_numpy(g['mc_score_fp64'])) < TOL

    A = lambda v: v[..., ::2, ::2]  # noqa: E731
    guided = sc.GaussianScore(torch.from_numpy(g['y']).cuda(), A=A, std=0.1, sde=sc.VPSDE(score, shape=()), gamma=1e-2).cuda()
    assert rel_l2(guided(x, t), torch.from_numpy(g['gaussian_score'])) < TOL

    from sda_b200.mcs import KolmogorovFlow

    A2 = lambda v: KolmogorovFlow.coarsen(v[:, ::2], 4)  # noqa: E731
    guided2 = sc.GaussianScore(torch.from_numpy(g['y2']).cuda(), A=A2, std=0.1, sde=sc.VPSDE(score, shape=()), gamma=1e-2).cuda()
    assert rel_l2(guided2(x, torch.tensor(0.8).cuda()), torch.from_numpy(g['gaussian_score_coarsen'])) < TOL


def test_sampler_steps_with_injected_noise(golden):
    r"""First denoising steps of VPSDE.sample (guided, corrections=1) against the reference's own
    loop, with the reference's recorded corrector noise injected."""

    import sda_b200.score as sc

    g = golden('net_small')
    score, k = build_score('net_small', 16, 'cuda')
    A = lambda v: v[..., ::2, ::2]  # noqa: E731
    guided = sc.GaussianScore(torch.from_numpy(g['y']).cuda(), A=A, std=0.1, sde=sc.VPSDE(score, shape=()), gamma=1e-2)
    sde = sc.VPSDE(guided, shape=tuple(g['x'].shape[1:])).cuda()
    full, n, corr = (int(v) for v in g['sample_meta'])
    noise = iter(torch.from_numpy(g['sample_noise']).cuda())
    sde.noise_source = lambda v: next(noise)
    x = torch.from_numpy(g['sample_x1']).cuda().clone()
    state = sde.sampler_state(x, full)

    for i in range(n):
        x = sde.denoise_step(x, i, state, corrections=corr, tau=0.5)

    assert rel_l2(x, torch.from_numpy(g['sample_after'])) < TOL


def test_sample_is_reproducible_and_finite():
    import sda_b200.score as sc

    score, k = build_score('net_small', 16, 'cuda')
    sde = sc.VPSDE(score, shape=(5, 2, 16, 16)).cuda()
    torch.manual_seed(5)
    a = sde.sample((2,), steps=4, corrections=1, tau=0.5)
    torch.manual_seed(5)
    b = sde.sample((2,), steps=4, corrections=1, tau=0.5)
    assert a.shape == (2, 5, 2, 16, 16) and torch.isfinite(a).all()
    assert torch.equal(a, b)
    torch.manual_seed(6)
    assert not torch.equal(a, sde.sample((2,), steps=4, corrections=1, tau=0.5))


def test_config_network_against_oracle_at_64():
    r"""BASELINE config 2 resolution (64 x 64, window 5, 96/192/384 net): guided score on 4 windows
    against the CPU oracle on the same seeded inputs."""

    import sda_b200.score as sc

    score, k = build_score('net_config', 64, 'cuda')
    state, _ = build_state('net_config', 64)
    x = randn((1, 8, 2, 64, 64), seed=7)
    y = randn((1, 8, 2, 32, 32), seed=8)
    t = torch.tensor(0.63)
    A = lambda v: v[..., ::2, ::2]  # noqa: E731
    ref_eps = so.mc_score(state, x, t, k)
    ref = so.gaussian_score(lambda a, b: so.mc_score(state, a, b, k), y, A, 0.1, x, t, gamma=1e-2)

    with torch.no_grad():
        eps = score(x.cuda(), t.cuda())

    guided = sc.GaussianScore(y.cuda(), A=A, std=0.1, sde=sc.VPSDE(score, shape=()), gamma=1e-2).cuda()
    out = guided(x.cuda(), t.cuda())
    assert rel_l2(eps, ref_eps) < TOL
    assert rel_l2(out, ref) < TOL


def test_batch_invariance_and_per_sample_time():
    r"""A window's score does not depend on which other windows share the launch (what makes the
    multi-GPU sharding bit-reproducible); per-sample t (training-style call) matches per-call t."""

    score, k = build_score('net_small', 32, 'cuda')
    kern = score.kernel
    x = torch.randn(6, 6, 32, 32, device='cuda')
    t = torch.tensor(0.3, device='cuda')

    with torch.no_grad():
        full = kern(x, t)
        parts = torch.cat([kern(x[:1], t), kern(x[1:4], t), kern(x[4:], t)])
        assert torch.equal(full, parts)
        ts = torch.tensor([0.3, 0.7, 0.3, 0.3, 0.9, 0.3], device='cuda')
        per = kern(x, ts)
        # the time embedding is a torch GEMM whose rounding depends on the batch: compare to tolerance
        assert torch.allclose(per[0], full[0], atol=1e-4) and torch.allclose(per[2], full[2], atol=1e-4)
        assert torch.allclose(per[1], kern(x[1:2], torch.tensor(0.7, device='cuda'))[0], atol=1e-4)


def test_input_gradient_adjoint_identity_at_256():
    r"""Full BASELINE resolution (256 x 256): <J v, g> == <v, J^T g> with J v from a central difference
    of the forward pass -- a size-independent check of dgrad."""

    score, k = build_score('net_config', 256, 'cuda')
    kern = score.kernel
    torch.manual_seed(0)
    x = torch.randn(1, 10, 256, 256, device='cuda')
    v = torch.randn_like(x)
    g = torch.randn_like(x)
    t = torch.tensor(0.5, device='cuda')
    xr = x.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad(kern(xr, t), xr, g)
    h = 1e-2

    with torch.no_grad():
        jv = (kern(x + h * v, t) - kern(x - h * v, t)) / (2 * h)

    lhs, rhs = float((jv.double() * g.double()).sum()), float((v.double() * gx.double()).sum())
    assert abs(lhs - rhs) <= 2e-3 * max(abs(lhs), abs(rhs))


@pytest.mark.parametrize('name, size, batch', [('net_small', 16, 3), ('net_config', 32, 2), ('net_config', 64, 1)])
def test_parameter_gradients_match_oracle(name, size, batch):
    r"""Training path: d loss / d (every parameter) and d loss / d x of
    the kernel with per-sample times, against torch.autograd through the fp64 oracle.  Tolerance:
    2e-4 relative L2 per tensor (fp32 CUDA-core weight gradients over bf16x3 activations)."""

    from oracle import score_oracle as so

    score, k = build_score(name, size, 'cuda')
    kern = score.kernel.train()
    state, _ = build_state(name, size)
    C = (2 * k + 1) * 2
    x = randn((batch, C, size, size), seed=11)
    t = torch.linspace(0.2, 0.9, batch)
    r = randn((batch, C, size, size), seed=12)

    ref_state = {kk: v.double().requires_grad_(v.is_floating_point() and kk != 'forcing') for kk, v in state.items()}
    xr = x.double().requires_grad_(True)
    ref_loss = (so.score_unet(ref_state, xr, t.double(), ref_state['forcing']) * r.double()).sum()
    ref_loss.backward()

    xg = x.cuda().requires_grad_(True)
    loss = (kern(xg, t.cuda()) * r.cuda()).sum()
    loss.backward()
    assert abs(float(loss.detach()) - float(ref_loss.detach())) <= 1e-4 * abs(float(ref_loss.detach())) + 1e-3
    assert rel_l2(xg.grad, xr.grad) < 2e-4
    worst = 0.0

    for key, p in kern.named_parameters():
        ref = ref_state[key].grad

        if ref is None:  # embedding.freqs is a buffer in the oracle state
            continue

        assert p.grad is not None, key
        err = rel_l2(p.grad, ref)
        worst = max(worst, err)
        assert err < 2e-4, f'{key}: rel-L2 {err:.2e}'

    assert worst > 0.0


def test_parameter_gradients_on_the_cuda_core_engine(monkeypatch):
    r"""SDAB_ENGINE=simt: fp32 CUDA-core convolutions AND weight gradients (every operand loader of
    csrc/wgrad.cu: stride-2 parity layout, tails at the high resolution) against the tensor-core engine,
    with a scalar time (one shared shift row)."""

    score, k = build_score('net_small', 16, 'cuda')
    kern = score.kernel.train()
    x = randn((2, 6, 16, 16), seed=21).cuda()
    r = randn((2, 6, 16, 16), seed=22).cuda()
    t = torch.tensor(0.4, device='cuda')
    grads = {}

    for engine in ('umma', 'simt'):
        monkeypatch.setenv('SDAB_ENGINE', engine)
        kern.zero_grad()
        (kern(x, t) * r).sum().backward()
        grads[engine] = {key: p.grad.clone() for key, p in kern.named_parameters()}

    for key in grads['umma']:
        assert rel_l2(grads['simt'][key], grads['umma'][key]) < 1e-4, key


def test_vpsde_loss_trains_through_the_native_unet():
    r"""VPSDE.loss (sda/score.py:265-276) + one AdamW step (sda/utils.py:136-143) lowers the loss."""

    import sda_b200.score as sc

    score, k = build_score('net_small', 16, 'cuda')
    sde = sc.VPSDE(score.kernel, shape=(6, 16, 16)).cuda().train()
    opt = torch.optim.AdamW(sde.parameters(), lr=1e-3)
    x = randn((8, 6, 16, 16), seed=5).cuda()
    torch.manual_seed(0)
    losses = []

    for _ in range(12):
        torch.manual_seed(1)  # same noise and times every step: the loss must go down
        l = sde.loss(x)
        opt.zero_grad()
        l.backward()
        opt.step()
        losses.append(float(l))

    assert all(map(lambda v: v == v, losses)) and losses[-1] < 0.9 * losses[0]


def test_guidance_stays_on_the_input_gradient_path():
    r"""GaussianScore differentiates w.r.t. x only although the parameters require gradients:
    no parameter gradient may be produced (or paid for) during guided sampling."""

    import sda_b200.score as sc

    score, k = build_score('net_small', 16, 'cuda')
    x = randn((1, 5, 2, 16, 16), seed=1).cuda()
    y = randn((1, 5, 2, 8, 8), seed=2).cuda()
    guided = sc.GaussianScore(y, A=lambda v: v[..., ::2, ::2], std=0.1, sde=sc.VPSDE(score, shape=())).cuda()
    guided(x, torch.tensor(0.5).cuda())
    assert score.kernel.network._saved_level == 1
    assert all(p.grad is None for p in score.parameters())


def test_fast_mode_error_is_reported_not_hidden(golden, monkeypatch):
    g = golden('net_config')
    score, k = build_score('net_config', 16, 'cuda')
    x, t = torch.from_numpy(g['x']).cuda(), torch.tensor(float(g['t'])).cuda()
    monkeypatch.setenv('SDAB_MODE', 'bf16')

    with torch.no_grad():
        err = rel_l2(score(x, t), torch.from_numpy(g['mc_score_fp64']))

    assert 1e-4 < err < 3e-2  # single-pass bf16 does NOT meet the parity bar; it is opt-in only


def test_guided_score_at_256_matches_oracle():
    r"""The BENCHMARKED configuration (bench.py: 256 x 256, window 5, 96/192/384 net, every 4th frame
    coarsened x8 observed, std 0.1, gamma 1e-2) on a 2-window slice (L = 6): guided and unguided
    score of the CUDA path against the CPU oracle in fp32 and in fp64, rel-L2 <= 1e-4."""

    import sda_b200.score as sc
    from sda_b200.mcs import KolmogorovFlow

    size, L = 256, 6
    score, k = build_score('net_config', size, 'cuda')
    state, _ = build_state('net_config', size)
    x = randn((1, L, 2, size, size), seed=31)
    y = randn((1, (L + 3) // 4, 2, size // 8, size // 8), seed=32)
    t = torch.tensor(0.37)
    A = lambda v: KolmogorovFlow.coarsen(v[:, ::4], 8)  # noqa: E731  (bench.py observation)
    A_ref = lambda v: so.coarsen(v[:, ::4], 8)  # noqa: E731

    with torch.no_grad():
        eps = score(x.cuda(), t.cuda()).cpu()

    guided = sc.GaussianScore(y.cuda(), A=A, std=0.1, sde=sc.VPSDE(score, shape=()), gamma=1e-2).cuda()
    out = guided(x.cuda(), t.cuda()).cpu()

    for dtype in (torch.float32, torch.float64):
        st = {kk: (v.to(dtype) if v.is_floating_point() else v) for kk, v in state.items()}
        xd, yd, td = x.to(dtype), y.to(dtype), t.to(dtype)
        ref_eps = so.mc_score(st, xd, td, k)
        ref = so.gaussian_score(lambda a, b: so.mc_score(st, a, b, k), yd, A_ref, 0.1, xd, td, gamma=1e-2)
        assert rel_l2(eps, ref_eps) < TOL, dtype
        assert rel_l2(out, ref) < TOL, dtype


def test_detached_guidance_matches_oracle():
    r"""GaussianScore(detach=True) (sda/score.py:378-379): no back-propagation through the network."""

    import sda_b200.score as sc

    score, k = build_score('net_small', 32, 'cuda')
    state, _ = build_state('net_small', 32)
    x = randn((2, 6, 2, 32, 32), seed=41)
    y = randn((2, 6, 2, 16, 16), seed=42)
    t = torch.tensor(0.55)
    A = lambda v: v[..., ::2, ::2]  # noqa: E731
    ref = so.gaussian_score(lambda a, b: so.mc_score(state, a, b, k), y, A, 0.1, x, t, gamma=1e-2, detach=True)
    guided = sc.GaussianScore(y.cuda(), A=A, std=0.1, sde=sc.VPSDE(score, shape=()), gamma=1e-2, detach=True).cuda()

    with torch.no_grad():  # as VPSDE.sample calls it (sda/score.py:249)
        out = guided(x.cuda(), t.cuda())

    assert rel_l2(out, ref) < TOL
    # nothing was saved for a backward that can never come
    assert score.kernel.network._saved_level == 0


def test_no_grad_forward_saves_nothing():
    r"""Unguided sampling / validation run under torch.no_grad() with trainable parameters: the native
    forward must not save activations (needs_input_grad ignores grad mode)."""

    score, k = build_score('net_small', 16, 'cuda')
    assert all(p.requires_grad for p in score.parameters())
    x = randn((1, 5, 2, 16, 16), seed=1).cuda()

    with torch.no_grad():
        score(x, torch.tensor(0.5).cuda())

    assert score.kernel.network._saved_level == 0
    score(x.requires_grad_(True), torch.tensor(0.5).cuda())
    assert score.kernel.network._saved_level == 2  # grad mode on, parameters trainable: training state


@pytest.mark.parametrize('kind', ['sub', 'subsub'])
def test_sub_vpsde_sampler_steps_match_oracle(kind):
    r"""SubVPSDE / SubSubVPSDE (sda/score.py:279-300) through the CUDA sampler: two predictor-corrector
    steps with injected corrector noise against the oracle loop with the same sigma(t)."""

    import sda_b200.score as sc

    score, k = build_score('net_small', 16, 'cuda')
    state, _ = build_state('net_small', 16)
    cls = {'sub': sc.SubVPSDE, 'subsub': sc.SubSubVPSDE}[kind]
    sde = cls(score, shape=(5, 2, 16, 16)).cuda()
    x0 = randn((1, 5, 2, 16, 16), seed=51)
    noise = [randn((1, 5, 2, 16, 16), seed=60 + i) for i in range(2)]
    steps = 16

    # oracle loop (sda/score.py:246-261) with this SDE's sigma
    def sig(tt):
        return so.sigma(tt, sde={'sub': 'subvp', 'subsub': 'subsubvp'}[kind])

    xr = x0.clone()
    time = torch.linspace(1, 0, steps + 1)
    dt = 1 / steps

    with torch.no_grad():
        for i in range(2):
            tt = time[i]
            r = so.mu(tt - dt) / so.mu(tt)
            xr = r * xr + (sig(tt - dt) - r * sig(tt)) * so.mc_score(state, xr, tt, k)
            e = so.mc_score(state, xr, tt - dt, k)
            delta = 0.5 / e.square().mean(dim=(-4, -3, -2, -1), keepdim=True)
            xr = xr - (delta * e + torch.sqrt(2 * delta) * noise[i]) * sig(tt - dt)

    it = iter(n.cuda() for n in noise)
    sde.noise_source = lambda v: next(it)
    x = x0.cuda().clone()
    st = sde.sampler_state(x, steps)

    for i in range(2):
        x = sde.denoise_step(x, i, st, corrections=1, tau=0.5)

    assert rel_l2(x, xr) < TOL


def test_dps_guidance_matches_the_reference_formula():
    r"""DPSGaussianScore.forward(x, t) (sda/score.py:331-344) called directly, against the same formula
    evaluated by autograd through the CPU oracle."""

    import sda_b200.score as sc

    score, k = build_score('net_small', 16, 'cuda')
    state, _ = build_state('net_small', 16)
    x = randn((1, 6, 2, 16, 16), seed=71)
    y = randn((1, 6, 2, 8, 8), seed=72)
    t = torch.tensor(0.4)
    A = lambda v: v[..., ::2, ::2]  # noqa: E731
    zeta = 0.7

    m, s = so.mu(t), so.sigma(t)
    xr = x.clone().requires_grad_(True)
    eps = so.mc_score(state, xr, t, k)
    err = (y - A((xr - s * eps) / m)).square().sum()
    (g,) = torch.autograd.grad(err, xr)
    ref = (eps - s * (-g * zeta / err.sqrt())).detach()

    dps = sc.DPSGaussianScore(y.cuda(), A=A, sde=sc.VPSDE(score, shape=()), zeta=zeta).cuda()
    assert rel_l2(dps(x.cuda(), t.cuda()), ref) < TOL


@pytest.mark.parametrize('H, W, activation, channels, blocks', [
    (32, 64, 'SiLU', (32, 64), (1, 2)),      # non-square images
    (64, 16, 'SiLU', (64, 32, 96), (2, 1, 1)),  # non-monotone widths, three levels, tall images
    (32, 32, 'ReLU', (32, 64), (2, 1)),      # ReLU blocks
])
def test_network_variants_against_oracle(H, W, activation, channels, blocks):
    r"""Constructor arguments beyond the two golden configurations (sda/nn.py:94-182): non-square images,
    ReLU, uneven block counts -- guided score and unguided score against the oracle."""

    import sda_b200.score as sc
    from oracle.testing import fill_state_

    k, C = 1, 2
    score = sc.MCScoreNet(C, order=k)
    score.kernel = sc.ScoreUNet(
        (2 * k + 1) * C, 0, embedding=32, hidden_channels=channels, hidden_blocks=blocks, kernel_size=3,
        activation=getattr(torch.nn, activation), spatial=2, padding_mode='circular',
    )
    fill_state_(score.state_dict(), seed=77)
    state = {kk[len('kernel.'):]: v.clone() for kk, v in score.state_dict().items()}
    score = score.cuda()
    x = randn((2, 5, C, H, W), seed=81)
    y = randn((2, 5, C, H // 2, W // 2), seed=82)
    t = torch.tensor(0.52)
    A = lambda v: v[..., ::2, ::2]  # noqa: E731
    mc = lambda a, b: so.fold(so.score_unet(state, so.unfold(a, k).flatten(0, 1), b, None, activation=activation)  # noqa: E731
                              .unflatten(0, (a.shape[0], -1)), k)
    ref_eps = mc(x, t)
    ref = so.gaussian_score(mc, y, A, 0.1, x, t, gamma=1e-2)

    with torch.no_grad():
        eps = score(x.cuda(), t.cuda())

    out = sc.GaussianScore(y.cuda(), A=A, std=0.1, sde=sc.VPSDE(score, shape=()), gamma=1e-2).cuda()(x.cuda(), t.cuda())
    assert rel_l2(eps, ref_eps) < TOL
    # ReLU's derivative is a step: a pre-activation within the 5e-6 forward error of zero flips its whole
    # gradient contribution (a fraction ~5e-6 of the units, i.e. ~sqrt(5e-6) in relative L2), so the guided score of
    # a ReLU network is compared at 5e-3.  The Kolmogorov configuration of the reference uses SiLU (smooth).
    assert rel_l2(out, ref) < (5e-3 if activation == 'ReLU' else TOL)


def test_packed_weights_follow_parameter_updates():
    r"""The bf16-packed weight copies are rebuilt after optimizer steps / load_state_dict / .to() on their own and
    after `.data` surgery once `invalidate_packed()` is called (`.data` updates do not bump the
    version counter the cache is keyed on)."""

    score, k = build_score('net_small', 16, 'cuda')
    net = score.kernel.network
    x = randn((1, 5, 2, 16, 16), seed=1).cuda()
    t = torch.tensor(0.5).cuda()

    with torch.no_grad():
        base = score(x, t).clone()
        first = net.heads[0].weight
        first.mul_(2.0)                       # in-place op on the parameter: version bump -> repacked
        bumped = score(x, t).clone()
        assert not torch.equal(base, bumped)
        first.data.mul_(0.5)                  # .data surgery (exact inverse): invisible to the cache key ...
        net.invalidate_packed()               # ... until the documented call
        restored = score(x, t).clone()
        assert torch.equal(restored, base)
        state = {kk: v.clone() for kk, v in score.state_dict().items()}
        state['kernel.network.heads.0.weight'] = state['kernel.network.heads.0.weight'] * 2
        score.load_state_dict(state)          # load_state_dict invalidates by itself
        assert not torch.equal(score(x, t), restored)


def test_packed_weights_follow_fused_optimizers():
    r"""torch.optim.AdamW(fused=True) updates parameters without bumping their version counters; the packed
    weights must follow anyway (the cache key counts optimizer steps): fused and default AdamW train alike."""

    import copy

    import sda_b200.score as sc

    score, _ = build_score('net_small', 16, 'cuda')
    sde_a = sc.VPSDE(score.kernel, shape=(6, 16, 16)).cuda().train()
    sde_b = copy.deepcopy(sde_a)
    opt_a = torch.optim.AdamW(sde_a.parameters(), lr=1e-3)
    opt_b = torch.optim.AdamW(sde_b.parameters(), lr=1e-3, fused=True)
    x = randn((8, 6, 16, 16), seed=4).cuda()
    losses = []

    for it in range(4):
        pair = []

        for sde_, opt_ in ((sde_a, opt_a), (sde_b, opt_b)):
            torch.manual_seed(it)
            l = sde_.loss(x)
            opt_.zero_grad()
            l.backward()
            opt_.step()
            pair.append(float(l.detach()))

        losses.append(pair)

    # with stale packed weights the fused run would see the initial network at every step
    assert all(abs(a - b) <= 1e-3 * abs(a) for a, b in losses), losses
    assert losses[-1][1] != losses[0][1]


def test_peer_adamw_matches_torch_adamw():
    r"""sda_b200.parallel.PeerAdamW (one GPU: the fused AdamW kernel over the flat parameter buffer) follows
    torch.optim.AdamW, the optimizer of the reference's training loop (sda/utils.py:125-143), step by step."""

    import copy

    import sda_b200.score as sc
    from sda_b200.parallel import PeerAdamW

    score, _ = build_score('net_small', 16, 'cuda')
    sde_a = sc.VPSDE(score.kernel, shape=(6, 16, 16)).cuda().train()
    sde_b = copy.deepcopy(sde_a)
    opt_a = torch.optim.AdamW(sde_a.parameters(), lr=1e-3, weight_decay=1e-2)
    opt_b = PeerAdamW(sde_b, lr=1e-3, weight_decay=1e-2)
    x = randn((8, 6, 16, 16), seed=3).cuda()

    for it in range(3):
        torch.manual_seed(it)
        la = sde_a.loss(x)
        opt_a.zero_grad()
        la.backward()
        opt_a.step()
        torch.manual_seed(it)
        lb = sde_b.loss(x)
        opt_b.zero_grad()
        lb.backward()
        opt_b.step()
        assert abs(float(la) - float(lb)) <= 1e-4 * abs(float(la)), (it, float(la), float(lb))

    for (name, pa), (_, pb) in zip(sde_a.named_parameters(), sde_b.named_parameters()):
        assert rel_l2(pb, pa) < 2e-5, name
