r"""Kolmogorov stepper on the GPU against the NumPy oracle (parity unpinned by the reference, see
oracle/kolmogorov_oracle.py) and through the invariants of the scheme.  fp32 tolerance: 1e-4
relative L2 on one transition (82 inner steps at 256 x 256) against the fp64 oracle."""

import random

import numpy as np
import pytest
import torch

from oracle import kolmogorov_oracle as ko

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('size', [64, 128, 256])  # 64 / 256: fused radix-8 / radix-16 path, 128: generic 5-kernel path
def test_transition_matches_oracle(size):
    from sda_b200.mcs import KolmogorovFlow

    chain = KolmogorovFlow(size=size, dt=0.2)
    assert chain.steps == ko.inner_steps(size, 0.2)
    x0 = ko.prior((3,), size, np.random.default_rng(0), np.float32)  # odd ensemble: exercises pair padding
    ref = ko.transition(x0.astype(np.float64), dt=0.2)
    out = chain.transition(torch.from_numpy(x0))
    assert out.device.type == 'cpu' and out.shape == x0.shape  # CPU in -> CPU out, like the reference
    out = out.numpy()
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 1e-4
    assert np.abs(ko.divergence(out.astype(np.float64))).max() < 1e-3
    # CUDA in -> CUDA out, same numbers
    out2 = chain.transition(torch.from_numpy(x0).cuda())
    assert out2.is_cuda and np.array_equal(out2.cpu().numpy(), out)


def test_trajectory_and_members_are_independent():
    from sda_b200.mcs import KolmogorovFlow

    chain = KolmogorovFlow(size=64, dt=0.2)
    x0 = torch.from_numpy(ko.prior((4,), 64, np.random.default_rng(1), np.float32))
    traj = chain.trajectory(x0, length=3)
    assert traj.shape == (3, 4, 2, 64, 64)
    step = x0

    for i in range(3):
        step = chain.transition(step)
        assert torch.equal(step, traj[i])

    assert torch.equal(chain.trajectory(x0, length=3, last=True), traj[-1])
    # a member's evolution depends on its partner in the complex-FFT pair only through rounding
    alone = chain.transition(x0[1:2])
    assert torch.allclose(alone[0], traj[0, 1], rtol=0, atol=2e-5)


def test_prior_distribution_and_seeding():
    from sda_b200.mcs import KolmogorovFlow

    chain = KolmogorovFlow(size=256, dt=0.2)
    random.seed(3)
    a = chain.prior((4,))
    random.seed(3)
    b = chain.prior((4,))
    assert torch.equal(a, b) and a.shape == (4, 2, 256, 256)
    speed = a.square().sum(dim=1).sqrt().amax(dim=(-2, -1))
    assert torch.allclose(speed, torch.full_like(speed, 3.0), atol=1e-3)  # maximum_velocity=3 (mcs.py:301)
    assert np.abs(ko.divergence(a.double().numpy())).max() < 1e-3
    # spectrum peaks near wavenumber 4 (peak_wavenumber=4, mcs.py:302), like the oracle's prior
    def peak(x):
        spec = np.abs(np.fft.fft2(x[:, 0].double().numpy())) ** 2
        k = np.fft.fftfreq(256, 1 / 256)
        kk = np.sqrt(k[:, None] ** 2 + k[None, :] ** 2).round().astype(int)
        e = np.bincount(kk.ravel(), weights=spec.mean(0).ravel())[:32]
        return int(e.argmax())

    ref = torch.from_numpy(ko.prior((4,), 256, np.random.default_rng(0), np.float32))
    assert abs(peak(a) - peak(ref)) <= 1 and 2 <= peak(a) <= 6


def test_observation_kernels(golden):
    from sda_b200 import _lib

    lib = _lib.load()
    g = golden('helpers')
    x = torch.from_numpy(g['x']).cuda()
    out = torch.empty(3, 2, 4, 4, device='cuda')
    _lib.check(lib.sdab_coarsen(x.data_ptr(), out.data_ptr(), 6, 16, 16, 4, _lib.stream_ptr()))
    assert torch.allclose(out.cpu(), torch.from_numpy(g['coarsen4']), atol=1e-6)
    w = torch.empty(3, 16, 16, device='cuda')
    _lib.check(lib.sdab_vorticity(x.data_ptr(), w.data_ptr(), 3, 16, 16, _lib.stream_ptr()))
    assert torch.allclose(w.cpu(), torch.from_numpy(g['vorticity']), atol=1e-6)


def test_generate_then_train_pipeline(tmp_path):
    r"""BASELINE config 4 -> config 5 on a toy scale: tools/generate_kolmogorov.py (the reference's
    simulate + aggregate, experiments/kolmogorov/generate.py:15-53) writes .npy splits that
    sda.utils.TrajectoryDataset reads and sda.utils.loop (sda/utils.py:89-165) trains the native U-Net on."""

    import subprocess
    import sys
    from pathlib import Path

    import sda_b200.score as sc
    from sda_b200.utils import TrajectoryDataset, loop

    root = Path(__file__).resolve().parents[1]
    subprocess.run([sys.executable, str(root / 'tools' / 'generate_kolmogorov.py'), '--out', str(tmp_path), '--members', '10',
                    '--size', '64', '--length', '6', '--keep', '4', '--coarsen', '4'], check=True, timeout=600)
    train = TrajectoryDataset(tmp_path / 'train.npy', window=3, flatten=True)
    valid = TrajectoryDataset(tmp_path / 'valid.npy', window=3, flatten=True)
    assert len(train) == 8 and len(valid) == 1 and train[0][0].shape == (6, 16, 16)
    assert np.isfinite(train.data).all() and 0.05 < train.data.std() < 5.0
    # members are distinct trajectories and each follows the seeding of the reference (random.seed(i))
    assert not np.allclose(train.data[0], train.data[1])

    torch.manual_seed(0)
    kernel = sc.ScoreUNet(6, 0, embedding=32, hidden_channels=(32, 64), hidden_blocks=(1, 1), kernel_size=3,
                          activation=torch.nn.SiLU, spatial=2, padding_mode='circular').cuda()
    sde = sc.VPSDE(kernel, shape=(6, 16, 16)).cuda()
    losses = [lt for lt, lv, lr in loop(sde, train, valid, epochs=6, batch_size=4, learning_rate=2e-3, device='cuda')]
    assert all(np.isfinite(losses)) and min(losses[3:]) < losses[0]


@pytest.mark.parametrize('size', [32, 64])
def test_kernels_reach_the_laminar_kolmogorov_flow(size):
    r"""An anchor that does not go through the (unpinned) oracle: at Re = 10 the forced flow converges to the
    exact steady state u = A sin(4 y), v = 0, A = 1 / (nu |lambda_4| + 0.1) (tests/test_oracle.py).  size = 64
    runs the fused fast path, 32 the generic kernels."""

    import math

    from sda_b200.mcs import KolmogorovFlow

    nu = 0.1
    chain = KolmogorovFlow(size=size, dt=0.2, reynolds=1 / nu)
    x = chain.trajectory(torch.zeros(2, 2, size, size), 60, last=True).cpu().numpy().astype(np.float64)
    h = 2 * math.pi / size
    amp = 1 / (nu * (2 - 2 * math.cos(4 * h)) / h ** 2 + 0.1)
    exact = amp * np.sin(4 * (np.arange(size) + 0.5) * h)[None, :] * np.ones((size, 1))
    assert np.abs(x[:, 0] - exact).max() < 2e-5 * amp
    assert np.abs(x[:, 1]).max() < 2e-5 * amp


def test_observation_operators_values_and_gradients(golden):
    r"""KolmogorovFlow.coarsen / vorticity / upsample on CUDA tensors run the libsdab kernels with analytic
    adjoints: values against the unmodified reference's recorded outputs, values
    and input-gradients against the reference's PyTorch formulas evaluated by autograd on the CPU."""

    import torch.nn.functional as F

    from sda_b200 import _lib
    from sda_b200.mcs import KolmogorovFlow as KF

    g = golden('helpers')
    x = torch.from_numpy(g['x'])
    assert torch.allclose(KF.coarsen(x.cuda(), 4).cpu(), torch.from_numpy(g['coarsen4']), atol=1e-6)
    assert torch.allclose(KF.vorticity(x.cuda()).cpu(), torch.from_numpy(g['vorticity']), atol=1e-6)

    if 'upsample2' in g:
        assert torch.allclose(KF.upsample(x.cuda(), 2).cpu(), torch.from_numpy(g['upsample2']), atol=1e-6)

    def ref_upsample(v, r):  # sda/mcs.py:349-359
        *batch, h, w = v.shape
        v = F.pad(v.reshape(-1, 1, h, w), pad=(1, 1, 1, 1), mode='circular')
        v = F.interpolate(v, scale_factor=(r, r), mode='bilinear')[..., r:-r, r:-r]
        return v.reshape(*batch, r * h, r * w)

    def ref_vorticity(v):  # sda/mcs.py:361-375
        *batch, _, h, w = v.shape
        y = F.pad(v.reshape(-1, 2, h, w), pad=(1, 1, 1, 1), mode='circular')
        (du,) = torch.gradient(y[:, 0], dim=-1)
        (dv,) = torch.gradient(y[:, 1], dim=-2)
        return (du - dv)[:, 1:-1, 1:-1].reshape(*batch, h, w)

    def ref_coarsen(v, r):  # sda/mcs.py:340-347
        *batch, h, w = v.shape
        return v.reshape(*batch, h // r, r, w // r, r).mean(dim=(-3, -1))

    torch.manual_seed(3)
    cases = [
        (lambda v: KF.coarsen(v, 4), lambda v: ref_coarsen(v, 4), (2, 3, 2, 16, 24)),
        (lambda v: KF.coarsen(v[:, ::2], 8), lambda v: ref_coarsen(v[:, ::2], 8), (1, 5, 2, 32, 32)),  # bench.py's A
        (KF.vorticity, ref_vorticity, (3, 2, 12, 20)),
        (lambda v: KF.upsample(v, 2), lambda v: ref_upsample(v, 2), (2, 2, 6, 10)),
        (lambda v: KF.upsample(v, 3), lambda v: ref_upsample(v, 3), (4, 5, 7)),
    ]

    for ours, ref, shape in cases:
        v = torch.randn(shape)
        vr = v.clone().requires_grad_(True)
        out_ref = ref(vr)
        cot = torch.randn_like(out_ref)
        (g_ref,) = torch.autograd.grad(out_ref, vr, cot)
        vc = v.cuda().requires_grad_(True)
        _lib.launch_count(reset=True)
        out = ours(vc)
        (g_ours,) = torch.autograd.grad(out, vc, cot.cuda())
        assert _lib.launch_count() == 2  # one forward and one adjoint kernel: the native path ran
        assert torch.allclose(out.detach().cpu(), out_ref.detach(), atol=2e-6, rtol=1e-5)
        assert torch.allclose(g_ours.cpu(), g_ref, atol=2e-6, rtol=1e-5)
