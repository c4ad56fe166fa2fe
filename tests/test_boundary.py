r"""Drop-in boundary: class surface, constructor signatures, state_dict
keys, and -- when /root/reference is present -- the reference's own experiments/ helper modules
running UNCHANGED on top of this repo's `sda` alias package.  CPU only."""

import inspect
import json
from pathlib import Path

import pytest
import torch

import sda
import sda_b200
from oracle import ref_import

ROOT = Path(__file__).resolve().parents[1]


def test_alias_package_exposes_the_reference_surface():
    assert sda.score is sda_b200.score and sda.nn is sda_b200.nn and sda.mcs is sda_b200.mcs
    from sda.score import (DPSGaussianScore, GaussianScore, MCScoreNet, MCScoreWrapper, ScoreNet, ScoreUNet,  # noqa: F401
                           SubSubVPSDE, SubVPSDE, TimeEmbedding, VPSDE)
    from sda.nn import ModResidualBlock, ResidualBlock, ResMLP, UNet  # noqa: F401
    from sda.mcs import KolmogorovFlow, Lorenz63, MarkovChain, NoisyLorenz63  # noqa: F401
    from sda.utils import ACTIVATIONS, TrajectoryDataset, load_config, loop, save_config, to  # noqa: F401


def test_signatures_match_the_reference_contract():
    sig = lambda f: list(inspect.signature(f).parameters)  # noqa: E731
    sc, nn_, mcs = sda_b200.score, sda_b200.nn, sda_b200.mcs
    assert sig(sc.MCScoreNet.__init__) == ['self', 'features', 'context', 'order', 'kwargs']
    assert sig(sc.MCScoreNet.forward) == ['self', 'x', 't', 'c']
    assert sig(sc.ScoreUNet.__init__) == ['self', 'channels', 'context', 'embedding', 'kwargs']
    assert sig(sc.VPSDE.__init__) == ['self', 'eps', 'shape', 'alpha', 'eta']
    assert sig(sc.VPSDE.sample) == ['self', 'shape', 'c', 'steps', 'corrections', 'tau']
    assert sig(sc.VPSDE.loss) == ['self', 'x', 'c', 'w']
    assert sig(sc.GaussianScore.__init__) == ['self', 'y', 'A', 'std', 'sde', 'gamma', 'detach']
    assert sig(sc.DPSGaussianScore.forward) == ['self', 'x', 't']
    assert sig(nn_.UNet.__init__) == ['self', 'in_channels', 'out_channels', 'mod_features', 'hidden_channels',
                                      'hidden_blocks', 'kernel_size', 'stride', 'activation', 'spatial', 'kwargs']
    assert sig(mcs.KolmogorovFlow.__init__)[:4] == ['self', 'size', 'dt', 'reynolds']
    assert sig(mcs.MarkovChain.trajectory) == ['self', 'x', 'length', 'last']


def test_state_dict_keys_match_the_reference_fixture():
    r"""tests/golden/state_keys.json was written from the reference's make_score(CONFIG)."""

    from helpers import build_score

    score, _ = build_score('net_config', 64)
    expected = json.loads((ROOT / 'tests' / 'golden' / 'state_keys.json').read_text())
    ours = {k: list(v.shape) for k, v in score.state_dict().items()}
    assert list(ours) == list(expected)
    assert ours == expected


def test_vpsde_device_buffer_and_schedule(golden):
    sde = sda_b200.score.VPSDE(None, shape=())
    assert 'device' in dict(sde.named_buffers())
    g = golden('schedule')
    t = torch.from_numpy(g['t'])
    assert torch.equal(sde.mu(t), torch.from_numpy(g['mu_cos']))
    assert torch.equal(sde.sigma(t), torch.from_numpy(g['sigma_cos']))
    assert torch.equal(sda_b200.score.SubVPSDE(None, shape=()).sigma(t), torch.from_numpy(g['sigma_subvp']))
    assert torch.equal(sda_b200.score.SubSubVPSDE(None, shape=(), alpha='lin').mu(t), torch.from_numpy(g['mu_lin']))


def test_observation_helpers_match_reference_vectors(golden):
    g = golden('helpers')
    x = torch.from_numpy(g['x'])
    K = sda_b200.mcs.KolmogorovFlow
    assert torch.allclose(K.coarsen(x, 2), torch.from_numpy(g['coarsen2']), atol=1e-7)
    assert torch.allclose(K.coarsen(x, 4), torch.from_numpy(g['coarsen4']), atol=1e-7)
    assert torch.allclose(K.vorticity(x), torch.from_numpy(g['vorticity']), atol=1e-6)
    assert K.upsample(x, 2).shape == (3, 2, 32, 32)


def test_cpu_window_maps_are_the_reference_views(golden):
    g = golden('maps')
    x = torch.arange(2 * 9 * 2 * 2 * 3, dtype=torch.float64).reshape(2, 9, 2, 2, 3)
    u = sda_b200.score.MCScoreNet.unfold(x, 2)
    assert torch.equal(u, torch.from_numpy(g['unfold_L9_k2']))
    assert torch.equal(sda_b200.score.MCScoreNet.fold(u, 2), torch.from_numpy(g['fold_L9_k2']))

    with pytest.raises(RuntimeError):
        sda_b200.score.MCScoreNet.unfold(torch.zeros(1, 4, 2, 2, 2), 2)


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference only exists in the build container')
def test_reference_experiment_helpers_run_unchanged():
    exp = ref_import.import_experiments(sda, 'kolmogorov')
    score = exp.make_score(window=5, embedding=64, hidden_channels=(96, 192, 384), hidden_blocks=(3, 3, 3),
                           kernel_size=3, activation='SiLU')
    assert type(score).__module__ == 'sda_b200.score'
    assert type(score.kernel).__name__ == 'LocalScoreUNet' and score.kernel.network._native

    ref = ref_import.import_reference()
    rexp = ref_import.import_experiments(ref, 'kolmogorov')
    theirs = rexp.make_score(window=5, embedding=64, hidden_channels=(96, 192, 384), hidden_blocks=(3, 3, 3),
                             kernel_size=3, activation='SiLU')
    a, b = score.state_dict(), theirs.state_dict()
    assert list(a) == list(b) and all(a[k].shape == b[k].shape for k in a)
    score.load_state_dict(b)  # strict

    lor = ref_import.import_experiments(sda, 'lorenz')
    local = lor.make_local_score(window=5, embedding=32, width=64, depth=2)
    glob = lor.make_global_score(embedding=32, hidden_channels=(64,), hidden_blocks=(3,))
    x = torch.randn(2, 16, 3)
    assert local(x, torch.tensor(0.5)).shape == x.shape
    assert glob(x, torch.tensor(0.5)).shape == x.shape  # 1-D U-Net: plain PyTorch plumbing
