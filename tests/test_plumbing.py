r"""BASELINE config 1: Lorenz-63, L=64, VPSDE posterior sampling on CPU -- API plumbing only (no GPU
kernel is involved: ResMLP score, plain PyTorch).  Compared with the oracle's sampler on the same
weights and injected noise."""

import torch

import sda_b200.mcs as mcs
import sda_b200.score as sc
from oracle import score_oracle as so
from oracle.testing import rel_l2


def test_lorenz_chain_and_local_score_sampling():
    torch.manual_seed(0)
    chain = mcs.NoisyLorenz63(dt=0.025)
    x = chain.prior((4,))
    traj = chain.trajectory(x, length=8)
    assert traj.shape == (8, 4, 3) and torch.isfinite(traj).all()
    assert torch.allclose(chain.postprocess(chain.preprocess(traj)), traj, atol=1e-4)
    assert chain.log_prob(traj[0], traj[1]).shape == (4,)

    score = sc.MCScoreNet(3, order=2, embedding=32, hidden_features=[64] * 3, activation=torch.nn.SiLU)
    y = torch.randn(8, 1)
    A = lambda v: v[..., ::8, :1]  # noqa: E731
    sde = sc.VPSDE(sc.GaussianScore(y, A=A, std=0.05, sde=sc.VPSDE(score, shape=()), gamma=3e-2), shape=(64, 3))

    drawn = []
    real = torch.randn_like

    def source(v):
        z = real(v)
        drawn.append(z)
        return z

    sde.noise_source = source
    torch.manual_seed(1)
    out = sde.sample((), steps=8, corrections=1, tau=0.25)
    assert out.shape == (64, 3) and torch.isfinite(out).all()

    # same run through the oracle's restatement of the sampler (score.py:246-261)
    torch.manual_seed(1)
    x1 = torch.randn(1, 64, 3)
    ref = so.pc_sample(
        lambda a, b: so.gaussian_score(lambda c, d: score(c, d), y, A, 0.05, a, b, gamma=3e-2),
        x1, steps=8, corrections=1, tau=0.25, noise=drawn, event_dims=2,
    )
    assert rel_l2(out, ref[0]) < 1e-4


def test_score_wrapper_and_loss_on_cpu():
    net = sc.MCScoreWrapper(sc.ScoreUNet(channels=3, embedding=32, hidden_channels=(32,), hidden_blocks=(2,),
                                         activation=torch.nn.SiLU, spatial=1))
    x = torch.randn(5, 16, 3)
    assert net(x, torch.tensor(0.3)).shape == x.shape
    sde = sc.VPSDE(net, shape=(16, 3))
    loss = sde.loss(x)
    loss.backward()  # plain PyTorch path trains normally
    assert torch.isfinite(loss) and any(p.grad is not None for p in net.parameters())
