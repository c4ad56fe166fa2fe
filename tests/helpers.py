r"""Shared test helpers: the two golden network configurations, built on our classes or as a
bare state_dict for the oracle, with weights from oracle.testing.fill_state_ (keyed by name)."""

import torch

from oracle.testing import fill_state_

CONFIGS = {
    # name: (window, hidden_channels, hidden_blocks, seed)     (tests/golden/make_golden.py)
    'net_small': (3, (32, 64), (1, 2), 100),
    'net_config': (5, (96, 192, 384), (3, 3, 3), 200),
}


def build_score(name, size, device='cpu'):
    r"""MCScoreNet with a LocalScoreUNet-style kernel (experiments/kolmogorov/utils.py:29-70)."""

    import sda_b200.score as sc

    window, channels, blocks, seed = CONFIGS[name]

    class LocalScoreUNet(sc.ScoreUNet):
        def __init__(self, channels, size=64, **kwargs):
            super().__init__(channels, 1, **kwargs)
            domain = 2 * torch.pi / size * (torch.arange(size) + 1 / 2)
            self.register_buffer('forcing', torch.sin(4 * domain).expand(1, size, size).clone())

        def forward(self, x, t, c=None):
            return super().forward(x, t, self.forcing)

    score = sc.MCScoreNet(2, order=window // 2)
    score.kernel = LocalScoreUNet(
        window * 2, size, embedding=64, hidden_channels=channels, hidden_blocks=blocks, kernel_size=3,
        activation=torch.nn.SiLU, spatial=2, padding_mode='circular',
    )
    fill_state_(score.state_dict(), seed=seed)

    return score.to(device), window // 2


def build_state(name, size):
    r"""The kernel's state_dict (keys embedding.*, network.*, forcing) for the oracle."""

    score, k = build_score(name, size)

    return {kk[len('kernel.'):]: v.clone() for kk, v in score.state_dict().items()}, k
