r"""Helpers -- host glue around the hot path (reference: /root/reference/sda/utils.py).

Only what the hot path's callers need is provided: the activation table, config JSON
round trip, device mover, the windowed trajectory dataset and the training loop
generator.  The Lorenz evaluation metrics of the reference (bpf / emd / mmd,
sda/utils.py:168-263) are out of scope.
"""

from __future__ import annotations

import json
import math
from pathlib import Path
from typing import Any, Dict, Iterator, Tuple

import numpy as np
import torch
from torch import Tensor
from torch.utils.data import DataLoader, Dataset

from .score import *  # noqa: F401,F403  (the reference re-exports sda.score from sda.utils)
from .score import VPSDE

ACTIVATIONS = {
    'ReLU': torch.nn.ReLU,
    'ELU': torch.nn.ELU,
    'GELU': torch.nn.GELU,
    'SELU': torch.nn.SELU,
    'SiLU': torch.nn.SiLU,
}


def save_config(config: Dict[str, Any], path: Path) -> None:
    r"""Writes `path/config.json`; refuses to overwrite.  Reference: sda/utils.py:35-37."""

    with open(Path(path) / 'config.json', mode='x') as f:
        json.dump(config, f)


def load_config(path: Path) -> Dict[str, Any]:
    r"""Reference: sda/utils.py:40-42."""

    with open(Path(path) / 'config.json', mode='r') as f:
        return json.load(f)


def to(x: Any, **kwargs) -> Any:
    r"""Recursively moves tensors in lists / tuples / dicts.  Reference: sda/utils.py:45-55."""

    if torch.is_tensor(x):
        return x.to(**kwargs)

    if isinstance(x, (list, tuple)):
        return type(x)(to(y, **kwargs) for y in x)

    if isinstance(x, dict):
        return {k: to(v, **kwargs) for k, v in x.items()}

    return x


class TrajectoryDataset(Dataset):
    r"""Random windows of stored trajectories.  Reference: sda/utils.py:58-86.

    Reads the reference's HDF5 layout (dataset 'x' of shape (n, L, ...)) when h5py is
    installed, or a `.npy` file of the same array (h5py is absent from this image).
    """

    def __init__(self, file: Path, window: int = None, flatten: bool = False):
        super().__init__()

        file = Path(file)

        if file.suffix == '.npy':
            self.data = np.load(file)
        else:
            import h5py  # raises if unavailable: there is no silent substitute for an .h5 file

            with h5py.File(file, mode='r') as f:
                self.data = f['x'][:]

        self.window = window
        self.flatten = flatten

    def __len__(self) -> int:
        return len(self.data)

    def __getitem__(self, i: int) -> Tuple[Tensor, Dict]:
        x = torch.from_numpy(self.data[i])

        if self.window is not None:
            start = torch.randint(0, len(x) - self.window + 1, size=())
            x = torch.narrow(x, dim=0, start=start, length=self.window)

        return (x.flatten(0, 1) if self.flatten else x), {}


def loop(
    sde: VPSDE,
    trainset: Dataset,
    validset: Dataset,
    epochs: int = 256,
    batch_size: int = 64,
    optimizer: str = 'AdamW',
    learning_rate: float = 1e-3,
    weight_decay: float = 1e-3,
    scheduler: float = 'linear',
    device: str = 'cpu',
    **absorb,
) -> Iterator:
    r"""Training loop generator yielding (train loss, valid loss, lr) per epoch.

    Reference: sda/utils.py:89-165.  The libsdab U-Net trains through sdab_unet_backward
    (tensor-core forward / input-gradient, fp32 CUDA-core weight gradients).
    """

    loaders = [
        DataLoader(ds, batch_size=batch_size, shuffle=True, num_workers=1, persistent_workers=True)
        for ds in (trainset, validset)
    ]

    if optimizer != 'AdamW':
        raise ValueError()

    opt = torch.optim.AdamW(sde.parameters(), lr=learning_rate, weight_decay=weight_decay)

    schedules = {
        'linear': lambda t: 1 - (t / epochs),
        'cosine': lambda t: (1 + math.cos(math.pi * t / epochs)) / 2,
        'exponential': lambda t: math.exp(-7 * (t / epochs) ** 2),
    }

    if scheduler not in schedules:
        raise ValueError()

    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=schedules[scheduler])

    for _ in range(epochs):
        sde.train()
        train_losses = []

        for batch in loaders[0]:
            x, kwargs = to(batch, device=device)
            loss = sde.loss(x, **kwargs)
            loss.backward()
            opt.step()
            opt.zero_grad()
            train_losses.append(loss.detach())

        sde.eval()
        valid_losses = []

        with torch.no_grad():
            for batch in loaders[1]:
                x, kwargs = to(batch, device=device)
                valid_losses.append(sde.loss(x, **kwargs))

        yield (
            torch.stack(train_losses).mean().item(),
            torch.stack(valid_losses).mean().item(),
            opt.param_groups[0]['lr'],
        )

        sched.step()


# --------------------------------------------------------------------------- Lorenz evaluation helpers
# Host-side glue of experiments/lorenz/*.py (`from sda.utils import *`), outside the accelerated path
# (out of scope as kernel targets); plain PyTorch so that those scripts import and run.
def random_config(configs: Dict[str, Any]) -> Dict[str, Any]:
    r"""One random choice per key.  Reference: sda/utils.py:28-32."""

    import random

    return {key: random.choice(list(values)) for key, values in configs.items()}


def bpf(x: Tensor, y: Tensor, transition, likelihood, step: int = 1) -> Tensor:
    r"""Bootstrap particle filter: (M, *) initial particles, (N, *) observations -> (M, N * step + 1, *)
    resampled trajectories.  Reference: sda/utils.py:168-202."""

    paths = x.unsqueeze(1)

    for obs in y:
        for _ in range(step):
            paths = torch.cat((paths, transition(paths[:, -1]).unsqueeze(1)), dim=1)

        weights = likelihood(obs, paths[:, -1])
        paths = paths[torch.multinomial(weights, weights.shape[0], replacement=True)]

    return paths


def emd(x: Tensor, y: Tensor) -> Tensor:
    r"""Earth mover's distance between two sample sets (POT, imported on use).  Reference: sda/utils.py:205-223."""

    import ot  # optional dependency of the Lorenz evaluation only

    return ot.emd2(x.new_tensor(()), y.new_tensor(()), torch.cdist(x.flatten(1), y.flatten(1)))


def mmd(x: Tensor, y: Tensor) -> Tensor:
    r"""Empirical maximum mean discrepancy with Gaussian kernels of bandwidths 1e-3 ... 1e3.
    Reference: sda/utils.py:226-263."""

    x, y = x.flatten(1), y.flatten(1)
    sq = lambda a, b: (a.square().sum(1, keepdim=True) + b.square().sum(1) - 2 * a @ b.T)  # noqa: E731
    dxx, dyy, dxy = sq(x, x), sq(y, y), sq(x, y)
    total = 0

    for exponent in range(-3, 4):
        bandwidth = 10.0 ** exponent
        total = total + torch.exp(-dxx / bandwidth).mean() + torch.exp(-dyy / bandwidth).mean() - 2 * torch.exp(-dxy / bandwidth).mean()

    return total
