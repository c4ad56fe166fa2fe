r"""Torch-CPU restatement of the reference score path -- TEST INFRASTRUCTURE ONLY.

A *functional* restatement (no nn.Module, explicit state_dict) of the algorithm in
``/root/reference/sda/score.py`` and ``/root/reference/sda/nn.py``.  It runs in
fp32 or fp64 on the CPU through plain ATen ops and is what the CUDA path is
compared with on the GPU box (where /root/reference does not exist).

PINNED: ``tests/golden/make_golden.py`` checks every function below against the
reference classes imported unmodified from /root/reference (zuko shim in
``oracle/ref_import.py``) and commits the vectors; ``tests/test_oracle.py``
re-checks the oracle against those vectors on every run.
"""

from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

State = Dict[str, Tensor]


# --------------------------------------------------------------------------- #
# Window maps                                                     score.py:146-164
# --------------------------------------------------------------------------- #

def unfold(x: Tensor, order: int) -> Tensor:
    r"""(B, L, C, ...) -> (B, L-2k, (2k+1) C, ...); window i = frames i..i+2k.

    Restates MCScoreNet.unfold (sda/score.py:146-153) with explicit indexing.
    """

    w = 2 * order + 1
    B, L = x.shape[:2]

    if L < w:
        raise RuntimeError(f'trajectory length {L} is smaller than the window {w}')

    wins = [torch.cat([x[:, i + s] for s in range(w)], dim=1) for i in range(L - w + 1)]

    return torch.stack(wins, dim=1)


def fold_map(L: int, order: int) -> Sequence[Tuple[int, int]]:
    r"""(window, slot) that output frame j is taken from.  sda/score.py:155-164."""

    k = order
    nw = L - 2 * k
    out = []

    for j in range(L):
        if j < k:
            out.append((0, j))
        elif j < L - k:
            out.append((j - k, k))
        else:
            out.append((nw - 1, j - (nw - 1)))

    return out


def fold(s: Tensor, order: int) -> Tensor:
    r"""(B, L-2k, (2k+1) C, ...) -> (B, L, C, ...).  MCScoreNet.fold, sda/score.py:155-164."""

    w = 2 * order + 1
    B, nw = s.shape[:2]
    C = s.shape[2] // w
    L = nw + 2 * order
    frames = [s[:, i, slot * C:(slot + 1) * C] for i, slot in fold_map(L, order)]

    return torch.stack(frames, dim=1)


def unfold_transpose(g: Tensor, order: int) -> Tensor:
    r"""Adjoint of `unfold` (overlap-add); autograd's UnfoldBackward0 of score.py:148."""

    w = 2 * order + 1
    B, nw = g.shape[:2]
    C = g.shape[2] // w
    L = nw + 2 * order
    out = g.new_zeros((B, L, C) + tuple(g.shape[3:]))

    for i in range(nw):
        for s in range(w):
            out[:, i + s] += g[:, i, s * C:(s + 1) * C]

    return out


# --------------------------------------------------------------------------- #
# Network pieces                                                         nn.py
# --------------------------------------------------------------------------- #

def layer_norm_c(x: Tensor, eps: float = 1e-5) -> Tensor:
    r"""zuko.nn.LayerNorm(dim=-3): standardise over channels, unbiased variance, no affine.

    Call sites sda/nn.py:137,163.
    """

    var, mean = torch.var_mean(x, dim=-3, keepdim=True)

    return (x - mean) / (var + eps).sqrt()


def conv3x3_circular(x: Tensor, weight: Tensor, bias: Optional[Tensor], stride: int = 1) -> Tensor:
    r"""nn.Conv2d(kernel_size=3, padding=1, padding_mode='circular', stride=stride).

    Built at sda/nn.py:125-128,138-140,151-174 with padding_mode from
    experiments/kolmogorov/utils.py:67.
    """

    x = F.pad(x, (1, 1, 1, 1), mode='circular')

    return F.conv2d(x, weight, bias, stride=stride)


ACTIVATIONS: Dict[str, Callable[[Tensor], Tensor]] = {
    'SiLU': F.silu,
    'ReLU': F.relu,
    'ELU': F.elu,
    'GELU': F.gelu,
    'SELU': F.selu,
}


def unet_depth(state: State, prefix: str = '') -> Tuple[Sequence[int], Sequence[int]]:
    r"""Recovers (hidden_channels, hidden_blocks) from the state_dict keys."""

    channels, blocks = [], []
    d = 0

    while f'{prefix}descent.{d}.0.project.0.weight' in state:
        channels.append(state[f'{prefix}descent.{d}.0.project.0.weight'].shape[0])
        b = 0

        while f'{prefix}descent.{d}.{b}.project.0.weight' in state:
            b += 1

        blocks.append(b)
        d += 1

    return channels, blocks


def mod_block(state: State, prefix: str, x: Tensor, y: Tensor, act) -> Tensor:
    r"""ModResidualBlock: x + conv2(act(conv1(LN_C(x + Linear(y)[:, :, None, None])))).

    sda/nn.py:27-28 with the constructor at :131-142.
    """

    p = F.linear(y, state[prefix + 'project.0.weight'], state[prefix + 'project.0.bias'])
    h = layer_norm_c(x + p[:, :, None, None])
    h = conv3x3_circular(h, state[prefix + 'residue.1.weight'], state[prefix + 'residue.1.bias'])
    h = act(h)
    h = conv3x3_circular(h, state[prefix + 'residue.3.weight'], state[prefix + 'residue.3.bias'])

    return x + h


def unet(state: State, x: Tensor, y: Tensor, prefix: str = '', activation: str = 'SiLU') -> Tensor:
    r"""UNet.forward (sda/nn.py:184-206), 2-D, kernel 3, stride 2, circular padding.

    x: (N, C_in, H, W), y: (Nt, mod_features) with Nt in {1, N}.
    Module order after the `reversed` at sda/nn.py:180,182: tails.0 / ascent.0 are
    the deepest level.
    """

    act = ACTIVATIONS[activation]
    channels, blocks = unet_depth(state, prefix)
    D = len(channels)
    memory = []

    for d in range(D):
        if d == 0:
            x = conv3x3_circular(x, state[f'{prefix}heads.0.weight'], state[f'{prefix}heads.0.bias'])
        else:
            x = conv3x3_circular(
                x, state[f'{prefix}heads.{d}.0.weight'], state[f'{prefix}heads.{d}.0.bias'], stride=2
            )

        for b in range(blocks[d]):
            x = mod_block(state, f'{prefix}descent.{d}.{b}.', x, y, act)

        memory.append(x)

    memory.pop()

    for i in range(D):
        d = D - 1 - i

        for b in range(blocks[d]):
            x = mod_block(state, f'{prefix}ascent.{i}.{b}.', x, y, act)

        if d > 0:
            h = layer_norm_c(x)
            h = h.repeat_interleave(2, dim=-2).repeat_interleave(2, dim=-1)  # nearest x2
            h = conv3x3_circular(h, state[f'{prefix}tails.{i}.2.weight'], state[f'{prefix}tails.{i}.2.bias'])
            x = h + memory.pop()
        else:
            x = conv3x3_circular(x, state[f'{prefix}tails.{i}.weight'], state[f'{prefix}tails.{i}.bias'])

    return x


def time_embedding(state: State, t: Tensor, prefix: str = 'embedding.') -> Tensor:
    r"""TimeEmbedding.forward, sda/score.py:15-35.  t: (Nt,) -> (Nt, features)."""

    freqs = state[prefix + 'freqs'].to(t.dtype)
    e = freqs * t.unsqueeze(-1)
    e = torch.cat((e.cos(), e.sin()), dim=-1)
    e = F.linear(e, state[prefix + '0.weight'], state[prefix + '0.bias'])
    e = F.silu(e)
    e = F.linear(e, state[prefix + '2.weight'], state[prefix + '2.bias'])

    return e


def score_unet(state: State, x: Tensor, t: Tensor, c: Optional[Tensor] = None, activation: str = 'SiLU') -> Tensor:
    r"""ScoreUNet.forward, sda/score.py:81-93.  x: (..., C, H, W); c: (C', H, W) or None.

    LocalScoreUNet (experiments/kolmogorov/utils.py:29-46) is this with
    c = state['forcing'].
    """

    if c is None:
        y = x
    else:
        cb = c.expand(x.shape[:-3] + c.shape[-3:])
        y = torch.cat((x, cb), dim=-3)

    y = y.reshape(-1, *y.shape[-3:])
    e = time_embedding(state, t.reshape(-1))

    return unet(state, y, e, prefix='network.', activation=activation).reshape(x.shape)


def mc_score(state: State, x: Tensor, t: Tensor, order: int, activation: str = 'SiLU') -> Tensor:
    r"""MCScoreNet.forward (sda/score.py:134-144) with kernel = LocalScoreUNet.

    `state` is the kernel's state_dict (keys embedding.*, network.*, forcing).
    """

    w = unfold(x, order)
    s = score_unet(state, w, t, state.get('forcing'), activation)

    return fold(s, order)


# --------------------------------------------------------------------------- #
# VPSDE                                                           score.py:167-276
# --------------------------------------------------------------------------- #

def alpha(t: Tensor, kind: str = 'cos', eta: float = 1e-3) -> Tensor:
    r"""sda/score.py:195-200."""

    if kind == 'lin':
        return 1 - (1 - eta) * t
    elif kind == 'cos':
        return torch.cos(math.acos(math.sqrt(eta)) * t) ** 2
    elif kind == 'exp':
        return torch.exp(math.log(eta) * t ** 2)

    raise ValueError(kind)


def mu(t: Tensor, kind: str = 'cos', eta: float = 1e-3) -> Tensor:
    r"""VPSDE.mu, sda/score.py:206-207."""

    return alpha(t, kind, eta)


def sigma(t: Tensor, kind: str = 'cos', eta: float = 1e-3, sde: str = 'vp') -> Tensor:
    r"""VPSDE.sigma :209-210, SubVPSDE.sigma :287-288, SubSubVPSDE.sigma :299-300."""

    a = alpha(t, kind, eta)

    if sde == 'vp':
        return (1 - a ** 2 + eta ** 2).sqrt()
    elif sde == 'subvp':
        return 1 - a ** 2 + eta
    elif sde == 'subsubvp':
        return 1 - a + eta

    raise ValueError(sde)


def gaussian_score(
    eps_fn: Callable[[Tensor, Tensor], Tensor],
    y: Tensor,
    A: Callable[[Tensor], Tensor],
    std,
    x: Tensor,
    t: Tensor,
    gamma=1e-2,
    detach: bool = False,
    kind: str = 'cos',
    eta: float = 1e-3,
) -> Tensor:
    r"""GaussianScore.forward, sda/score.py:375-396 (returns eps - sigma * grad_x log p(y | x))."""

    m, s = mu(t, kind, eta), sigma(t, kind, eta)
    std = torch.as_tensor(std, dtype=x.dtype)
    gamma = torch.as_tensor(gamma, dtype=x.dtype)

    if detach:
        with torch.no_grad():
            eps = eps_fn(x, t)

    with torch.enable_grad():
        x = x.detach().requires_grad_(True)

        if not detach:
            eps = eps_fn(x, t)

        x_ = (x - s * eps) / m
        err = y - A(x_)
        var = std ** 2 + gamma * (s / m) ** 2
        log_p = -(err ** 2 / var).sum() / 2

    g, = torch.autograd.grad(log_p, x)

    return (eps - s * g).detach()


def pc_sample(
    eps_fn: Callable[[Tensor, Tensor], Tensor],
    x: Tensor,
    steps: int,
    corrections: int = 0,
    tau: float = 1.0,
    noise: Optional[Sequence[Tensor]] = None,
    event_dims: int = 4,
    kind: str = 'cos',
    eta: float = 1e-3,
    n_steps: Optional[int] = None,
) -> Tensor:
    r"""The loop of VPSDE.sample, sda/score.py:246-261, from a given x(1).

    `noise` injects the corrector draws (one tensor per corrector update, in
    order) so that two implementations can be compared step by step;
    `n_steps` stops after that many denoising steps.
    """

    time = torch.linspace(1, 0, steps + 1).to(x)
    dt = 1 / steps
    dims = tuple(range(-event_dims, 0))
    it = iter(noise) if noise is not None else None

    with torch.no_grad():
        for i, t in enumerate(time[:-1]):
            if n_steps is not None and i >= n_steps:
                break

            r = mu(t - dt, kind, eta) / mu(t, kind, eta)
            x = r * x + (sigma(t - dt, kind, eta) - r * sigma(t, kind, eta)) * eps_fn(x, t)

            for _ in range(corrections):
                z = next(it) if it is not None else torch.randn_like(x)
                eps = eps_fn(x, t - dt)
                delta = tau / eps.square().mean(dim=dims, keepdim=True)
                x = x - (delta * eps + torch.sqrt(2 * delta) * z) * sigma(t - dt, kind, eta)

    return x


# --------------------------------------------------------------------------- #
# Observation helpers                                               mcs.py:340-375
# --------------------------------------------------------------------------- #

def coarsen(x: Tensor, r: int = 2) -> Tensor:
    r"""KolmogorovFlow.coarsen, sda/mcs.py:340-347: mean over r x r blocks."""

    *batch, h, w = x.shape

    return x.reshape(*batch, h // r, r, w // r, r).mean(dim=(-3, -1))


def vorticity(x: Tensor) -> Tensor:
    r"""KolmogorovFlow.vorticity, sda/mcs.py:361-375: central d(u)/d(axis -1) - d(v)/d(axis -2)."""

    u, v = x[..., 0, :, :], x[..., 1, :, :]
    du = (torch.roll(u, -1, dims=-1) - torch.roll(u, 1, dims=-1)) / 2
    dv = (torch.roll(v, -1, dims=-2) - torch.roll(v, 1, dims=-2)) / 2

    return du - dv
