r"""Markov chains -- drop-in for ``sda.mcs`` (reference: /root/reference/sda/mcs.py).

``KolmogorovFlow`` keeps the reference's constructor and methods (prior, transition,
trajectory, coarsen, upsample, vorticity) but steps the flow with libsdab's CUDA
kernels (``sdab_kolmogorov_*``) instead of jax / jax-cfd: no jax import, no
torch->numpy->jax hop per transition.  There is no CPU implementation of the stepper
in this package (oracle/kolmogorov_oracle.py is test infrastructure).

The low-dimensional chains (config 1 plumbing) are plain PyTorch like the reference's.
"""

from __future__ import annotations

import abc
import ctypes
import math
import random
from typing import Callable, Tuple

import torch
from torch import Size, Tensor
from torch.distributions import MultivariateNormal, Normal

from . import _lib

__all__ = [
    'MarkovChain', 'DampedSpring', 'DiscreteODE', 'Lorenz63', 'NoisyLorenz63', 'Lorenz96', 'LotkaVolterra',
    'KolmogorovFlow',
]


class MarkovChain(abc.ABC):
    r"""Abstract first-order time-invariant Markov chain.  Reference: sda/mcs.py:22-57."""

    @abc.abstractmethod
    def prior(self, shape: Size = ()) -> Tensor:
        r"""x_0 ~ p(x_0)"""

    @abc.abstractmethod
    def transition(self, x: Tensor) -> Tensor:
        r"""x_i ~ p(x_i | x_{i-1})"""

    def trajectory(self, x: Tensor, length: int, last: bool = False) -> Tensor:
        r"""(x_1, ..., x_n) ~ prod_i p(x_i | x_{i-1}); only x_n when `last`."""

        states = []

        for _ in range(length):
            x = self.transition(x)

            if not last:
                states.append(x)

        return x if last else torch.stack(states)


class DampedSpring(MarkovChain):
    r"""Linear-Gaussian mass-spring system.  Reference: sda/mcs.py:60-82."""

    def __init__(self, dt: float = 0.01):
        super().__init__()

        self.mu_0 = torch.tensor([1.0, 0.0, 0.0, 0.0])
        self.Sigma_0 = torch.eye(4)
        self.A = torch.tensor([
            [1.0, dt, dt**2 / 2, 0.0],
            [0.0, 1.0, dt, 0.0],
            [-0.5, -0.1, 0.0, 0.2],
            [0.0, 0.0, 0.0, 0.99],
        ])
        self.b = torch.zeros(4)
        self.Sigma_x = torch.tensor([0.1, 0.1, 0.1, 1.0]).diag() * dt

    def prior(self, shape: Size = ()) -> Tensor:
        return MultivariateNormal(self.mu_0, self.Sigma_0).sample(shape)

    def transition(self, x: Tensor) -> Tensor:
        return MultivariateNormal(x @ self.A.T + self.b, self.Sigma_x).sample()


class DiscreteODE(MarkovChain):
    r"""ODE discretised with `steps` RK4 sub-steps per transition.  Reference: sda/mcs.py:85-122."""

    def __init__(self, dt: float = 0.01, steps: int = 1):
        super().__init__()

        self.dt, self.steps = dt, steps

    @staticmethod
    def rk4(f: Callable[[Tensor], Tensor], x: Tensor, dt: float) -> Tensor:
        k1 = f(x)
        k2 = f(x + dt * k1 / 2)
        k3 = f(x + dt * k2 / 2)
        k4 = f(x + dt * k3)

        return x + dt * (k1 + 2 * k2 + 2 * k3 + k4) / 6

    @abc.abstractmethod
    def f(self, x: Tensor) -> Tensor:
        r"""f(x) = dx/dt"""

    def transition(self, x: Tensor) -> Tensor:
        for _ in range(self.steps):
            x = self.rk4(self.f, x, self.dt / self.steps)

        return x


class Lorenz63(DiscreteODE):
    r"""Lorenz 1963 system.  Reference: sda/mcs.py:125-172."""

    def __init__(self, sigma: float = 10.0, rho: float = 28.0, beta: float = 8 / 3, **kwargs):
        super().__init__(**kwargs)

        self.sigma, self.rho, self.beta = sigma, rho, beta

    def prior(self, shape: Size = ()) -> Tensor:
        mean = torch.tensor([0.0, 0.0, 25.0])
        cov = torch.tensor([[64.0, 50.0, 0.0], [50.0, 81.0, 0.0], [0.0, 0.0, 75.0]])

        return MultivariateNormal(mean, cov).sample(shape)

    def f(self, x: Tensor) -> Tensor:
        a, b, c = x[..., 0], x[..., 1], x[..., 2]

        return torch.stack((self.sigma * (b - a), a * (self.rho - c) - b, a * b - self.beta * c), dim=-1)

    @staticmethod
    def preprocess(x: Tensor) -> Tensor:
        return (x - x.new_tensor([0.0, 0.0, 25.0])) / x.new_tensor([8.0, 9.0, 8.6])

    @staticmethod
    def postprocess(x: Tensor) -> Tensor:
        return x.new_tensor([0.0, 0.0, 25.0]) + x.new_tensor([8.0, 9.0, 8.6]) * x


class NoisyLorenz63(Lorenz63):
    r"""Lorenz 1963 with additive Gaussian transition noise.  Reference: sda/mcs.py:175-185."""

    def moments(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        return super().transition(x), self.dt**0.5

    def transition(self, x: Tensor) -> Tensor:
        return Normal(*self.moments(x)).sample()

    def log_prob(self, x1: Tensor, x2: Tensor) -> Tensor:
        return Normal(*self.moments(x1)).log_prob(x2).sum(dim=-1)


class Lorenz96(DiscreteODE):
    r"""Lorenz 1996 system.  Reference: sda/mcs.py:188-211."""

    def __init__(self, n: int = 32, F: float = 16.0, **kwargs):
        super().__init__(**kwargs)

        self.n, self.F = n, F

    def prior(self, shape: Size = ()) -> Tensor:
        return torch.randn(*shape, self.n)

    def f(self, x: Tensor) -> Tensor:
        return (torch.roll(x, 1, -1) - torch.roll(x, -2, -1)) * torch.roll(x, -1, -1) - x + self.F


class LotkaVolterra(DiscreteODE):
    r"""Lotka-Volterra system in log space.  Reference: sda/mcs.py:214-241."""

    def __init__(self, alpha: float = 1.0, beta: float = 1.0, delta: float = 1.0, gamma: float = 1.0, **kwargs):
        super().__init__(**kwargs)

        self.alpha, self.beta = alpha, beta
        self.delta, self.gamma = delta, gamma

    def prior(self, shape: Size = ()) -> Tensor:
        return torch.rand(*shape, 2)

    def f(self, x: Tensor) -> Tensor:
        prey, pred = x[..., 0], x[..., 1]

        return torch.stack((self.alpha - self.beta * pred.exp(), self.delta * prey.exp() - self.gamma), dim=-1)


class KolmogorovFlow(MarkovChain):
    r"""2-D incompressible flow with Kolmogorov forcing.  Reference: sda/mcs.py:244-375.

    Same arguments as the reference.  `transition` accepts a tensor on any device and
    returns the result on that device (the reference always returns a CPU tensor; a CPU
    input therefore gives the reference's behaviour).  `device` selects the GPU that steps
    CPU inputs (default: the current CUDA device).
    """

    def __init__(self, size: int = 256, dt: float = 0.01, reynolds: int = 1e3, device=None):
        super().__init__()

        self.size, self.dt, self.reynolds = size, dt, reynolds
        self.device = device
        lib = _lib.load()
        handle = ctypes.c_void_p()
        _lib.check(lib.sdab_kolmogorov_create(size, float(dt), float(reynolds), ctypes.byref(handle)))
        self._handle = handle
        self.steps = lib.sdab_kolmogorov_inner_steps(handle)
        self._workspace = None

    def __del__(self):
        handle = getattr(self, '_handle', None)

        if handle is not None:
            try:
                _lib.load().sdab_kolmogorov_destroy(handle)
            except Exception:
                pass

    # ------------------------------------------------------------------ helpers
    def _device(self, x: Tensor = None) -> torch.device:
        if x is not None and x.is_cuda:
            return x.device

        if not torch.cuda.is_available():
            raise RuntimeError('sda_b200.mcs.KolmogorovFlow steps the flow on an sm_100 GPU only (no CPU fallback)')

        return torch.device(self.device) if self.device is not None else torch.device('cuda', torch.cuda.current_device())

    def _get_workspace(self, E: int, device) -> Tensor:
        nbytes = _lib.load().sdab_kolmogorov_workspace_bytes(self._handle, E) + 1024

        if self._workspace is None or self._workspace.numel() < nbytes or self._workspace.device != device:
            self._workspace = None
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=device)

        return self._workspace

    def _run(self, x: Tensor, n: int, keep: bool) -> Tensor:
        if x.shape[-3:] != (2, self.size, self.size):
            raise RuntimeError(f'expected a state of shape (..., 2, {self.size}, {self.size}), got {tuple(x.shape)}')

        device = self._device(x)
        batch = x.shape[:-3]
        uv = x.detach().to(device=device, dtype=torch.float32, copy=True).reshape(-1, 2, self.size, self.size).contiguous()
        E = uv.shape[0]
        lib = _lib.load()

        with torch.cuda.device(device):
            ws = self._get_workspace(E, device)
            base = (ws.data_ptr() + 1023) // 1024 * 1024
            traj = torch.empty((n,) + uv.shape, dtype=torch.float32, device=device) if keep else None
            _lib.check(
                lib.sdab_kolmogorov_transition(
                    self._handle, uv.data_ptr(), E, n, None if traj is None else traj.data_ptr(), base,
                    ws.numel() - (base - ws.data_ptr()), _lib.stream_ptr(),
                )
            )

        out = traj.reshape((n,) + tuple(batch) + (2, self.size, self.size)) if keep else uv.reshape(x.shape)

        return out.to(x.device)

    # ------------------------------------------------------------------ MarkovChain API
    def prior(self, shape: Size = ()) -> Tensor:
        r"""Filtered random velocity field (max speed 3, peak wavenumber 4), sda/mcs.py:321-331.
        Seeded from Python's `random` like the reference (experiments/kolmogorov/generate.py:20)."""

        seed = random.randrange(2**32)
        shape = tuple(shape)
        device = self._device()
        E = max(1, math.prod(shape))
        uv = torch.empty((E, 2, self.size, self.size), dtype=torch.float32, device=device)
        lib = _lib.load()

        with torch.cuda.device(device):
            ws = self._get_workspace(E, device)
            base = (ws.data_ptr() + 1023) // 1024 * 1024
            _lib.check(
                lib.sdab_kolmogorov_prior(
                    self._handle, uv.data_ptr(), E, seed, base, ws.numel() - (base - ws.data_ptr()), _lib.stream_ptr()
                )
            )

        return uv.reshape(shape + (2, self.size, self.size)).cpu()

    def transition(self, x: Tensor) -> Tensor:
        return self._run(x, 1, keep=False)

    def trajectory(self, x: Tensor, length: int, last: bool = False) -> Tensor:
        r"""All `length` transitions in one library call (the state never leaves the GPU)."""

        return self._run(x, length, keep=not last)

    # ------------------------------------------------------------------ observation helpers
    # On CUDA fp32 tensors these run as single libsdab kernels with analytic adjoints (they are the A(x) of the
    # guided sampler, differentiated at every score evaluation: sda/score.py:389-394); anything else takes the
    # reference's PyTorch formulas.
    @staticmethod
    def coarsen(x: Tensor, r: int = 2) -> Tensor:
        r"""Mean over r x r blocks.  Reference: sda/mcs.py:340-347."""

        *batch, h, w = x.shape

        if _native_image(x) and h % r == 0 and w % r == 0:
            return _Coarsen.apply(x, r)

        return x.reshape(*batch, h // r, r, w // r, r).mean(dim=(-3, -1))

    @staticmethod
    def upsample(x: Tensor, r: int = 2, mode: str = 'bilinear') -> Tensor:
        r"""Circular-padded interpolation.  Reference: sda/mcs.py:349-359."""

        *batch, h, w = x.shape

        if mode == 'bilinear' and _native_image(x) and h >= 3 and w >= 3 and isinstance(r, int):
            return _Upsample.apply(x, r)

        x = x.reshape(-1, 1, h, w)
        x = torch.nn.functional.pad(x, pad=(1, 1, 1, 1), mode='circular')
        x = torch.nn.functional.interpolate(x, scale_factor=(r, r), mode=mode)
        x = x[..., r:-r, r:-r]

        return x.reshape(*batch, r * h, r * w)

    @staticmethod
    def vorticity(x: Tensor) -> Tensor:
        r"""Central-difference curl with circular wrap.  Reference: sda/mcs.py:361-375."""

        if _native_image(x) and x.dim() >= 3 and x.shape[-3] == 2:
            return _Vorticity.apply(x)

        u, v = x[..., 0, :, :], x[..., 1, :, :]
        du = (torch.roll(u, -1, dims=-1) - torch.roll(u, 1, dims=-1)) / 2
        dv = (torch.roll(v, -1, dims=-2) - torch.roll(v, 1, dims=-2)) / 2

        return du - dv


def _native_image(x: Tensor) -> bool:
    return x.is_cuda and x.dtype == torch.float32 and x.dim() >= 2 and x.numel() > 0


def _call(name: str, *args) -> None:
    _lib.check(getattr(_lib.load(), name)(*args, _lib.stream_ptr()))


class _Coarsen(torch.autograd.Function):
    r"""sdab_coarsen / sdab_coarsen_adjoint."""

    @staticmethod
    def forward(ctx, x: Tensor, r: int) -> Tensor:
        *batch, h, w = x.shape
        x = x.contiguous()
        out = torch.empty((*batch, h // r, w // r), dtype=x.dtype, device=x.device)
        ctx.meta = (tuple(x.shape), r)

        with torch.cuda.device(x.device):
            _call('sdab_coarsen', x.data_ptr(), out.data_ptr(), x.numel() // (h * w), h, w, r)

        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        shape, r = ctx.meta
        h, w = shape[-2:]
        g = g.contiguous()
        gx = torch.empty(shape, dtype=g.dtype, device=g.device)

        with torch.cuda.device(g.device):
            _call('sdab_coarsen_adjoint', g.data_ptr(), gx.data_ptr(), gx.numel() // (h * w), h, w, r)

        return gx, None


class _Upsample(torch.autograd.Function):
    r"""sdab_upsample_bilinear / sdab_upsample_bilinear_adjoint."""

    @staticmethod
    def forward(ctx, x: Tensor, r: int) -> Tensor:
        *batch, h, w = x.shape
        x = x.contiguous()
        out = torch.empty((*batch, r * h, r * w), dtype=x.dtype, device=x.device)
        ctx.meta = (tuple(x.shape), r)

        with torch.cuda.device(x.device):
            _call('sdab_upsample_bilinear', x.data_ptr(), out.data_ptr(), x.numel() // (h * w), h, w, r)

        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        shape, r = ctx.meta
        h, w = shape[-2:]
        g = g.contiguous()
        gx = torch.empty(shape, dtype=g.dtype, device=g.device)

        with torch.cuda.device(g.device):
            _call('sdab_upsample_bilinear_adjoint', g.data_ptr(), gx.data_ptr(), gx.numel() // (h * w), h, w, r)

        return gx, None


class _Vorticity(torch.autograd.Function):
    r"""sdab_vorticity / sdab_vorticity_adjoint."""

    @staticmethod
    def forward(ctx, x: Tensor) -> Tensor:
        *batch, _, h, w = x.shape
        x = x.contiguous()
        out = torch.empty((*batch, h, w), dtype=x.dtype, device=x.device)
        ctx.shape = tuple(x.shape)

        with torch.cuda.device(x.device):
            _call('sdab_vorticity', x.data_ptr(), out.data_ptr(), x.numel() // (2 * h * w), h, w)

        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        h, w = ctx.shape[-2:]
        g = g.contiguous()
        gx = torch.empty(ctx.shape, dtype=g.dtype, device=g.device)

        with torch.cuda.device(g.device):
            _call('sdab_vorticity_adjoint', g.data_ptr(), gx.data_ptr(), gx.numel() // (2 * h * w), h, w)

        return gx
