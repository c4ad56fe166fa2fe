r"""Deterministic test inputs shared by the golden generator, tests, smoke and bench.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Weights are filled per state_dict key from a generator seeded by crc32(key), so
the values depend neither on module construction order
nor on which implementation (reference / oracle / sda_b200) owns the tensors.
"""

from __future__ import annotations

import math
import zlib
from typing import Dict

import torch
from torch import Tensor

SKIP_SUFFIXES = ('freqs', 'forcing', 'device')


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device='cpu')
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def fill_state_(state: Dict[str, Tensor], seed: int = 0, gain: float = 1.0) -> Dict[str, Tensor]:
    r"""In-place PyTorch-default-like init U(-1/sqrt(fan_in), 1/sqrt(fan_in)), keyed by name."""

    for key in sorted(state):
        t = state[key]

        if not t.is_floating_point() or key.endswith(SKIP_SUFFIXES):
            continue

        if key.endswith('bias'):
            w = state.get(key[: -len('bias')] + 'weight')
            fan_in = math.prod(w.shape[1:]) if w is not None else t.numel()
        else:
            fan_in = math.prod(t.shape[1:]) if t.dim() > 1 else t.numel()

        bound = gain / math.sqrt(max(fan_in, 1))
        v = (torch.rand(t.shape, generator=_gen(seed, key), dtype=torch.float64) * 2 - 1) * bound

        with torch.no_grad():
            t.copy_(v.to(t.dtype))

    return state


def randn(shape, seed: int, dtype=torch.float32) -> Tensor:
    r"""Seeded CPU standard normal draw (fp64 generator, cast)."""

    g = torch.Generator(device='cpu')
    g.manual_seed(seed)

    return torch.randn(shape, generator=g, dtype=torch.float64).to(dtype)


def rel_l2(a: Tensor, b: Tensor) -> float:
    r"""||a - b|| / ||b|| in fp64."""

    a = a.detach().double().cpu()
    b = b.detach().double().cpu()

    return float((a - b).norm() / b.norm().clamp_min(1e-300))
