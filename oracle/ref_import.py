r"""Import the UNMODIFIED reference package from /root/reference (build container only).

TEST INFRASTRUCTURE.  The reference needs two third-party packages that are not
installed here and cannot be fetched (no network):

* ``zuko==0.1.4`` (``environment.yml:23``) -- only two symbols are used:
  ``zuko.utils.broadcast`` (``sda/score.py:10,57,60,87``) and
  ``zuko.nn.LayerNorm`` (``sda/nn.py:8,61,137,163``).  They are restated below
  from the published zuko 0.1.4 behaviour: LayerNorm standardises over ``dim``
  with the *unbiased* variance, ``eps`` inside the square root and no affine
  parameters; ``broadcast(*ts, ignore=n)`` broadcasts all but the last ``n``
  dimensions.  (Assumption recorded in DESIGN.md; nothing in the reference tree
  pins it.)
* ``jax`` / ``jax_cfd`` / ``h5py`` / ``ot`` -- imported at module import time by
  ``sda/mcs.py:4-6`` and ``sda/utils.py:3,6`` but not needed by the score path;
  empty stub modules are registered so the package imports.

The GPU box has no /root/reference: nothing executed there may call this.
"""

from __future__ import annotations

import sys
import types
from pathlib import Path

import torch
import torch.nn as nn

REFERENCE_ROOT = Path('/root/reference')


class _LayerNorm(nn.Module):
    r"""zuko.nn.LayerNorm(dim=-1, eps=1e-5): (x - mean) / sqrt(var_unbiased + eps)."""

    def __init__(self, dim=-1, eps: float = 1e-5):
        super().__init__()
        self.dim = dim if isinstance(dim, int) else tuple(dim)
        self.eps = eps

    def forward(self, x):
        variance, mean = torch.var_mean(x, dim=self.dim, keepdim=True)
        return (x - mean) / (variance + self.eps).sqrt()


def _broadcast(*tensors, ignore=0):
    r"""zuko.utils.broadcast: broadcast all but the last `ignore` dimensions."""

    if isinstance(ignore, int):
        ignore = [ignore] * len(tensors)

    dims = [t.dim() - i for t, i in zip(tensors, ignore)]
    common = torch.broadcast_shapes(*(t.shape[:d] for t, d in zip(tensors, dims)))

    return [torch.broadcast_to(t, common + t.shape[d:]) for t, d in zip(tensors, dims)]


def available() -> bool:
    return (REFERENCE_ROOT / 'sda' / 'score.py').exists()


def import_reference():
    r"""Returns the reference `sda` package (modules score, nn, mcs, utils)."""

    if not available():
        raise RuntimeError('/root/reference is not present on this machine')

    if 'zuko' not in sys.modules:
        zuko = types.ModuleType('zuko')
        zuko.nn = types.ModuleType('zuko.nn')
        zuko.utils = types.ModuleType('zuko.utils')
        zuko.nn.LayerNorm = _LayerNorm
        zuko.utils.broadcast = _broadcast
        sys.modules['zuko'] = zuko
        sys.modules['zuko.nn'] = zuko.nn
        sys.modules['zuko.utils'] = zuko.utils

    for name in ('jax', 'jax.numpy', 'jax.random', 'h5py', 'ot'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)

    sys.modules['jax'].numpy = sys.modules['jax.numpy']
    sys.modules['jax'].random = sys.modules['jax.random']

    # the repo ships its own top-level `sda` alias package: load the reference
    # under a private name so the two never collide.
    import importlib.util

    name = '_sda_reference'

    if name in sys.modules:
        return sys.modules[name]

    spec = importlib.util.spec_from_file_location(
        name,
        REFERENCE_ROOT / 'sda' / '__init__.py',
        submodule_search_locations=[str(REFERENCE_ROOT / 'sda')],
    )
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    spec.loader.exec_module(module)

    return module


def import_experiments(sda_package, which: str = 'kolmogorov'):
    r"""Loads /root/reference/experiments/<which>/utils.py UNCHANGED against `sda_package`.

    `sda_package` is either the reference package (golden generation) or this
    repo's drop-in `sda` alias (boundary test: "the experiments/ scripts run
    unchanged").  `seaborn` (plotting, absent here) is stubbed.
    """

    import importlib.util

    if 'seaborn' not in sys.modules:
        try:
            import seaborn  # noqa: F401
        except Exception:
            sys.modules['seaborn'] = types.ModuleType('seaborn')

    saved = {k: sys.modules.get(k) for k in ('sda', 'sda.mcs', 'sda.score', 'sda.utils', 'sda.nn')}

    try:
        sys.modules['sda'] = sda_package
        for sub in ('mcs', 'score', 'utils', 'nn'):
            sys.modules[f'sda.{sub}'] = getattr(sda_package, sub)

        name = f'_ref_experiments_{which}_{sda_package.__name__}'
        spec = importlib.util.spec_from_file_location(name, REFERENCE_ROOT / 'experiments' / which / 'utils.py')
        module = importlib.util.module_from_spec(spec)

        import os

        cwd = os.getcwd()
        os.chdir('/tmp')  # the script mkdirs PATH='.' at import time

        try:
            spec.loader.exec_module(module)
        finally:
            os.chdir(cwd)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    return module
