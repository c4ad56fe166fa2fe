r"""The convolution engines through the C ABI (sdab_conv3x3) against fp64 circular convolution on
the layer shapes of the U-Net (SURVEY.md section 8a layer table), forward, stride 2, and the
transposed (input-gradient) form.  Tolerances: 5e-5 relative L2 in the bf16x3 parity mode (the
north_star's per-step bar is 1e-4), 5e-3 in the single-pass bf16 mode."""

import pytest
import torch
import torch.nn.functional as F

from oracle.testing import rel_l2

pytestmark = pytest.mark.gpu

CASES = [
    # N, Cin, Cout, H, W, stride
    (2, 11, 96, 16, 16, 1),    # head conv, C_in padded 11 -> 32
    (1, 96, 96, 32, 32, 1),    # level-0 block conv
    (3, 96, 96, 4, 4, 1),      # tiles spanning several images, partially out of range
    (2, 96, 192, 32, 32, 2),   # stride-2 head (parity layout)
    (1, 192, 192, 16, 16, 1),
    (1, 384, 384, 16, 16, 1),  # two N halves, single TMEM accumulator stage
    (1, 384, 192, 16, 16, 1),  # tail conv
    (2, 96, 10, 16, 16, 1),    # final conv, C_out padded 10 -> 16
    (1, 96, 96, 2, 256, 1),    # row tiles of 128 pixels
    (5, 64, 64, 8, 8, 2),
    (3, 96, 96, 16, 8, 1),     # patch kernel: odd number of 8 x 16 tiles (ghost tile in the last CTA pair)
    (2, 192, 96, 32, 64, 1),   # patch kernel: several tiles per image in both directions
]


def reference(x, w, b, stride):
    x = F.pad(x.double(), (1, 1, 1, 1), mode='circular')
    return F.conv2d(x, w.double(), None if b is None else b.double(), stride=stride)


def run(x, w, b, stride, transpose, mode, engine):
    from sda_b200 import _lib

    lib = _lib.load()
    N, _, H, W = x.shape
    Cout, Cin = w.shape[:2]
    nbytes = lib.sdab_conv3x3_workspace_bytes(N, Cin, Cout, H, W, stride, transpose)
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device='cuda')
    base = (ws.data_ptr() + 1023) // 1024 * 1024
    out = torch.full((N, Cin if transpose else Cout, H // stride, W // stride), float('nan'), device='cuda')
    _lib.check(lib.sdab_conv3x3(x.data_ptr(), w.data_ptr(), None if b is None else b.data_ptr(), out.data_ptr(), N, Cin,
                                Cout, H, W, stride, transpose, mode, engine, base, nbytes, _lib.stream_ptr()))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize('engine', [0, 1], ids=['umma', 'simt'])
@pytest.mark.parametrize('case', CASES, ids=lambda c: 'x'.join(map(str, c)))
def test_conv_matches_fp64(case, engine):
    N, Cin, Cout, H, W, s = case
    torch.manual_seed(sum(case))
    x = torch.randn(N, Cin, H, W, device='cuda')
    w = torch.randn(Cout, Cin, 3, 3, device='cuda') / (9 * Cin) ** 0.5
    b = torch.randn(Cout, device='cuda')
    ref = reference(x, w, b, s)
    assert rel_l2(run(x, w, b, s, 0, 0, engine), ref) < 5e-5
    assert rel_l2(run(x, w, b, s, 0, 1, engine), ref) < 5e-3

    if s == 1:
        g = torch.randn(N, Cout, H, W, device='cuda')
        xr = x.double().requires_grad_(True)
        (gref,) = torch.autograd.grad(reference(xr, w, None, 1), xr, g.double())
        assert rel_l2(run(g, w, None, 1, 1, 0, engine), gref) < 5e-5


def test_engines_agree_at_full_resolution():
    r"""256 x 256 (BASELINE config 3 resolution): tcgen05 engine vs the CUDA-core engine."""

    torch.manual_seed(0)
    x = torch.randn(1, 96, 256, 256, device='cuda')
    w = torch.randn(96, 96, 3, 3, device='cuda') / (9 * 96) ** 0.5
    b = torch.randn(96, device='cuda')
    a, c = run(x, w, b, 1, 0, 0, 0), run(x, w, b, 1, 0, 0, 1)
    assert rel_l2(a, c) < 2e-5
    # linearity of the operator (size-independent property)
    x2 = torch.randn_like(x)
    lhs = run(x + 2 * x2, w, None, 1, 0, 0, 0)
    rhs = run(x, w, None, 1, 0, 0, 0) + 2 * run(x2, w, None, 1, 0, 0, 0)
    assert rel_l2(lhs, rhs) < 2e-5


def test_unsupported_shapes_are_explicit_errors():
    x = torch.zeros(1, 32, 24, 24, device='cuda')
    w = torch.zeros(32, 32, 3, 3, device='cuda')

    with pytest.raises(RuntimeError, match='power of two'):
        run(x, w, None, 1, 0, 0, 0)


def test_patch_kernel_agrees_with_per_tap_kernel():
    r"""The patch kernel (one haloed input patch per K-block, taps as shifted descriptors, K-block-major
    accumulation) against the per-tap kernel (SDAB_UMMA_PATCH=0, read once at library load, hence a
    subprocess): same products, different summation order."""

    import subprocess
    import sys
    from pathlib import Path

    code = """
import os, sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from test_gpu_conv import run
torch.manual_seed(3)
x = torch.randn(2, 192, 32, 64, device='cuda'); w = torch.randn(192, 192, 3, 3, device='cuda') / (9 * 192) ** 0.5
b = torch.randn(192, device='cuda')
torch.save(run(x, w, b, 1, 0, 0, 0).cpu(), sys.argv[1])
"""
    root = Path(__file__).resolve().parents[1]
    outs = []

    for flag in ('1', '0'):
        path = root / f'.pytest_patch_{flag}.pt'
        env = dict(**__import__('os').environ, SDAB_UMMA_PATCH=flag)
        subprocess.run([sys.executable, '-c', code.format(root=str(root), tests=str(root / 'tests')), str(path)], check=True, env=env, timeout=300)
        outs.append(torch.load(path))
        path.unlink()

    assert rel_l2(outs[0], outs[1]) < 5e-6
