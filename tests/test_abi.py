r"""The C-ABI library: builds, loads, exports every symbol include/sdab.h declares, and fails
loudly (no fallback) without a GPU.  No compute calls here."""

import ctypes
import re
from pathlib import Path

import pytest
import torch

from sda_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / 'include' / 'sdab.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sdab_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30

    for name in names:
        assert hasattr(lib, name), f'{name} is declared in include/sdab.h but not exported by libsdab.so'


def test_binding_covers_the_header():
    assert sorted(_lib.SYMBOLS) == declared_symbols()


def test_version_and_handles_without_gpu():
    lib = _lib.load()
    assert lib.sdab_version() >= 100
    desc = _lib.UNetDesc()
    desc.in_channels, desc.out_channels, desc.mod_features, desc.depth = 11, 10, 64, 3
    for i, (c, b) in enumerate(zip((96, 192, 384), (3, 3, 3))):
        desc.hidden_channels[i], desc.hidden_blocks[i] = c, b
    h = ctypes.c_void_p()
    assert lib.sdab_unet_create(ctypes.byref(desc), ctypes.byref(h)) == 0
    assert lib.sdab_unet_num_convs(h) == 42  # SURVEY.md section 8a layer table
    assert lib.sdab_unet_num_blocks(h) == 18
    co, ci = ctypes.c_int(), ctypes.c_int()
    assert lib.sdab_unet_conv_shape(h, 0, ctypes.byref(co), ctypes.byref(ci)) == 0 and (co.value, ci.value) == (96, 11)
    assert lib.sdab_unet_conv_shape(h, 41, ctypes.byref(co), ctypes.byref(ci)) == 0 and (co.value, ci.value) == (10, 96)
    assert lib.sdab_unet_packed_bytes(h) > 2 * 2 * 2 * 22_000_000  # hi/lo bf16, forward + transposed
    assert lib.sdab_unet_workspace_bytes(h, 4, 64, 64, 1) > lib.sdab_unet_workspace_bytes(h, 4, 64, 64, 0) > 0
    lib.sdab_unet_destroy(h)

    # unsupported configurations are explicit errors
    desc.hidden_channels[0] = 50
    assert lib.sdab_unet_create(ctypes.byref(desc), ctypes.byref(h)) != 0
    assert b'multiples of 32' in lib.sdab_last_error()

    k = ctypes.c_void_p()
    assert lib.sdab_kolmogorov_create(256, 0.2, 1e3, ctypes.byref(k)) == 0
    assert lib.sdab_kolmogorov_inner_steps(k) == 82  # sda/mcs.py:274-284 at size 256, dt 0.2
    lib.sdab_kolmogorov_destroy(k)
    assert lib.sdab_kolmogorov_create(100, 0.2, 1e3, ctypes.byref(k)) != 0


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_no_cpu_fallback():
    lib = _lib.load()
    assert lib.sdab_device_check() != 0
    assert b'no CPU fallback' in lib.sdab_last_error()

    from helpers import build_score

    score, _ = build_score('net_small', 16)

    with pytest.raises(RuntimeError, match='no CPU fallback'):
        score(torch.zeros(1, 5, 2, 16, 16), torch.tensor(0.5))

    from sda_b200.mcs import KolmogorovFlow

    with pytest.raises(RuntimeError, match='no CPU fallback'):
        KolmogorovFlow(size=64, dt=0.2).transition(torch.zeros(2, 64, 64))
