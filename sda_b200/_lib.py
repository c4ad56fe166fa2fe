r"""ctypes binding of libsdab (include/sdab.h).

The library is built in-tree by ``sda_b200.build`` (nvcc, sm_100a).  There is no
fallback of any kind: if the shared object cannot be built or loaded, or no
sm_100 device is current, the operators raise.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_uint64, c_void_p
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / 'libsdab.so'

MAX_DEPTH = 8
MODE_BF16X3, MODE_BF16 = 0, 1
ENGINE_UMMA, ENGINE_SIMT = 0, 1
ACT_SILU, ACT_RELU = 0, 1


class UNetDesc(ctypes.Structure):
    _fields_ = [
        ('in_channels', c_int),
        ('out_channels', c_int),
        ('mod_features', c_int),
        ('depth', c_int),
        ('hidden_channels', c_int * MAX_DEPTH),
        ('hidden_blocks', c_int * MAX_DEPTH),
        ('activation', c_int),
    ]


_PROTOTYPES = {
    'sdab_last_error': (c_char_p, []),
    'sdab_version': (c_int, []),
    'sdab_device_check': (c_int, []),
    'sdab_launch_count': (c_longlong, [c_int]),
    'sdab_conv_profile': (c_int, [c_int]),
    'sdab_conv_profile_read': (c_int, [POINTER(c_double), POINTER(c_double), POINTER(c_longlong)]),
    # U-Net
    'sdab_unet_create': (c_int, [POINTER(UNetDesc), POINTER(c_void_p)]),
    'sdab_unet_destroy': (None, [c_void_p]),
    'sdab_unet_num_convs': (c_int, [c_void_p]),
    'sdab_unet_num_blocks': (c_int, [c_void_p]),
    'sdab_unet_conv_shape': (c_int, [c_void_p, c_int, POINTER(c_int), POINTER(c_int)]),
    'sdab_unet_packed_bytes': (c_size_t, [c_void_p]),
    'sdab_unet_set_weights': (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                      POINTER(c_void_p), c_void_p, c_size_t, c_void_p]),
    'sdab_unet_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    'sdab_unet_forward': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_size_t, c_int, c_int, c_int, c_void_p]),
    'sdab_unet_dgrad': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    'sdab_unet_backward': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int,
                                   c_int, c_void_p]),
    'sdab_unet_shift_rows': (c_int, [c_void_p]),
    'sdab_mcscore_forward': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_size_t, c_int, c_int, c_int, c_void_p]),
    'sdab_mcscore_dgrad': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    'sdab_conv3x3_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    'sdab_conv3x3': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_int, c_int, c_void_p, c_size_t, c_void_p]),
    # window maps
    'sdab_unfold_cat': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'sdab_fold': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'sdab_fold_transpose': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'sdab_frames_assemble': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'sdab_unfold_transpose_add': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    # exchange over peer memory
    'sdab_peer_header_bytes': (c_size_t, []),
    'sdab_peer_alloc': (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    'sdab_peer_open': (c_int, [c_void_p, POINTER(c_void_p)]),
    'sdab_peer_close': (c_int, [c_void_p]),
    'sdab_peer_free': (c_int, [c_void_p]),
    'sdab_peer_allgather': (c_int, [POINTER(c_void_p), c_int, c_int, c_size_t, c_size_t, c_uint64, c_void_p]),
    'sdab_peer_signal_wait': (c_int, [POINTER(c_void_p), c_int, c_int, c_uint64, c_void_p]),
    'sdab_peer_adamw': (c_int, [POINTER(c_void_p), POINTER(c_void_p), c_void_p, c_void_p, c_size_t, c_size_t, c_int, c_int,
                                c_float, c_float, c_float, c_float, c_float, c_int, c_uint64, c_void_p]),
    # sampler
    'sdab_vpsde_predict': (c_int, [c_void_p, c_void_p, c_float, c_float, c_size_t, c_void_p]),
    'sdab_vpsde_correct_scratch_floats': (c_size_t, [c_int]),
    'sdab_vpsde_correct': (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_uint64, c_uint64, c_int, c_size_t,
                                   c_void_p, c_void_p]),
    'sdab_randn': (c_int, [c_void_p, c_size_t, c_uint64, c_uint64, c_void_p]),
    'sdab_tweedie': (c_int, [c_void_p, c_void_p, c_float, c_float, c_void_p, c_size_t, c_void_p]),
    'sdab_tweedie_dev': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'sdab_axpy': (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_size_t, c_void_p]),
    # Kolmogorov
    'sdab_kolmogorov_create': (c_int, [c_int, c_double, c_double, POINTER(c_void_p)]),
    'sdab_kolmogorov_destroy': (None, [c_void_p]),
    'sdab_kolmogorov_inner_steps': (c_int, [c_void_p]),
    'sdab_kolmogorov_workspace_bytes': (c_size_t, [c_void_p, c_int]),
    'sdab_kolmogorov_transition': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'sdab_kolmogorov_prior': (c_int, [c_void_p, c_void_p, c_int, c_uint64, c_void_p, c_size_t, c_void_p]),
    'sdab_coarsen': (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_void_p]),
    'sdab_vorticity': (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    'sdab_coarsen_adjoint': (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_void_p]),
    'sdab_vorticity_adjoint': (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    'sdab_upsample_bilinear': (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_void_p]),
    'sdab_upsample_bilinear_adjoint': (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_void_p]),
}

SYMBOLS = tuple(_PROTOTYPES)

_lib = None


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    r"""Loads (building first if needed) libsdab.so.  Raises if that is impossible."""

    global _lib

    if _lib is not None:
        return _lib

    # build() returns at once when the source digest matches its stamp, and rebuilds a stale library
    # after an edit of csrc/ or a pull (a stale .so with changed signatures would otherwise be loaded)
    if build_if_missing and not os.environ.get('SDAB_NO_BUILD'):
        from . import build as _build

        _build.build()
    elif not LIB_PATH.exists():
        raise RuntimeError(f'{LIB_PATH} is missing: run `python -m sda_b200.build` (there is no CPU fallback)')

    # SDAB_LIB: developer A/B switch (another build of the SAME ABI, e.g. sda_b200/build/libsdab_alt.so)
    lib = ctypes.CDLL(os.environ.get('SDAB_LIB') or str(LIB_PATH))

    for name, (restype, argtypes) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the binding diverge
        fn.restype = restype
        fn.argtypes = argtypes

    _lib = lib

    return lib


def check(status: int) -> None:
    r"""Raises RuntimeError(sdab_last_error()) on a non-zero status."""

    if status != 0:
        msg = load().sdab_last_error()
        raise RuntimeError(f'libsdab error {status}: {msg.decode() if msg else "unknown"}')


def stream_ptr() -> int:
    r"""cudaStream_t of torch's current stream."""

    import torch

    return torch.cuda.current_stream().cuda_stream


def launch_count(reset: bool = False) -> int:
    return int(load().sdab_launch_count(1 if reset else 0))
