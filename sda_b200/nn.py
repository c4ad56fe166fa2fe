r"""Neural networks -- drop-in for ``sda.nn`` (reference: /root/reference/sda/nn.py).

Same class names, constructor arguments, ``forward`` signatures and ``state_dict``
keys as the reference.  ``UNet`` keeps its parameters in ordinary ``nn.Conv2d`` /
``nn.Linear`` modules (so ``load_state_dict``, ``.cuda()`` and optimizers work) but,
for the Kolmogorov configuration (2-D, kernel 3, stride 2, circular padding --
experiments/kolmogorov/utils.py:59-68), ``forward`` runs the whole network through
libsdab: hand-written sm_100a kernels behind the C ABI of ``include/sdab.h``.
That path has no fallback: CPU tensors, non-sm_100 devices or a missing library
raise.

Configurations outside the hot path (1-D U-Net of the Lorenz global score,
``ResMLP``; SURVEY.md section 8a rows a6 and config 1 "plumbing, no GPU") are plain
PyTorch by design and never touch the library.
"""

from __future__ import annotations

import contextlib
import ctypes
import os
import threading
from typing import Callable, Sequence, Union

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib

__all__ = ['LayerNorm', 'ResidualBlock', 'ModResidualBlock', 'ResMLP', 'UNet']


class LayerNorm(nn.Module):
    r"""Standardisation over `dim`, unbiased variance, no affine parameters.

    Restates ``zuko.nn.LayerNorm`` (zuko 0.1.4), imported by the reference at
    sda/nn.py:8 and used at :61, :137, :163.
    """

    def __init__(self, dim: Union[int, Sequence[int]] = -1, eps: float = 1e-5):
        super().__init__()

        self.dim = dim if isinstance(dim, int) else tuple(dim)
        self.eps = eps

    def forward(self, x: Tensor) -> Tensor:
        variance, mean = torch.var_mean(x, dim=self.dim, keepdim=True)

        return (x - mean) / (variance + self.eps).sqrt()

    def extra_repr(self) -> str:
        return f'dim={self.dim}'


class ResidualBlock(nn.Sequential):
    r"""x + f(x).  Reference: sda/nn.py:11-15."""

    def forward(self, x: Tensor) -> Tensor:
        return x + super().forward(x)


class ModResidualBlock(nn.Module):
    r"""x + residue(x + project(y)).  Reference: sda/nn.py:18-28."""

    def __init__(self, project: nn.Module, residue: nn.Module):
        super().__init__()

        self.project = project
        self.residue = residue

    def forward(self, x: Tensor, y: Tensor) -> Tensor:
        return x + self.residue(x + self.project(y))


class ResMLP(nn.Sequential):
    r"""Residual MLP (plain PyTorch; Lorenz plumbing).  Reference: sda/nn.py:31-71."""

    def __init__(
        self,
        in_features: int,
        out_features: int,
        hidden_features: Sequence[int] = (64, 64),
        activation: Callable[[], nn.Module] = nn.ReLU,
        **kwargs,
    ):
        layers = []
        widths = [in_features, *hidden_features, out_features]

        for before, after in zip(widths[:-1], widths[1:]):
            if after != before:
                layers.append(nn.Linear(before, after, **kwargs))

            layers.append(
                ResidualBlock(
                    LayerNorm(),
                    nn.Linear(after, after, **kwargs),
                    activation(),
                    nn.Linear(after, after, **kwargs),
                )
            )

        super().__init__(*layers)

        self.in_features = in_features
        self.out_features = out_features


_ACTIVATION_CODES = {nn.SiLU: _lib.ACT_SILU, nn.ReLU: _lib.ACT_RELU}


def _mode() -> int:
    r"""Arithmetic mode of the tensor-core convolutions (env SDAB_MODE: bf16x3 | bf16)."""

    name = os.environ.get('SDAB_MODE', 'bf16x3').lower()

    if name not in ('bf16x3', 'bf16'):
        raise ValueError(f'SDAB_MODE must be bf16x3 or bf16, not {name}')

    return _lib.MODE_BF16X3 if name == 'bf16x3' else _lib.MODE_BF16


def _engine() -> int:
    name = os.environ.get('SDAB_ENGINE', 'umma').lower()

    if name not in ('umma', 'simt'):
        raise ValueError(f'SDAB_ENGINE must be umma or simt, not {name}')

    return _lib.ENGINE_UMMA if name == 'umma' else _lib.ENGINE_SIMT


_scope = threading.local()


@contextlib.contextmanager
def input_gradient_only():
    r"""Inside this context the native U-Net differentiates w.r.t. its input only, even though its
    parameters require gradients: the likelihood guidance (GaussianScore / DPSGaussianScore,
    sda/score.py:381-394) asks autograd for d log p / d x alone, and a custom Function cannot see
    which of its inputs a torch.autograd.grad call is after."""

    prev = getattr(_scope, 'input_only', False)
    _scope.input_only = True

    try:
        yield
    finally:
        _scope.input_only = prev


class _UNetFunction(torch.autograd.Function):
    r"""UNet.forward and its backward through libsdab.

    Guided sampling differentiates w.r.t. the input only (sdab_unet_forward(save=1) /
    sdab_unet_dgrad); when a parameter or the modulation vector requires a gradient (training,
    VPSDE.loss -> sda/utils.py:loop) the forward saves the training state (save=2) and the
    backward is sdab_unet_backward: input, convolution weight / bias and shift-table gradients
    from the library, the gradients of the projection Linears and of y derived from the latter.
    `params` are the parameters in the library's order ((weight, bias) of every convolution, then
    of every projection) so that autograd routes their gradients."""

    @staticmethod
    def forward(ctx, x: Tensor, y: Tensor, net: 'UNet', *params: Tensor) -> Tensor:
        # grad mode is always off inside Function.forward: UNet.forward recorded the caller's
        grad = getattr(_scope, 'grad_enabled', True)
        train = grad and any(ctx.needs_input_grad[1:]) and not getattr(_scope, 'input_only', False)
        save = 2 if train else int(grad and bool(ctx.needs_input_grad[0]))
        out = net._native_forward(x, y, save)
        ctx.net = net
        ctx.save = save
        ctx.token = net._forward_token

        if train:
            ctx.save_for_backward(y.detach())

        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        net = ctx.net
        nparams = len(ctx.needs_input_grad) - 3

        if not ctx.save:
            return (None,) * (3 + nparams)

        if ctx.token != net._forward_token:
            raise RuntimeError(
                'sda_b200.nn.UNet: backward through a forward pass whose saved activations were '
                'overwritten by a later forward of the same module'
            )

        if ctx.save == 1:
            return (net._native_dgrad(g), None, None) + (None,) * nparams

        (y,) = ctx.saved_tensors
        gx, conv_grads, dshift = net._native_backward(g, y.reshape(-1, y.shape[-1]).shape[0])
        convs, projs = net._ordered_parameters()
        # the 18 projection Linears (sda/nn.py:132-135) share their input y, so their gradients are slices of
        # three GEMMs on the concatenated shift table instead of 54 small ones
        y2 = y.reshape(-1, y.shape[-1]).to(torch.float32)
        gw_all = dshift.t() @ y2            # (sum C, mod)
        gb_all = dshift.sum(dim=0)          # (sum C)
        proj_grads, off = [], 0

        for m in projs:
            proj_grads += [gw_all[off:off + m.out_features], gb_all[off:off + m.out_features]]
            off += m.out_features

        gy = dshift @ torch.cat([m.weight.detach() for m in projs]) if ctx.needs_input_grad[1] else None
        grads = conv_grads + proj_grads
        grads = [gr if need else None for gr, need in zip(grads, ctx.needs_input_grad[3:])]

        return (
            gx if ctx.needs_input_grad[0] else None,
            gy.reshape(y.shape) if gy is not None else None,
            None,
            *grads,
        )


_optimizer_steps = [0]


def _count_optimizer_step(optimizer, args, kwargs) -> None:
    _optimizer_steps[0] += 1


try:  # global post-step hook of torch.optim (every optimizer instance, present and future)
    from torch.optim.optimizer import register_optimizer_step_post_hook

    register_optimizer_step_post_hook(_count_optimizer_step)
except ImportError:  # pragma: no cover -- older torch: fused optimizers then need UNet.invalidate_packed()
    pass


class UNet(nn.Module):
    r"""U-Net with additive time modulation.  Reference: sda/nn.py:74-206.

    Arguments are the reference's: `in_channels, out_channels, mod_features,
    hidden_channels, hidden_blocks, kernel_size, stride, activation, spatial` and
    `**kwargs` forwarded to the convolutions (e.g. `padding_mode='circular'`).
    """

    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        mod_features: int,
        hidden_channels: Sequence[int] = (32, 64, 128),
        hidden_blocks: Sequence[int] = (2, 3, 5),
        kernel_size: Union[int, Sequence[int]] = 3,
        stride: Union[int, Sequence[int]] = 2,
        activation: Callable[[], nn.Module] = nn.ReLU,
        spatial: int = 2,
        **kwargs,
    ):
        super().__init__()

        self.in_channels = in_channels
        self.out_channels = out_channels
        self.spatial = spatial

        conv = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}[spatial]
        ks = [kernel_size] * spatial if isinstance(kernel_size, int) else list(kernel_size)
        st = [stride] * spatial if isinstance(stride, int) else list(stride)
        kwargs.update(kernel_size=ks, padding=[k // 2 for k in ks])

        def block(c: int) -> ModResidualBlock:
            return ModResidualBlock(
                project=nn.Sequential(nn.Linear(mod_features, c), nn.Unflatten(-1, (-1,) + (1,) * spatial)),
                residue=nn.Sequential(LayerNorm(-(spatial + 1)), conv(c, c, **kwargs), activation(), conv(c, c, **kwargs)),
            )

        heads, tails, descent, ascent = [], [], [], []

        for i, n in enumerate(hidden_blocks):
            c = hidden_channels[i]

            if i == 0:
                heads.append(conv(in_channels, c, **kwargs))
                tails.append(conv(c, out_channels, **kwargs))
            else:
                below = hidden_channels[i - 1]
                heads.append(nn.Sequential(conv(below, c, stride=st, **kwargs)))
                tails.append(
                    nn.Sequential(
                        LayerNorm(-(spatial + 1)),
                        nn.Upsample(scale_factor=tuple(st), mode='nearest'),
                        conv(c, below, **kwargs),
                    )
                )

            descent.append(nn.ModuleList(block(c) for _ in range(n)))
            ascent.append(nn.ModuleList(block(c) for _ in range(n)))

        # same registration order and naming as the reference (state_dict compatibility)
        self.heads = nn.ModuleList(heads)
        self.tails = nn.ModuleList(reversed(tails))
        self.descent = nn.ModuleList(descent)
        self.ascent = nn.ModuleList(reversed(ascent))

        # ---- native path bookkeeping (not part of the state_dict)
        act_code = _ACTIVATION_CODES.get(activation if isinstance(activation, type) else type(activation()))
        self._native = (
            spatial == 2
            and ks == [3, 3]
            and st == [2, 2]
            and kwargs.get('padding_mode', 'zeros') == 'circular'
            and act_code is not None
            and all(c % 32 == 0 and 32 <= c <= 512 for c in hidden_channels)
            and len(hidden_channels) == len(hidden_blocks) <= _lib.MAX_DEPTH
            and kwargs.get('bias', True)
            and kwargs.get('dilation', 1) in (1, [1, 1], (1, 1))
            and kwargs.get('groups', 1) == 1
        )
        self._desc = (in_channels, out_channels, mod_features, tuple(hidden_channels), tuple(hidden_blocks), act_code)
        self._handle = None
        self._packed = None
        self._packed_key = None
        self._workspace = None
        self._forward_token = 0
        self._saved_level = 0
        self._buffers_mc = {}  # persistent gather buffers of the window-sharded evaluation
        self._grad_flat = None  # flat buffer behind the convolution gradients of the last training backward

    # ------------------------------------------------------------------ reference module-tree forward
    def _module_forward(self, x: Tensor, y: Tensor) -> Tensor:
        r"""Literal module-tree evaluation (sda/nn.py:184-206) for configurations outside the
        native path (1-D / non-circular U-Nets of the Lorenz experiments)."""

        memory = []

        for head, blocks in zip(self.heads, self.descent):
            x = head(x)

            for blk in blocks:
                x = blk(x, y)

            memory.append(x)

        memory.pop()

        for blocks, tail in zip(self.ascent, self.tails):
            for blk in blocks:
                x = blk(x, y)

            x = tail(x) + memory.pop() if memory else tail(x)

        return x

    # ------------------------------------------------------------------ native path
    def _ordered_parameters(self):
        r"""Parameters in the library's canonical order (include/sdab.h: sdab_unet_set_weights)."""

        D = len(self.heads)
        convs, projs = [], []

        def add_blocks(mods):
            for blk in mods:
                convs.extend((blk.residue[1], blk.residue[3]))
                projs.append(blk.project[0])

        for d in range(D):
            convs.append(self.heads[d] if d == 0 else self.heads[d][0])
            add_blocks(self.descent[d])

        for i in range(D):  # ascent[i] / tails[i]: i = 0 is the deepest level
            add_blocks(self.ascent[i])
            convs.append(self.tails[i][2] if i < D - 1 else self.tails[i])

        return convs, projs

    def __getstate__(self):
        state = self.__dict__.copy()

        for k in ('_handle', '_packed', '_packed_key', '_workspace'):
            state[k] = None

        state['_buffers_mc'] = {}
        state['_grad_target'] = None
        state['_grad_flat'] = None

        return state

    def invalidate_packed(self) -> None:
        r"""Forces the bf16-packed weight copies to be rebuilt at the next forward.  The cache is keyed on
        (data_ptr, version) of every parameter, which in-place updates through `.data` (EMA, weight
        surgery) do not change: call this after such an update.  `load_state_dict`, `.to()`, `.cuda()`
        and friends call it themselves, and every `torch.optim.Optimizer.step()` invalidates as well (fused
        optimizers do not bump version counters).  SDAB_ALWAYS_REPACK=1 repacks at every forward."""

        self._packed_key = None

    def _apply(self, fn, *args, **kwargs):
        self._packed_key = None

        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self._packed_key = None

        return super()._load_from_state_dict(*args, **kwargs)

    def _ensure_handle(self, device: torch.device):
        lib = _lib.load()

        if self._handle is None:
            in_c, out_c, mod, channels, blocks, act = self._desc
            desc = _lib.UNetDesc()
            desc.in_channels, desc.out_channels, desc.mod_features = in_c, out_c, mod
            desc.depth = len(channels)

            for i, (c, b) in enumerate(zip(channels, blocks)):
                desc.hidden_channels[i] = c
                desc.hidden_blocks[i] = b

            desc.activation = act
            handle = ctypes.c_void_p()
            _lib.check(lib.sdab_unet_create(ctypes.byref(desc), ctypes.byref(handle)))
            self._handle = handle

        convs, projs = self._ordered_parameters()
        params = [p for m in convs + projs for p in (m.weight, m.bias)]
        # (data_ptr, version) of every parameter -- and the count of optimizer steps taken in this process: fused
        # optimizers (torch.optim.AdamW(fused=True), torch._fused_adamw_) update parameters WITHOUT bumping their
        # version counters, so every Optimizer.step() anywhere makes the packed copies stale
        key = (device, _optimizer_steps[0], tuple((p.data_ptr(), p._version) for p in params))

        if key != self._packed_key or os.environ.get('SDAB_ALWAYS_REPACK'):
            for p in params:
                if p.device != device or p.dtype != torch.float32:
                    raise RuntimeError('sda_b200.nn.UNet: parameters must be float32 on the same CUDA device as the input')

            assert len(convs) == lib.sdab_unet_num_convs(self._handle)
            assert len(projs) == lib.sdab_unet_num_blocks(self._handle)

            nbytes = lib.sdab_unet_packed_bytes(self._handle)

            if self._packed is None or self._packed.numel() < nbytes or self._packed.device != device:
                self._packed = torch.empty(nbytes, dtype=torch.uint8, device=device)

            def ptrs(tensors):
                # .contiguous() copies must outlive the call: keep them in `keep`
                arr = (ctypes.c_void_p * len(tensors))()
                keep = []
                for i, t in enumerate(tensors):
                    t = t.detach().contiguous()
                    keep.append(t)
                    arr[i] = t.data_ptr()
                return arr, keep

            cw, k1 = ptrs([m.weight for m in convs])
            cb, k2 = ptrs([m.bias for m in convs])
            pw, k3 = ptrs([m.weight for m in projs])
            pb, k4 = ptrs([m.bias for m in projs])
            _lib.check(
                lib.sdab_unet_set_weights(
                    self._handle, cw, cb, pw, pb, self._packed.data_ptr(), self._packed.numel(), _lib.stream_ptr()
                )
            )
            del k1, k2, k3, k4
            self._packed_key = key

        return lib

    def _get_workspace(self, lib, N: int, H: int, W: int, save: bool, device) -> Tensor:
        nbytes = lib.sdab_unet_workspace_bytes(self._handle, N, H, W, int(save))

        if nbytes == 0:
            raise RuntimeError('sda_b200.nn.UNet: invalid workspace query')

        ws = self._workspace

        if ws is None or ws.numel() < nbytes + 1024 or ws.device != device:
            self._workspace = None
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            self._workspace = ws

        return ws

    def _native_forward(self, x: Tensor, y: Tensor, save: bool) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError(
                'sda_b200.nn.UNet: the 2-D circular U-Net runs on sm_100 CUDA devices only (no CPU fallback); '
                f'got a tensor on {x.device}'
            )

        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise RuntimeError(f'expected input of shape (N, {self.in_channels}, H, W), got {tuple(x.shape)}')

        with torch.cuda.device(x.device):
            lib = self._ensure_handle(x.device)
            x = x.detach().to(torch.float32).contiguous()
            y = y.detach().to(torch.float32).reshape(-1, y.shape[-1]).contiguous()
            N, _, H, W = x.shape
            Nt = y.shape[0]

            if Nt not in (1, N):
                raise RuntimeError(f'the modulation batch ({Nt}) must be 1 or match the input batch ({N})')

            ws = self._get_workspace(lib, N, H, W, save, x.device)
            base = (ws.data_ptr() + 1023) // 1024 * 1024
            out = torch.empty((N, self.out_channels, H, W), dtype=torch.float32, device=x.device)
            self._forward_token += 1
            self._saved_mode = (_mode(), _engine())
            self._saved_level = int(save)
            _lib.check(
                lib.sdab_unet_forward(
                    self._handle, x.data_ptr(), y.data_ptr(), Nt, N, H, W, out.data_ptr(), base,
                    ws.numel() - (base - ws.data_ptr()), int(save), self._saved_mode[0], self._saved_mode[1],
                    _lib.stream_ptr(),
                )
            )

        return out

    def _native_dgrad(self, g: Tensor) -> Tensor:
        with torch.cuda.device(g.device):
            lib = _lib.load()
            g = g.detach().to(torch.float32).contiguous()
            N, _, H, W = g.shape
            ws = self._workspace
            base = (ws.data_ptr() + 1023) // 1024 * 1024
            gx = torch.empty((N, self.in_channels, H, W), dtype=torch.float32, device=g.device)
            _lib.check(
                lib.sdab_unet_dgrad(
                    self._handle, g.data_ptr(), gx.data_ptr(), base, ws.numel() - (base - ws.data_ptr()),
                    self._saved_mode[0], self._saved_mode[1], _lib.stream_ptr(),
                )
            )

        return gx

    def _native_backward(self, g: Tensor, Nt: int):
        r"""sdab_unet_backward: (gx, [dW_0, db_0, dW_1, ...] in the library's convolution order, dshift)."""

        with torch.cuda.device(g.device):
            lib = _lib.load()
            g = g.detach().to(torch.float32).contiguous()
            N, _, H, W = g.shape
            ws = self._workspace
            base = (ws.data_ptr() + 1023) // 1024 * 1024
            gx = torch.empty((N, self.in_channels, H, W), dtype=torch.float32, device=g.device)
            convs, _ = self._ordered_parameters()
            # one flat buffer for all convolution gradients: the tensors autograd receives are views of it, so a
            # data-parallel step all-reduces 99 % of the parameters with ONE in-place collective and no bucket
            # copies (sda_b200.parallel.allreduce_gradients)
            sizes = [m.weight.numel() for m in convs] + [m.bias.numel() for m in convs]
            # (sda_b200.parallel.PeerAdamW points _grad_target at its peer-visible gradient buffer)
            flat = getattr(self, '_grad_target', None)

            if flat is None or flat.numel() != sum(sizes) or flat.device != g.device:
                flat = torch.empty(sum(sizes), dtype=torch.float32, device=g.device)

            views = list(flat.split(sizes))
            dws = [v.view_as(m.weight) for v, m in zip(views[:len(convs)], convs)]
            dbs = views[len(convs):]
            self._grad_flat = flat
            dshift = torch.empty((Nt, lib.sdab_unet_shift_rows(self._handle)), dtype=torch.float32, device=g.device)
            pw = (ctypes.c_void_p * len(convs))(*[t.data_ptr() for t in dws])
            pb = (ctypes.c_void_p * len(convs))(*[t.data_ptr() for t in dbs])
            _lib.check(
                lib.sdab_unet_backward(
                    self._handle, g.data_ptr(), gx.data_ptr(), pw, pb, dshift.data_ptr(), base,
                    ws.numel() - (base - ws.data_ptr()), self._saved_mode[0], self._saved_mode[1], _lib.stream_ptr(),
                )
            )

        return gx, [t for pair in zip(dws, dbs) for t in pair], dshift

    # ------------------------------------------------------------------ trajectory-level entry (MCScoreNet)
    def _native_mcscore_forward(self, x: Tensor, y: Tensor, ctx, order: int, begin: int, end: int, out: Tensor,
                                per: int, cap: int, save: int) -> None:
        r"""sdab_mcscore_forward: the windows [begin, end) of trajectory `x` (B, L, C, H, W) through the network
        with unfold / context concat / fold as addressing; frames are written into `out` (include/sdab.h)."""

        with torch.cuda.device(x.device):
            lib = self._ensure_handle(x.device)
            B, L, C, H, W = x.shape
            Cc = 0 if ctx is None else ctx.shape[0]
            y = y.detach().to(torch.float32).reshape(-1, y.shape[-1]).contiguous()

            if y.shape[0] != 1:
                raise RuntimeError('the fused window path takes one diffusion time per call')

            ws = self._get_workspace(lib, end - begin, H, W, save, x.device)
            base = (ws.data_ptr() + 1023) // 1024 * 1024
            self._forward_token += 1
            self._saved_mode = (_mode(), _engine())
            self._saved_level = int(save)
            _lib.check(
                lib.sdab_mcscore_forward(
                    self._handle, x.data_ptr(), None if ctx is None else ctx.data_ptr(), y.data_ptr(), B, L, C, Cc, H, W,
                    order, begin, end, out.data_ptr(), per, cap, base, ws.numel() - (base - ws.data_ptr()), int(save),
                    self._saved_mode[0], self._saved_mode[1], _lib.stream_ptr(),
                )
            )

    def _native_mcscore_dgrad(self, g: Tensor, gwin: Tensor, Cc: int, order: int, begin: int, end: int) -> None:
        r"""sdab_mcscore_dgrad: cotangent of the folded score (B, L, C, H, W) -> window input-gradients of the
        windows [begin, end), written at gwin[begin:end] ((2k+1) C channels each)."""

        with torch.cuda.device(g.device):
            lib = _lib.load()
            B, L, C, H, W = g.shape
            ws = self._workspace
            base = (ws.data_ptr() + 1023) // 1024 * 1024
            _lib.check(
                lib.sdab_mcscore_dgrad(
                    self._handle, g.data_ptr(), gwin[begin:].data_ptr(), B, L, C, Cc, H, W, order, begin, end, base,
                    ws.numel() - (base - ws.data_ptr()), self._saved_mode[0], self._saved_mode[1], _lib.stream_ptr(),
                )
            )

    def forward(self, x: Tensor, y: Tensor) -> Tensor:
        if not self._native:
            return self._module_forward(x, y)

        convs, projs = self._ordered_parameters()
        params = [p for m in convs + projs for p in (m.weight, m.bias)]
        # under torch.no_grad() (unguided sampling, GaussianScore(detach=True), validation) nothing can
        # ever back-propagate: the forward must not save activations although needs_input_grad is set
        prev = getattr(_scope, 'grad_enabled', True)
        _scope.grad_enabled = torch.is_grad_enabled()

        try:
            return _UNetFunction.apply(x, y, self, *params)
        finally:
            _scope.grad_enabled = prev

    def __del__(self):
        handle = getattr(self, '_handle', None)

        if handle is not None:
            try:
                _lib.load().sdab_unet_destroy(handle)
            except Exception:
                pass
