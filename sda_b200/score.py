r"""Score modules -- drop-in for ``sda.score`` (reference: /root/reference/sda/score.py).

Same classes, constructor arguments and call signatures as the reference.  What
changes is where the arithmetic runs:

* ``MCScoreNet``    window unfold / fold and their adjoints are libsdab kernels
                    (sdab_unfold_cat, sdab_fold, sdab_fold_transpose,
                    sdab_unfold_transpose_add) wrapped in autograd Functions;
* ``ScoreUNet``     its ``network`` is ``sda_b200.nn.UNet`` (tcgen05 convolutions);
* ``VPSDE.sample``  schedule scalars are computed once on the host, the predictor and
                    corrector updates are single fused kernels with a counter-based
                    Philox stream (identical on every rank);
* ``GaussianScore`` the Tweedie estimate is a fused kernel with an analytic adjoint;
                    the user's observation operator ``A`` stays ordinary PyTorch.

On CPU tensors (Lorenz plumbing) the window maps, sampler and
guidance run the reference's plain PyTorch formulas; the 2-D U-Net itself never
runs on the CPU.
"""

from __future__ import annotations

import math
import os
from typing import Callable, Optional, Union

import torch
import torch.nn as nn
from torch import Size, Tensor

from . import _lib
from .nn import *  # noqa: F401,F403  (the reference re-exports sda.nn from sda.score)
from .nn import ResMLP, UNet, input_gradient_only
from . import nn as _nn

try:  # progress bar as in the reference (sda/score.py:250); optional here
    from tqdm import tqdm
except Exception:  # pragma: no cover
    tqdm = None


def broadcast(*tensors: Tensor, ignore: Union[int, list] = 0):
    r"""Broadcasts all but the last `ignore` dimensions (zuko.utils.broadcast, used at
    sda/score.py:57,60,87)."""

    if isinstance(ignore, int):
        ignore = [ignore] * len(tensors)

    dims = [t.dim() - i for t, i in zip(tensors, ignore)]
    common = torch.broadcast_shapes(*(t.shape[:d] for t, d in zip(tensors, dims)))

    return [torch.broadcast_to(t, common + t.shape[d:]) for t, d in zip(tensors, dims)]


class TimeEmbedding(nn.Sequential):
    r"""cos/sin features of pi * (1..16) * t through a 2-layer MLP.  Reference: sda/score.py:15-35."""

    def __init__(self, features: int):
        super().__init__(nn.Linear(32, 256), nn.SiLU(), nn.Linear(256, features))

        self.register_buffer('freqs', torch.pi * torch.arange(1, 16 + 1))

    def forward(self, t: Tensor) -> Tensor:
        t = self.freqs * t.unsqueeze(dim=-1)
        t = torch.cat((t.cos(), t.sin()), dim=-1)

        return super().forward(t)


class ScoreNet(nn.Module):
    r"""MLP score network (Lorenz plumbing, plain PyTorch).  Reference: sda/score.py:38-63."""

    def __init__(self, features: int, context: int = 0, embedding: int = 16, **kwargs):
        super().__init__()

        self.embedding = TimeEmbedding(embedding)
        self.network = ResMLP(features + context + embedding, features, **kwargs)

    def forward(self, x: Tensor, t: Tensor, c: Tensor = None) -> Tensor:
        t = self.embedding(t)
        parts = (x, t) if c is None else (x, t, c)

        return self.network(torch.cat(broadcast(*parts, ignore=1), dim=-1))


class ScoreUNet(nn.Module):
    r"""U-Net score network.  Reference: sda/score.py:66-93.

    Subclassable exactly like the reference's (experiments/kolmogorov/utils.py:29-46
    overrides `forward` and registers a `forcing` buffer).
    """

    def __init__(self, channels: int, context: int = 0, embedding: int = 64, **kwargs):
        super().__init__()

        self.embedding = TimeEmbedding(embedding)
        self.network = UNet(channels + context, channels, embedding, **kwargs)

    def forward(self, x: Tensor, t: Tensor, c: Tensor = None) -> Tensor:
        if isinstance(x, WindowBatch):
            return _score_windows(self, x, t, c)

        dims = self.network.spatial + 1

        if c is None:
            y = x
        else:
            y = torch.cat(broadcast(x, c, ignore=dims), dim=-dims)

        y = y.reshape(-1, *y.shape[-dims:])
        t = self.embedding(t.reshape(-1))

        return self.network(y, t).reshape(x.shape)


class MCScoreWrapper(nn.Module):
    r"""Disguises a `ScoreUNet` over time as a Markov-chain score.  Reference: sda/score.py:96-110."""

    def __init__(self, score: nn.Module):
        super().__init__()

        self.score = score

    def forward(self, x: Tensor, t: Tensor, c: Tensor = None) -> Tensor:
        return self.score(x.transpose(1, 2), t, c).transpose(1, 2)


# --------------------------------------------------------------------------- window maps
def _is_native(x: Tensor) -> bool:
    return x.is_cuda and x.dtype == torch.float32 and x.dim() == 5


class _Unfold(torch.autograd.Function):
    r"""(B, L, C, H, W) -> (B, L-2k, (2k+1) C, H, W), materialised by sdab_unfold_cat;
    backward = deterministic overlap-add (sdab_unfold_transpose_add)."""

    @staticmethod
    def forward(ctx, x: Tensor, order: int) -> Tensor:
        B, L, C, H, W = x.shape
        ctx.dims = (B, L, C, H, W, order)
        x = x.contiguous()
        out = torch.empty((B, L - 2 * order, (2 * order + 1) * C, H, W), dtype=x.dtype, device=x.device)

        with torch.cuda.device(x.device):
            _lib.check(_lib.load().sdab_unfold_cat(x.data_ptr(), None, out.data_ptr(), B, L, C, 0, H, W, order, _lib.stream_ptr()))

        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        B, L, C, H, W, order = ctx.dims
        g = g.contiguous()
        gx = torch.empty((B, L, C, H, W), dtype=g.dtype, device=g.device)

        with torch.cuda.device(g.device):
            _lib.check(_lib.load().sdab_unfold_transpose_add(g.data_ptr(), gx.data_ptr(), B, L, C, 0, H, W, order, _lib.stream_ptr()))

        return gx, None


class _Fold(torch.autograd.Function):
    r"""(B, L-2k, (2k+1) C, H, W) -> (B, L, C, H, W) by sdab_fold; backward = sdab_fold_transpose."""

    @staticmethod
    def forward(ctx, s: Tensor, order: int) -> Tensor:
        B, nw, CH, H, W = s.shape
        C = CH // (2 * order + 1)
        L = nw + 2 * order
        ctx.dims = (B, L, C, H, W, order)
        s = s.contiguous()
        out = torch.empty((B, L, C, H, W), dtype=s.dtype, device=s.device)

        with torch.cuda.device(s.device):
            _lib.check(_lib.load().sdab_fold(s.data_ptr(), out.data_ptr(), B, L, C, H, W, order, _lib.stream_ptr()))

        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        B, L, C, H, W, order = ctx.dims
        g = g.contiguous()
        gs = torch.empty((B, L - 2 * order, (2 * order + 1) * C, H, W), dtype=g.dtype, device=g.device)

        with torch.cuda.device(g.device):
            _lib.check(_lib.load().sdab_fold_transpose(g.data_ptr(), gs.data_ptr(), B, L, C, H, W, order, _lib.stream_ptr()))

        return gs, None


class WindowBatch:
    r"""Stand-in for the unfolded window tensor (B, L - 2k, (2k+1) C, H, W) that `MCScoreNet.forward` hands to
    a `ScoreUNet` kernel on the native path: it carries the TRAJECTORY, and the network reads its windows (and
    writes the folded score) by addressing (sdab_mcscore_forward) instead of through materialised unfold / cat /
    fold tensors.  A `ScoreUNet` subclass that only forwards `x` to `super().forward` (as the reference's
    LocalScoreUNet does, experiments/kolmogorov/utils.py:45-46) works unchanged; one that computes on `x` should
    call `x.materialize()` first, or set `MCScoreNet.fuse_windows = False`."""

    def __init__(self, x: Tensor, order: int, group=None):
        self.x, self.order, self.group = x, order, group

    @property
    def shape(self):
        B, L, C, H, W = self.x.shape
        return torch.Size((B, L - 2 * self.order, (2 * self.order + 1) * C, H, W))

    def materialize(self) -> Tensor:
        return MCScoreNet.unfold(self.x, self.order)


def shard_geometry(n_windows: int, windows_per_trajectory: int, order: int, rank: int, world: int):
    r"""Window range [begin, end) of `rank`, windows per rank `per` and frames per shard `cap` of the
    window-sharded evaluation (include/sdab.h: sdab_mcscore_forward)."""

    per = -(-n_windows // world)
    begin = min(rank * per, n_windows)
    end = min(begin + per, n_windows)
    touched = (per + windows_per_trajectory - 2) // windows_per_trajectory + 1

    return begin, end, per, per + 2 * order * touched


class _RawCuda:
    r"""Device memory owned by libsdab, presented to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 2}


class PeerBuffer:
    r"""One buffer of `numel` fp32 per rank in device memory owned by the library, exported by CUDA IPC and mapped
    by every other rank of `group` (csrc/peer.cu; all ranks on one box).  `view` is the own payload as a tensor,
    `ptrs` the ctypes array of every rank's base pointer as mapped here.  Construction is collective; `error` is
    set instead of raising so that the caller can agree on a fallback with the other ranks."""

    def __init__(self, numel: int, group, device):
        import ctypes

        import torch.distributed as dist

        lib = _lib.load()
        sharded = dist.is_available() and dist.is_initialized() and group is not False
        self.world = dist.get_world_size(group) if sharded else 1
        self.rank = dist.get_rank(group) if sharded else 0
        self.header = int(lib.sdab_peer_header_bytes())
        self._own, self._opened, self.error, self.view = None, [], None, None
        nbytes = 4 * int(numel)
        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        self.ptrs = (ctypes.c_void_p * self.world)()

        with torch.cuda.device(device):
            try:
                _lib.check(lib.sdab_peer_alloc(nbytes, ctypes.byref(ptr), handle))
                self._own = ptr.value
            except RuntimeError as e:
                self.error = str(e)

            handles = [handle.raw if self.error is None else None]

            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, None if self.error else handle.raw, group=group)

            for p, h in enumerate(handles):
                if p == self.rank:
                    self.ptrs[p] = ptr.value
                elif h is not None and self.error is None:
                    q = ctypes.c_void_p()

                    try:
                        _lib.check(lib.sdab_peer_open(h, ctypes.byref(q)))
                        self._opened.append(q.value)
                        self.ptrs[p] = q.value
                    except RuntimeError as e:
                        self.error = str(e)
                else:
                    self.error = self.error or f'rank {p} could not allocate its peer buffer'

            if self.world > 1:
                errors = [None] * self.world
                dist.all_gather_object(errors, self.error, group=group)

                if any(errors):
                    self.error = '; '.join(f'rank {p}: {e}' for p, e in enumerate(errors) if e)

            if self.error is None:
                self.view = torch.as_tensor(_RawCuda(ptr.value + self.header, nbytes), device=device).view(torch.float32)
            else:
                self.close()

    def close(self) -> None:
        lib = _lib.load()

        for q in self._opened:
            lib.sdab_peer_close(q)

        if self._own:
            lib.sdab_peer_free(self._own)

        self._opened, self._own, self.view = [], None, None


class PeerExchange:
    r"""The exchange step of the window-sharded score over NVLink peer memory (csrc/peer.cu): two alternating
    `PeerBuffer`s of `shape` fp32 per rank.  `acquire()` hands out the rank's buffer of this turn, `gather()` is ONE
    kernel that pushes the rank's shard to all peers, signals, and retires when every peer's shard has arrived.
    Construction is collective; it raises on every rank if any rank cannot map its peers."""

    def __init__(self, shape, group, device):
        numel = int(torch.Size(shape).numel())
        self.buffers = [PeerBuffer(numel, group, device) for _ in range(2)]
        errors = [b.error for b in self.buffers if b.error]

        if errors:
            self.close()
            raise RuntimeError('peer-memory exchange unavailable: ' + errors[0])

        self.world, self.rank = self.buffers[0].world, self.buffers[0].rank
        self.views = [b.view.view(shape) for b in self.buffers]
        self.uses, self.turn = [0, 0], 0

    def acquire(self) -> Tensor:
        self.turn ^= 1
        self.uses[self.turn] += 1

        return self.views[self.turn]

    def gather(self, shard_offset_bytes: int, shard_bytes: int) -> None:
        r"""Completes the buffer handed out by the last `acquire()` (enqueued on the current stream)."""

        ptrs = self.buffers[self.turn].ptrs
        _lib.check(_lib.load().sdab_peer_allgather(ptrs, self.rank, self.world, shard_offset_bytes, shard_bytes, self.uses[self.turn], _lib.stream_ptr()))

    def close(self) -> None:
        for b in self.buffers:
            b.close()

        self.views = []


def _exchange(net: UNet, key, shape, group, device):
    r"""Persistent exchange buffers of `net` for `key`: a PeerExchange (transport 'peer', the default of
    sda_b200.parallel.shard_windows) or a plain tensor for NCCL's all_gather_into_tensor (transport 'nccl',
    SDAB_PEER_GATHER=0, or after the peer mapping failed -- with a warning, on every rank alike)."""

    ex = net._buffers_mc.get(key)

    if ex is None:
        import os

        transport = getattr(net, '_shard_transport', 'peer')

        if os.environ.get('SDAB_PEER_GATHER', '1') == '0':
            transport = 'nccl'

        if transport == 'peer':
            try:
                ex = PeerExchange(shape, group, device)
            except RuntimeError as e:
                import warnings

                warnings.warn(f'{e} -- falling back to NCCL all-gather')
                net._shard_transport = 'nccl'

        if ex is None:
            ex = torch.empty(shape, dtype=torch.float32, device=device)

        net._buffers_mc[key] = ex

    return ex


class _MCScore(torch.autograd.Function):
    r"""MCScoreNet.forward (sda/score.py:134-144) as one library call per rank: unfold, the context concat and
    fold are addressing inside the network's first and last layer.  Window-sharded (group of > 1 ranks): every
    rank evaluates a contiguous range of the flattened windows and ONE all-gather (NCCL over NVLink) moves the
    frames each range feeds (32 MiB at 256 x 256, L = 64, whatever the number of ranks); the backward all-gathers
    the window input-gradients and overlap-adds them locally in fixed order, so the result is bit-identical for
    any number of ranks.  Differentiable w.r.t. the trajectory only."""

    @staticmethod
    def forward(ctx, x: Tensor, y: Tensor, cvals, net: UNet, order: int, group) -> Tensor:
        import torch.distributed as dist

        B, L, C, H, W = x.shape
        nw = L - 2 * order
        sharded = group is not False
        world = dist.get_world_size(group) if sharded else 1
        rank = dist.get_rank(group) if sharded else 0
        begin, end, per, cap = shard_geometry(B * nw, nw, order, rank, world)
        save = int(getattr(_nn._scope, 'grad_enabled', True) and ctx.needs_input_grad[0])
        x = x.detach().contiguous()
        out = torch.empty_like(x)
        lib = _lib.load()

        if world == 1:
            net._native_mcscore_forward(x, y, cvals, order, 0, B * nw, out, 0, 0, save)
        else:
            ex = _exchange(net, ('fwd', world, cap, C, H, W, x.device), (world * cap, C, H, W), group, x.device)
            buf = ex.acquire() if isinstance(ex, PeerExchange) else ex

            if end > begin:
                net._native_mcscore_forward(x, y, cvals, order, begin, end, buf[rank * cap:], per, cap, save)

            if isinstance(ex, PeerExchange):
                with torch.cuda.device(x.device):
                    ex.gather(rank * cap * C * H * W * 4, cap * C * H * W * 4)
            else:
                dist.all_gather_into_tensor(buf, buf[rank * cap:(rank + 1) * cap], group=group)

            with torch.cuda.device(x.device):
                _lib.check(lib.sdab_frames_assemble(buf.data_ptr(), out.data_ptr(), B, L, C, H, W, order, per, cap, _lib.stream_ptr()))

        ctx.net, ctx.save, ctx.token = net, save, net._forward_token
        ctx.meta = (order, group, world, rank, begin, end, per, 0 if cvals is None else cvals.shape[0])

        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        import torch.distributed as dist

        net = ctx.net
        order, group, world, rank, begin, end, per, Cc = ctx.meta

        if not ctx.save:
            return (None,) * 6

        if ctx.token != net._forward_token:
            raise RuntimeError(
                'sda_b200.score.MCScoreNet: backward through a forward pass whose saved activations were '
                'overwritten by a later forward of the same network'
            )

        g = g.detach().to(torch.float32).contiguous()
        B, L, C, H, W = g.shape
        key, shape = ('bwd', world, per, C, H, W, g.device), (world * per, (2 * order + 1) * C, H, W)

        if world > 1:
            ex = _exchange(net, key, shape, group, g.device)
        else:
            ex = net._buffers_mc.get(key)

            if ex is None:
                ex = net._buffers_mc[key] = torch.empty(shape, dtype=torch.float32, device=g.device)

        gwin = ex.acquire() if isinstance(ex, PeerExchange) else ex

        if end > begin:
            net._native_mcscore_dgrad(g, gwin, Cc, order, begin, end)

        if isinstance(ex, PeerExchange):
            shard = per * (2 * order + 1) * C * H * W * 4

            with torch.cuda.device(g.device):
                ex.gather(rank * shard, shard)
        elif world > 1:
            dist.all_gather_into_tensor(gwin, gwin[rank * per:(rank + 1) * per], group=group)

        gx = torch.empty_like(g)

        with torch.cuda.device(g.device):
            _lib.check(_lib.load().sdab_unfold_transpose_add(gwin.data_ptr(), gx.data_ptr(), B, L, C, 0, H, W, order, _lib.stream_ptr()))

        return gx, None, None, None, None, None


def _score_windows(kernel: 'ScoreUNet', wb: WindowBatch, t: Tensor, c) -> Tensor:
    r"""ScoreUNet.forward on a WindowBatch: returns the FOLDED score (B, L, C, H, W)."""

    x = wb.x
    H, W = x.shape[-2:]

    if c is not None:
        # the context must be one stack of planes shared by every window (the forcing channel of
        # LocalScoreUNet); anything else goes through the materialised windows
        shared = c.dim() >= 3 and tuple(c.shape[-2:]) == (H, W) and all(d == 1 for d in c.shape[:-3])

        if not shared or c.requires_grad:
            xw = wb.materialize()
            return MCScoreNet.fold(ScoreUNet.forward(kernel, xw, t, c), wb.order)

        c = c.detach().to(torch.float32).reshape(-1, H, W).contiguous()

    y = kernel.embedding(t.reshape(-1))
    prev = getattr(_nn._scope, 'grad_enabled', True)
    _nn._scope.grad_enabled = torch.is_grad_enabled()

    try:
        return _MCScore.apply(x, y, c, kernel.network, wb.order, wb.group)
    finally:
        _nn._scope.grad_enabled = prev


class MCScoreNet(nn.Module):
    r"""Score of a Markov chain composed from window scores.  Reference: sda/score.py:113-164.

    `window_range = (begin, end)` (optional attribute, default all) restricts the evaluated
    windows to a contiguous range: the hook used by `sda_b200.parallel` to shard windows over GPUs.
    """

    def __init__(self, features: int, context: int = 0, order: int = 1, **kwargs):
        super().__init__()

        self.order = order

        if kwargs.get('spatial', 0) > 0:
            build = ScoreUNet
        else:
            build = ScoreNet

        self.kernel = build(features * (2 * order + 1), context, **kwargs)

    fuse_windows = True  # native path: unfold / context concat / fold as addressing inside the network
    shard_group = False  # False: not sharded; None / a process group: window-sharded (sda_b200.parallel)
    _sdab_sharded = False

    def forward(self, x: Tensor, t: Tensor, c: Tensor = None) -> Tensor:
        if self._fusable(x, t):
            return self.kernel(WindowBatch(x, self.order, self._group()), t, c)

        x = self.unfold(x, self.order)
        s = self.kernel(x, t, c)
        s = self.fold(s, self.order)

        return s

    def _group(self):
        import torch.distributed as dist

        if self._sdab_sharded and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.shard_group) > 1:
            return self.shard_group

        return False

    def _fusable(self, x: Tensor, t: Tensor) -> bool:
        r"""The fused window path serves evaluations that are differentiated w.r.t. the trajectory at most
        (sampling, likelihood guidance): CUDA fp32 trajectory, one diffusion time, native U-Net kernel."""

        if not (self.fuse_windows and self.order >= 1 and _is_native(x) and isinstance(self.kernel, ScoreUNet)):
            return False

        net = self.kernel.network

        if not (isinstance(net, UNet) and net._native and t.dim() == 0 and not t.requires_grad):
            return False

        if x.shape[1] < 2 * self.order + 1 or (2 * self.order + 1) * x.shape[2] != net.out_channels:
            return False

        if torch.is_grad_enabled() and not getattr(_nn._scope, 'input_only', False):
            return not any(p.requires_grad for p in self.kernel.parameters())

        return True

    @staticmethod
    def unfold(x: Tensor, order: int) -> Tensor:
        if order >= 1 and _is_native(x):
            if x.shape[1] < 2 * order + 1:
                raise RuntimeError(
                    f'maximum size for tensor at dimension 1 is {x.shape[1]} but size is {2 * order + 1}'
                )

            return _Unfold.apply(x, order)

        # reference formula (zero-copy view), sda/score.py:148-153
        return x.unfold(1, 2 * order + 1, 1).movedim(-1, 2).flatten(2, 3)

    @staticmethod
    def fold(x: Tensor, order: int) -> Tensor:
        if order >= 1 and _is_native(x) and x.shape[2] % (2 * order + 1) == 0:
            return _Fold.apply(x, order)

        # reference formula, sda/score.py:157-164
        x = x.unflatten(2, (2 * order + 1, -1))

        return torch.cat((x[:, 0, :order], x[:, :, order], x[:, -1, -order:]), dim=1)


# --------------------------------------------------------------------------- VPSDE
class VPSDE(nn.Module):
    r"""Variance-preserving SDE noise schedule, sampler and loss.  Reference: sda/score.py:167-276.

    mu(t) = alpha(t), sigma(t)^2 = 1 - alpha(t)^2 + eta^2.
    """

    def __init__(self, eps: nn.Module, shape: Size, alpha: str = 'cos', eta: float = 1e-3):
        super().__init__()

        self.eps = eps
        self.shape = shape
        self.dims = tuple(range(-len(shape), 0))
        self.eta = eta

        if alpha == 'lin':
            self.alpha = lambda t: 1 - (1 - eta) * t
        elif alpha == 'cos':
            self.alpha = lambda t: torch.cos(math.acos(math.sqrt(eta)) * t) ** 2
        elif alpha == 'exp':
            self.alpha = lambda t: torch.exp(math.log(eta) * t**2)
        else:
            raise ValueError()

        self.register_buffer('device', torch.empty(()))

        # test hook: callable(x) -> z replacing the corrector's Philox draw (noise injection)
        self.noise_source: Optional[Callable[[Tensor], Tensor]] = None

    def mu(self, t: Tensor) -> Tensor:
        return self.alpha(t)

    def sigma(self, t: Tensor) -> Tensor:
        return (1 - self.alpha(t) ** 2 + self.eta**2).sqrt()

    def forward(self, x: Tensor, t: Tensor, train: bool = False) -> Tensor:
        r"""Samples from the perturbation kernel p(x(t) | x)."""

        t = t.reshape(t.shape + (1,) * len(self.shape))

        eps = torch.randn_like(x)
        x = self.mu(t) * x + self.sigma(t) * eps

        if train:
            return x, eps
        else:
            return x

    def sample(
        self,
        shape: Size = (),
        c: Tensor = None,
        steps: int = 64,
        corrections: int = 0,
        tau: float = 1.0,
    ) -> Tensor:
        r"""Predictor-corrector sampling of p(x(0)).  Reference: sda/score.py:225-263."""

        shape = tuple(shape)

        # initial noise drawn on the CPU then moved, as the reference does (score.py:243)
        x = torch.randn(shape + tuple(self.shape)).to(self.device)
        x = x.reshape(-1, *self.shape).contiguous()
        # window-sharded evaluation (sda_b200.parallel) needs a bit-identical state on every rank:
        # rank 0's draw wins, whatever the ranks' own generators hold
        x = self._from_shard_root(x)

        state = self.sampler_state(x, steps)
        iterator = range(steps)

        if tqdm is not None and not os.environ.get('SDAB_NO_TQDM'):
            iterator = tqdm(iterator, ncols=88)

        for i in iterator:
            x = self.denoise_step(x, i, state, c=c, corrections=corrections, tau=tau)

        return x.reshape(shape + tuple(self.shape))

    def sampler_state(self, x: Tensor, steps: int) -> dict:
        r"""Per-run constants of the sampling loop: the time grid (device), the schedule scalars
        (computed once on the host in the reference's fp32 arithmetic, score.py:252-253) and the
        Philox stream of the corrector noise."""

        dt = 1 / steps
        time_host = torch.linspace(1, 0, steps + 1).to(self.device.dtype)
        mu_t, mu_n = self.mu(time_host), self.mu(time_host - dt)
        sg_t, sg_n = self.sigma(time_host), self.sigma(time_host - dt)
        ratio = mu_n / mu_t
        state = {
            'dt': dt,
            'time': torch.linspace(1, 0, steps + 1).to(x.device),
            'ratio': ratio,
            'coef': sg_n - ratio * sg_t,
            'sigma_next': sg_n,
            'native': x.is_cuda and x.dtype == torch.float32,
            'draws': 0,
        }

        if state['native']:
            lib = _lib.load()
            state['scratch'] = torch.empty(
                lib.sdab_vpsde_correct_scratch_floats(x.shape[0]), dtype=torch.float32, device=x.device
            )
            seed = torch.randint(0, 2**62, (1,), dtype=torch.int64)  # torch's CPU generator
            state['seed'] = int(self._from_shard_root(seed.to(x.device)))

        return state

    def _from_shard_root(self, v: Tensor) -> Tensor:
        r"""Broadcasts `v` from the first rank of the shard group when `eps` contains a window-sharded
        MCScoreNet (identity otherwise): the ranks of a sharded sampler must agree on the initial noise
        and on the Philox seed, and torchrun does not seed them identically."""

        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            return v

        for m in self.eps.modules():
            if getattr(m, '_sdab_sharded', False) and dist.get_world_size(m.shard_group) > 1:
                src = 0 if m.shard_group is None else dist.get_global_rank(m.shard_group, 0)
                v = v.contiguous()
                dist.broadcast(v, src=src, group=m.shard_group)
                break

        return v

    def denoise_step(self, x: Tensor, i: int, state: dict, c: Tensor = None, corrections: int = 0, tau: float = 1.0) -> Tensor:
        r"""One iteration of the sampling loop (score.py:250-261): predictor + `corrections` Langevin
        corrections, i.e. (1 + corrections) score evaluations.  On CUDA `x` is updated in place."""

        dt, t = state['dt'], state['time'][i]
        ratio, coef, sigma_next = state['ratio'][i], state['coef'][i], state['sigma_next'][i]

        with torch.no_grad():
            # Predictor
            eps = self.eps(x, t, c)

            if state['native']:
                lib = _lib.load()
                B = x.shape[0]
                eps = eps.contiguous()

                with torch.cuda.device(x.device):
                    _lib.check(lib.sdab_vpsde_predict(x.data_ptr(), eps.data_ptr(), float(ratio), float(coef), x.numel(), _lib.stream_ptr()))
            else:
                x = ratio * x + coef * eps

            # Corrector
            for _ in range(corrections):
                if state['native']:
                    z = self.noise_source(x).contiguous() if self.noise_source is not None else None
                    eps = self.eps(x, t - dt, c).contiguous()

                    with torch.cuda.device(x.device):
                        _lib.check(
                            lib.sdab_vpsde_correct(
                                x.data_ptr(), eps.data_ptr(), None if z is None else z.data_ptr(), float(tau),
                                float(sigma_next), state['seed'], state['draws'] * ((x.numel() + 3) // 4 + B), B,
                                x.numel(), state['scratch'].data_ptr(), _lib.stream_ptr(),
                            )
                        )

                    state['draws'] += 1
                else:
                    # (a window-sharded sampler on this plain-PyTorch path takes rank 0's draw)
                    z = self.noise_source(x) if self.noise_source is not None else self._from_shard_root(torch.randn_like(x))
                    eps = self.eps(x, t - dt, c)
                    delta = tau / eps.square().mean(dim=self.dims, keepdim=True)

                    x = x - (delta * eps + torch.sqrt(2 * delta) * z) * sigma_next

        return x

    def loss(self, x: Tensor, c: Tensor = None, w: Tensor = None) -> Tensor:
        r"""Denoising loss.  Reference: sda/score.py:265-276."""

        t = torch.rand(x.shape[0], dtype=x.dtype, device=x.device)
        x, eps = self.forward(x, t, train=True)

        err = (self.eps(x, t, c) - eps).square()

        if w is None:
            return err.mean()
        else:
            return (err * w).mean() / w.mean()


class SubVPSDE(VPSDE):
    r"""sigma(t) = 1 - alpha(t)^2 + eta.  Reference: sda/score.py:279-288."""

    def sigma(self, t: Tensor) -> Tensor:
        return 1 - self.alpha(t) ** 2 + self.eta


class SubSubVPSDE(VPSDE):
    r"""sigma(t) = 1 - alpha(t) + eta.  Reference: sda/score.py:291-300."""

    def sigma(self, t: Tensor) -> Tensor:
        return 1 - self.alpha(t) + self.eta


# --------------------------------------------------------------------------- guidance
class _Tweedie(torch.autograd.Function):
    r"""x_hat = (x - sigma eps) / mu as one kernel (sdab_tweedie_dev) with its analytic adjoint.  mu and sigma
    stay 0-d device tensors: nothing on this path synchronises the host with the GPU."""

    @staticmethod
    def forward(ctx, x: Tensor, eps: Tensor, mu: Tensor, sigma: Tensor) -> Tensor:
        mu = mu.detach().to(device=x.device, dtype=torch.float32).contiguous()
        sigma = sigma.detach().to(device=x.device, dtype=torch.float32).contiguous()
        ctx.save_for_backward(mu, sigma)
        x, eps = x.contiguous(), eps.contiguous()
        out = torch.empty_like(x)

        with torch.cuda.device(x.device):
            _lib.check(_lib.load().sdab_tweedie_dev(x.data_ptr(), eps.data_ptr(), mu.data_ptr(), sigma.data_ptr(), out.data_ptr(), x.numel(), _lib.stream_ptr()))

        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        mu, sigma = ctx.saved_tensors
        gx = g / mu if ctx.needs_input_grad[0] else None
        ge = g * (-sigma / mu) if ctx.needs_input_grad[1] else None

        return gx, ge, None, None


def _tweedie(x: Tensor, eps: Tensor, mu: Tensor, sigma: Tensor) -> Tensor:
    if x.is_cuda and x.dtype == torch.float32 and eps.shape == x.shape and mu.dim() == 0 and sigma.dim() == 0:
        return _Tweedie.apply(x, eps, mu, sigma)

    return (x - sigma * eps) / mu


class DPSGaussianScore(nn.Module):
    r"""Diffusion posterior sampling guidance.  Reference: sda/score.py:303-344.

    Returns -sigma(t) s(x(t), t | y).  The signature `forward(x, t)` is the reference's (it
    does not accept the context that `VPSDE.sample` passes; kept as is).
    """

    def __init__(self, y: Tensor, A: Callable[[Tensor], Tensor], sde: VPSDE, zeta: float = 1.0):
        super().__init__()

        self.register_buffer('y', y)

        self.A = A
        self.sde = sde
        self.zeta = zeta

    def forward(self, x: Tensor, t: Tensor) -> Tensor:
        mu, sigma = self.sde.mu(t), self.sde.sigma(t)

        with torch.enable_grad():
            x = x.detach().requires_grad_(True)

            with input_gradient_only():
                eps = self.sde.eps(x, t)

            x_ = _tweedie(x, eps, mu, sigma)
            err = (self.y - self.A(x_)).square().sum()

        (s,) = torch.autograd.grad(err, x)
        s = -s * self.zeta / err.sqrt()

        return eps - sigma * s


class GaussianScore(nn.Module):
    r"""Likelihood guidance for p(y | x) = N(y | A(x), Sigma).  Reference: sda/score.py:347-396.

    Returns -sigma(t) s(x(t), t | y).
    """

    def __init__(
        self,
        y: Tensor,
        A: Callable[[Tensor], Tensor],
        std: Union[float, Tensor],
        sde: VPSDE,
        gamma: Union[float, Tensor] = 1e-2,
        detach: bool = False,
    ):
        super().__init__()

        self.register_buffer('y', y)
        self.register_buffer('std', torch.as_tensor(std))
        self.register_buffer('gamma', torch.as_tensor(gamma))

        self.A = A
        self.sde = sde
        self.detach = detach

    def forward(self, x: Tensor, t: Tensor, c: Tensor = None) -> Tensor:
        mu, sigma = self.sde.mu(t), self.sde.sigma(t)

        if self.detach:
            eps = self.sde.eps(x, t, c)

        with torch.enable_grad():
            x = x.detach().requires_grad_(True)

            if not self.detach:
                with input_gradient_only():
                    eps = self.sde.eps(x, t, c)

            x_ = _tweedie(x, eps, mu, sigma)

            err = self.y - self.A(x_)
            var = self.std**2 + self.gamma * (sigma / mu) ** 2

            log_p = -(err**2 / var).sum() / 2

        (s,) = torch.autograd.grad(log_p, x)

        return (eps - sigma * s).detach()
