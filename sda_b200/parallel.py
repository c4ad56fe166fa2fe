r"""Window sharding of the Markov-blanket score over the GPUs of one box.

The reference has no distributed code; its long-trajectory
scaling is algorithmic: the score of a trajectory is composed from independent window
scores (sda/score.py:134-144).  That independence is the data-parallel axis used here:

* one process per GPU (torchrun), every rank holds the full trajectory `x`, advances
  it with the same counter-based Philox noise, and therefore stays bit-identical
  (`VPSDE.sample` / `sampler_state` broadcast the initial noise and the seed from the
  first rank of the group);
* per score evaluation each rank runs the U-Net on a contiguous range of the flattened
  (B, L - 2k) windows -- it reads its windows straight from the trajectory and writes
  only the frames they feed (`fold` keeps one of the 2k + 1 slots of an interior
  window) into its shard of ONE in-place all-gather (NCCL over NVLink / NVSwitch):
  32 MiB in total at 256 x 256, L = 64, instead of the 157 MB of window scores;
* when the evaluation is differentiated (GaussianScore), the backward pass runs the
  input-VJP of the local windows only and all-gathers the window input-gradients in
  place; the overlap-add that follows (`unfold` adjoint) is local and in fixed order,
  so the result does not depend on the number of ranks.

`shard_windows(score)` switches a `MCScoreNet` to this mode; nothing else changes for
the caller.  Evaluations the fused window path does not serve (CPU tensors, per-sample
times, a kernel that is not a native `ScoreUNet`, parameter gradients) fall back to
`_ShardedKernel`: the windows are materialised, the local range goes through the kernel
and the window scores / window input-gradients are all-gathered.
"""

from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from .score import MCScoreNet, shard_geometry


def window_range(n_windows: int, rank: int, world: int):
    r"""Contiguous range [begin, end) of flattened windows owned by `rank`, and the padded
    per-rank count (equal on all ranks, as all_gather needs)."""

    per = -(-n_windows // world)
    begin = min(rank * per, n_windows)
    end = min(begin + per, n_windows)

    return begin, end, per


class _ShardedKernel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xw: Tensor, kernel, t: Tensor, c: Optional[Tensor], group) -> Tensor:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        B, nw = xw.shape[:2]
        flat = xw.reshape(B * nw, *xw.shape[2:])
        begin, end, per = window_range(B * nw, rank, world)
        local = flat[begin:end]
        need_grad = ctx.needs_input_grad[0]

        if end > begin:
            if need_grad:
                with torch.enable_grad():
                    local = local.detach().requires_grad_(True)
                    out_local = kernel(local.unsqueeze(0), t, c).squeeze(0)
            else:
                out_local = kernel(local.unsqueeze(0), t, c).squeeze(0)
        else:
            out_local = flat.new_zeros((0,) + tuple(flat.shape[1:]))

        ctx.saved = (local, out_local) if need_grad else None
        ctx.meta = (B, nw, begin, end, per, group, tuple(flat.shape[1:]))

        return _gather(out_local.detach(), per, B * nw, group).reshape(xw.shape)

    @staticmethod
    def backward(ctx, g: Tensor):
        B, nw, begin, end, per, group, tail = ctx.meta
        local, out_local = ctx.saved
        g_local = g.reshape(B * nw, *tail)[begin:end]

        if end > begin:
            (gx_local,) = torch.autograd.grad(out_local, local, g_local.contiguous())
        else:
            gx_local = g.new_zeros((0,) + tail)

        gx = _gather(gx_local, per, B * nw, group).reshape(g.shape)

        return gx, None, None, None, None


def _gather(local: Tensor, per: int, total: int, group) -> Tensor:
    r"""all-gather of equally padded shards; returns the first `total` rows."""

    world = dist.get_world_size(group)
    padded = local.new_zeros((per,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    out = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)

    return out[:total]


class ShardedMCScoreNet(MCScoreNet):
    r"""`MCScoreNet` whose kernel evaluations are sharded over `shard_group` (see module docstring)."""

    _sdab_sharded = True
    shard_group = None

    def forward(self, x: Tensor, t: Tensor, c: Tensor = None) -> Tensor:
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.shard_group) == 1 or t.dim() > 0:
            return super().forward(x, t, c)

        if self._fusable(x, t):
            return super().forward(x, t, c)  # fused window path; the shard group travels in the WindowBatch

        xw = self.unfold(x, self.order)
        s = _ShardedKernel.apply(xw, self.kernel, t, c, self.shard_group)

        return self.fold(s, self.order)


def shard_windows(score: MCScoreNet, group=None, transport: str = 'peer') -> MCScoreNet:
    r"""Switches `score` (in place) to window-sharded evaluation over `group`.

    transport: how the fused window path exchanges its shards -- 'peer' (default): one kernel per rank over NVLink
    peer memory (`sda_b200.score.PeerExchange`, all ranks on one box); 'nccl': `all_gather_into_tensor`.  Results
    are bit-identical either way."""

    if not isinstance(score, MCScoreNet):
        raise TypeError('shard_windows expects a MCScoreNet')

    if transport not in ('peer', 'nccl'):
        raise ValueError("transport must be 'peer' or 'nccl'")

    score.__class__ = ShardedMCScoreNet
    score.shard_group = group
    network = getattr(score.kernel, 'network', None)

    if network is not None:
        network._shard_transport = transport

    return score


def allreduce_gradients(module: torch.nn.Module, group=None) -> None:
    r"""Data-parallel training step (BASELINE config 5; the reference trains on one GPU, sda/utils.py:136-143):
    averages the parameter gradients of `module` over `group` after `loss.backward()`.  The convolution gradients
    of every native `UNet` inside live in one flat buffer (`UNet._grad_flat`, written by sdab_unet_backward), which
    is all-reduced in place with ONE collective; the remaining few gradients (projection Linears, time embedding:
    about 1 % of the parameters) go through one coalesced all-reduce.  Same result as DistributedDataParallel's
    averaging, without its bucket copies."""

    from .nn import UNet

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return

    world = dist.get_world_size(group)
    covered = set()

    for m in module.modules():
        flat = getattr(m, '_grad_flat', None) if isinstance(m, UNet) else None

        if flat is None:
            continue

        convs, _ = m._ordered_parameters()
        params = [c.weight for c in convs] + [c.bias for c in convs]
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4

        # the views are only trusted while every gradient still points into the flat buffer (autograd keeps
        # the tensors it is handed when .grad was None: optimizer.zero_grad(set_to_none=True), the default)
        if all(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in params):
            dist.all_reduce(flat, group=group)
            flat.div_(world)
            covered.update(id(p) for p in params)

    rest = [p.grad for p in module.parameters() if p.grad is not None and id(p) not in covered]

    if rest:
        packed = torch.cat([g.reshape(-1) for g in rest])
        dist.all_reduce(packed, group=group)
        packed.div_(world)

        for g, v in zip(rest, packed.split([g.numel() for g in rest])):
            g.copy_(v.view_as(g))


class PeerAdamW(torch.optim.Optimizer):
    r"""`torch.optim.AdamW` for data-parallel training of a module with native `UNet`s (BASELINE config 5; the
    reference trains on one GPU with `torch.optim.AdamW`, sda/utils.py:125-143) as ONE kernel per rank and step over
    NVLink peer memory (csrc/peer.cu: sdab_peer_adamw).

    All parameters live in one flat buffer per rank, all gradients in another (the convolution gradients are
    written there by sdab_unet_backward, the few others accumulate there through autograd); both are mapped by
    every rank of `group`.  `step()` lets rank r read the gradients of its 1 / world slice from all ranks (peer
    loads, added in rank order: deterministic), apply AdamW to that slice -- it holds the moments of its slice
    only -- and store the new parameters into every rank's parameter buffer.  No all-reduce, no bucket copies, no
    separate broadcast; the parameters stay bit-identical across the ranks.  With one rank it is a plain fused AdamW.

        opt = PeerAdamW(sde, lr=2e-4, weight_decay=1e-3)      # collective; broadcasts rank 0's parameters
        loss.backward(); opt.step(); opt.zero_grad()

    `lr` is read from `param_groups[0]` at every step, so `torch.optim.lr_scheduler` works as usual."""

    def __init__(self, module: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, group=None):
        from . import _lib
        from .nn import UNet
        from .score import PeerBuffer

        sharded = dist.is_available() and dist.is_initialized()
        self.group = group if sharded else False
        self.world = dist.get_world_size(group) if sharded else 1
        self.rank = dist.get_rank(group) if sharded else 0

        # order: per native UNet its convolution weights, then its convolution biases (the order of the library's
        # flat gradient), then every other trainable parameter of the module
        ordered, seen, self._nets = [], set(), []

        for m in module.modules():
            if isinstance(m, UNet) and m._native:
                convs, _ = m._ordered_parameters()
                ps = [c.weight for c in convs] + [c.bias for c in convs]

                if all(p.requires_grad and id(p) not in seen for p in ps):
                    self._nets.append((m, len(ordered), sum(p.numel() for p in ps), ps))
                    ordered += ps
                    seen.update(id(p) for p in ps)

        self._convs = set(seen)
        ordered += [p for p in module.parameters() if p.requires_grad and id(p) not in seen]

        if not ordered or any(p.dtype != torch.float32 or not p.is_cuda for p in ordered):
            raise ValueError('PeerAdamW expects trainable fp32 CUDA parameters')

        device = ordered[0].device
        n = sum(p.numel() for p in ordered)
        quantum = 4 * self.world
        self.numel = -(-n // quantum) * quantum
        self.params_buf = PeerBuffer(self.numel, self.group, device)
        self.grads_buf = PeerBuffer(self.numel, self.group, device)

        for b in (self.params_buf, self.grads_buf):
            if b.error:
                raise RuntimeError('PeerAdamW: peer-memory buffers unavailable: ' + b.error)

        flat_p, flat_g = self.params_buf.view, self.grads_buf.view
        flat_p.zero_(), flat_g.zero_()
        offset, self._views = 0, []

        with torch.no_grad():
            for p in ordered:
                k = p.numel()
                flat_p[offset:offset + k].copy_(p.reshape(-1))
                p.data = flat_p[offset:offset + k].view_as(p)
                self._views.append((p, flat_g[offset:offset + k].view_as(p)))
                offset += k

            if self.world > 1:
                dist.broadcast(flat_p, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)

        for net, start, size, _ in self._nets:
            net._grad_target = flat_g[start:start + size]
            net.invalidate_packed()

        per = self.numel // self.world
        self.begin, self.end = self.rank * per, (self.rank + 1) * per
        self.m = torch.zeros(per, dtype=torch.float32, device=device)
        self.v = torch.zeros(per, dtype=torch.float32, device=device)
        self.steps, self._tail = 0, sum(size for _, _, size, _ in self._nets)
        self._lib = _lib
        super().__init__(ordered, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.zero_grad()

    def zero_grad(self, set_to_none: bool = True) -> None:
        r"""`.grad = None` for every parameter: the native backward hands autograd views of the flat gradient
        buffer for the convolutions, which autograd keeps as they are; the few other gradients are gathered into
        the flat buffer by `step()` with one multi-tensor copy."""

        for p, _ in self._views:
            p.grad = None

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise ValueError('PeerAdamW does not take a closure')

        flat_g = self.grads_buf.view
        lo, hi = flat_g.data_ptr(), flat_g.data_ptr() + 4 * flat_g.numel()
        dst, src = [], []

        for p, gview in self._views:
            if p.grad is None:
                gview.zero_()  # took no part in this backward
            elif not (lo <= p.grad.data_ptr() < hi):
                dst.append(gview), src.append(p.grad)  # not written in place: projection Linears, time embedding

        if dst:
            torch._foreach_copy_(dst, src)

        g = self.param_groups[0]
        self.steps += 1
        lib = self._lib.load()

        with torch.cuda.device(flat_g.device):
            # every rank's gradients are complete before anyone reads them
            self._lib.check(lib.sdab_peer_signal_wait(self.grads_buf.ptrs, self.rank, self.world, self.steps, self._lib.stream_ptr()))
            self._lib.check(
                lib.sdab_peer_adamw(
                    self.grads_buf.ptrs, self.params_buf.ptrs, self.m.data_ptr(), self.v.data_ptr(), self.begin, self.end,
                    self.rank, self.world, float(g['lr']), float(g['betas'][0]), float(g['betas'][1]), float(g['eps']),
                    float(g['weight_decay']), self.steps, self.steps, self._lib.stream_ptr(),
                )
            )

        # the parameters changed behind autograd's back (no version bump): the packed tensor-core weights are stale
        for net, *_ in self._nets:
            net.invalidate_packed()


__all__ = ['shard_windows', 'ShardedMCScoreNet', 'window_range', 'shard_geometry', 'allreduce_gradients', 'PeerAdamW']
