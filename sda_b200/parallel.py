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


__all__ = ['shard_windows', 'ShardedMCScoreNet', 'window_range', 'shard_geometry', 'allreduce_gradients']
