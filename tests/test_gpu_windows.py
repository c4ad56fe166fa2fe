r"""The fused window path (sdab_mcscore_forward / sdab_mcscore_dgrad: MCScoreNet.unfold, the context concat
and fold of sda/score.py:87,146-164 as addressing inside the network's first and last layer) against the
materialised path, bit for bit -- unsharded, and with the shard layout of every world size emulated on ONE
GPU (virtual ranks evaluated one after the other into the gather buffers, no collective), so that the
geometry the multi-GPU path relies on is pinned even where only one GPU is available."""

import pytest
import torch

from helpers import build_score
from oracle.testing import randn

pytestmark = pytest.mark.gpu


def _guided(score, x, y, t, A):
    import sda_b200.score as sc

    return sc.GaussianScore(y, A=A, std=0.1, sde=sc.VPSDE(score, shape=()), gamma=1e-2).cuda()(x, t)


@pytest.mark.parametrize('name, B, L', [('net_small', 1, 3), ('net_small', 2, 7), ('net_config', 1, 5), ('net_config', 2, 6)])
def test_fused_windows_equal_materialised_windows(name, B, L):
    score, k = build_score(name, 16, 'cuda')
    x = randn((B, L, 2, 16, 16), seed=1).cuda()
    y = randn((B, L, 2, 8, 8), seed=2).cuda()
    t = torch.tensor(0.45).cuda()
    A = lambda v: v[..., ::2, ::2]  # noqa: E731
    outs = {}

    for fuse in (True, False):
        score.fuse_windows = fuse

        with torch.no_grad():
            eps = score(x, t)

        outs[fuse] = (eps, _guided(score, x, y, t, A))

    assert torch.equal(outs[True][0], outs[False][0])
    assert torch.equal(outs[True][1], outs[False][1])


def test_fused_path_is_taken_and_falls_back():
    import sda_b200.score as sc
    from sda_b200 import _lib

    score, k = build_score('net_small', 16, 'cuda')
    x = randn((1, 6, 2, 16, 16), seed=3).cuda()
    t = torch.tensor(0.3).cuda()
    with torch.no_grad():
        assert score._fusable(x, t)
        assert not score._fusable(x, torch.tensor([0.3]).cuda())  # per-trajectory times: materialised path
        assert not score._fusable(x.cpu(), t.cpu())

    with torch.enable_grad():  # trainable parameters and grad mode on: the training path keeps autograd's route
        assert not score._fusable(x, t)

        with sc.input_gradient_only():
            assert score._fusable(x, t)

    # a per-window context cannot be addressed as shared planes: same numbers through the materialised windows
    kern = score.kernel
    plain = sc.ScoreUNet.forward(kern, sc.MCScoreNet.unfold(x, k), t, kern.forcing)
    ctx = kern.forcing.expand(1, x.shape[1] - 2 * k, 1, 16, 16).contiguous()

    with torch.no_grad():
        via_proxy = sc.ScoreUNet.forward(kern, sc.WindowBatch(x, k, False), t, ctx)

    assert torch.equal(via_proxy, sc.MCScoreNet.fold(plain.detach(), k))
    _lib.launch_count(reset=True)

    with torch.no_grad():
        score(x, t)

    fused = _lib.launch_count()
    score.fuse_windows = False
    _lib.launch_count(reset=True)

    with torch.no_grad():
        score(x, t)

    assert fused < _lib.launch_count()  # no unfold_cat / fold launches on the fused path


@pytest.mark.parametrize('B, L, world', [(1, 9, 2), (1, 9, 3), (1, 12, 8), (2, 7, 4), (3, 6, 4), (2, 5, 5), (1, 5, 2)])
def test_shard_layout_of_every_world_size_on_one_gpu(B, L, world):
    r"""Virtual ranks: each writes its frames into its shard of the gather buffer and its window input-gradients
    into its slice, then the local kernels (frames assemble, ordered overlap-add) finish -- bit-identical to
    the unsharded evaluation, for even and uneven splits, ranks without windows and ranks spanning trajectories."""

    import sda_b200.score as sc
    from sda_b200 import _lib
    from sda_b200.nn import input_gradient_only

    score, k = build_score('net_small', 16, 'cuda')
    kern, net = score.kernel, score.kernel.network
    lib = _lib.load()
    C, H, W = 2, 16, 16
    nw = L - 2 * k
    x = randn((B, L, C, H, W), seed=5).cuda()
    g = randn((B, L, C, H, W), seed=6).cuda()
    t = torch.tensor(0.6).cuda()

    xr = x.clone().requires_grad_(True)

    with input_gradient_only():
        ref = score(xr, t)

    (ref_gx,) = torch.autograd.grad(ref, xr, g)

    with torch.no_grad():
        y = kern.embedding(t.reshape(-1))

    ctx = kern.forcing.reshape(-1, H, W).contiguous()
    geo = [sc.shard_geometry(B * nw, nw, k, r, world) for r in range(world)]
    per, cap = geo[0][2], geo[0][3]
    assert all(gm[2:] == (per, cap) for gm in geo) and geo[-1][1] == B * nw
    buf = torch.full((world * cap, C, H, W), float('nan'), device='cuda')
    gwin = torch.full((world * per, (2 * k + 1) * C, H, W), float('nan'), device='cuda')

    for r, (begin, end, _, _) in enumerate(geo):
        if end > begin:
            net._native_mcscore_forward(x, y, ctx, k, begin, end, buf[r * cap:], per, cap, 1)
            net._native_mcscore_dgrad(g, gwin, 1, k, begin, end)

    out, gx = torch.empty_like(x), torch.empty_like(x)
    _lib.check(lib.sdab_frames_assemble(buf.data_ptr(), out.data_ptr(), B, L, C, H, W, k, per, cap, _lib.stream_ptr()))
    _lib.check(lib.sdab_unfold_transpose_add(gwin.data_ptr(), gx.data_ptr(), B, L, C, 0, H, W, k, _lib.stream_ptr()))
    assert torch.equal(out, ref.detach())
    assert torch.equal(gx, ref_gx)
