"""Gradients of the hot path: the hand-written adjoint kernels against torch autograd
run over the CPU oracle (the reference obtains these gradients from autograd over the same ops)."""
import numpy as np
import pytest
import torch

from conftest import ATOL, RTOL
from oracle import fgn_oracle as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def close(got, want, atol=ATOL, rtol=RTOL, what=""):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = (got - want).abs()
    bad = err > atol + rtol * want.abs()
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} outside tol, max abs err {float(err.max()):.3e}"


def test_roi_align_multilevel_backward():
    from fgn_b200 import autograd as A
    from fgn_b200.episodes import synth_rois
    g = torch.Generator().manual_seed(301)
    strides, B, C = [4, 8, 16, 32], 2, 64
    feats = [torch.randn(B, C, 256 // s, 320 // s, generator=g) for s in strides]
    rois = synth_rois(g, 120, 256, 320, B, smin=4.0)
    rois[0, 1:] = torch.tensor([-20., -10., 200., 150.])
    gout = torch.randn(120, C, 7, 7, generator=g)
    fc = [f.clone().requires_grad_(True) for f in feats]
    want, _ = O.single_roi_extractor(fc, rois, strides, 7, 0, True, 56.0, "tv")
    want.backward(gout)
    fd = [f.to(dev()).contiguous(memory_format=torch.channels_last).requires_grad_(True) for f in feats]
    got = A.roi_align_multilevel(fd, rois.to(dev()), [1 / s for s in strides], 7, 0, True)
    close(got, want, what="forward")
    got.backward(gout.to(dev()))
    assert fc[0].grad is not None and fc[1].grad is not None
    for l in range(4):
        if fc[l].grad is None:                     # no RoI was assigned to this level
            assert float(fd[l].grad.abs().max()) == 0.0
        else:
            close(fd[l].grad, fc[l].grad, atol=2e-4, what=f"grad level {l}")  # atomics: order differs, sums of ~100 terms


def test_attention_backward():
    from fgn_b200 import autograd as A
    g = torch.Generator().manual_seed(302)
    B, N, K, C = 2, 3, 2, 64
    q = torch.randn(B, C, 12, 20, generator=g)
    s = torch.randn(B * N * K, C, 8, 8, generator=g)
    gout = torch.randn(B * N, C, 12, 20, generator=g)
    qc, sc = q.clone().requires_grad_(True), s.clone().requires_grad_(True)
    _, mod = O.agrpn_attention(qc, sc, N, K)
    mod.backward(gout)
    qd = q.to(dev()).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    sd = s.to(dev()).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    vec = A.attention_vectors(sd, N, K)
    out = A.channel_attention(qd, vec)
    close(out, mod, what="forward")
    out.backward(gout.to(dev()))
    close(qd.grad, qc.grad, what="grad qry")
    close(sd.grad, sc.grad, what="grad spp")


def test_support_pool_backward():
    from fgn_b200 import autograd as A
    g = torch.Generator().manual_seed(303)
    B, N, K, C = 2, 3, 2, 64
    f = torch.randn(B * N * K, C, 7, 7, generator=g)
    m = torch.rand(B * N * K, 1, 7, 7, generator=g)
    g_cat, g_gap = torch.randn(B, N, C, 7, 7, generator=g), torch.randn(B, N, C, 1, 1, generator=g)
    fc = f.clone().requires_grad_(True)
    cat = fc.view(B, N, K, C, 7, 7).mean(2)
    gap = (fc * m).view(B, N, K, C, 7, 7).mean((2, 4, 5)).view(B, N, C, 1, 1)
    (cat * g_cat).sum().backward(retain_graph=True)
    (gap * g_gap).sum().backward()
    fd = f.to(dev()).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    cd, gd = A.support_pool(fd, m.to(dev()), N, K)
    close(cd, cat, what="cat fwd")
    close(gd, gap, what="gap fwd")
    ((cd * g_cat.to(dev())).sum() + (gd * g_gap.to(dev())).sum()).backward()
    close(fd.grad, fc.grad, what="grad f")
