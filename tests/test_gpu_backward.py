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


@pytest.mark.parametrize("N,C,R,B", [(1, 64, 24, 1), (3, 64, 30, 2), (2, 256, 16, 1)])
def test_relation_fusion_backward(N, C, R, B):
    """fgn_relation_fusion_bwd against torch autograd over the oracle's line-for-line restatement of
    count_one_roi_by_n_spp + BBoxHead.forward + count_modified_cls_bbox (fgn_roi_head.py:253-279,338,302-326):
    gradients w.r.t. the RoI features, the class maps and all eight parameter tensors."""
    from fgn_b200 import autograd as A
    from fgn_b200.episodes import make_weights
    g = torch.Generator().manual_seed(500 + N + C)
    w = make_weights(C, 2)
    feats = torch.randn(R, C, 7, 7, generator=g)
    cat = torch.randn(B, N, C, 7, 7, generator=g)
    rois = torch.zeros(R, 5)
    rois[:, 0] = (torch.arange(R) * B // R).float()
    g_cls = torch.randn(R, N + 1, generator=g)
    g_reg = torch.randn(R, 4 * N, generator=g)
    names = ["conv_w", "conv_b", "gn_w", "gn_b", "fc_cls_w", "fc_cls_b", "fc_reg_w", "fc_reg_b"]
    # oracle + autograd (CPU)
    fo, co = feats.clone().requires_grad_(True), cat.clone().requires_grad_(True)
    po = {k: w[k].clone().requires_grad_(True) for k in names}
    _, fused = O.count_one_roi_by_n_spp(fo, rois, co, N, po["conv_w"], po["conv_b"], po["gn_w"], po["gn_b"])
    rc, rr = O.bbox_head_forward(fused, po["fc_cls_w"], po["fc_cls_b"], po["fc_reg_w"], po["fc_reg_b"])
    wc, wr = O.count_modified_cls_bbox(R, rc, rr, N)
    (wc * g_cls).sum().add((wr * g_reg).sum()).backward()
    # CUDA path
    fd = feats.to(dev()).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    cd = cat.to(dev()).requires_grad_(True)
    pd = {k: w[k].to(dev()).requires_grad_(True) for k in names}
    c, r = A.relation_fusion(fd, rois[:, 0].to(dev()), cd, N, *[pd[k] for k in names])
    close(c, wc, what="cls")
    close(r, wr, what="reg")
    (c * g_cls.to(dev())).sum().add((r * g_reg.to(dev())).sum()).backward()
    close(fd.grad, fo.grad, atol=2e-4, rtol=1e-3, what="d roi_feat")
    close(cd.grad, co.grad, atol=5e-4, rtol=1e-3, what="d spp_cat_mean")
    for k in names:
        scale = float(po[k].grad.abs().max()) + 1e-6
        close(pd[k].grad.reshape(po[k].grad.shape), po[k].grad, atol=2e-4 * max(1.0, scale), rtol=1e-3, what=f"d {k}")


def test_guided_heads_train_through_the_adjoint_kernels():
    """FGNRoIHead with autograd recording (forward_train's use, fgn_roi_head.py:344-358,451-529): count_spp ->
    _bbox_forward -> _mask_forward on maps that require grad; the loss gradient w.r.t. the query / support pyramids and
    the relation head's parameters equals autograd over the oracle."""
    from fgn_b200.episodes import CONFIGS, build_heads, episode_to_device, make_episode, make_weights
    cfg = CONFIGS["tiny_fpn"]
    ep = make_episode(cfg, seed=4)
    _, head = build_heads(cfg, dev(), seed=0)
    head.train()
    n_ext = len(cfg.strides)
    epd = episode_to_device(ep, dev())
    qd = [q.clone().requires_grad_(True) for q in epd["qry"][:n_ext]]
    sd = [s.clone().requires_grad_(True) for s in epd["spp"][:n_ext]]
    head.count_spp(sd, epd["spp_bboxes"].clone(), epd["spp_masks"])
    res = head._bbox_forward(qd, epd["rois"])
    assert res["cls_score"].grad_fn is not None and res["bbox_feats"] is not None
    head.gather_mask_vectors(epd["det_labels_list"])
    mres = head._mask_forward(qd, epd["det_rois"])
    g = torch.Generator().manual_seed(9)
    gc, gr = torch.randn(res["cls_score"].shape, generator=g), torch.randn(res["bbox_pred"].shape, generator=g)
    gm = torch.randn(mres["mask_feats"].shape, generator=g) * 0.1
    loss = (res["cls_score"] * gc.to(dev())).sum() + (res["bbox_pred"] * gr.to(dev())).sum() + \
        (mres["mask_feats"] * gm.to(dev())).sum()
    loss.backward()
    # oracle + autograd on the CPU
    w = {k: v.clone().requires_grad_(True) for k, v in make_weights(cfg.channels, 0).items()}
    qo = [q.clone().requires_grad_(True) for q in ep["qry"][:n_ext]]
    so = [s.clone().requires_grad_(True) for s in ep["spp"][:n_ext]]
    cat_mean, mp, _, _ = O.count_spp_fpn(so, cfg.strides, ep["spp_bboxes"].clone(), ep["spp_masks"], cfg.n_ways, cfg.k_shots)
    want = O.bbox_forward(qo, cfg.strides, ep["rois"], cat_mean, cfg.n_ways, w)
    mf = O.mask_attention(qo, cfg.strides, ep["det_rois"], mp, ep["det_labels_list"], cfg.n_ways, cfg.mask_size)
    close(res["cls_score"], want["cls_score"], what="cls (train mode)")
    close(mres["mask_feats"], mf, what="mask_feats (train mode)")
    ((want["cls_score"] * gc).sum() + (want["bbox_pred"] * gr).sum() + (mf * gm).sum()).backward()
    for l in range(n_ext):
        for got, ref, what in ((qd[l].grad, qo[l].grad, "qry"), (sd[l].grad, so[l].grad, "spp")):
            if ref is None:
                assert got is None or float(got.abs().max()) == 0.0
            else:
                close(got, ref, atol=5e-4, rtol=2e-3, what=f"d {what} level {l}")
    close(head.cls_reg_shared_conv.weight.grad.reshape(w["conv_w"].shape), w["conv_w"].grad, atol=2e-3, rtol=2e-3, what="d conv_w")
    close(head.cls_reg_shared_conv_norm.weight.grad, w["gn_w"].grad, atol=2e-3, rtol=2e-3, what="d gn_w")
    close(head.bbox_head.fc_reg.weight.grad, w["fc_reg_w"].grad, atol=2e-3, rtol=2e-3, what="d fc_reg_w")


def test_agrpn_train_mode_builds_the_per_class_loss_inputs_and_backpropagates():
    """AGRPNHead.forward_single(train_mode=True) (fgn_ag_rpn_head.py:58-79): per-(image, class) GT lists, the loss
    called on the un-selected [B*N,...] maps, the 1/N balancer, and gradients reaching the query / support maps through
    the attention adjoints."""
    from fgn_b200 import AGRPNHead
    g = torch.Generator().manual_seed(77)
    B, N, K, C = 2, 3, 2, 32
    seen = {}

    def loss_fn(cls_scores, bbox_preds, gt_bboxes, img_metas):
        seen["gts"], seen["metas"] = gt_bboxes, img_metas
        return dict(loss_rpn_cls=[cls_scores[0].square().mean()], loss_rpn_bbox=[bbox_preds[0].abs().mean()])

    head = AGRPNHead(in_channels=C, feat_channels=C, n_ways=N, k_shots=K, loss_fn=loss_fn).to(dev()).train()
    q = torch.randn(B, C, 10, 14, generator=g).to(dev()).requires_grad_(True)
    s = torch.randn(B * N * K, C, 6, 6, generator=g).to(dev()).requires_grad_(True)
    boxes = [torch.rand(4, 4, generator=g).to(dev()) * 100, torch.rand(2, 4, generator=g).to(dev()) * 100]
    cats = [torch.tensor([0, 2, 2, 1]).to(dev()), torch.tensor([1, 1]).to(dev())]
    metas = [dict(img_shape=(160, 224, 3)), dict(img_shape=(160, 224, 3))]
    cls, reg, losses = head.forward_single(q, s, qry_bboxes=boxes, qry_cat_ids=cats, img_metas_cpu=metas, train_mode=True)
    assert cls.shape == (B, head.num_anchors, 10, 14) and reg.shape == (B, 4 * head.num_anchors, 10, 14)
    assert [int(t.shape[0]) for t in seen["gts"]] == [1, 1, 2, 0, 2, 0] and len(seen["metas"]) == B * N
    assert torch.equal(seen["gts"][2], boxes[0][[1, 2]])
    # the balancer, and the same loss through torch ops on the oracle's attention
    _, mod = O.agrpn_attention(q.detach().cpu(), s.detach().cpu(), N, K)
    x = torch.relu(torch.nn.functional.conv2d(mod.to(dev()), head.rpn_conv.weight, head.rpn_conv.bias, padding=1))
    want = torch.nn.functional.conv2d(x, head.rpn_cls.weight, head.rpn_cls.bias).square().mean() / N
    assert abs(float(losses["loss_rpn_cls"][0].detach()) - float(want.detach())) <= 1e-5 * max(1.0, abs(float(want.detach())))
    (losses["loss_rpn_cls"][0] + losses["loss_rpn_bbox"][0]).backward()
    assert q.grad is not None and s.grad is not None and float(q.grad.abs().max()) > 0 and float(s.grad.abs().max()) > 0
    assert head.rpn_conv.weight.grad is not None


def test_roi_align_backward_deterministic_option():
    """fgn_roi_align_ml_bwd_det: 64-bit fixed-point accumulation -- the gradients of many heavily overlapping RoIs are
    bit-identical from run to run (the float-atomic default is allowed not to be) and match autograd over the oracle."""
    from fgn_b200 import autograd as A
    from fgn_b200.episodes import synth_rois
    g = torch.Generator().manual_seed(311)
    strides, B, C = [4, 8, 16, 32], 2, 64
    feats = [torch.randn(B, C, 256 // s, 320 // s, generator=g) for s in strides]
    rois = synth_rois(g, 600, 256, 320, B, smin=24.0)                     # large boxes: hundreds of terms per cell
    gout = torch.randn(600, C, 7, 7, generator=g) * 3.0
    fc = [f.clone().requires_grad_(True) for f in feats]
    want, _ = O.single_roi_extractor(fc, rois, strides, 7, 0, True, 56.0, "tv")
    want.backward(gout)
    runs = []
    A.set_deterministic_backward(True)
    try:
        for _ in range(3):
            fd = [f.to(dev()).contiguous(memory_format=torch.channels_last).requires_grad_(True) for f in feats]
            got = A.roi_align_multilevel(fd, rois.to(dev()), [1 / s for s in strides], 7, 0, True)
            got.backward(gout.to(dev()))
            runs.append([f.grad.clone() for f in fd])
    finally:
        A.set_deterministic_backward(None)
    for l in range(4):
        assert torch.equal(runs[0][l], runs[1][l]) and torch.equal(runs[0][l], runs[2][l]), f"level {l} differs between runs"
        if fc[l].grad is not None:
            close(runs[0][l], fc[l].grad, atol=5e-4, rtol=1e-5, what=f"deterministic grad level {l}")
    # the switch follows torch.use_deterministic_algorithms when left at None
    assert A.deterministic_backward() == torch.are_deterministic_algorithms_enabled()
