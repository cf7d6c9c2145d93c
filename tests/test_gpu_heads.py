"""Post-RoI heads on tcgen05: the 3x3 implicit-GEMM convolution, FCNMaskHead's fused
deconv + ReLU + logits tail, the whole FCNMaskHead and the res5 bottleneck, against plain PyTorch fp32 on the CPU
(the operators the reference calls: fgn_roi_head.py:202-233,380; fgn_r50_c4_densecl.py:115-129)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
# fp32 route = 3xTF32 error-compensated passes: 1e-4 + 1e-5|b| on O(1) outputs; tf32 route = one pass (10-bit mantissas),
# the precision cuDNN gives the same convolutions under allow_tf32 -- stated tolerance 2e-2 on O(1) outputs
TOL = {"fp32": (1e-4, 1e-5), "tf32": (2e-2, 1e-2)}


def close(got, want, prec, what):
    atol, rtol = TOL[prec]
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = (got - want).abs()
    bad = err > atol + rtol * want.abs()
    assert not bad.any(), f"{what} [{prec}]: {int(bad.sum())}/{bad.numel()} outside tol, max abs err {float(err.max()):.3e}"


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
@pytest.mark.parametrize("r,cin,cout,h,w", [
    (37, 64, 48, 7, 7),        # two RoIs per 98-row tile, odd RoI count (half-empty last tile)
    (5, 32, 256, 14, 14),      # nine-row boxes of one RoI (126 rows) + a five-row remainder
    (3, 64, 512, 7, 7),        # two column tiles
    (1, 16, 16, 5, 9),         # non-square tile, a single k-block per tap
    (131, 16, 32, 3, 3),       # fourteen RoIs per tile
    (2, 48, 64, 10, 20),       # wide rows: six rows per box, ragged last box
])
def test_conv3x3_against_torch(prec, r, cin, cout, h, w):
    from fgn_b200 import ops
    g = torch.Generator().manual_seed(r * 1000 + cin + cout)
    x = torch.randn(r, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    b = torch.randn(cout, generator=g)
    res = torch.randn(r, cout, h, w, generator=g)
    taps = ops.conv_taps(wt.to(DEV))
    assert taps.shape == (9, cout, cin)
    got = ops.conv3x3(x.to(DEV), taps, precision=prec)
    assert got.is_contiguous(memory_format=torch.channels_last) or min(h, w, cout) == 1
    close(got, F.conv2d(x, wt, padding=1), prec, "conv3x3 plain")
    got = ops.conv3x3(x.to(DEV).contiguous(memory_format=torch.channels_last), taps, b.to(DEV), residual=res.to(DEV), relu=True,
                      precision=prec, w_split=ops.conv_split_weights(taps) if prec == "fp32" else None)
    close(got, torch.relu(F.conv2d(x, wt, b, padding=1) + res), prec, "conv3x3 + bias + residual + relu")


def test_conv3x3_zero_rois_and_unsupported_shapes():
    from fgn_b200 import ops, FgnError
    taps = torch.zeros(9, 32, 32, device=DEV)
    assert ops.conv3x3(torch.zeros(0, 32, 7, 7, device=DEV), taps).shape == (0, 32, 7, 7)
    with pytest.raises(FgnError):
        ops.conv3x3(torch.zeros(2, 24, 7, 7, device=DEV), torch.zeros(9, 32, 24, device=DEV))      # Cin % 16
    with pytest.raises(FgnError):
        ops.conv3x3(torch.zeros(2, 32, 7, 7), taps.cpu())                                           # no CPU path


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
@pytest.mark.parametrize("r,cin,cout,ncls,h,w", [(7, 64, 256, 1, 14, 14), (3, 32, 48, 3, 7, 7), (1, 16, 16, 4, 2, 3)])
def test_deconv2x2_logits_against_torch(prec, r, cin, cout, ncls, h, w):
    from fgn_b200 import ops
    g = torch.Generator().manual_seed(77 + r + cin)
    x = torch.randn(r, cin, h, w, generator=g)
    up = torch.nn.ConvTranspose2d(cin, cout, 2, stride=2)
    lg = torch.nn.Conv2d(cout, ncls, 1)
    with torch.no_grad():
        up.weight.copy_(torch.randn(up.weight.shape, generator=g) / cin ** 0.5)
        up.bias.copy_(torch.randn(cout, generator=g) * 0.3)
        lg.weight.copy_(torch.randn(lg.weight.shape, generator=g) / cout ** 0.5)
        lg.bias.copy_(torch.randn(ncls, generator=g))
        want = lg(torch.relu(up(x)))
    got = ops.deconv2x2_logits(x.to(DEV), ops.conv_taps(up.weight.detach().to(DEV), transposed=True), up.bias.detach().to(DEV),
                               lg.weight.detach().to(DEV), lg.bias.detach().to(DEV), precision=prec)
    assert got.shape == (r, ncls, 2 * h, 2 * w) and got.is_contiguous()
    close(got, want, prec, "deconv2x2 + relu + logits")


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_fcn_mask_head_fused_route_against_torch_modules(prec):
    """The config's head (num_convs=4, conv_out_channels=256, class_agnostic) at a reduced input width: five library
    launches, state-dict names of mmdet's FCNMaskHead, values against the same modules run by PyTorch on the CPU."""
    from fgn_b200 import FCNMaskHead, _lib
    torch.manual_seed(5)
    head = FCNMaskHead(num_convs=4, in_channels=64, conv_out_channels=256, num_classes=1, class_agnostic=True, precision=prec)
    keys = set(head.state_dict().keys())
    assert {"convs.0.conv.weight", "convs.3.conv.bias", "upsample.weight", "upsample.bias", "conv_logits.weight",
            "conv_logits.bias"} <= keys
    with torch.no_grad():
        for m in head.convs:
            m.conv.bias.copy_(torch.randn_like(m.conv.bias) * 0.1)
        head.upsample.bias.copy_(torch.randn_like(head.upsample.bias) * 0.1)
    head.eval()
    x = torch.randn(9, 64, 14, 14)
    with torch.no_grad():
        want = head(x)                                                   # CPU tensors -> the torch modules
        hd = head.to(DEV)
        xd = x.to(DEV).contiguous(memory_format=torch.channels_last)
        hd(xd)                                                           # first call prepares the tap-major / split weights
        before = _lib.load().fgn_launch_count()
        got = hd(xd)
        launches = _lib.load().fgn_launch_count() - before
    assert want.shape == (9, 1, 28, 28)
    assert launches == 5, launches                                       # four convolutions + the fused tail, nothing per call
    close(got, want, prec, "FCNMaskHead")
    # training / autograd: the torch modules, with a grad_fn
    hd.train()
    out = hd(x.to(DEV).requires_grad_(True))
    assert out.grad_fn is not None


@pytest.mark.parametrize("prec_tf32", [False, True])
def test_bottleneck_three_convs_on_tcgen05(prec_tf32):
    from fgn_b200 import _lib
    from fgn_b200.roi_head import _Bottleneck
    g = torch.Generator().manual_seed(909)
    blk = _Bottleneck(256, 128)
    with torch.no_grad():
        for bn in (blk.bn1, blk.bn2, blk.bn3):
            bn.running_mean.copy_(torch.randn(bn.num_features, generator=g) * 0.2)
            bn.running_var.copy_(torch.rand(bn.num_features, generator=g) + 0.5)
            bn.weight.copy_(1 + 0.2 * torch.randn(bn.num_features, generator=g))
            bn.bias.copy_(0.1 * torch.randn(bn.num_features, generator=g))
    blk.eval()
    blk.tc_3x3 = True                                                    # ("auto" keeps the 3x3 on cuDNN under strict fp32)
    x = torch.randn(61, 256, 7, 7, generator=g)
    old = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = prec_tf32
        with torch.no_grad():
            want = blk(x)                                                # CPU: plain torch modules in fp32
            bd = blk.to(DEV)
            before = _lib.load().fgn_launch_count()
            got = bd(x.to(DEV))
            assert _lib.load().fgn_launch_count() - before >= 3
    finally:
        torch.backends.cudnn.allow_tf32 = old
    prec = "tf32" if prec_tf32 else "fp32"
    atol = 5e-2 if prec_tf32 else 5e-4                                   # three chained convolutions of O(1..10) activations
    err = (got.cpu() - want).abs()
    assert bool((err <= atol + (1e-2 if prec_tf32 else 1e-4) * want.abs()).all()), f"bottleneck [{prec}] max abs err {float(err.max()):.3e}"


@pytest.mark.parametrize("name", ["tiny_fpn", "tiny_c4", "tiny_fpn_big_support", "cfg3_coco2voc_n1k1_fpn", "cfg4_coco2voc_n20k5_fpn"])
def test_support_prologue_one_launch_against_the_separate_ops_and_the_oracle(name):
    """count_spp as ONE launch (fgn_support_prologue_fwd) against the four separate library ops it replaces and, for the
    class maps / vectors, against the oracle's count_spp (fgn_roi_head.py:419-449); its class term against the exact
    fp32 expression cat_mean Ws^T + b."""
    from fgn_b200 import _lib
    from fgn_b200.episodes import CONFIGS, build_heads, episode_to_device, make_episode
    from oracle import fgn_oracle as O
    dev = torch.device(DEV)
    if name == "tiny_fpn_big_support":
        # a 640-px support: the mask bin's adaptive grid is 74 x 74 samples, past the 64 staged per axis (on-the-fly path)
        import dataclasses
        cfg = dataclasses.replace(CONFIGS["tiny_fpn"], name=name, spp_size=640, n_ways=2, k_shots=1)
    else:
        cfg = CONFIGS[name]
    ep = make_episode(cfg, seed=3)
    _, head = build_heads(cfg, dev, seed=0, shared_head=None)
    epd = episode_to_device(ep, dev)
    n_ext = len(cfg.strides)
    spp = epd["spp"][:n_ext] if cfg.mode == "fpn" else epd["spp"][0]
    with torch.no_grad():
        head.fused_prologue = False
        head.count_spp(spp, epd["spp_bboxes"].clone(), epd["spp_masks"])
        cat0, mp0 = head.spp_fmaps_roi_aligned_cat_mean, head.spp_fvecs_roi_aligned_cat_mean_mp
        assert head._class_term is None
        head.fused_prologue = "always"                                   # (cfg4's hundred supports are past the size rule)
        head.relation_params()                                           # (packs / splits the weights once: not part of count_spp)
        before = _lib.load().fgn_launch_count()
        head.count_spp(spp, epd["spp_bboxes"].clone(), epd["spp_masks"])
        assert _lib.load().fgn_launch_count() - before == 1
        cat1, mp1 = head.spp_fmaps_roi_aligned_cat_mean, head.spp_fvecs_roi_aligned_cat_mean_mp
        term = head._valid_class_term(head.relation_params())
    assert term is not None and cat1.shape == cat0.shape and mp1.shape == mp0.shape
    close(cat1, cat0, "fp32", "cat_mean fused vs separate")              # (same sampling order: differences are the mean's only)
    close(mp1, mp0, "fp32", "masked_gap fused vs separate")
    if cfg.mode == "fpn":
        want_cat, want_mp, _, _ = O.count_spp_fpn(ep["spp"][:n_ext], cfg.strides, ep["spp_bboxes"].clone(), ep["spp_masks"],
                                                  cfg.n_ways, cfg.k_shots)
        close(cat1, want_cat.view(cat1.shape), "fp32", "cat_mean vs oracle")
        close(mp1, want_mp.view(mp1.shape), "fp32", "masked_gap vs oracle")
    c = cfg.channels
    w = head.cls_reg_shared_conv.weight.detach().reshape(c, 2 * c).double().cpu()
    rows = cat1.permute(0, 1, 3, 4, 2).reshape(-1, c).double().cpu()     # [B*N*49, C] in bin order
    want_term = rows @ w[:, c:].t() + head.cls_reg_shared_conv.bias.detach().double().cpu()
    close(term, want_term.float(), "fp32", "class term")
    # the box head gives the same logits with and without the precomputed class term
    with torch.no_grad():
        qry = epd["qry"][:n_ext] if cfg.mode == "fpn" else epd["qry"][0]
        a = head._bbox_forward(qry, epd["rois"])
        head._class_term = None
        b = head._bbox_forward(qry, epd["rois"])
    close(a["cls_score"], b["cls_score"], "fp32", "cls_score with / without precomputed class term")
    close(a["bbox_pred"], b["bbox_pred"], "fp32", "bbox_pred with / without precomputed class term")


def _guarded(numel, device, pad=4096):
    """fp32 buffer of `numel` elements between two canary regions (compute-sanitizer is not available on the GPU pool)"""
    buf = torch.full((numel + 2 * pad,), 777.25, device=device, dtype=torch.float32)
    return buf, buf[pad:pad + numel], pad


def _canaries_intact(buf, numel, pad):
    return bool((buf[:pad] == 777.25).all()) and bool((buf[pad + numel:] == 777.25).all())


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("r,cin,cout,h,w", [(37, 64, 64, 7, 7), (5, 32, 256, 14, 14), (7, 64, 512, 7, 7), (9, 16, 32, 3, 3),
                                            (3, 48, 64, 10, 20), (1, 16, 48, 7, 7)])
def test_conv_kernels_write_nothing_outside_their_output(prec, r, cin, cout, h, w):
    """Raw C-ABI calls with the output between canaries: the single-CTA, CTA-pair and two-row-tile kernels pad their last
    tiles (pairs / quads of row tiles, 98- and 126-row boxes) -- none of the padding may reach memory."""
    import ctypes
    from fgn_b200 import _lib, ops
    lib = _lib.load()
    dev = torch.device(DEV)
    g = torch.Generator().manual_seed(r + cin)
    x = torch.randn(r, h, w, cin, generator=g).to(dev)
    taps = (torch.randn(9, cout, cin, generator=g) / (3 * cin ** 0.5)).to(dev)
    n = r * h * w * cout
    buf, out, pad = _guarded(n, dev)
    wsb = int(lib.fgn_conv_split_weights_bytes(9, cout, cin))
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.fgn_conv3x3_nhwc(x.data_ptr(), taps.data_ptr(), None, None, None, 0, out.data_ptr(), r, h, w, cin, cout, prec,
                                    ws.data_ptr(), wsb, st), "fgn_conv3x3_nhwc")
    torch.cuda.synchronize()
    assert _canaries_intact(buf, n, pad)
    want = F.conv2d(x.permute(0, 3, 1, 2).cpu(), taps.cpu().view(3, 3, cout, cin).permute(2, 3, 0, 1), padding=1)
    close(out.view(r, h, w, cout).permute(0, 3, 1, 2), want, "fp32" if prec == 0 else "tf32", "conv3x3 through the raw ABI")
    # deconv tail on the same activations
    ncls = 2
    up = (torch.randn(4, 32, cin, generator=g) / cin ** 0.5).to(dev)
    wl, bl, bd = torch.randn(ncls, 32, generator=g).to(dev) / 6, torch.randn(ncls, generator=g).to(dev), torch.randn(32, generator=g).to(dev)
    m = r * ncls * 4 * h * w
    buf2, out2, pad2 = _guarded(m, dev)
    wsb2 = int(lib.fgn_conv_split_weights_bytes(4, 32, cin))
    ws2 = torch.empty(wsb2, device=dev, dtype=torch.uint8)
    _lib.check(lib.fgn_deconv2x2_logits_nhwc(x.data_ptr(), up.data_ptr(), None, bd.data_ptr(), wl.data_ptr(), bl.data_ptr(), out2.data_ptr(),
                                             r, h, w, cin, 32, ncls, prec, ws2.data_ptr(), wsb2, st), "fgn_deconv2x2_logits_nhwc")
    torch.cuda.synchronize()
    assert _canaries_intact(buf2, m, pad2)
    assert bool(torch.isfinite(out2).all())


@pytest.mark.parametrize("name", ["tiny_fpn", "tiny_c4"])
def test_image_batched_call_equals_one_call_per_episode_bit_for_bit(name):
    """The reference runs B episodes per call (main.py:492-499); episodes.batch_episodes builds that call.  Per-image results
    must not depend on the batching: every output row of the batched call equals the single-episode call's, bitwise."""
    from fgn_b200.episodes import CONFIGS, batch_episodes, build_heads, episode_to_device, make_episode, run_guided_path
    import dataclasses
    dev = torch.device(DEV)
    # (24 mask RoIs per image: launches of <= 16 RoIs take the reference-order small-launch kernel, whose summation order
    #  differs from the window kernel's in the last bits -- the same kernel must serve both sides of a bitwise comparison)
    cfg = dataclasses.replace(CONFIGS[name], mask_rois=24)
    eps = [episode_to_device(make_episode(cfg, seed=40 + i), dev) for i in range(3)]
    rpn, head = build_heads(cfg, dev, seed=0, shared_head=None)
    with torch.no_grad():
        singles = []
        for ep in eps:
            o = run_guided_path(rpn, head, ep)
            singles.append({k: (v.clone() if torch.is_tensor(v) else [t.clone() for t in v]) for k, v in o.items()
                            if k in ("cls_score", "bbox_pred", "mask_feats", "qry_fmap_mod")})
        ob = run_guided_path(rpn, head, batch_episodes(eps))
    r, d, n = cfg.num_rois, cfg.mask_rois, cfg.n_ways
    for i, s in enumerate(singles):
        assert torch.equal(ob["cls_score"][i * r:(i + 1) * r], s["cls_score"]), f"cls_score of image {i}"
        assert torch.equal(ob["bbox_pred"][i * r:(i + 1) * r], s["bbox_pred"]), f"bbox_pred of image {i}"
        assert torch.equal(ob["mask_feats"][i * d:(i + 1) * d], s["mask_feats"]), f"mask_feats of image {i}"
        qb = ob["qry_fmap_mod"] if isinstance(ob["qry_fmap_mod"], (list, tuple)) else [ob["qry_fmap_mod"]]
        qs = s["qry_fmap_mod"] if isinstance(s["qry_fmap_mod"], (list, tuple)) else [s["qry_fmap_mod"]]
        for lb, ls in zip(qb, qs):
            assert torch.equal(lb[i * n:(i + 1) * n], ls), f"attended maps of image {i}"
