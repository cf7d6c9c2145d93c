"""Pin the oracle (CPU) against the committed golden fixtures -- runs without a GPU.

fgn_reference_*.npz hold outputs of the reference's own unmodified methods (tests/golden/make_golden.py);
roi_align_kat.npz / map_roi_levels_kat.npz hold torchvision / torch known-answer vectors.
"""
import glob
import os

import math

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fgn_oracle as O

REF_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "fgn_reference_*.npz")))


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _weights(z):
    c = int(z["C"])
    return dict(conv_w=_t(z["w.cls_reg_shared_conv.weight"]).view(c, 2 * c), conv_b=_t(z["w.cls_reg_shared_conv.bias"]),
                gn_w=_t(z["w.cls_reg_shared_conv_norm.weight"]), gn_b=_t(z["w.cls_reg_shared_conv_norm.bias"]),
                fc_cls_w=_t(z["w.bbox_head.fc_cls.weight"]), fc_cls_b=_t(z["w.bbox_head.fc_cls.bias"]),
                fc_reg_w=_t(z["w.bbox_head.fc_reg.weight"]), fc_reg_b=_t(z["w.bbox_head.fc_reg.bias"]))


def _shared_head(z):
    if "w.shared_head.0.weight" not in z.files:
        return None
    w, b = _t(z["w.shared_head.0.weight"]), _t(z["w.shared_head.0.bias"])
    return lambda x: torch.relu(torch.nn.functional.conv2d(x, w, b, padding=1))


def test_fixtures_exist():
    assert len(REF_FIXTURES) >= 4
    for f in ("roi_align_kat.npz", "map_roi_levels_kat.npz"):
        assert os.path.exists(os.path.join(GOLDEN, f))


# ---- [3P] arithmetic: C restatement vs torchvision known answers ---------------------------------
def test_c_roi_align_is_bit_exact_against_torchvision_kats():
    z = np.load(os.path.join(GOLDEN, "roi_align_kat.npz"))
    names = sorted({k.split(".")[0] for k in z.files})
    assert len(names) >= 6
    for n in names:
        out = O.roi_align_c(_t(z[f"{n}.feat"]), _t(z[f"{n}.rois"]), float(z[f"{n}.scale"]), int(z[f"{n}.P"]),
                            int(z[f"{n}.sr"]), bool(z[f"{n}.aligned"]))
        assert np.array_equal(out.numpy(), z[f"{n}.out"]), n           # bit-exact, incl. degenerate RoIs


def test_c_roi_align_matches_live_torchvision_op():
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(2, 6, 40, 56, generator=g)
    from fgn_b200.episodes import synth_rois
    rois = synth_rois(g, 128, 640, 896, 2, smin=4.0)
    for aligned, sr in ((True, 0), (False, -1), (True, 3)):
        a = O.roi_align_c(feat, rois, 1 / 16, 7, sr, aligned)
        b = O.roi_align_tv(feat, rois, 1 / 16, 7, sr, aligned)
        assert torch.equal(a, b)


def test_map_roi_levels_c_matches_torch_expression_on_boundaries():
    z = np.load(os.path.join(GOLDEN, "map_roi_levels_kat.npz"))
    rois = _t(z["rois"])
    for L in (1, 2, 4, 5):
        want = z[f"L{L}"]
        assert np.array_equal(O.map_roi_levels(rois, L).numpy(), want)         # torch expression, live
        got = O.map_roi_levels_c(rois, L).numpy()
        # NaN scale (x2<x1 xor y2<y1): torch's NaN->long cast is platform-defined; contract = level 0
        area = (z["rois"][:, 3] - z["rois"][:, 1]) * (z["rois"][:, 4] - z["rois"][:, 2])
        ok = ~(area < 0)
        assert np.array_equal(got[ok], want[ok])
        assert (got[~ok] == 0).all()


def test_log2_contract_next_to_powers_of_two():
    # the C/CUDA contract (float)log2((double)v) must agree with torch.log2 on fp32 at every value
    # within 8 ulp of 2^k, k = -2..6 (this is where floor() can flip)
    vals = []
    for k in range(-2, 7):
        v = np.float32(2.0 ** k)
        lo = v
        for _ in range(8):
            lo = np.nextafter(lo, np.float32(0), dtype=np.float32)
        x = lo
        for _ in range(17):
            vals.append(x)
            x = np.nextafter(x, np.float32(np.inf), dtype=np.float32)
    v = np.asarray(vals, np.float32)
    torch_l2 = torch.log2(torch.from_numpy(v)).numpy()
    mine = np.log2(v.astype(np.float64)).astype(np.float32)
    assert np.array_equal(np.floor(torch_l2), np.floor(mine))


def test_index_tables_are_consistent_with_values():
    # rebuild RoIAlign from the exported per-axis (valid, low, high) tables with float64 weights and
    # compare with the op: checks the tables really are the indices the values were gathered from
    g = torch.Generator().manual_seed(9)
    H, W, P, scale = 20, 28, 7, 1 / 8
    feat = torch.randn(1, 2, H, W, generator=g)
    from fgn_b200.episodes import synth_rois
    rois = synth_rois(g, 24, H / scale, W / scale, 1, smin=4.0)
    rois[0, 1:] = torch.tensor([-40., -30., 60., 50.])
    lv = torch.zeros(24, dtype=torch.long)
    grid, ytab, xtab = O.roi_align_indices_c(rois, lv, [(H, W)], [scale], P, 0, True, 16)
    ref = O.roi_align_tv(feat, rois, scale, P, 0, True).numpy()
    f = feat.numpy().astype(np.float64)
    for r in range(rois.shape[0]):
        gh, gw = grid[r].tolist()
        assert gh <= 16 and gw <= 16
        x1, y1, x2, y2 = (rois[r, 1:].numpy().astype(np.float64) * scale - 0.5)
        bh, bw = (y2 - y1) / P, (x2 - x1) / P
        for ph in range(P):
            for pw in range(P):
                acc = 0.0
                for iy in range(gh):
                    vy, yl, yh = ytab[r, ph, iy].tolist()
                    y = min(max(y1 + ph * bh + (iy + .5) * bh / gh, 0.0), H - 1)
                    for ix in range(gw):
                        vx, xl, xh = xtab[r, pw, ix].tolist()
                        if not (vy and vx):
                            continue
                        x = min(max(x1 + pw * bw + (ix + .5) * bw / gw, 0.0), W - 1)
                        ly, lx = y - yl, x - xl
                        acc += ((1 - ly) * (1 - lx) * f[0, :, yl, xl] + (1 - ly) * lx * f[0, :, yl, xh]
                                + ly * (1 - lx) * f[0, :, yh, xl] + ly * lx * f[0, :, yh, xh])
                acc = acc / max(gh * gw, 1)
                assert np.allclose(acc, ref[r, :, ph, pw], atol=2e-5), (r, ph, pw)


# ---- FGN-owned arithmetic: restatement vs the reference's own code -------------------------------
@pytest.mark.parametrize("path", REF_FIXTURES, ids=[os.path.basename(p)[14:-4] for p in REF_FIXTURES])
def test_oracle_matches_reference_fixture(path):
    z = np.load(path)
    B, N, K, C, stride = (int(z[k]) for k in ("B", "N", "K", "C", "stride"))
    sh = _shared_head(z)
    qry, spp, rois = _t(z["qry"]), _t(z["spp"]), _t(z["rois"])
    boxes = _t(z["spp_bboxes"].copy())
    for impl in ("tv", "c"):
        b = boxes.clone()
        cat_mean, mp, _, _ = O.count_spp(spp, b, _t(z["spp_masks"]), N, K, stride, sh, impl)
        assert torch.equal(b, _t(z["spp_bboxes_after"]))                      # in-place /= 16 (:430)
        assert torch.equal(cat_mean, _t(z["cat_mean"])), impl
        assert torch.equal(mp, _t(z["masked_gap"])), impl
    vec, mod = O.agrpn_attention(qry, spp, N, K)
    assert torch.equal(mod, _t(z["rpn_qry_fmap_mod"]))
    cls, reg = O.best_class_selection(_t(z["rpn_cls_raw"]), _t(z["rpn_reg_raw"]), B, N)
    assert torch.equal(cls, _t(z["rpn_cls"])) and torch.equal(reg, _t(z["rpn_reg"]))
    if "cls_score" not in z.files:
        return
    w = _weights(z)
    res = O.bbox_forward([qry], [stride], rois, cat_mean, N, w, sh, "tv")
    assert torch.equal(res["bbox_feats"], _t(z["bbox_feats"]))
    assert torch.equal(res["cls_score"], _t(z["cls_score"]))
    assert torch.equal(res["bbox_pred"], _t(z["bbox_pred"]))
    # chunked evaluation (used for the N=20 stress config) only changes GEMM blocking
    res_c = O.bbox_forward([qry], [stride], rois, cat_mean, N, w, sh, "tv", chunk=7)
    assert torch.allclose(res_c["cls_score"], _t(z["cls_score"]), atol=1e-5, rtol=1e-5)
    det_rois, labels = _t(z["det_rois"]), _t(z["det_labels"])
    det_labels = [labels[det_rois[:, 0] == b] for b in range(B)]
    mf = O.mask_attention([qry], [stride], det_rois, mp, det_labels, N, 7, sh, "tv")
    assert torch.equal(mf, _t(z["mask_feats"]))


def test_count_modified_cls_bbox_generalisation():
    g = torch.Generator().manual_seed(3)
    R = 50
    cls, reg = torch.randn(R * 3, 2, generator=g), torch.randn(R * 3, 4, generator=g)
    c, r = O.count_modified_cls_bbox(R, cls, reg, 3)
    t = cls.view(R, 6)
    top = t[:, [1, 3, 5]].argmax(-1) * 2                                      # reference's literal form
    assert torch.equal(c, torch.cat((t[:, [1, 3, 5]], t[torch.arange(R), top].view(R, 1)), 1))
    assert torch.equal(r, reg.view(R, 12))
    c1, r1 = O.count_modified_cls_bbox(R, cls[:R], reg[:R], 1)
    assert torch.equal(c1, cls[:R][:, [1, 0]]) and torch.equal(r1, reg[:R])


def test_split_weight_identity():
    # conv1x1(cat(q,s)) == Wq q + Ws s + b : the algebra the CUDA path relies on
    g = torch.Generator().manual_seed(4)
    C = 32
    q, s = torch.randn(5, C, 7, 7, generator=g), torch.randn(5, C, 7, 7, generator=g)
    w, b = torch.randn(C, 2 * C, generator=g) / 8, torch.randn(C, generator=g)
    full = torch.nn.functional.conv2d(torch.cat((q, s), 1), w.view(C, 2 * C, 1, 1), b)
    split = torch.einsum("oc,rchw->rohw", w[:, :C], q) + torch.einsum("oc,rchw->rohw", w[:, C:], s) + b.view(1, C, 1, 1)
    assert torch.allclose(full, split, atol=1e-5)


# ---- test-time box post-processing (BBoxHead.get_bboxes [3P]) ---------------------------------------------------
def _random_boxes(g, n, size=200.0):
    c = torch.rand(n, 2, generator=g) * size
    wh = torch.rand(n, 2, generator=g) * 60 + 2
    return torch.cat([c - wh / 2, c + wh / 2], 1)


def test_nms_restatement_matches_torchvision_op():
    """The plain-loop NMS restatement (fp32 IoU, stable descending order) selects exactly what the compiled
    torchvision CPU op selects -- the op mmcv's nms is equivalent to -- including on class-offset boxes."""
    import torchvision
    g = torch.Generator().manual_seed(5)
    for n, thr in ((1, 0.5), (60, 0.5), (300, 0.7), (300, 0.3)):
        boxes, scores = _random_boxes(g, n), torch.rand(n, generator=g)
        scores[::7] = scores[0]                                  # ties: stable order decides
        labels = torch.randint(0, 4, (n,), generator=g)
        shifted = boxes + (labels.float() * (boxes.max() + 1))[:, None]
        for b in (boxes, shifted):
            assert torch.equal(O.nms_greedy(b, scores, thr), torchvision.ops.nms(b, scores, thr))


def test_multiclass_nms_is_class_aware_and_score_ordered():
    import torchvision
    g = torch.Generator().manual_seed(6)
    R, N = 120, 3
    boxes = _random_boxes(g, R * N).view(R, N * 4)
    scores = torch.softmax(torch.randn(R, N + 1, generator=g) * 2, -1)
    dets, labels, flat = O.multiclass_nms(boxes, scores, 0.05, 0.5, 50)
    assert dets.shape[0] == labels.shape[0] == flat.shape[0] <= 50
    assert (dets[:-1, 4] >= dets[1:, 4]).all()
    assert torch.equal(labels, flat % N) and torch.equal(dets[:, :4], boxes.view(R, N, 4).reshape(-1, 4)[flat])
    # same selection as independent per-class NMS
    want = []
    for c in range(N):
        sc = scores[:, c]
        idx = (sc > 0.05).nonzero().squeeze(1)
        keep = idx[torchvision.ops.nms(boxes.view(R, N, 4)[idx, c], sc[idx], 0.5)]
        want += [(float(sc[i]), int(i) * N + c) for i in keep]
    want.sort(key=lambda t: (-t[0], t[1]))
    assert [w[1] for w in want[:50]] == flat.tolist()
    # the loop restatement and the torchvision-backed path agree
    d2, l2, f2 = O.multiclass_nms(boxes, scores, 0.05, 0.5, 50, nms_impl="loop")
    assert torch.equal(f2, flat) and torch.equal(d2, dets)


def test_delta2bbox_known_values():
    rois = torch.tensor([[10., 20., 50., 100.]])
    z = O.delta2bbox(rois, torch.zeros(1, 8), (0, 0, 0, 0), (0.1, 0.1, 0.2, 0.2), max_shape=(90, 45))
    assert torch.equal(z, torch.tensor([[10., 20., 45., 90., 10., 20., 45., 90.]]))      # identity, then clipped
    d = torch.tensor([[1.0, -1.0, math.log(2.0) / 0.2, 100.0]])                         # dw = log 2, dh clamped
    out = O.delta2bbox(rois, d, (0, 0, 0, 0), (0.1, 0.1, 0.2, 0.2))
    cx, cy, w, h = 30 + 40 * 0.1, 60 - 80 * 0.1, 80.0, 80 * 1000 / 16
    assert torch.allclose(out, torch.tensor([[cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2]]), rtol=1e-5)


def test_anchor_grid_matches_mmdet_docstring_example():
    """Known answer published in mmdet's AnchorGenerator docstring [3P]:
    AnchorGenerator([16], [1.], [1.], [9]).grid_anchors([(2, 2)]) ->
    [[-4.5,-4.5,4.5,4.5],[11.5,-4.5,20.5,4.5],[-4.5,11.5,4.5,20.5],[11.5,11.5,20.5,20.5]]."""
    base = O.anchor_base(9, [1.0], [1.0])
    got = O.anchor_grid(base, 2, 2, 16)
    want = torch.tensor([[-4.5, -4.5, 4.5, 4.5], [11.5, -4.5, 20.5, 4.5], [-4.5, 11.5, 4.5, 20.5], [11.5, 11.5, 20.5, 20.5]])
    assert torch.equal(got, want)
    # ratio-major / scale-minor ordering of the base anchors, h/w = ratio
    b = O.anchor_base(16, [2, 4], [0.5, 2.0])
    wh = torch.stack([b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], 1)
    assert torch.allclose(wh[:, 1] / wh[:, 0], torch.tensor([0.5, 0.5, 2.0, 2.0]))
    assert torch.allclose((wh[:, 0] * wh[:, 1]).sqrt(), torch.tensor([32., 64., 32., 64.]))


def test_rpn_restatement_small_case():
    """_get_bboxes_single restatement on a hand-checkable case: zero deltas return the (clipped) anchors, NMS keeps
    the best of coincident anchors, results are score-ordered and capped."""
    base = O.anchor_base(16, [2], [1.0])                       # one 32x32 anchor per cell
    cls = torch.tensor([[[2.0, 1.0], [0.5, -1.0]]])            # [A=1, H=2, W=2]
    reg = torch.zeros(4, 2, 2)
    dets, ids = O.rpn_get_bboxes_single([cls], [reg], [O.anchor_grid(base, 2, 2, 16)], (32, 32, 3), 6000, 0.7, 3, 0.0)
    assert dets.shape == (3, 5) and torch.equal(ids, torch.zeros(3, dtype=torch.long))
    assert torch.allclose(dets[:, 4], torch.tensor([2.0, 1.0, 0.5]).sigmoid())
    assert torch.equal(dets[0, :4], torch.tensor([0., 0., 16., 16.]))          # anchor (-16,-16,16,16) clipped


# ---- mask pasting + RLE (get_seg_masks / encode_mask_results restatements) ---------------------------------
_PASTE_BOXES = np.array([[10.3, 5.2, 50.7, 40.1], [0, 0, 83, 61], [-5.5, -3.2, 20.1, 70.3], [30, 30, 30, 45.5],
                         [70.2, 50.1, 82.9, 60.9], [40.5, 10.5, 12.5, 33.0], [100, 100, 120, 130]], np.float32)


def _torch_paste(logits, boxes, h, w):
    """mmdet _do_paste_mask(skip_empty=False) [3P] written with the very torch calls it makes (CPU)."""
    import torch.nn.functional as F
    mt, bt = torch.from_numpy(logits).sigmoid(), torch.from_numpy(boxes)
    d = mt.shape[0]
    x0, y0, x1, y1 = torch.split(bt, 1, dim=1)
    iy = (torch.arange(0, h).float() + 0.5 - y0) / (y1 - y0) * 2 - 1
    ix = (torch.arange(0, w).float() + 0.5 - x0) / (x1 - x0) * 2 - 1
    ix[torch.isinf(ix)] = 0
    iy[torch.isinf(iy)] = 0
    gx = ix[:, None, :].expand(d, h, w)
    gy = iy[:, :, None].expand(d, h, w)
    return F.grid_sample(mt, torch.stack([gx, gy], 3), align_corners=False)[:, 0].numpy()


def test_paste_restatement_matches_torch_grid_sample():
    rng = np.random.default_rng(7)
    h, w, m = 61, 83, 28
    logits = rng.normal(0, 3, (len(_PASTE_BOXES), 1, m, m)).astype(np.float32)
    got = O.paste_values(logits, _PASTE_BOXES, h, w)
    want = _torch_paste(logits, _PASTE_BOXES, h, w)
    assert np.nanmax(np.abs(got - want)) < 5e-6
    differ = (got >= 0.5) != (want >= 0.5)
    assert differ.sum() == 0 or np.abs(want[differ] - 0.5).max() < 5e-6


def test_rle_restatement_round_trip_and_known_runs():
    # known runs, column-major: a 3x4 mask whose columns are 010 / 111 / 000 / 001
    mk = np.array([[0, 1, 0, 0], [1, 1, 0, 0], [0, 1, 0, 1]], dtype=bool)
    assert O.rle_counts(mk) == [1, 1, 1, 3, 5, 1]
    assert O.rle_counts(np.ones((2, 2), bool)) == [0, 4]
    assert O.rle_counts(np.zeros((2, 3), bool)) == [6]
    # string form: small values are one character (value + 48); from the fourth on, the difference to two back
    assert O.rle_to_string([1, 1, 1, 3, 5, 1]) == bytes([49, 49, 49, 50, 52, 48 + (-2 & 0x1F)])
    rng = np.random.default_rng(3)
    for h, w, p in [(1, 1, 0.5), (7, 5, 0.5), (64, 48, 0.02), (33, 200, 0.9), (480, 640, 0.001)]:
        mk = rng.random((h, w)) < p
        c = O.rle_counts(mk)
        assert sum(c) == h * w
        s = O.rle_to_string(c)
        assert O.rle_from_string(s) == c
        assert (O.rle_decode(c, h, w) == mk).all()
    big = [0, 1066400] + [3, 70000, 5, 2] * 3                   # multi-character values, negative differences
    assert O.rle_from_string(O.rle_to_string(big)) == big


def _paste_golden():
    z = np.load(os.path.join(GOLDEN, "mask_paste_kat.npz"))
    lens = z["counts_len"].tolist()
    flat = z["counts_flat"].tolist()
    counts, k = [], 0
    for n in lens:
        counts.append(flat[k:k + n])
        k += n
    return z, counts


def test_paste_and_rle_restatements_reproduce_the_golden_fixture():
    """tests/golden/mask_paste_kat.npz (minted with torch's own sigmoid + F.grid_sample, make_golden_paste.py): the
    restatement's masks and run lengths are the fixture's."""
    z, counts = _paste_golden()
    h, w = [int(v) for v in z["img_hw"]]
    got = O.paste_values(z["logits"], z["boxes"], h, w)
    assert np.abs(got - z["values"]).max() < 5e-6
    masks = got >= np.float32(z["thr"])
    assert np.array_equal(masks, z["masks"])
    for m, c in zip(masks, counts):
        assert O.rle_counts(m) == c
        assert O.rle_from_string(O.rle_to_string(c)) == c
