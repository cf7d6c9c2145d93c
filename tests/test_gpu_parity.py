"""Parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Bars (BASELINE.json north_star): bit-exact for RoI level assignment and RoIAlign sample indices;
fp32 features/logits within |a-b| <= 1e-4 + 1e-5*|b| of the reference.
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import ATOL, GOLDEN, RTOL
from oracle import fgn_oracle as O

pytestmark = pytest.mark.gpu

REF_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "fgn_reference_*.npz")))


def dev():
    return torch.device("cuda:0")


def _t(a):
    return torch.from_numpy(np.asarray(a))


def close(got, want, atol=ATOL, rtol=RTOL, what=""):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = (got - want).abs()
    bound = atol + rtol * want.abs()
    bad = err > bound
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} outside tol, max abs err {float(err.max()):.3e}"


@pytest.fixture(scope="module", autouse=True)
def _exact_convs():
    # the adjacent torch/cuDNN modules (RPN convs, shared_head) must not run in TF32 during parity
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_library_is_native_and_on_blackwell():
    import ctypes
    from fgn_b200 import _lib
    lib = _lib.load()
    sm, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.fgn_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)) == 0
    assert major.value == 10, "libfgn_b200 is built for sm_100a only"
    assert sm.value > 0


# ---- a4: map_roi_levels, bit-exact --------------------------------------------------------------
def test_map_roi_levels_bit_exact_on_boundaries():
    from fgn_b200 import ops
    z = np.load(os.path.join(GOLDEN, "map_roi_levels_kat.npz"))
    rois = _t(z["rois"])
    area = (z["rois"][:, 3] - z["rois"][:, 1]) * (z["rois"][:, 4] - z["rois"][:, 2])
    ok = ~(area < 0)
    for L in (1, 2, 4, 5):
        got = ops.map_roi_levels(rois.to(dev()), L).cpu().numpy()
        assert got.dtype == np.int64
        assert np.array_equal(got[ok], z[f"L{L}"][ok])
        assert np.array_equal(got, O.map_roi_levels_c(rois, L).numpy())
        # torch's own CUDA evaluation of the mmdet expression: libdevice log2f is not correctly rounded,
        # so it may disagree with torch-CPU (our contract) ONLY on the crafted rows within a few ulp of a
        # level boundary -- never on ordinary proposals.  The count is printed for DESIGN.md.
        r = rois.to(dev())
        scale = torch.sqrt((r[:, 3] - r[:, 1]) * (r[:, 4] - r[:, 2]))
        tc = torch.floor(torch.log2(scale / 56 + 1e-6)).clamp(min=0, max=L - 1).long().cpu().numpy()
        diff = np.nonzero((got != tc) & ok)[0]
        n_crafted = z["rois"].shape[0] - 4000
        print(f"L={L}: torch-CUDA vs torch-CPU level disagreements on boundary rows: {len(diff)} of {n_crafted}")
        assert (diff < n_crafted).all()


# ---- a5: RoIAlign sample indices, bit-exact -----------------------------------------------------
@pytest.mark.parametrize("aligned,sr,P", [(True, 0, 7), (False, -1, 7), (True, 0, 14), (True, 2, 7)])
def test_sample_indices_bit_exact(aligned, sr, P):
    from fgn_b200 import ops
    from fgn_b200.episodes import synth_rois
    g = torch.Generator().manual_seed(11)
    hw = [(200, 336), (100, 168), (50, 84), (25, 42)]
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    rois = synth_rois(g, 1500, 800, 1344, 1, smin=2.0)
    z = np.load(os.path.join(GOLDEN, "map_roi_levels_kat.npz"))
    edge = torch.tensor([[0, 5., 5., 5., 5.], [0, 10., 10., 4., 4.], [0, -500., -500., -400., -400.],
                         [0, -50., -50., 1400., 860.], [0, 0., 0., 1344., 800.], [0, 1343., 799., 1344., 800.],
                         [0, -4., -4., 0., 0.], [0, 1344., 800., 1360., 816.], [0, 0., 0., 3000., 10.]])
    rois = torch.cat([edge, _t(z["rois"][:200]), rois], 0)
    ok = ~(((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2])) < 0)
    rois = rois[ok]
    G = 64
    lvl, grid, ytab, xtab = ops.roi_align_sample_indices(hw, rois.to(dev()), scales, P, sr, aligned, 56.0, G)
    lv_o = O.map_roi_levels_c(rois, 4)
    assert torch.equal(lvl.cpu().long(), lv_o)
    grid_o, ytab_o, xtab_o = O.roi_align_indices_c(rois, lv_o, hw, scales, P, sr, aligned, G)
    assert torch.equal(grid.cpu(), grid_o)
    assert torch.equal(ytab.cpu(), ytab_o)
    assert torch.equal(xtab.cpu(), xtab_o)
    if sr <= 0:
        assert int(grid_o.max()) > 4       # the adaptive grid really varies


# ---- a5: RoIAlign values -------------------------------------------------------------------------
def test_roi_align_kats_all_kernels():
    from fgn_b200 import ops
    z = np.load(os.path.join(GOLDEN, "roi_align_kat.npz"))
    for n in sorted({k.split(".")[0] for k in z.files}):
        feat, rois, want = _t(z[f"{n}.feat"]), _t(z[f"{n}.rois"]), _t(z[f"{n}.out"])
        scale, P, sr, aligned = float(z[f"{n}.scale"]), int(z[f"{n}.P"]), int(z[f"{n}.sr"]), bool(z[f"{n}.aligned"])
        f, r = feat.to(dev()), rois.to(dev())
        # direct kernel, reference layout: the reference's own summation order -> bit-exact
        d = ops.roi_align_multilevel([f], r, [scale], P, sr, aligned, force_direct=True)
        assert torch.equal(d.cpu(), want), n
        d2 = ops.roi_align_multilevel([f.contiguous(memory_format=torch.channels_last)], r, [scale], P, sr, aligned,
                                      force_direct=True, out_format="nhwc")
        assert torch.equal(d2.cpu().contiguous(), want), n
        if feat.shape[1] % 4 == 0 and P in (7, 14):
            for fmt in ("nchw", "nhwc"):
                s = ops.roi_align_multilevel([f], r, [scale], P, sr, aligned, out_format=fmt)
                close(s, want, what=f"{n}/{fmt}")
            cl = ops.roi_align_multilevel([f.contiguous(memory_format=torch.channels_last)], r, [scale], P, sr, aligned)
            close(cl, want, what=n)


@pytest.mark.parametrize("C,P", [(64, 7), (256, 7), (1024, 7), (256, 14), (132, 7)])
def test_roi_align_multilevel_vs_oracle(C, P):
    from fgn_b200 import ops
    from fgn_b200.episodes import synth_rois
    g = torch.Generator().manual_seed(20 + C + P)
    strides = [4, 8, 16, 32]
    B, img_h, img_w = 2, 256, 384
    feats = [torch.randn(B, C, img_h // s, img_w // s, generator=g) for s in strides]
    rois = synth_rois(g, 300, img_h, img_w, B, smin=6.0)
    want, lv = O.single_roi_extractor(feats, rois, strides, P, 0, True, 56.0, "tv")
    fd = [f.to(dev()).contiguous(memory_format=torch.channels_last) for f in feats]
    got, lvl = ops.roi_align_multilevel(fd, rois.to(dev()), [1 / s for s in strides], P, 0, True, return_levels=True)
    assert torch.equal(lvl.cpu(), lv)
    close(got, want, what="multilevel")
    got2 = ops.roi_align_multilevel(fd, rois.to(dev()), [1 / s for s in strides], P, 0, True, out_format="nhwc")
    assert torch.equal(got2.contiguous(), got)                                   # layouts agree bitwise
    got3 = ops.roi_align_multilevel([f.to(dev()) for f in feats], rois.to(dev()), [1 / s for s in strides], P, 0, True)
    assert torch.equal(got3, got)                                                # NCHW input (repacked) too


@pytest.mark.parametrize("env", [{"FGN_RA_IMPL": "2"}, {"FGN_RA_IMPL": "2", "FGN_RA_NS": "2"},
                                 {"FGN_RA_IMPL": "2", "FGN_RA_VEC": "1"}, {"FGN_RA_IMPL": "2", "FGN_RA_VEC": "3"},
                                 {"FGN_RA_IMPL": "2", "FGN_RA_CLASSES": "3"}, {"FGN_RA_SPLIT": "0"},
                                 {"FGN_RA_SPLIT": "40"}, {"FGN_RA_NS": "2"}],
                         ids=["stream", "stream-ns2", "cb128", "sliced", "three-class",
                              "window-nosplit", "window-split40", "window-ns2"])
def test_roi_align_kernel_variants_agree(env, monkeypatch):
    """Every RoIAlign kernel variant the library can dispatch to (selected through its tuning
    environment knobs) reproduces the oracle and, among the streaming variants, each other bitwise."""
    from fgn_b200 import ops
    from fgn_b200.episodes import synth_rois
    g = torch.Generator().manual_seed(77)
    strides, B, C = [4, 8, 16, 32], 2, 256
    feats = [torch.randn(B, C, 256 // s, 320 // s, generator=g) for s in strides]
    rois = synth_rois(g, 400, 256, 320, B, smin=4.0)
    rois[0, 1:] = torch.tensor([-30., -30., 400., 300.])            # wider than 32 cells on level 3: segmented rows
    rois[1, 1:] = torch.tensor([0., 100., 320., 104.])              # 80 cells wide, 1 cell tall on level 0
    want, lv = O.single_roi_extractor(feats, rois, strides, 7, 0, True, 56.0, "tv")
    fd = [f.to(dev()).contiguous(memory_format=torch.channels_last) for f in feats]
    scales = [1 / s for s in strides]
    base = ops.roi_align_multilevel(fd, rois.to(dev()), scales, 7, 0, True, out_format="nhwc")
    close(base, want, what="default")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got, lvl = ops.roi_align_multilevel(fd, rois.to(dev()), scales, 7, 0, True, out_format="nhwc", return_levels=True)
    assert torch.equal(lvl.cpu(), lv)
    close(got, want, what=str(env))
    assert torch.equal(got, base), "streaming variants share one summation order"


@pytest.mark.parametrize("P,C,B", [(7, 256, 2), (14, 256, 2), (7, 128, 1), (7, 64, 1), (7, 1024, 1), (14, 128, 1), (7, 320, 1)])
def test_roi_align_window_kernel_chunked_and_ragged(P, C, B, monkeypatch):
    """The default (persistent rotating-window) kernel on the shapes that exercise its planner: RoIs
    smaller than the bin grid (split into bin-row chunks), huge and out-of-image RoIs, P=14, channel
    counts below / above / not a multiple of its channel block, the fused channel attention.  Bitwise
    equal to the row-streaming kernel (same summation order), within tolerance of the oracle, and the
    planner's window self-check stays at zero."""
    from fgn_b200 import _lib, ops
    from fgn_b200.episodes import synth_rois
    g = torch.Generator().manual_seed(100 + P + C)
    strides = [4, 8, 16, 32]
    feats = [torch.randn(B, C, 192 // s, 256 // s, generator=g) for s in strides]
    rois = synth_rois(g, 300, 192, 256, B, smin=4.0)
    n = rois.shape[0]
    # tiny boxes (bins far smaller than a cell), slivers, boxes hanging over every border, full image
    rois[0, 1:] = torch.tensor([10.2, 11.7, 12.9, 13.1])
    rois[1, 1:] = torch.tensor([100., 50., 100.5, 120.])
    rois[2, 1:] = torch.tensor([-40., -40., 30., 20.])
    rois[3, 1:] = torch.tensor([200., 150., 300., 260.])
    rois[4, 1:] = torch.tensor([0., 0., 256., 192.])
    rois[5, 1:] = torch.tensor([-500., -500., -400., -300.])
    rois[6, 1:] = torch.tensor([3., 3., 9., 30.])
    rois[7, 1:] = torch.tensor([60., 60., 60., 60.])
    for i in range(8, 40):                                    # a run of very small boxes
        cx, cy = float(torch.rand(1, generator=g)) * 256, float(torch.rand(1, generator=g)) * 192
        w, h = 1 + 11 * float(torch.rand(1, generator=g)), 1 + 11 * float(torch.rand(1, generator=g))
        rois[i, 1:] = torch.tensor([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
    vec = torch.randn(5, C, generator=g)
    idx = torch.randint(0, 5, (n,), generator=g)
    want, lv = O.single_roi_extractor(feats, rois, strides, P, 0, True, 56.0, "tv")
    want_s = want * vec[idx][:, :, None, None]
    fd = [f.to(dev()).contiguous(memory_format=torch.channels_last) for f in feats]
    scales = [1 / s for s in strides]
    kw = dict(out_format="nhwc", return_levels=True)
    before = _lib.load().fgn_launch_count()
    got, lvl = ops.roi_align_multilevel(fd, rois.to(dev()), scales, P, 0, True, **kw)
    got_s, _ = ops.roi_align_multilevel(fd, rois.to(dev()), scales, P, 0, True, chan_scale=vec.to(dev()),
                                        scale_index=idx.to(dev()), **kw)
    assert _lib.load().fgn_launch_count() == before + 2      # one kernel per call
    assert torch.equal(lvl.cpu(), lv)
    close(got, want, what="window vs oracle")
    close(got_s, want_s, what="window + channel attention vs oracle")
    # fixed sampling grids and the torchvision (aligned=False) convention
    for sr, al in ((2, True), (-1, False), (3, False)):
        w2, _ = O.single_roi_extractor(feats, rois, strides, P, sr, al, 56.0, "tv")
        g2 = ops.roi_align_multilevel(fd, rois.to(dev()), scales, P, sr, al, out_format="nhwc")
        close(g2, w2, what=f"window sr={sr} aligned={al}")
    monkeypatch.setenv("FGN_RA_IMPL", "2")
    ref, _ = ops.roi_align_multilevel(fd, rois.to(dev()), scales, P, 0, True, **kw)
    ref_s, _ = ops.roi_align_multilevel(fd, rois.to(dev()), scales, P, 0, True, chan_scale=vec.to(dev()),
                                        scale_index=idx.to(dev()), **kw)
    assert torch.equal(got, ref) and torch.equal(got_s, ref_s), "window and streaming kernels share one summation order"
    torch.cuda.synchronize()
    assert _lib.load().fgn_debug_roi_window_violations() == 0


@pytest.mark.parametrize("P,C,B", [(7, 256, 2), (14, 256, 1), (7, 64, 1), (7, 512, 1), (7, 328, 1)])
def test_roi_align_window_kernel_bf16_cells(P, C, B, monkeypatch):
    """bf16 variant (reported separately): the rotating-window kernel with half-width cells (bf16 NHWC pyramid, bf16 or
    fp32 out) on the planner-stressing RoIs of the fp32 test, rows wider than a 64-cell stage, channel counts below /
    above / not a multiple of the 256-channel block, the fused channel attention.  Levels bit-exact; bitwise equal to
    the one-CTA-per-RoI bf16 kernel where both use the same rounding order (no channel attention); the oracle on the
    SAME bf16-rounded maps within fp32 summation order (fp32 out) or one bf16 rounding (bf16 out)."""
    from fgn_b200 import _lib, ops
    from fgn_b200.episodes import synth_rois
    g = torch.Generator().manual_seed(900 + P + C)
    strides = [4, 8, 16, 32]
    feats = [torch.randn(B, C, 192 // s, 320 // s, generator=g).bfloat16() for s in strides]
    rois = synth_rois(g, 300, 192, 320, B, smin=4.0)
    n = rois.shape[0]
    rois[0, 1:] = torch.tensor([10.2, 11.7, 12.9, 13.1])
    rois[1, 1:] = torch.tensor([100., 50., 100.5, 120.])
    rois[2, 1:] = torch.tensor([-40., -40., 30., 20.])
    rois[3, 1:] = torch.tensor([0., 100., 320., 104.])              # 80 cells wide on level 0: segmented rows
    rois[4, 1:] = torch.tensor([0., 0., 320., 192.])
    rois[5, 1:] = torch.tensor([-500., -500., -400., -300.])
    rois[6, 1:] = torch.tensor([3., 3., 9., 30.])
    rois[7, 1:] = torch.tensor([60., 60., 60., 60.])
    for i in range(8, 40):
        cx, cy = float(torch.rand(1, generator=g)) * 320, float(torch.rand(1, generator=g)) * 192
        w, h = 1 + 11 * float(torch.rand(1, generator=g)), 1 + 11 * float(torch.rand(1, generator=g))
        rois[i, 1:] = torch.tensor([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
    vec = torch.randn(5, C, generator=g)
    idx = torch.randint(0, 5, (n,), generator=g)
    fr = [f.float().contiguous() for f in feats]                    # the rounded maps, as fp32, for the oracle
    want, lv = O.single_roi_extractor(fr, rois, strides, P, 0, True, 56.0, "tv")
    want_s = want * vec[idx][:, :, None, None]
    fd = [f.to(dev()).contiguous(memory_format=torch.channels_last) for f in feats]
    scales = [1 / s for s in strides]
    rd = rois.to(dev())
    sc = dict(chan_scale=vec.to(dev()), scale_index=idx.to(dev()))
    before = _lib.load().fgn_launch_count()
    got16, lvl = ops.roi_align_multilevel(fd, rd, scales, P, 0, True, return_levels=True)
    got32 = ops.roi_align_multilevel(fd, rd, scales, P, 0, True, out_dtype=torch.float32)
    got16_s = ops.roi_align_multilevel(fd, rd, scales, P, 0, True, **sc)
    got32_s = ops.roi_align_multilevel(fd, rd, scales, P, 0, True, out_dtype=torch.float32, **sc)
    assert _lib.load().fgn_launch_count() == before + 4      # one kernel per call
    assert got16.dtype == torch.bfloat16 and got32.dtype == torch.float32 and torch.equal(lvl.cpu(), lv)
    close(got32, want, what="bf16 cells, fp32 out vs oracle")
    close(got32_s, want_s, what="bf16 cells + channel attention, fp32 out vs oracle")
    close(got16.float(), want, atol=1e-3, rtol=2 ** -8, what="bf16 cells, bf16 out vs oracle")
    close(got16_s.float(), want_s, atol=1e-3, rtol=2 ** -8, what="bf16 cells + channel attention, bf16 out vs oracle")
    assert torch.equal(got16, got32.bfloat16()), "bf16 out = round-to-nearest-even of the fp32 out"
    monkeypatch.setenv("FGN_RA_NS", "2")                      # a shallower ring gives the same answer
    assert torch.equal(ops.roi_align_multilevel(fd, rd, scales, P, 0, True, out_dtype=torch.float32), got32)
    monkeypatch.delenv("FGN_RA_NS")
    monkeypatch.setenv("FGN_RA_IMPL", "2")
    ref16 = ops.roi_align_multilevel(fd, rd, scales, P, 0, True)
    ref32 = ops.roi_align_multilevel(fd, rd, scales, P, 0, True, out_dtype=torch.float32)
    assert torch.equal(got32, ref32) and torch.equal(got16, ref16), "window and streaming bf16 kernels share one summation order"
    torch.cuda.synchronize()
    assert _lib.load().fgn_debug_roi_window_violations() == 0


def test_roi_align_window_ticket_schemes():
    """The window kernel's ticket schemes give the same answer: R above its sort capacity (plain tickets in
    index order: per RoI when a CTA gets many items, per bin-row chunk otherwise), R below the number of resident CTAs (every RoI is some CTA's first, ticket-less item), the
    sorted scheme switched off, and the default -- all bitwise equal to the row-streaming kernel."""
    from fgn_b200 import _lib, ops
    from fgn_b200.episodes import synth_rois
    g = torch.Generator().manual_seed(4242)
    strides, B, C = [4, 8, 16, 32], 2, 256
    feats = [torch.randn(B, C, 160 // s, 224 // s, generator=g) for s in strides]
    fd = [f.to(dev()).contiguous(memory_format=torch.channels_last) for f in feats]
    scales = [1 / s for s in strides]
    for R in (10000, 2500, 1100, 40):                          # 10000: one ticket per RoI (many items per CTA)
        rois = synth_rois(g, R, 160, 224, B, smin=4.0 if R == 10000 else 6.0)
        rd = rois.to(dev())
        got, lvl = ops.roi_align_multilevel(fd, rd, scales, 7, 0, True, out_format="nhwc", return_levels=True)
        assert torch.equal(lvl.cpu(), O.map_roi_levels_c(rois, 4))
        want, _ = O.single_roi_extractor(feats, rois[:: max(1, R // 200)], strides, 7, 0, True, 56.0, "tv")
        close(got[:: max(1, R // 200)], want, what=f"window R={R} vs oracle")
        os.environ["FGN_RA_DEBUG"] = "64"                      # sorted tickets off
        try:
            plain = ops.roi_align_multilevel(fd, rd, scales, 7, 0, True, out_format="nhwc")
        finally:
            del os.environ["FGN_RA_DEBUG"]
        os.environ["FGN_RA_IMPL"] = "2"
        try:
            ref = ops.roi_align_multilevel(fd, rd, scales, 7, 0, True, out_format="nhwc")
        finally:
            del os.environ["FGN_RA_IMPL"]
        assert torch.equal(got, ref) and torch.equal(plain, ref), f"R={R}"
    torch.cuda.synchronize()
    assert _lib.load().fgn_debug_roi_window_violations() == 0


def test_roi_align_window_concurrent_launches_on_two_streams():
    """P=7 and P=14 launches of the window kernel in flight at the same time on two streams (the episode runner's
    multi-stream mode: bbox + mask branches of different episodes).  Every call owns its ticket counter and plans in
    its own workspace, so overlapping persistent kernels cannot share tickets; results equal the direct kernel's
    within tolerance and the serial results bitwise."""
    from fgn_b200 import ops
    from fgn_b200.episodes import synth_rois
    g = torch.Generator().manual_seed(777)
    strides, B, C = [4, 8, 16, 32], 2, 256
    feats = [torch.randn(B, C, 320 // s, 448 // s, generator=g) for s in strides]
    fd = [f.to(dev()).contiguous(memory_format=torch.channels_last) for f in feats]
    scales = [1 / s for s in strides]
    r7 = synth_rois(g, 1500, 320, 448, B, smin=8.0).to(dev())
    r14 = synth_rois(g, 600, 320, 448, B, smin=8.0).to(dev())
    serial7 = ops.roi_align_multilevel(fd, r7, scales, 7, 0, True, out_format="nhwc")
    serial14 = ops.roi_align_multilevel(fd, r14, scales, 14, 0, True, out_format="nhwc")
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs7, outs14 = [], []
    for _ in range(6):
        with torch.cuda.stream(s1):
            outs7.append(ops.roi_align_multilevel(fd, r7, scales, 7, 0, True, out_format="nhwc"))
        with torch.cuda.stream(s2):
            outs14.append(ops.roi_align_multilevel(fd, r14, scales, 14, 0, True, out_format="nhwc"))
    torch.cuda.synchronize()
    for o in outs7:
        assert torch.equal(o, serial7)
    for o in outs14:
        assert torch.equal(o, serial14)
    d7 = ops.roi_align_multilevel(fd, r7[::10].contiguous(), scales, 7, 0, True, out_format="nhwc", force_direct=True)
    close(serial7[::10], d7, what="window P=7 vs direct kernel")
    d14 = ops.roi_align_multilevel(fd, r14[::10].contiguous(), scales, 14, 0, True, out_format="nhwc", force_direct=True)
    close(serial14[::10], d14, what="window P=14 vs direct kernel")


def test_roi_align_edge_cases():
    from fgn_b200 import ops
    f = torch.randn(1, 8, 12, 12, device=dev()).contiguous(memory_format=torch.channels_last)
    empty = ops.roi_align_multilevel([f], torch.zeros(0, 5, device=dev()), [1 / 16])
    assert empty.shape == (0, 8, 7, 7)
    rois = torch.tensor([[0, -900., -900., -800., -800.], [0, 50., 50., 50., 50.], [0, 80., 80., 20., 20.]], device=dev())
    out = ops.roi_align_multilevel([f], rois, [1 / 16])
    want = O.roi_align_tv(f.cpu().contiguous(), rois.cpu(), 1 / 16, 7, 0, True)
    close(out, want, what="degenerate")
    assert (out[0] == 0).all()


def test_roi_batch_index_outside_the_batch_pools_zeros():
    """A stale / corrupt rois[:,0] (negative, == B, huge) never reads or writes out of bounds: every forward kernel
    returns zeros for that RoI, the backward adds nothing, the other RoIs are untouched."""
    from fgn_b200 import autograd as A
    from fgn_b200 import ops
    from fgn_b200.episodes import synth_rois
    g = torch.Generator().manual_seed(5)
    B, C = 2, 256
    strides = [4, 8, 16, 32]
    feats = [torch.randn(B, C, 128 // s, 160 // s, generator=g).to(dev()).contiguous(memory_format=torch.channels_last) for s in strides]
    scales = [1 / s for s in strides]
    rois = synth_rois(g, 64, 128, 160, B, smin=8.0).to(dev())
    good = ops.roi_align_multilevel(feats, rois, scales, 7, 0, True, out_format="nhwc")
    bad = rois.clone()
    bad_rows = torch.tensor([3, 17, 40, 63], device=dev())
    bad[bad_rows, 0] = torch.tensor([-1.0, float(B), 1.0e6, -3.0e9], device=dev())
    for kw in (dict(out_format="nhwc"), dict(out_format="nchw"), dict(out_format="nhwc", force_direct=True)):
        got = ops.roi_align_multilevel(feats, bad, scales, 7, 0, True, **kw)
        assert (got[bad_rows] == 0).all(), kw
        keep = torch.ones(64, dtype=torch.bool, device=dev())
        keep[bad_rows] = False
        want = good if kw.get("out_format") == "nhwc" and not kw.get("force_direct") else \
            ops.roi_align_multilevel(feats, rois, scales, 7, 0, True, **kw)
        assert torch.equal(got[keep], want[keep]), kw
    fr = [f.clone().requires_grad_(True) for f in feats]
    A.roi_align_multilevel(fr, bad, scales, 7, 0, True).sum().backward()
    fr2 = [f.clone().requires_grad_(True) for f in feats]
    keep_rois = bad[torch.tensor([i for i in range(64) if i not in (3, 17, 40, 63)], device=dev())]
    A.roi_align_multilevel(fr2, keep_rois, scales, 7, 0, True).sum().backward()
    for a, b in zip(fr, fr2):
        assert torch.allclose(a.grad, b.grad, atol=1e-5, rtol=1e-5)


# ---- full-size, size-independent properties (cfg3 pyramid) ---------------------------------------
def test_full_size_properties_cfg3():
    from fgn_b200 import ops
    from fgn_b200.episodes import CONFIGS, make_episode, episode_to_device
    ep = episode_to_device(make_episode(CONFIGS["cfg3_coco2voc_n1k1_fpn"], seed=3), dev())
    feats, rois = ep["qry"][:4], ep["rois"]
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    a, lv = ops.roi_align_multilevel(feats, rois, scales, 7, 0, True, return_levels=True)
    assert torch.equal(lv.cpu(), O.map_roi_levels_c(rois.cpu(), 4))
    # (1) fast kernel == direct kernel (reference summation order) within tolerance, all 1000 RoIs
    d = ops.roi_align_multilevel(feats, rois, scales, 7, 0, True, force_direct=True)
    close(a, d, what="separable vs direct")
    # (2) linearity: RA(2x + y) == 2 RA(x) + RA(y)
    g = torch.Generator(device="cpu").manual_seed(8)
    other = [torch.randn(f.shape, generator=g).to(dev()).contiguous(memory_format=torch.channels_last) for f in feats]
    b = ops.roi_align_multilevel(other, rois, scales, 7, 0, True)
    mix = [2 * x + y for x, y in zip(feats, other)]
    c = ops.roi_align_multilevel(mix, rois, scales, 7, 0, True)
    close(c, 2 * a + b, atol=2e-4, what="linearity")
    # (3) a constant map pools to the constant wherever every sample is inside the map
    ones = [torch.ones_like(f) for f in feats]
    o = ops.roi_align_multilevel(ones, rois, scales, 7, 0, True)
    inside = (rois[:, 1] > 40) & (rois[:, 2] > 40) & (rois[:, 3] < 1344 - 40) & (rois[:, 4] < 800 - 40)
    assert inside.sum() > 100
    assert (o[inside] - 1).abs().max() < 1e-5
    # (4) channel-permutation equivariance, bitwise
    perm = torch.randperm(256, generator=g).to(dev())
    pf = [f[:, perm].contiguous(memory_format=torch.channels_last) for f in feats]
    p = ops.roi_align_multilevel(pf, rois, scales, 7, 0, True)
    assert torch.equal(p, a[:, perm])
    # (5) an oracle spot check on a 64-RoI subset of the full-size episode
    sub = rois[::16]
    want, _ = O.single_roi_extractor([f.cpu().contiguous() for f in feats], sub.cpu(), [4, 8, 16, 32], 7, 0, True, 56.0, "tv")
    close(a[::16], want, what="cfg3 subset vs oracle")


# ---- a2: support branch -------------------------------------------------------------------------
def test_support_mask_pool_vs_oracle():
    from fgn_b200 import ops
    from fgn_b200.episodes import synth_support
    g = torch.Generator().manual_seed(31)
    for S in (64, 128, 256):
        boxes, masks = synth_support(g, 6, S)
        boxes[0, 0] = torch.tensor([-10., -10., S + 5., S + 5.])
        boxes[1, 0] = torch.tensor([3., 4., 3.5, 4.2])
        m = boxes.shape[0]
        rois = torch.cat([torch.arange(m).float().view(m, 1), boxes.view(m, 4)], 1)
        want = O.roi_align_tv(masks.float(), rois, 1.0, 7, -1, False)
        got = ops.support_mask_pool(masks.to(dev()), boxes.to(dev()), 7)
        close(got, want, what=f"mask pool S={S}")


@pytest.mark.parametrize("path", REF_FIXTURES, ids=[os.path.basename(p)[14:-4] for p in REF_FIXTURES])
def test_reference_fixture_end_to_end(path):
    """The reference's own outputs (tests/golden/fgn_reference_*.npz) through the mirror modules."""
    from fgn_b200 import AGRPNHead, FGNRoIHead, ops
    z = np.load(path)
    B, N, K, C, stride = (int(z[k]) for k in ("B", "N", "K", "C", "stride"))
    shared = None
    if "w.shared_head.0.weight" in z.files:
        shared = torch.nn.Sequential(torch.nn.Conv2d(C, C, 3, padding=1), torch.nn.ReLU())
        shared[0].weight.data.copy_(_t(z["w.shared_head.0.weight"]))
        shared[0].bias.data.copy_(_t(z["w.shared_head.0.bias"]))
    head = FGNRoIHead(bbox_roi_extractor=dict(type="SingleRoIExtractor",
                                              roi_layer=dict(type="RoIAlign", output_size=7, sampling_ratio=0),
                                              out_channels=C, featmap_strides=[stride]),
                      shared_head=shared, channels=C, n_ways=N, k_shots=K)
    sd = {k[2:]: _t(z[k]) for k in z.files if k.startswith("w.") and not k.startswith("w.shared_head")}
    missing = head.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if not k.startswith("shared_head")]
    head = head.to(dev()).eval()
    head.subsampling_ratio = stride
    qry, spp = _t(z["qry"]).to(dev()), _t(z["spp"]).to(dev())
    boxes = _t(z["spp_bboxes"].copy()).to(dev())
    with torch.no_grad():
        head.count_spp(spp, boxes, _t(z["spp_masks"]).to(dev()))
        assert torch.equal(boxes.cpu(), _t(z["spp_bboxes_after"]))              # side effect of :430 kept
        close(head.spp_fmaps_roi_aligned_cat_mean, _t(z["cat_mean"]), what="cat_mean")
        close(head.spp_fvecs_roi_aligned_cat_mean_mp, _t(z["masked_gap"]), what="masked_gap")
        assert head.spp_fmaps_roi_aligned_cat_mean.shape == (B, N, C, 7, 7)
        assert head.spp_fvecs_roi_aligned_cat_mean_mp.shape == (B, N, C, 1, 1)

        rpn = AGRPNHead(in_channels=C, feat_channels=C, num_anchors=15, n_ways=N, k_shots=K)
        rpn.load_state_dict({k[5:]: _t(z[k]) for k in z.files if k.startswith("rpnw.")})
        rpn = rpn.to(dev()).eval()
        vec, mod = rpn.attention(qry, spp)
        close(mod, _t(z["rpn_qry_fmap_mod"]), what="qry_fmap_mod")
        cls, reg = rpn.forward_single(qry, spp)
        close(cls, _t(z["rpn_cls"]), atol=2e-4, what="rpn_cls")               # includes cuDNN convs
        close(reg, _t(z["rpn_reg"]), atol=2e-4, what="rpn_reg")
        if N > 1:   # selection alone, on the reference's own conv outputs: exact
            c2, r2 = ops.best_class_select(_t(z["rpn_cls_raw"]).to(dev()), _t(z["rpn_reg_raw"]).to(dev()), B, N)
            assert torch.equal(c2.cpu(), _t(z["rpn_cls"])) and torch.equal(r2.cpu(), _t(z["rpn_reg"]))
        if "cls_score" not in z.files:
            return
        rois = _t(z["rois"]).to(dev())
        res = head._bbox_forward(qry, rois, need_feats=True)
        close(res["bbox_feats"], _t(z["bbox_feats"]), what="bbox_feats")
        close(res["cls_score"], _t(z["cls_score"]), what="cls_score")
        close(res["bbox_pred"], _t(z["bbox_pred"]), what="bbox_pred")
        assert res["cls_score"].shape == (rois.shape[0], N + 1) and res["bbox_pred"].shape == (rois.shape[0], 4 * N)
        # relation fusion alone on the reference's RoI features, with the raw head outputs
        c3, r3, rawc, rawr = ops.relation_fusion(_t(z["bbox_feats"]).to(dev()), rois[:, 0], head.spp_fmaps_roi_aligned_cat_mean,
                                                 N, head.relation_params(), return_raw=True)
        close(rawc, _t(z["raw_cls"]), what="raw_cls")
        close(rawr, _t(z["raw_reg"]), what="raw_reg")
        c4, r4 = head.count_modified_cls_bbox(rois.shape[0], _t(z["raw_cls"]).to(dev()), _t(z["raw_reg"]).to(dev()))
        assert torch.equal(c4.cpu(), _t(z["cls_score"])) and torch.equal(r4.cpu(), _t(z["bbox_pred"]))
        if shared is None:      # FPN-style single call (no shared head between RoIAlign and fusion)
            fused = head._bbox_forward(qry, rois, need_feats=False)
            assert fused["bbox_feats"] is None
            close(fused["cls_score"], _t(z["cls_score"]), what="fused cls")
            close(fused["bbox_pred"], _t(z["bbox_pred"]), what="fused reg")
        # mask branch
        det_rois, labels = _t(z["det_rois"]).to(dev()), _t(z["det_labels"]).to(dev())
        det_labels = [labels[det_rois[:, 0] == b] for b in range(B)]
        head.gather_mask_vectors(det_labels)
        mres = head._mask_forward(qry, det_rois)
        close(mres["mask_feats"], _t(z["mask_feats"]), what="mask_feats")
        # train-time form: pos_inds + bbox_feats (fgn_roi_head.py:371-374)
        pos = torch.zeros(rois.shape[0], dtype=torch.uint8, device=dev())
        pos[: det_rois.shape[0]] = 1
        m2 = head._mask_forward(qry, pos_inds=pos, bbox_feats=res["bbox_feats"])
        close(m2["mask_feats"], _t(z["mask_feats"]), what="mask_feats (pos_inds)")


# ---- FPN-mode configs vs the oracle ---------------------------------------------------------------
@pytest.mark.parametrize("name,R", [("tiny_fpn", 96), ("tiny_c4", 64)])
def test_guided_path_small_configs_vs_oracle(name, R):
    from fgn_b200.episodes import CONFIGS, build_heads, episode_to_device, make_episode, make_weights, run_guided_path
    cfg = CONFIGS[name]
    ep = make_episode(cfg, seed=2)
    rpn, head = build_heads(cfg, dev(), seed=0)
    with torch.no_grad():
        out = run_guided_path(rpn, head, episode_to_device(ep, dev()))
    w = make_weights(cfg.channels, 0)
    n_ext = len(cfg.strides)
    for lvl in range(len(cfg.rpn_strides)):
        _, mod = O.agrpn_attention(ep["qry"][lvl], ep["spp"][lvl], cfg.n_ways, cfg.k_shots)
        close(out["qry_fmap_mod"][lvl], mod, what=f"attention level {lvl}")
    if cfg.mode == "fpn":
        cat_mean, mp, _, _ = O.count_spp_fpn(ep["spp"][:n_ext], cfg.strides, ep["spp_bboxes"].clone(), ep["spp_masks"],
                                             cfg.n_ways, cfg.k_shots)
    else:
        cat_mean, mp, _, _ = O.count_spp(ep["spp"][0], ep["spp_bboxes"].clone(), ep["spp_masks"], cfg.n_ways, cfg.k_shots, 16)
    close(head.spp_fmaps_roi_aligned_cat_mean, cat_mean, what="cat_mean")
    close(head.spp_fvecs_roi_aligned_cat_mean_mp, mp, what="masked_gap")
    res = O.bbox_forward(ep["qry"][:n_ext], cfg.strides, ep["rois"], cat_mean, cfg.n_ways, w)
    close(out["cls_score"], res["cls_score"], what="cls_score")
    close(out["bbox_pred"], res["bbox_pred"], what="bbox_pred")
    det = ep["det_rois"]
    labels = [ep["det_labels"][det[:, 0] == b] for b in range(cfg.batch)]
    mf = O.mask_attention(ep["qry"][:n_ext], cfg.strides, det, mp, labels, cfg.n_ways, cfg.mask_size)
    close(out["mask_feats"], mf, what="mask_feats")


@pytest.mark.timeout(180)
@pytest.mark.parametrize("R,images,order", [(600, 1, "grouped"), (1000, 3, "grouped"), (2311, 5, "grouped"),
                                            (1500, 4, "interleaved"), (12000, 12, "grouped")])
def test_relation_epilogue_ring_kernel_bitwise(R, images, order, monkeypatch):
    """The persistent, bulk-copy-fed epilogue of the headline shape (N = 1, C = 256, P = 7; launches of 512+ RoIs) gives
    bit for bit the results of the one-CTA-per-RoI kernels it stands in for -- RoI counts that leave CTAs with ragged or
    empty ranges, image changes inside a CTA's range (the class-term reload behind the two teams' barrier), RoIs NOT
    grouped by image (a reload on every RoI), one consumer team instead of two -- and those agree with the oracle."""
    from fgn_b200 import ops
    from fgn_b200.episodes import CONFIGS, build_heads, make_weights
    cfg = CONFIGS["cfg3_coco2voc_n1k1_fpn"]
    rpn, head = build_heads(cfg, dev(), seed=0)
    params = head.relation_params()
    g = torch.Generator().manual_seed(R + images)
    C = cfg.channels
    feat = torch.randn(R, 7, 7, C, generator=g).permute(0, 3, 1, 2)
    rb = (torch.arange(R) % images) if order == "interleaved" else (torch.arange(R) * images // R)
    spp = torch.randn(images, 1, C, 7, 7, generator=g)
    fd, rbd, sd = feat.to(dev()), rb.to(dev()), spp.to(dev())
    got = ops.relation_fusion(fd, rbd, sd, 1, params, return_raw=True)
    monkeypatch.setenv("FGN_EPI_TEAMS", "1")
    one_team = ops.relation_fusion(fd, rbd, sd, 1, params, return_raw=True)
    monkeypatch.setenv("FGN_EPI_RING", "0")
    per_roi = ops.relation_fusion(fd, rbd, sd, 1, params, return_raw=True)
    monkeypatch.setenv("FGN_EPI_ONE", "0")
    general = ops.relation_fusion(fd, rbd, sd, 1, params, return_raw=True)
    for a, b, c, d in zip(got, one_team, per_roi, general):
        assert torch.equal(a, d) and torch.equal(b, d) and torch.equal(c, d)
    if R <= 1000:
        w = make_weights(C, 0)
        sub = slice(0, R, 7)
        rois = torch.zeros(R, 5)
        rois[:, 0] = rb.float()
        n_r, fused = O.count_one_roi_by_n_spp(feat[sub].contiguous(), rois[sub], spp, 1, w["conv_w"], w["conv_b"], w["gn_w"], w["gn_b"])
        rc, rr = O.bbox_head_forward(fused, w["fc_cls_w"], w["fc_cls_b"], w["fc_reg_w"], w["fc_reg_b"])
        wc, wr = O.count_modified_cls_bbox(n_r, rc, rr, 1)
        close(got[0][sub], wc, what="ring epilogue cls vs oracle")
        close(got[1][sub], wr, what="ring epilogue reg vs oracle")


def test_relation_stress_n20_k5_vs_oracle():
    """cfg4's shape family (N=20, K=5, C=256) at a RoI count the CPU oracle finishes in seconds."""
    from fgn_b200 import ops
    from fgn_b200.episodes import make_weights
    g = torch.Generator().manual_seed(44)
    N, C, R, B = 20, 256, 40, 2
    w = make_weights(C, 1)
    feats = torch.randn(R, C, 7, 7, generator=g)
    cat_mean = torch.randn(B, N, C, 7, 7, generator=g)
    rois = torch.zeros(R, 5)
    rois[R // 2:, 0] = 1
    _, fused = O.count_one_roi_by_n_spp(feats, rois, cat_mean, N, w["conv_w"], w["conv_b"], w["gn_w"], w["gn_b"])
    rc, rr = O.bbox_head_forward(fused, w["fc_cls_w"], w["fc_cls_b"], w["fc_reg_w"], w["fc_reg_b"])
    wc, wr = O.count_modified_cls_bbox(R, rc, rr, N)
    params = ops.RelationParams(*[w[k].to(dev()) for k in ("conv_w", "conv_b", "gn_w", "gn_b", "fc_cls_w", "fc_cls_b",
                                                            "fc_reg_w", "fc_reg_b")])
    for fmt in (torch.contiguous_format, torch.channels_last):
        c, r, rawc, rawr = ops.relation_fusion(feats.to(dev()).contiguous(memory_format=fmt), rois[:, 0].to(dev()),
                                               cat_mean.to(dev()), N, params, return_raw=True)
        close(rawc, rc, what="raw cls")
        close(rawr, rr, what="raw reg")
        close(c, wc, what="cls")
        close(r, wr, what="reg")
        assert c.shape == (R, N + 1) and r.shape == (R, 4 * N)


def test_mask_branch_p14_vs_oracle():
    """cfg5's op: 14x14 multi-level RoIAlign with the AG-FCN vector multiply fused in."""
    from fgn_b200 import ops
    from fgn_b200.episodes import synth_rois
    g = torch.Generator().manual_seed(55)
    strides, B, C, N = [4, 8, 16, 32], 4, 256, 3
    feats = [torch.randn(B, C, 160 // s, 224 // s, generator=g) for s in strides]
    rois = synth_rois(g, 64, 160, 224, B, smin=6.0)
    mp = torch.randn(B, N, C, 1, 1, generator=g)
    labels = torch.randint(0, N, (64,), generator=g)
    det_labels = [labels[rois[:, 0] == b] for b in range(B)]
    want = O.mask_attention(feats, strides, rois, mp, det_labels, N, 14)
    gather = torch.cat([det_labels[i] + N * i for i in range(B)]).to(torch.int32)
    fd = [f.to(dev()).contiguous(memory_format=torch.channels_last) for f in feats]
    got = ops.roi_align_multilevel(fd, rois.to(dev()), [1 / s for s in strides], 14, 0, True,
                                   chan_scale=mp.view(B * N, C).to(dev()), scale_index=gather.to(dev()))
    close(got, want, what="p14 mask feats")


@pytest.mark.parametrize("M,N,K", [(49 * 40, 256, 256), (49 * 3, 64, 64), (1000, 256, 256), (128 * 5 + 17, 1024, 1024), (300, 48, 32)])
def test_relation_contraction_tcgen05_vs_fp64(M, N, K):
    """3xTF32 tcgen05 contraction holds fp32 parity (vs an fp64 matmul); the SIMT kernel and the strided
    weight views used by the relation head (conv_w[:, :C], conv_w[:, C:]) go through the same entry."""
    from fgn_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, 2 * K, generator=g) * (2.0 / (2 * K)) ** 0.5
    bias = torch.randn(N, generator=g)
    wd = w.to(dev())
    for half in (0, 1):
        bview = wd[:, half * K:(half + 1) * K]
        want = (a.double() @ w[:, half * K:(half + 1) * K].double().t() + bias.double()).float()
        got = ops.gemm_nt(a.to(dev()), bview, bias.to(dev()), "fp32")
        # tensor-core fp32 accumulation truncates: error grows ~K*6e-8 (4e-5 at K=1024), inside the 1e-4 bar
        close(got, want, what=f"3xTF32 half={half}")
        assert float((got.cpu() - want).abs().max()) < (2e-5 if K <= 256 else 6e-5)
        simt = ops.gemm_nt(a.to(dev()), bview, bias.to(dev()), "fp32", use_workspace=False)
        close(simt, want, atol=2e-5, rtol=1e-5, what="simt")
    if K % 32 == 0 and N % 16 == 0:
        fast = ops.gemm_nt(a.to(dev()), wd[:, :K], bias.to(dev()), "tf32")
        err = (fast.cpu() - (a.double() @ w[:, :K].double().t() + bias.double()).float()).abs().max()
        assert 1e-6 < float(err) < 2e-2, float(err)        # single-pass TF32: visibly lower precision


def test_c4_full_size_cfg2_subset_vs_oracle():
    """cfg2 (OMNIISEG N3K1, C4 1024 channels, 32x32 map) at full tensor sizes; the oracle is run on a
    60-RoI subset so the CPU side stays within seconds."""
    from fgn_b200.episodes import CONFIGS, build_heads, episode_to_device, make_episode, make_weights
    cfg = CONFIGS["cfg2_omniiseg_n3k1_c4"]
    ep = make_episode(cfg, seed=5)
    rpn, head = build_heads(cfg, dev(), seed=0, shared_head=None)
    epd = episode_to_device(ep, dev())
    with torch.no_grad():
        head.count_spp(epd["spp"][0], epd["spp_bboxes"].clone(), epd["spp_masks"])
        res = head._bbox_forward(epd["qry"][0], epd["rois"], need_feats=True)
    w = make_weights(cfg.channels, 0)
    cat_mean, mp, _, _ = O.count_spp(ep["spp"][0], ep["spp_bboxes"].clone(), ep["spp_masks"], cfg.n_ways, cfg.k_shots, 16)
    close(head.spp_fmaps_roi_aligned_cat_mean, cat_mean, what="cat_mean")
    close(head.spp_fvecs_roi_aligned_cat_mean_mp, mp, what="masked_gap")
    sub = ep["rois"][::5]
    want = O.bbox_forward([ep["qry"][0]], cfg.strides, sub, cat_mean, cfg.n_ways, w)
    close(res["bbox_feats"][::5], want["bbox_feats"], what="bbox_feats")
    close(res["cls_score"][::5], want["cls_score"], what="cls_score")
    close(res["bbox_pred"][::5], want["bbox_pred"], what="bbox_pred")


def test_support_pool_layouts_agree():
    from fgn_b200 import ops
    g = torch.Generator().manual_seed(91)
    N, K, C = 4, 3, 96
    f = torch.randn(2 * N * K, C, 7, 7, generator=g)
    m = torch.rand(2 * N * K, 1, 7, 7, generator=g)
    want_cat = f.view(2, N, K, C, 7, 7).mean(2)
    want_gap = (f * m).view(2, N, K, C, 7, 7).mean((2, 4, 5)).view(2, N, C, 1, 1)
    for fmt in (torch.contiguous_format, torch.channels_last):
        for out_fmt in ("nchw", "nhwc"):
            cat, gap = ops.support_pool(f.to(dev()).contiguous(memory_format=fmt), m.to(dev()), N, K, out_format=out_fmt)
            close(cat, want_cat, what=f"cat {fmt} {out_fmt}")
            close(gap, want_gap, what=f"gap {fmt} {out_fmt}")
            assert cat.shape == (2, N, C, 7, 7) and gap.shape == (2, N, C, 1, 1)


def test_attention_layouts_and_multilevel_agree():
    from fgn_b200 import ops
    g = torch.Generator().manual_seed(92)
    B, N, K, C = 2, 3, 2, 64
    qs = [torch.randn(B, C, h, w, generator=g) for h, w in ((24, 36), (12, 18), (6, 9), (3, 5))]
    ss = [torch.randn(B * N * K, C, h, h, generator=g) for h in (16, 8, 4, 2)]
    want = [O.agrpn_attention(q, s, N, K) for q, s in zip(qs, ss)]
    cl = lambda t: t.to(dev()).contiguous(memory_format=torch.channels_last)
    vecs, mods = ops.attention_multilevel([cl(q) for q in qs], [cl(s) for s in ss], N, K)
    for l in range(4):
        close(vecs[l], want[l][0], what=f"vec ml {l}")
        close(mods[l], want[l][1], what=f"mod ml {l}")
        v = ops.attention_vectors(ss[l].to(dev()), N, K)                     # NCHW single-level path
        close(v, want[l][0], what=f"vec nchw {l}")
        close(ops.channel_attention(qs[l].to(dev()), v), want[l][1], what=f"mod nchw {l}")
        v2 = ops.attention_vectors(cl(ss[l]), N, K)                          # NHWC single-level path
        close(ops.channel_attention(cl(qs[l]), v2), want[l][1], what=f"mod nhwc {l}")


@pytest.mark.parametrize("M,N,K", [(49 * 40, 256, 256), (1000, 64, 64), (700, 1024, 1024)])
def test_bf16_contraction_variant(M, N, K):
    """bf16 variant (reported separately): exact products of bf16 operands, fp32 accumulation.  Against an
    fp64 matmul of the SAME bf16-rounded operands the error is fp32-accumulation sized; against the
    un-rounded fp32 operands it is the bf16 input rounding, ~2^-9 relative per term (stated tolerance 3e-2)."""
    from fgn_b200 import ops
    g = torch.Generator().manual_seed(M + K)
    a, w, bias = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * K ** -0.5, torch.randn(N, generator=g)
    ab, wb = a.bfloat16(), w.bfloat16()
    got = ops.gemm_nt(ab.to(dev()), wb.to(dev()), bias.to(dev()))
    want_same_inputs = (ab.double() @ wb.double().t() + bias.double()).float()
    close(got, want_same_inputs, atol=1e-4, rtol=1e-5, what="bf16 operands, fp32 accumulate")
    want_fp32 = (a.double() @ w.double().t() + bias.double()).float()
    assert float((got.cpu() - want_fp32).abs().max()) < 3e-2


def test_bf16_guided_path_variant():
    """bf16 variant of the FPN path (reported separately): bf16 NHWC pyramid, bf16 RoI features and
    contraction operands, fp32 everywhere else.  Levels stay bit-exact; RoI features are compared with
    the fp32 oracle evaluated on the SAME bf16-rounded maps (tolerance = one bf16 rounding of the
    output, 2^-8 relative); logits carry the stated bf16 tolerance of 3e-2 absolute."""
    from fgn_b200 import ops
    from fgn_b200.episodes import CONFIGS, build_heads, episode_to_device, make_episode, make_weights
    cfg = CONFIGS["tiny_fpn"]
    ep = make_episode(cfg, seed=4)
    rpn, head = build_heads(cfg, dev(), seed=0)
    epd = episode_to_device(ep, dev())
    q16 = [q.bfloat16().contiguous(memory_format=torch.channels_last) for q in epd["qry"][:4]]
    s16 = [s.bfloat16().contiguous(memory_format=torch.channels_last) for s in epd["spp"][:4]]
    qr = [q.float().cpu().contiguous() for q in q16]                         # the rounded maps, as fp32, for the oracle
    sr = [s.float().cpu().contiguous() for s in s16]
    scales = [1 / s for s in cfg.strides]
    feats, lvl = ops.roi_align_multilevel(q16, epd["rois"], scales, 7, 0, True, return_levels=True)
    want, lv = O.single_roi_extractor(qr, ep["rois"], cfg.strides, 7, 0, True, 56.0, "tv")
    assert feats.dtype == torch.bfloat16 and torch.equal(lvl.cpu(), lv)
    close(feats.float(), want, atol=1e-3, rtol=2 ** -8, what="bf16 roi feats")
    f32out = ops.roi_align_multilevel(q16, epd["rois"], scales, 7, 0, True, out_dtype=torch.float32)
    close(f32out, want, what="bf16 maps, fp32 accumulate + fp32 out")       # only summation order differs
    with torch.no_grad():
        head.count_spp(s16, epd["spp_bboxes"].clone(), epd["spp_masks"])
        res = head._bbox_forward(q16, epd["rois"], need_feats=False)
        head.gather_mask_vectors(epd["det_labels_list"])
        mres = head._mask_forward(q16, epd["det_rois"])
    vecs, mods = rpn.attention_multilevel(q16, s16)
    for l in range(4):
        wv, wm = O.agrpn_attention(qr[l], sr[l], cfg.n_ways, cfg.k_shots)
        close(vecs[l], wv, what=f"bf16 attention vec {l}")                     # fp32 accumulation of bf16 maps
        assert mods[l].dtype == torch.bfloat16
        close(mods[l].float(), wm, atol=1e-3, rtol=2 ** -8, what=f"bf16 attention out {l}")
    w = make_weights(cfg.channels, 0)
    cat_mean, mp, _, _ = O.count_spp_fpn(sr, cfg.strides, ep["spp_bboxes"].clone(), ep["spp_masks"], cfg.n_ways, cfg.k_shots)
    close(head.spp_fmaps_roi_aligned_cat_mean, cat_mean, what="cat_mean from bf16 maps (fp32 out)")
    ref = O.bbox_forward(qr, cfg.strides, ep["rois"], cat_mean, cfg.n_ways, w)
    for k in ("cls_score", "bbox_pred"):
        err = float((res[k].cpu() - ref[k]).abs().max())
        assert err < 3e-2, (k, err)
    mf = O.mask_attention(qr, cfg.strides, ep["det_rois"], mp, ep["det_labels_list"], cfg.n_ways, 7)
    assert mres["mask_feats"].dtype == torch.bfloat16
    close(mres["mask_feats"].float(), mf, atol=1e-3, rtol=2 ** -7, what="bf16 mask feats")


def test_full_size_properties_cfg4_relation():
    """cfg4 at BASELINE size (N=20 ways, K=5 shots, R=1000, C=256): size-independent properties of the
    relation fusion -- (1) each class's (bg,fg) logits and box deltas equal a 1-way run with that class
    alone, (2) permuting the classes permutes the outputs, (3) the background column follows the
    first-max foreground class, (4) chunking the RoIs changes nothing."""
    from fgn_b200 import ops
    from fgn_b200.episodes import make_weights
    g = torch.Generator().manual_seed(404)
    N, C, R = 20, 256, 1000
    w = make_weights(C, 3)
    params = ops.RelationParams(*[w[k].to(dev()) for k in ("conv_w", "conv_b", "gn_w", "gn_b", "fc_cls_w", "fc_cls_b",
                                                            "fc_reg_w", "fc_reg_b")])
    feats = torch.randn(R, C, 7, 7, generator=g).to(dev()).contiguous(memory_format=torch.channels_last)
    cat = torch.randn(1, N, C, 7, 7, generator=g).to(dev())
    rb = torch.zeros(R, device=dev())
    cls, reg, rawc, rawr = ops.relation_fusion(feats, rb, cat, N, params, return_raw=True)
    assert cls.shape == (R, N + 1) and reg.shape == (R, 4 * N) and torch.isfinite(cls).all()
    for n in (0, 7, 19):                                                     # (1)
        c1, r1 = ops.relation_fusion(feats, rb, cat[:, n:n + 1].contiguous(), 1, params)
        # (the class term of a 1-way call is a 49-row contraction and runs on the exact fp32 SIMT kernel, the 20-way
        #  call's 980 rows on the 3xTF32 tcgen05 kernel: same numbers up to the contraction's rounding)
        close(c1[:, 0], cls[:, n], atol=2e-5, what="1-way fg vs 20-way column")
        close(r1, reg[:, 4 * n:4 * n + 4], atol=2e-5, what="1-way deltas vs 20-way columns")
        close(c1[:, 1], rawc.view(R, N, 2)[:, n, 0], atol=2e-5, what="1-way bg vs 20-way raw")
    perm = torch.randperm(N, generator=g).to(dev())                          # (2)
    cp, rp = ops.relation_fusion(feats, rb, cat[:, perm].contiguous(), N, params)
    assert torch.equal(cp[:, :N], cls[:, perm]) and torch.equal(rp.view(R, N, 4), reg.view(R, N, 4)[:, perm])
    top = cls[:, :N].argmax(1)                                               # (3)
    assert torch.equal(cls[:, N], rawc.view(R, N, 2)[torch.arange(R, device=dev()), top, 0])
    ca, ra = ops.relation_fusion(feats[:300].contiguous(memory_format=torch.channels_last), rb[:300], cat, N, params)   # (4)
    assert torch.equal(ca, cls[:300]) and torch.equal(ra, reg[:300])


def test_full_size_properties_cfg5_mask_branch():
    """cfg5 at BASELINE size (16 images/GPU, P=14, C=256, 100 detections per image): batched extraction
    equals per-image extraction bitwise, and the fused AG-FCN multiply equals RoIAlign followed by the
    channel-attention kernel."""
    from fgn_b200 import ops
    from fgn_b200.episodes import CONFIGS, episode_to_device, make_episode
    cfg = CONFIGS["cfg5_coco2voc_mask_fpn"]
    ep = episode_to_device(make_episode(cfg, seed=1), dev())
    feats, det = ep["qry"][:4], ep["det_rois"]
    scales = [1 / s for s in cfg.strides]
    D, Cc = det.shape[0], cfg.channels
    vec = torch.randn(D, Cc, device=dev())
    fused = ops.roi_align_multilevel(feats, det, scales, 14, 0, True, chan_scale=vec, out_format="nhwc")
    plain = ops.roi_align_multilevel(feats, det, scales, 14, 0, True, out_format="nhwc")
    two_step = ops.channel_attention(plain, vec.view(D, 1, Cc, 1, 1))
    assert fused.shape == (D, Cc, 14, 14)
    assert torch.equal(fused, two_step)
    for b in (0, 9, 15):
        sel = det[:, 0] == b
        rois_b = det[sel].clone()
        rois_b[:, 0] = 0
        one = ops.roi_align_multilevel([f[b:b + 1] for f in feats], rois_b, scales, 14, 0, True, out_format="nhwc")
        assert torch.equal(one, plain[sel])
    lv = ops.map_roi_levels(det, 4)
    assert torch.equal(lv.cpu(), O.map_roi_levels_c(det.cpu(), 4))


def test_empty_proposals():
    from fgn_b200.episodes import CONFIGS, build_heads, episode_to_device, make_episode
    cfg = CONFIGS["tiny_fpn"]
    ep = episode_to_device(make_episode(cfg, seed=0), dev())
    rpn, head = build_heads(cfg, dev())
    with torch.no_grad():
        head.count_spp(ep["spp"][:4], ep["spp_bboxes"], ep["spp_masks"])
        res = head._bbox_forward(ep["qry"][:4], torch.zeros(0, 5, device=dev()))
    assert res["cls_score"].shape == (0, cfg.n_ways + 1) and res["bbox_pred"].shape == (0, 4 * cfg.n_ways)


def test_launches_are_counted():
    from fgn_b200 import ops
    before = ops.launch_count()
    ops.map_roi_levels(torch.rand(10, 5, device=dev()) * 100, 4)
    assert ops.launch_count() == before + 1


# ---- test-time box post-processing between a8 and a9 (BBoxHead.get_bboxes [3P]) ---------------------------------
def _det_case(g, n_per, N, img_h=800, img_w=1344, spread=2.0):
    from fgn_b200.episodes import synth_rois
    B = len(n_per)
    rois = torch.cat([torch.cat([torch.full((n, 1), float(b)), synth_rois(g, n, img_h, img_w, 1)[:, 1:]], 1)
                      for b, n in enumerate(n_per)]) if sum(n_per) else torch.zeros(0, 5)
    R = rois.shape[0]
    cls = torch.randn(R, N + 1, generator=g) * spread
    reg = torch.randn(R, 4 * N, generator=g) * 0.8
    return rois, cls, reg


@pytest.mark.parametrize("n_per,N,rescale", [((300,), 1, False), ((1000,), 1, True), ((300,), 3, True),
                                             ((1000,), 20, False), ((257, 0, 401), 3, True), ((64, 64), 5, False)])
def test_det_postprocess_vs_oracle(n_per, N, rescale):
    """Softmax + delta decode + clip / rescale + multiclass NMS + top-k through the C ABI against the restated
    mmdet semantics: the kept (RoI, class) pairs, their order and labels are identical; boxes and scores within
    the fp32 bar."""
    from fgn_b200 import ops
    g = torch.Generator().manual_seed(900 + N + len(n_per))
    rois, cls, reg = _det_case(g, n_per, N)
    shapes = [(800, 1344, 3)] * len(n_per)
    sfs = [(1.6, 1.5, 1.6, 1.5)] * len(n_per)
    det, lab, cnt = ops.det_postprocess(rois.to(dev()), cls.to(dev()), reg.to(dev()), n_per, shapes,
                                        sfs if rescale else None, score_thr=0.05, iou_thr=0.5, max_per_img=100)
    cnt = cnt.cpu().tolist()
    o = 0
    for b, n in enumerate(n_per):
        if n == 0:                                               # fgn_roi_head.py:596-603: no proposal in this image
            assert cnt[b] == 0
            continue
        wd, wl, _ = O.bbox_head_get_bboxes(rois[o:o + n], cls[o:o + n], reg[o:o + n], shapes[b], sfs[b], rescale,
                                           0.05, 0.5, 100)
        o += n
        assert cnt[b] == wd.shape[0], (b, cnt[b], wd.shape[0])
        assert torch.equal(lab[b, :cnt[b]].cpu().long(), wl)
        close(det[b, :cnt[b]], wd, what=f"image {b}")


def test_det_postprocess_edge_cases():
    from fgn_b200 import ops
    g = torch.Generator().manual_seed(77)
    rois, cls, reg = _det_case(g, (50,), 2)
    d = dev()
    # nothing passes the score threshold
    _, _, cnt = ops.det_postprocess(rois.to(d), cls.to(d), reg.to(d), (50,), [(800, 1344)], None, score_thr=2.0)
    assert cnt.tolist() == [0]
    # no clipping, no proposals in one image, identical boxes (ties -> lower RoI index first, one survivor)
    rois2 = rois.clone(); rois2[1:10, 1:] = rois2[0, 1:]
    cls2 = cls.clone(); cls2[:10] = cls2[0]
    reg2 = reg.clone(); reg2[:10] = reg2[0]
    det, lab, cnt = ops.det_postprocess(rois2.to(d), cls2.to(d), reg2.to(d), (50, 0), None, None, max_per_img=7)
    wd, wl, wf = O.bbox_head_get_bboxes(rois2, cls2, reg2, None, None, False, 0.05, 0.5, 7)
    assert cnt.tolist() == [wd.shape[0], 0] and torch.equal(lab[0, :cnt[0]].cpu().long(), wl)
    close(det[0, :wd.shape[0]], wd, what="ties / no clip")
    # R == 0
    _, _, cnt = ops.det_postprocess(torch.zeros(0, 5, device=d), torch.zeros(0, 3, device=d), torch.zeros(0, 8, device=d), (0,))
    assert cnt.tolist() == [0]


def test_simple_test_bboxes_returns_detections():
    """FGNRoIHead.simple_test_bboxes with test_cfg.rcnn: per-image (det_bboxes [D,5], det_labels [D]) like the
    reference (fgn_roi_head.py:531-616), equal to the oracle's get_bboxes on the head's own raw outputs."""
    from fgn_b200.episodes import CONFIGS, build_heads, episode_to_device, make_episode
    cfg = CONFIGS["cfg2_omniiseg_n3k1_c4"]
    rpn, head = build_heads(cfg, dev(), shared_head=None)
    ep = episode_to_device(make_episode(cfg, seed=1), dev())
    qry = ep["qry"][0]
    head.count_spp(ep["spp"][0], ep["spp_bboxes"].clone(), ep["spp_masks"])
    props = [ep["rois"][:, 1:]]
    metas = [dict(img_shape=(cfg.img_h, cfg.img_w, 3), scale_factor=(1.0, 1.0, 1.0, 1.0))]
    raw_cls, raw_reg = head.simple_test_bboxes(qry, metas, props, None)
    rcnn = dict(score_thr=0.05, nms=dict(type="nms", iou_threshold=0.5), max_per_img=100)
    dets, labels = head.simple_test_bboxes(qry, metas, props, rcnn)
    wd, wl, _ = O.bbox_head_get_bboxes(ep["rois"].cpu(), raw_cls[0].cpu(), raw_reg[0].cpu(), metas[0]["img_shape"],
                                       None, False, 0.05, 0.5, 100)
    assert labels[0].dtype == torch.long and torch.equal(labels[0].cpu(), wl)
    close(dets[0], wd, what="simple_test_bboxes")


def test_simple_test_detections_feed_the_mask_branch():
    """The a8 -> get_bboxes -> a9 chain of simple_test (fgn_roi_head.py:691-719): detections, the per-detection
    support vector gather and the attended mask-branch RoI features equal the oracle's on the same detections."""
    from fgn_b200.episodes import CONFIGS, build_heads, episode_to_device, make_episode
    cfg = CONFIGS["cfg2_omniiseg_n3k1_c4"]
    rpn, head = build_heads(cfg, dev(), shared_head=None)
    ep = episode_to_device(make_episode(cfg, seed=2), dev())
    qry = ep["qry"][0]
    head.count_spp(ep["spp"][0], ep["spp_bboxes"].clone(), ep["spp_masks"])
    metas = [dict(img_shape=(cfg.img_h, cfg.img_w, 3), scale_factor=(1.0, 1.0, 1.0, 1.0))]
    rcnn = dict(score_thr=0.05, nms=dict(type="nms", iou_threshold=0.5), max_per_img=100)
    dets, labels = head.simple_test_bboxes(qry, metas, [ep["rois"][:, 1:]], rcnn)
    assert 0 < dets[0].shape[0] <= 100 and int(labels[0].max()) < cfg.n_ways
    res = head.simple_test_mask(qry, dets, labels, metas, rescale=False)
    mask_rois = torch.cat([dets[0].new_zeros((dets[0].shape[0], 1)), dets[0][:, :4]], 1)
    want = O.mask_attention([qry.cpu().contiguous()], [16], mask_rois.cpu(), head.spp_fvecs_roi_aligned_cat_mean_mp.cpu(),
                            [labels[0].cpu()], cfg.n_ways, 7)
    close(res["mask_feats"], want, what="simple_test mask feats")


# ---- RPN proposals (RPNHead.get_bboxes [3P], fgn.py:229-235) ---------------------------------------------------
@pytest.mark.parametrize("levels,B,nms_pre,max_per_img", [([(32, 32, 16)], 1, 6000, 300), ([(50, 84, 16)], 2, 6000, 300),
                                                           ([(48, 64, 4), (24, 32, 8), (12, 16, 16), (6, 8, 32), (3, 4, 64)], 2, 1000, 1000),
                                                           ([(8, 8, 16)], 1, 6000, 300),
                                                           ([(50, 84, 16)], 1, 12000, 2000)],    # train-time rpn_proposal cfg: library sort path
                         ids=["c4-32x32", "c4-50x84-b2", "fpn5-b2", "tiny", "nms_pre-12000"])
def test_rpn_proposals_vs_oracle(levels, B, nms_pre, max_per_img):
    """Sigmoid + per-level top-k + anchor decode + clip + min-size filter + per-level NMS + top-k through the C ABI
    against the restated mmdet RPNHead._get_bboxes_single: same proposals in the same order."""
    from fgn_b200 import ops
    g = torch.Generator().manual_seed(31 + len(levels) + B)
    scales, ratios = [2, 4, 8, 16, 32] if len(levels) == 1 else [8], [0.5, 1.0, 2.0]
    A = len(scales) * len(ratios)
    cls = [torch.randn(B, A, h, w, generator=g) * 2 for h, w, _ in levels]
    reg = [torch.randn(B, 4 * A, h, w, generator=g) * 0.3 for h, w, _ in levels]
    strides = [s for _, _, s in levels]
    img_h, img_w = levels[0][0] * strides[0] - 5, levels[0][1] * strides[0] - 3
    base = [O.anchor_base(s, scales, ratios) for s in strides]
    assert torch.equal(torch.stack(base), torch.stack([ops.base_anchors(s, scales, ratios) for s in strides]))
    prop, lvl, cnt = ops.rpn_proposals([c.to(dev()) for c in cls], [r.to(dev()) for r in reg], strides,
                                       torch.stack(base), [(img_h, img_w, 3)] * B, nms_pre=nms_pre, iou_thr=0.7,
                                       max_per_img=max_per_img, min_bbox_size=0)
    cnt = cnt.cpu().tolist()
    for b in range(B):
        anchors = [O.anchor_grid(base[l], levels[l][0], levels[l][1], strides[l]) for l in range(len(levels))]
        wd, wid = O.rpn_get_bboxes_single([c[b] for c in cls], [r[b] for r in reg], anchors, (img_h, img_w, 3),
                                          nms_pre, 0.7, max_per_img, 0.0)
        assert cnt[b] == wd.shape[0], (b, cnt[b], wd.shape[0])
        assert torch.equal(lvl[b, :cnt[b]].cpu().long(), wid)
        close(prop[b, :cnt[b]], wd, what=f"proposals image {b}")


def test_agrpn_get_bboxes_feeds_the_roi_head():
    """AGRPNHead.forward_single -> get_bboxes -> FGNRoIHead.simple_test_bboxes: the reference's test-time chain
    (fgn.py:226-240) runs end to end on the device and yields detections."""
    from fgn_b200.episodes import CONFIGS, build_heads, episode_to_device, make_episode
    cfg = CONFIGS["cfg2_omniiseg_n3k1_c4"]
    rpn, head = build_heads(cfg, dev(), shared_head=None)
    ep = episode_to_device(make_episode(cfg, seed=4), dev())
    qry = ep["qry"][0]
    with torch.no_grad():
        cls, reg = rpn.forward_single(qry, ep["spp"][0])
    metas = [dict(img_shape=(cfg.img_h, cfg.img_w, 3), scale_factor=(1.0, 1.0, 1.0, 1.0))]
    rpn_cfg = dict(nms_pre=6000, nms=dict(type="nms", iou_threshold=0.7), max_per_img=300, min_bbox_size=0)
    props = rpn.get_bboxes([cls], [reg], img_metas=metas, cfg=rpn_cfg)
    assert len(props) == 1 and props[0].shape[1] == 5 and 0 < props[0].shape[0] <= 300
    assert (props[0][:-1, 4] >= props[0][1:, 4]).all()
    base = O.anchor_base(16, rpn.anchor_scales, rpn.anchor_ratios)
    wd, _ = O.rpn_get_bboxes_single([cls[0].cpu()], [reg[0].cpu()], [O.anchor_grid(base, cls.shape[2], cls.shape[3], 16)],
                                    metas[0]["img_shape"], 6000, 0.7, 300, 0.0)
    close(props[0], wd, what="AGRPNHead.get_bboxes")
    head.count_spp(ep["spp"][0], ep["spp_bboxes"].clone(), ep["spp_masks"])
    rcnn = dict(score_thr=0.05, nms=dict(type="nms", iou_threshold=0.5), max_per_img=100)
    dets, labels = head.simple_test_bboxes(qry, metas, [p[:, :4] for p in props], rcnn)
    assert dets[0].shape[1] == 5 and dets[0].shape[0] == labels[0].shape[0] <= 100


def test_fgn_detector_simple_test_chain():
    """fgn_b200.FGN.simple_test (fgn.py:186-240) with a toy stride-16 backbone: query image + N*K support images in,
    per-image detections out; every stage after the backbone is the device path, and the result equals running
    the stages by hand (rpn forward -> oracle proposals -> head with those proposals)."""
    import fgn_b200
    from fgn_b200.episodes import CONFIGS, build_heads
    cfg = CONFIGS["cfg2_omniiseg_n3k1_c4"]
    rpn, head = build_heads(cfg, dev(), shared_head=None)
    torch.manual_seed(3)
    backbone = torch.nn.Sequential(torch.nn.Conv2d(3, cfg.channels, 16, stride=16), torch.nn.ReLU()).to(dev())
    test_cfg = dict(rpn=dict(nms_pre=6000, nms=dict(type="nms", iou_threshold=0.7), max_per_img=300, min_bbox_size=0),
                    rcnn=dict(score_thr=0.05, nms=dict(type="nms", iou_threshold=0.5), max_per_img=100))
    det = fgn_b200.FGN(cfg.n_ways, cfg.k_shots, backbone, rpn, head, test_cfg).eval()
    g = torch.Generator().manual_seed(11)
    S = 128
    qry = torch.randn(1, 3, 256, 320, generator=g).to(dev())
    spp = torch.randn(1, cfg.n_ways, cfg.k_shots, 3, S, S, generator=g).to(dev())
    spp_boxes_yxyx = torch.tensor([12.0, 10.0, 110.0, 100.0]).repeat(1, cfg.n_ways, cfg.k_shots, 1).to(dev())
    masks = (torch.rand(1, cfg.n_ways, cfg.k_shots, S, S, generator=g) > 0.5).to(dev())
    dets, labels = det.simple_test(qry, spp, spp_boxes_yxyx, masks, img_shape=[(256, 320, 3)])
    assert len(dets) == 1 and dets[0].shape[1] == 5 and dets[0].shape[0] == labels[0].shape[0] <= 100
    assert labels[0].dtype == torch.long and int(labels[0].max()) < cfg.n_ways
    # by hand, with the oracle's proposals in the middle
    with torch.no_grad():
        qf, sf = backbone(qry), backbone(spp.reshape(-1, 3, S, S))
        cls, reg = rpn.forward_single(qf, sf)
    base = O.anchor_base(16, rpn.anchor_scales, rpn.anchor_ratios)
    props, _ = O.rpn_get_bboxes_single([cls[0].cpu()], [reg[0].cpu()], [O.anchor_grid(base, cls.shape[2], cls.shape[3], 16)],
                                       (256, 320, 3), 6000, 0.7, 300, 0.0)
    head.count_spp(sf, spp_boxes_yxyx[..., [1, 0, 3, 2]].reshape(-1, 1, 4).clone(), masks.reshape(-1, 1, S, S))
    metas = det.get_img_metas([(256, 320, 3)])
    d2, l2 = head.simple_test_bboxes(qf, metas, [props[:, :4].to(dev())], test_cfg["rcnn"])
    assert torch.equal(labels[0], l2[0])
    close(dets[0], d2[0], what="FGN.simple_test vs stages by hand")


# ---- mask pasting + RLE (fgn_mask_paste / fgn_mask_paste_rle) vs the oracle ---------------------------------
def _paste_case(seed, d, h, w, m=28, sharp=3.0):
    rng = np.random.default_rng(seed)
    logits = rng.normal(0, sharp, (d, 1, m, m)).astype(np.float32)
    cx, cy = rng.uniform(0, w, d), rng.uniform(0, h, d)
    bw, bh = rng.uniform(2, w * 0.7, d), rng.uniform(2, h * 0.7, d)
    boxes = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2,  rng.uniform(0, 1, d)], 1).astype(np.float32)
    return logits, boxes


_PASTE_EDGE_BOXES = np.array([[10.3, 5.2, 50.7, 40.1, .9], [0, 0, 83, 61, .9], [-5.5, -3.2, 20.1, 70.3, .9],
                              [30, 30, 30, 45.5, .9], [70.2, 50.1, 82.9, 60.9, .9], [40.5, 10.5, 12.5, 33.0, .9],
                              [100, 100, 120, 130, .9], [-40, -40, -10, -5, .9], [0.5, 0.5, 82.5, 60.5, .9]], np.float32)


def _assert_masks_match(got, values, thr):
    """bool masks equal except where the reference value is within fp32 rounding of the threshold"""
    want = values >= np.float32(thr)
    differ = got != want
    if differ.any():
        assert np.abs(values[differ] - thr).max() < 2e-6, f"{int(differ.sum())} pixels differ away from the threshold"
        assert differ.sum() <= max(2, got.size // 100000)


@pytest.mark.parametrize("case", ["edge", "random", "thr0", "m14"])
def test_mask_paste_dense_and_rle_match_oracle(case):
    from fgn_b200 import ops
    h, w, thr, m = 61, 83, 0.5, 28
    if case == "edge":
        boxes = _PASTE_EDGE_BOXES
        logits = np.random.default_rng(5).normal(0, 3, (len(boxes), 1, m, m)).astype(np.float32)
    elif case == "random":
        h, w = 120, 161
        logits, boxes = _paste_case(11, 40, h, w)
    elif case == "thr0":                                  # threshold 0: every pixel no mask cell reaches is "on"
        thr = 0.0
        logits, boxes = _paste_case(12, 5, h, w)
    else:
        m = 14
        logits, boxes = _paste_case(13, 12, h, w, m=m)
    values = O.paste_values(logits, boxes[:, :4], h, w)
    lt, bt = _t(logits).to(dev()), _t(boxes).to(dev())
    dense = ops.mask_paste(lt, bt, h, w, thr).cpu().numpy()
    _assert_masks_match(dense, values, thr)
    rles, counts = ops.mask_paste_rle(lt, bt, [(h, w)], mask_thr_binary=thr, cap=64, return_counts=True)   # cap grows
    for i in range(len(boxes)):
        # the fused kernel encodes exactly the mask the dense kernel writes (same device functions) ...
        assert counts[i] == O.rle_counts(dense[i]), (case, i)
        assert rles[i]["size"] == [h, w]
        assert rles[i]["counts"] == O.rle_to_string(counts[i])
        # ... and decodes back to it
        assert (O.rle_decode(O.rle_from_string(rles[i]["counts"]), h, w) == dense[i]).all()


def test_mask_paste_rle_full_size_round_trip():
    """cfg3 size (800x1344, 100 detections, two images of different shape in one launch): the RLE decodes to the
    dense kernel's masks, the runs sum to h*w, empty input returns empty."""
    from fgn_b200 import ops
    hw = [(800, 1344), (640, 960)]
    logits, boxes = _paste_case(21, 100, 640, 960)
    img = (np.arange(100) % 2).astype(np.int32)
    lt, bt = _t(logits).to(dev()), _t(boxes).to(dev())
    rles, counts = ops.mask_paste_rle(lt, bt, hw, det_img=_t(img).to(dev()), return_counts=True)
    for k in (0, 1):
        sel = np.flatnonzero(img == k)
        dense = ops.mask_paste(lt[sel], bt[sel], hw[k][0], hw[k][1]).cpu().numpy()
        for j, i in enumerate(sel):
            assert sum(counts[i]) == hw[k][0] * hw[k][1]
            assert rles[i]["size"] == list(hw[k])
            assert counts[i] == O.rle_counts(dense[j])
            assert O.rle_from_string(rles[i]["counts"]) == counts[i]
    assert ops.mask_paste_rle(lt[:0], bt[:0], hw) == []
    # sampled oracle check at full size: three detections against the CPU restatement
    sel = np.array([0, 2, 4])
    values = O.paste_values(logits[sel], boxes[sel, :4], 800, 1344)
    _assert_masks_match(ops.mask_paste(lt[sel], bt[sel], 800, 1344).cpu().numpy(), values, 0.5)


def test_roi_head_get_seg_masks_and_rles():
    """FGNRoIHead.get_seg_masks mirrors FCNMaskHead.get_seg_masks' call (fgn_roi_head.py:668-671) incl. rescale."""
    from fgn_b200 import FGNRoIHead
    head = FGNRoIHead(bbox_roi_extractor=dict(type="SingleRoIExtractor", roi_layer=dict(type="RoIAlign", output_size=7,
                      sampling_ratio=0), out_channels=64, featmap_strides=[16]), shared_head=None, channels=64, n_ways=1, k_shots=1,
                      test_cfg=dict(rcnn=dict(score_thr=0.05, nms=dict(iou_threshold=0.5), max_per_img=100,
                                              mask_thr_binary=0.4))).to(dev())
    logits, boxes = _paste_case(31, 9, 90, 70)
    lt, bt = _t(logits).to(dev()), _t(boxes).to(dev())
    sf = (1.5, 2.0, 1.5, 2.0)
    segms = head.get_seg_masks(lt, bt, None, None, ori_shape=(60, 70, 3), scale_factor=sf, rescale=True)
    rles = head.get_seg_masks(lt, bt, None, None, ori_shape=(60, 70, 3), scale_factor=sf, rescale=True, encode=True)
    values = O.paste_values(logits, boxes[:, :4] / np.array(sf, np.float32), 60, 70)
    _assert_masks_match(np.stack(segms[0]), values, 0.4)
    assert [r["counts"] for r in rles[0]] == [r["counts"] for r in O.encode_mask_results(np.stack(segms[0]))]
    segms2 = head.get_seg_masks(lt, bt, None, None, ori_shape=(60, 70, 3), scale_factor=sf, rescale=False)
    assert segms2[0][0].shape == (120, 105)              # round(ori * scale): (60*2.0, 70*1.5)


def test_fgn_detector_chain_to_result_dict():
    """FGN.simple_test with a mask head, down to the per-image result dict of fgn.py:262-303: detections, the RLE of
    every pasted mask (get_seg_masks + encode_mask_results in one kernel) and the YXYX boxes."""
    import fgn_b200
    from fgn_b200.episodes import CONFIGS, build_heads
    cfg = CONFIGS["cfg2_omniiseg_n3k1_c4"]
    rpn, head = build_heads(cfg, dev(), shared_head=None)
    torch.manual_seed(5)
    head.mask_head = torch.nn.ConvTranspose2d(cfg.channels, 1, 4, stride=4).to(dev())      # toy 7x7 -> 28x28 mask head
    backbone = torch.nn.Sequential(torch.nn.Conv2d(3, cfg.channels, 16, stride=16), torch.nn.ReLU()).to(dev())
    test_cfg = dict(rpn=dict(nms_pre=6000, nms=dict(type="nms", iou_threshold=0.7), max_per_img=300, min_bbox_size=0),
                    rcnn=dict(score_thr=0.05, nms=dict(type="nms", iou_threshold=0.5), max_per_img=100, mask_thr_binary=0.5))
    head.test_cfg = test_cfg["rcnn"]
    det = fgn_b200.FGN(cfg.n_ways, cfg.k_shots, backbone, rpn, head, test_cfg).eval()
    g = torch.Generator().manual_seed(12)
    S, H, W = 128, 256, 320
    qry = torch.randn(1, 3, H, W, generator=g).to(dev())
    spp = torch.randn(1, cfg.n_ways, cfg.k_shots, 3, S, S, generator=g).to(dev())
    spp_boxes_yxyx = torch.tensor([12.0, 10.0, 110.0, 100.0]).repeat(1, cfg.n_ways, cfg.k_shots, 1).to(dev())
    masks = (torch.rand(1, cfg.n_ways, cfg.k_shots, S, S, generator=g) > 0.5).to(dev())
    out = det.simple_test(qry, spp, spp_boxes_yxyx, masks, img_shape=[(H, W, 3)])
    dets, labels, res = out
    d = dets[0].shape[0]
    assert d > 0 and res["mask_pred"].shape == (d, 1, 28, 28) and len(res["segm_rles"]) == 1 and len(res["segm_rles"][0]) == d
    values = O.paste_values(res["mask_pred"].cpu().numpy(), dets[0][:, :4].cpu().numpy(), H, W)
    got = np.stack([O.rle_decode(O.rle_from_string(r["counts"]), H, W) for r in res["segm_rles"][0]])
    _assert_masks_match(got, values, 0.5)
    assert all(r["size"] == [H, W] for r in res["segm_rles"][0])
    fmt = det.format_results(out)
    assert set(fmt[0]) == {"dt_scores", "dt_bboxes", "dt_cat_ids", "dt_isegmaps_rle"}
    db = dets[0].cpu().numpy()
    assert np.array_equal(fmt[0]["dt_bboxes"], db[:, [1, 0, 3, 2]]) and np.array_equal(fmt[0]["dt_scores"], db[:, 4])
    assert np.array_equal(fmt[0]["dt_cat_ids"], labels[0].cpu().numpy()) and fmt[0]["dt_isegmaps_rle"] is res["segm_rles"][0]


# ---- AG-RPN attention folded into the RPN conv's weights (fgn_fold_attention_weights) ---------------------------
def test_fold_attention_weights_is_the_broadcast_product():
    from fgn_b200 import ops
    g = torch.Generator().manual_seed(4)
    w = torch.randn(24, 16, 3, 3, generator=g)
    v = torch.randn(2, 3, 16, 1, 1, generator=g)
    got = ops.fold_attention_weights(w.to(dev()), v.to(dev())).cpu()
    want = w[None] * v.reshape(6, 1, 16, 1, 1)
    assert got.shape == (6, 24, 16, 3, 3) and torch.equal(got, want)
    assert ops.fold_attention_weights(w.to(dev()), v[:0].to(dev())).shape == (0, 24, 16, 3, 3)


@pytest.mark.parametrize("batch,n_ways,k_shots", [(1, 1, 1), (2, 3, 2), (1, 5, 1)])
def test_agrpn_forward_single_with_folded_attention(batch, n_ways, k_shots):
    """conv(q * v) == conv(q, W * v): AGRPNHead.forward_single with fold_attention=True (qry_fmap_mod never
    materialised, one grouped conv) against the reference's order of operations on the same weights, and against the
    CPU restatement (fgn_ag_rpn_head.py:33-48 + RPNHead convs)."""
    from fgn_b200 import AGRPNHead
    c, h, w, s = 32, 20, 28, 8
    torch.manual_seed(21)
    ref = AGRPNHead(in_channels=c, feat_channels=c, n_ways=n_ways, k_shots=k_shots).to(dev()).eval()
    fold = AGRPNHead(in_channels=c, feat_channels=c, n_ways=n_ways, k_shots=k_shots,
                     fold_attention="auto" if n_ways == 5 else True).to(dev()).eval()       # 20x28 map, Cf=32: auto folds
    assert fold.fold_attention in (True, "auto")
    fold.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(22)
    q = torch.randn(batch, c, h, w, generator=g)
    sp = torch.randn(batch * n_ways * k_shots, c, s, s, generator=g).abs()
    with torch.no_grad():
        a = ref.forward_single(q.to(dev()), sp.to(dev()))
        b = fold.forward_single(q.to(dev()), sp.to(dev()))
        _, qmod = O.agrpn_attention(q, sp, n_ways, k_shots)
        cpu = ref.to("cpu")
        x = torch.relu(cpu.rpn_conv(qmod))
        cls_cpu, reg_cpu = cpu.rpn_cls(x), cpu.rpn_reg(x)
        sure = torch.ones(batch, cls_cpu.shape[1], h, w, dtype=torch.bool)
        if n_ways > 1:
            # anchors whose two best classes score within rounding of each other may pick either class
            top2 = cls_cpu.view(batch, n_ways, -1, h, w).topk(2, dim=1).values
            sure = (top2[:, 0] - top2[:, 1]) > 1e-4
            cls_cpu, reg_cpu = O.best_class_selection(cls_cpu, reg_cpu, batch, n_ways)
        else:
            cls_cpu, reg_cpu = cls_cpu.reshape(batch, -1, h, w), reg_cpu.reshape(batch, -1, h, w)
    assert sure.float().mean() > 0.99
    sure4 = sure.repeat_interleave(4, dim=1)
    for got, want, cpuw, m, what in ((b[0], a[0], cls_cpu, sure, "cls"), (b[1], a[1], reg_cpu, sure4, "reg")):
        close(got.cpu()[m], want.cpu()[m], what=f"folded vs materialised {what}")
        close(got.cpu()[m], cpuw[m], what=f"folded vs CPU restatement {what}")


def test_mask_paste_rle_against_the_golden_fixture():
    """tests/golden/mask_paste_kat.npz: masks pasted by torch's own grid_sample, run lengths counted with numpy.  The
    CUDA path reproduces the run lengths wherever no pixel value lies within rounding of the threshold."""
    from fgn_b200 import ops
    z = np.load(os.path.join(GOLDEN, "mask_paste_kat.npz"))
    h, w = [int(v) for v in z["img_hw"]]
    thr = float(z["thr"])
    lens, flat = z["counts_len"].tolist(), z["counts_flat"].tolist()
    lt, bt = _t(z["logits"]).to(dev()), _t(z["boxes"]).to(dev())
    dense = ops.mask_paste(lt, bt, h, w, thr).cpu().numpy()
    _assert_masks_match(dense, z["values"], thr)
    rles, counts = ops.mask_paste_rle(lt, bt, [(h, w)], mask_thr_binary=thr, return_counts=True)
    k = 0
    for i, n in enumerate(lens):
        want = flat[k:k + n]
        k += n
        if np.array_equal(dense[i], z["masks"][i]):            # (no near-threshold pixel flipped in this mask)
            assert counts[i] == want, i
            assert O.rle_from_string(rles[i]["counts"]) == want
    assert sum(np.array_equal(dense[i], z["masks"][i]) for i in range(len(lens))) >= len(lens) - 1


def test_conv1x1_and_bottleneck_on_the_contraction_kernel():
    """SURVEY 8f row 3, first piece: the res5 bottleneck's 1x1 convolutions (BatchNorm folded, ReLU / identity branch in
    the epilogue) on the tcgen05 contraction against the plain torch modules, and ops.conv1x1 against F.conv2d."""
    from fgn_b200 import ops
    from fgn_b200.roi_head import _Bottleneck
    g = torch.Generator().manual_seed(808)
    x = torch.randn(60, 256, 7, 7, generator=g)
    w = torch.randn(128, 256, generator=g) / 16
    b = torch.randn(128, generator=g)
    res = torch.randn(60, 128, 7, 7, generator=g)
    want = torch.relu(torch.nn.functional.conv2d(x, w.view(128, 256, 1, 1), b) + res)
    got = ops.conv1x1(x.to(dev()), w.to(dev()), b.to(dev()), residual=res.to(dev()), relu=True)
    assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)
    close(got, want, atol=2e-4, what="conv1x1 + residual + relu")
    close(ops.conv1x1(x.to(dev()).contiguous(memory_format=torch.channels_last), w.to(dev())),
          torch.nn.functional.conv2d(x, w.view(128, 256, 1, 1)), atol=2e-4, what="conv1x1 plain")
    blk = _Bottleneck(256, 128)
    with torch.no_grad():
        for bn in (blk.bn1, blk.bn2, blk.bn3):                       # non-trivial running statistics and affine
            bn.running_mean.copy_(torch.randn(bn.num_features, generator=g) * 0.2)
            bn.running_var.copy_(torch.rand(bn.num_features, generator=g) + 0.5)
            bn.weight.copy_(1 + 0.2 * torch.randn(bn.num_features, generator=g))
            bn.bias.copy_(0.1 * torch.randn(bn.num_features, generator=g))
    blk = blk.to(dev()).eval()
    xd = x.to(dev())
    with torch.no_grad():
        from fgn_b200 import _lib
        before = _lib.load().fgn_launch_count()
        fast = blk(xd)
        assert _lib.load().fgn_launch_count() > before             # the library's kernels ran
        blk.tc_1x1 = False
        ref = blk(xd)
    close(fast, ref, atol=5e-4, rtol=1e-4, what="bottleneck: tcgen05 1x1 convs vs torch modules")
