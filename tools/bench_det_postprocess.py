"""Timing of fgn_det_postprocess (softmax + decode + class-aware NMS + top-k) next to the CPU restatement
(torch softmax/decode + torchvision CPU nms) on the shapes of BASELINE.json's configs.  Reporting tool."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200 import ops
from fgn_b200.episodes import synth_rois
from oracle import fgn_oracle as O

dev = torch.device("cuda:0")
for name, n_per, N in (("cfg2 (R=300, N=3)", (300,), 3), ("cfg3 (R=1000, N=1)", (1000,), 1),
                       ("cfg4 (R=1000, N=20)", (1000,), 20), ("cfg5 (16 img x 512, N=1)", (512,) * 16, 1)):
    g = torch.Generator().manual_seed(3)
    rois = torch.cat([torch.cat([torch.full((n, 1), float(b)), synth_rois(g, n, 800, 1344, 1)[:, 1:]], 1)
                      for b, n in enumerate(n_per)])
    R = rois.shape[0]
    cls = torch.randn(R, N + 1, generator=g) * 2
    reg = torch.randn(R, 4 * N, generator=g) * 0.5
    shapes = [(800, 1344, 3)] * len(n_per)
    rd, cd, gd = rois.to(dev), cls.to(dev), reg.to(dev)
    for _ in range(3):
        det, lab, cnt = ops.det_postprocess(rd, cd, gd, n_per, shapes)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        det, lab, cnt = ops.det_postprocess(rd, cd, gd, n_per, shapes)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 50
    t0 = time.perf_counter()
    o = 0
    for b, n in enumerate(n_per):
        O.bbox_head_get_bboxes(rois[o:o + n], cls[o:o + n], reg[o:o + n], shapes[b], None, False)
        o += n
    cpu_us = (time.perf_counter() - t0) * 1e6
    print(json.dumps({"case": name, "gpu_us_per_call": round(us, 1), "kept": cnt.tolist()[:4],
                      "cpu_oracle_us": round(cpu_us, 1), "cpu_threads": torch.get_num_threads()}), flush=True)

# ---- RPN proposals (fgn_rpn_proposals) on the C4 and FPN shapes ---------------------------------------------
for name, levels, B, scales, nms_pre, maxp in (("cfg2 C4 32x32, A=15", [(32, 32, 16)], 1, [2, 4, 8, 16, 32], 6000, 300),
                                               ("cfg3 C4-exact 50x84, A=15", [(50, 84, 16)], 1, [2, 4, 8, 16, 32], 6000, 300),
                                               ("cfg3 FPN P2-P6, A=3", [(200, 336, 4), (100, 168, 8), (50, 84, 16), (25, 42, 32), (13, 21, 64)], 1, [8], 1000, 1000)):
    g = torch.Generator().manual_seed(5)
    ratios = [0.5, 1.0, 2.0]
    A = len(scales) * len(ratios)
    cls = [torch.randn(B, A, h, w, generator=g) * 2 for h, w, _ in levels]
    reg = [torch.randn(B, 4 * A, h, w, generator=g) * 0.3 for h, w, _ in levels]
    strides = [s for _, _, s in levels]
    base = torch.stack([ops.base_anchors(s, scales, ratios) for s in strides])
    shape = (levels[0][0] * strides[0], levels[0][1] * strides[0], 3)
    cd, rd = [c.to(dev) for c in cls], [r.to(dev) for r in reg]
    kw = dict(nms_pre=nms_pre, iou_thr=0.7, max_per_img=maxp, min_bbox_size=0)
    for _ in range(3):
        prop, lvl, cnt = ops.rpn_proposals(cd, rd, strides, base, [shape] * B, **kw)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        prop, lvl, cnt = ops.rpn_proposals(cd, rd, strides, base, [shape] * B, **kw)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 30
    t0 = time.perf_counter()
    anchors = [O.anchor_grid(base[l], levels[l][0], levels[l][1], strides[l]) for l in range(len(levels))]
    O.rpn_get_bboxes_single([c[0] for c in cls], [r[0] for r in reg], anchors, shape, nms_pre, 0.7, maxp, 0.0)
    cpu_us = (time.perf_counter() - t0) * 1e6
    print(json.dumps({"case": "rpn " + name, "gpu_us_per_call": round(us, 1), "proposals": cnt.tolist(),
                      "cpu_oracle_us": round(cpu_us, 1), "cpu_threads": torch.get_num_threads()}), flush=True)
