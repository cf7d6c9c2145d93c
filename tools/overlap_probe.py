"""Do two kernels of the path overlap when issued on two streams?  Times N launches of A on one stream and N of B on another
(issued together) against the same launches back to back on one stream.  Development tool (FGN_ATT_LEAN / FGN_ATT_GRID)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200 import ops
from fgn_b200.episodes import CONFIGS, build_heads, episode_to_device, make_episode
cfg = CONFIGS["cfg3_coco2voc_n1k1_fpn"]
dev = torch.device("cuda:0")
eps = [episode_to_device(make_episode(cfg, seed=i), dev) for i in range(4)]
rpn, head = build_heads(cfg, dev, seed=0, shared_head=None)
n_ext = len(cfg.strides)
scales = [1.0 / s for s in cfg.strides]
vec = torch.rand(len(cfg.rpn_strides), cfg.n_ways, cfg.channels, device=dev) + 0.5
a_q = torch.randn(49000, 256, device=dev)
w_q = torch.randn(256, 256, device=dev) / 16
w_split = ops.conv_split_weights(w_q[None])


def roi(i):
    ep = eps[i % 4]
    ops.roi_align_multilevel(ep["qry"][:n_ext], ep["rois"], scales, 7, 0, True, out_format="nhwc")


def att(i):
    ep = eps[(i + 2) % 4]
    rpn.attention_multilevel(ep["qry"], ep["spp"]) if hasattr(rpn, "attention_multilevel") else None


def gemm(i):
    ops.gemm_nt(a_q, w_q, None, "fp32", b_split=w_split)


def graph_of(f, n=4):
    """n launches of f captured once: replay has no host work, no allocation"""
    f(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            f(i)
    return g


def timed(ga, gb, n=10, two_streams=True, per=4):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for g in (ga, gb):
        if g is not None:
            g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for i in range(n):
        if ga is not None:
            with torch.cuda.stream(s1):
                ga.replay()
        if gb is not None:
            with torch.cuda.stream(s2 if two_streams else s1):
                gb.replay()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * per)


with torch.no_grad():
    graphs = {"roi_align": graph_of(roi), "attention": graph_of(att), "contraction": graph_of(gemm)}
    for a, b in (("roi_align", "attention"), ("contraction", "attention"), ("roi_align", "contraction")):
        ga, gb = graphs[a], graphs[b]
        print(json.dumps({"pair": a + " + " + b, "a_alone_us": round(timed(ga, None), 1), "b_alone_us": round(timed(None, gb), 1),
                          "same_stream_us": round(timed(ga, gb, two_streams=False), 1), "two_streams_us": round(timed(ga, gb), 1),
                          "lean": os.environ.get("FGN_ATT_LEAN", "2"), "att_grid": os.environ.get("FGN_ATT_GRID", "0")}), flush=True)
