"""Host-side logic that needs no GPU: layout detection, episode sharding, the N>1 gather path on
world_size-2 gloo, synthetic generators, module construction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fgn_b200 import episodes as E
from fgn_b200 import ops
from fgn_b200._lib import LAYOUT_NCHW, LAYOUT_NHWC
from oracle import fgn_oracle as O


def test_storage_layout_detection():
    x = torch.zeros(2, 8, 5, 6)
    assert ops.storage_layout(x) == LAYOUT_NCHW
    assert ops.storage_layout(x.contiguous(memory_format=torch.channels_last)) == LAYOUT_NHWC
    assert ops.storage_layout(x[:, ::2]) is None
    assert ops.storage_layout(torch.zeros(3, 7, 7, 8).permute(0, 3, 1, 2)) == LAYOUT_NHWC
    # ambiguous shapes resolve to NCHW (both descriptions are valid)
    assert ops.storage_layout(torch.zeros(2, 1, 5, 6)) == LAYOUT_NCHW


def test_shard_range_partitions_contiguously():
    for n in (0, 1, 7, 8, 64, 1000):
        for w in (1, 2, 4, 8):
            blocks = [E.shard_range(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c and b - a >= d - c >= 0
            assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1


def test_bbox2roi_matches_oracle():
    from fgn_b200 import bbox2roi
    g = torch.Generator().manual_seed(0)
    boxes = [torch.rand(5, 5, generator=g), torch.zeros(0, 5), torch.rand(3, 4, generator=g)]
    assert torch.equal(bbox2roi(boxes), O.bbox2roi(boxes))


def test_synthetic_episode_shapes():
    cfg = E.CONFIGS["cfg3_coco2voc_n1k1_fpn"]
    assert [E.level_hw(cfg.img_h, cfg.img_w, s) for s in cfg.rpn_strides] == [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    ep = E.make_episode(E.CONFIGS["tiny_fpn"], seed=1)
    cfg = ep["cfg"]
    assert len(ep["qry"]) == 5 and ep["qry"][0].shape == (1, 64, 32, 48)
    assert ep["spp"][0].shape == (cfg.n_ways * cfg.k_shots, 64, 16, 16)
    assert ep["spp_masks"].dtype == torch.bool and ep["spp_bboxes"].shape == (6, 1, 4)
    r = ep["rois"]
    assert r.shape == (96, 5) and (r[:, 3] >= r[:, 1]).all() and (r[:, 1] >= 0).all() and (r[:, 3] <= cfg.img_w).all()
    ep2 = E.make_episode(E.CONFIGS["tiny_fpn"], seed=1)
    assert torch.equal(ep["rois"], ep2["rois"]) and torch.equal(ep["qry"][2], ep2["qry"][2])
    # every FPN level receives RoIs under the stated size distribution
    big = E.synth_rois(torch.Generator().manual_seed(1), 1000, 800, 1344, 1)
    lv = O.map_roi_levels(big, 4)
    assert all((lv == l).sum() > 20 for l in range(4))


def test_modules_construct_with_reference_kwargs():
    from fgn_b200 import AGRPNHead, FGNRoIHead, SingleRoIExtractor
    rpn = AGRPNHead(in_channels=64, feat_channels=64,
                    anchor_generator=dict(type="AnchorGenerator", scales=[2, 4, 8, 16, 32], ratios=[0.5, 1.0, 2.0], strides=[16]),
                    loss_cls=dict(type="CrossEntropyLoss", use_sigmoid=True, loss_weight=1.0))
    assert rpn.num_anchors == 15 and rpn.rpn_cls.out_channels == 15 and rpn.rpn_reg.out_channels == 60
    head = FGNRoIHead(bbox_roi_extractor=dict(type="SingleRoIExtractor",
                                              roi_layer=dict(type="RoIAlign", output_size=7, sampling_ratio=0),
                                              out_channels=64, featmap_strides=[16]),
                      bbox_head=dict(type="FGNBBoxHead", with_avg_pool=True, roi_feat_size=7, in_channels=64, num_classes=1,
                                     reg_class_agnostic=False), channels=64, shared_head="c4")
    assert head.bbox_roi_extractor.num_inputs == 1 and head.with_shared_head and head.share_roi_extractor
    assert head.cls_reg_shared_conv.weight.shape == (64, 128, 1, 1)
    assert head.cls_reg_shared_conv_norm.num_groups == 32
    assert len(head.shared_head) == 3
    ext = SingleRoIExtractor(dict(type="RoIAlign", output_size=7, sampling_ratio=0), 256, [4, 8, 16, 32])
    assert ext.num_inputs == 4 and ext.finest_scale == 56 and ext.roi_layers[2].spatial_scale == 1 / 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_episodes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = E.shard_range(num_episodes, world, rank)
    # per-episode "result" is a deterministic function of the global episode id
    local = torch.stack([torch.full((3, 2), float(e)) + torch.arange(6).view(3, 2) for e in range(lo, hi)]) \
        if hi > lo else torch.zeros(0, 3, 2)
    full = E.gather_results(local, num_episodes)
    q.put((rank, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("num_episodes", [5, 8])
def test_gather_results_world_size_2_gloo(num_episodes):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_episodes, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([np.full((3, 2), float(e)) + np.arange(6).reshape(3, 2) for e in range(num_episodes)])
    for r in range(2):
        assert np.array_equal(got[r], want)      # world-size-2 result == unsharded result, on every rank


def _gatherer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = E.ResultGatherer((3, 4, 2), "cpu", depth=2)
    outs = []
    for k in range(5):                                     # five steps through two rotating buffers
        buf = g.local(k)
        buf.copy_(torch.arange(24, dtype=torch.float32).view(3, 4, 2) + 100 * rank + 1000 * k)
        g.submit(k)
        outs.append(g.result(k).clone())
    g.drain()
    q.put((rank, torch.stack(outs).numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_result_gatherer_world_size_2_gloo():
    """The overlapped gatherer (double-buffered, preallocated, no pad / cat for equal blocks) returns rank-major
    blocks on every rank, step after step."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gatherer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    base = np.arange(24, dtype=np.float32).reshape(3, 4, 2)
    want = np.stack([np.concatenate([base + 100 * r + 1000 * k for r in range(2)]) for k in range(5)])
    for r in range(2):
        assert np.array_equal(got[r], want)


def test_gather_results_force_in_a_one_rank_group():
    """world size 1 with the collective forced (the harness-overhead measurement of SURVEY 8e) is the identity."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_force_worker, args=(_free_port(), q))
    p.start()
    out = q.get(timeout=120)
    p.join(timeout=60)
    assert p.exitcode == 0
    assert np.array_equal(out, np.arange(12, dtype=np.float32).reshape(2, 3, 2))


def _force_worker(port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=0, world_size=1)
    x = torch.arange(12, dtype=torch.float32).view(2, 3, 2)
    q.put(E.gather_results(x, 2, force=True).numpy())
    dist.destroy_process_group()


def test_chunked_result_writer_matches_the_eval_hook_layout(tmp_path):
    """main.py:285-309: results are dumped as ResultsChunked/NN.pkl every 1000 items and once more at the end."""
    import pickle
    from fgn_b200.detector import ChunkedResultWriter
    w = ChunkedResultWriter(str(tmp_path), chunk=4)
    for i in range(10):
        w.add([dict(qry_img_id=i, dt_scores=np.arange(i, dtype=np.float32))])       # one simple_test call = a list
    paths = w.close()
    assert [os.path.basename(p) for p in paths] == ["00.pkl", "01.pkl", "02.pkl"]
    assert os.path.dirname(paths[0]).endswith("ResultsChunked")
    got = [r for p in paths for r in pickle.load(open(p, "rb"))]
    assert [len(pickle.load(open(p, "rb"))) for p in paths] == [4, 4, 2]
    assert [r["qry_img_id"] for r in got] == list(range(10))
    assert w.close() == paths                                                       # nothing left to flush


def test_roi_head_builds_the_configs_mask_head():
    """fgn_r50_c4_densecl.py:115-129: mask_head given as a config dict becomes the package's FCNMaskHead with mmdet's
    module names (a reference checkpoint's keys load)."""
    from fgn_b200 import FCNMaskHead, FGNRoIHead
    head = FGNRoIHead(shared_head=None, channels=64,
                      mask_head=dict(type="FCNMaskHead", init_cfg=None, num_convs=4, in_channels=64, conv_out_channels=32,
                                     num_classes=1, class_agnostic=True,
                                     loss_mask=dict(type="CrossEntropyLoss", use_mask=True, loss_weight=1.0)))
    assert isinstance(head.mask_head, FCNMaskHead) and head.with_mask
    keys = set(head.mask_head.state_dict())
    assert {"convs.0.conv.weight", "convs.3.conv.bias", "upsample.weight", "conv_logits.weight"} <= keys
    assert head.mask_head.conv_logits.out_channels == 1 and head.mask_head.convs[0].conv.in_channels == 64
    import torch
    x = torch.randn(2, 64, 14, 14)
    assert head.mask_head(x).shape == (2, 1, 28, 28)                     # CPU tensors: the plain torch modules


def test_count_spp_class_term_is_bound_to_its_class_maps_and_weights():
    """The class half of the relation conv that the one-launch support branch leaves on the head is used only while it
    belongs to the class maps on ``self`` and to the packed weights about to be used (host logic, no device work)."""
    import torch
    from fgn_b200 import FGNRoIHead
    head = FGNRoIHead(shared_head=None, channels=64)
    assert head.fused_prologue is True and head.fused_prologue_max_supports == 32
    cat, params, other = torch.zeros(1, 1, 64, 7, 7), object(), object()
    term = torch.zeros(49, 64)
    head.spp_fmaps_roi_aligned_cat_mean = cat
    head._class_term = (term, cat, params)
    assert head._valid_class_term(params) is term
    assert head._valid_class_term(other) is None                         # re-packed weights
    head.spp_fmaps_roi_aligned_cat_mean = torch.zeros_like(cat)           # class maps assigned by hand since
    assert head._valid_class_term(params) is None
    head._class_term = None
    assert head._valid_class_term(params) is None


def test_batch_episodes_is_the_references_image_batch():
    """episodes.batch_episodes: B episodes as one call -- maps concatenated along the image axis, support sets in image
    order, the image index in column 0 of the RoIs (bbox2roi's layout), per-image labels kept as a list."""
    import torch
    from fgn_b200.episodes import CONFIGS, batch_episodes, make_episode
    cfg = CONFIGS["tiny_fpn"]
    eps = [make_episode(cfg, seed=i) for i in range(3)]
    b = batch_episodes(eps)
    m = cfg.n_ways * cfg.k_shots
    assert b["cfg"].batch == 3 and b["cfg"].num_rois == cfg.num_rois
    assert b["qry"][0].shape[0] == 3 and b["spp"][0].shape[0] == 3 * m
    assert b["qry"][0].is_contiguous(memory_format=torch.channels_last)
    assert tuple(b["spp_bboxes"].shape) == (3 * m, 1, 4) and b["spp_masks"].shape[0] == 3 * m
    r = cfg.num_rois
    for i in range(3):
        assert torch.equal(b["qry"][1][i], eps[i]["qry"][1][0])
        assert torch.equal(b["spp"][0][i * m:(i + 1) * m], eps[i]["spp"][0])
        rows = b["rois"][i * r:(i + 1) * r]
        assert bool((rows[:, 0] == i).all()) and torch.equal(rows[:, 1:], eps[i]["rois"][:, 1:])
        assert torch.equal(b["det_labels_list"][i], eps[i]["det_labels_list"][0])
    assert bool((eps[1]["rois"][:, 0] == 0).all())                        # the inputs are left alone
