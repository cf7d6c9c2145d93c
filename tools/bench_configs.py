"""Throughput of the whole hot path for every BASELINE.json config (graphs + multi-stream replay).
Development / reporting tool: prints one JSON line per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200.episodes import CONFIGS, EpisodeRunner, build_heads, episode_to_device, make_episode

dev = torch.device("cuda:0")
FOLD = os.environ.get("BENCH_FOLD") == "1"     # AG-RPN attention folded into the RPN conv's weights instead of materialised
names = sys.argv[1:] or ["cfg1_mnistiseg_n1k1_c4", "cfg2_omniiseg_n3k1_c4", "cfg3_coco2voc_n1k1_fpn",
                         "cfg4_coco2voc_n20k5_fpn", "cfg5_coco2voc_mask_fpn", "cfg3_c4_exact"]
for name in names:
    cfg = CONFIGS[name]
    E = 2 if cfg.batch > 1 or cfg.n_ways > 8 else 8
    eps = [episode_to_device(make_episode(cfg, seed=i), dev) for i in range(min(E, 2))]
    eps = [eps[i % len(eps)] for i in range(E)]
    # C4 mode: by default the res5 shared_head is left out so the line measures the guided path's own kernels;
    # BENCH_SHARED_HEAD=tc / cudnn puts the reference's real res5 head in (the library's tcgen05 convolutions / the plain
    # torch modules on cuDNN, both under torch's default allow_tf32) -- the reference's actual C4 mode end to end
    sh = os.environ.get("BENCH_SHARED_HEAD", "")
    rpn, head = build_heads(cfg, dev, shared_head="c4" if (sh and cfg.mode == "c4") else None)
    if sh == "cudnn" and head.with_shared_head:
        for m in head.shared_head:
            m.tc_1x1, m.tc_3x3 = False, False
    streams = min(E, 8)
    runner = EpisodeRunner(rpn, head, eps, use_graphs=True, n_streams=streams, with_attention="fold" if FOLD else True)

    def step():
        runner.begin()
        for i in range(E):
            runner.run(i)
        runner.end()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rois = E * cfg.num_rois * cfg.batch
    print(json.dumps({"config": name, "shared_head": (sh if head.with_shared_head else "none"), "attention": "folded into rpn_conv weights" if FOLD else "materialised", "mode": cfg.mode, "N": cfg.n_ways, "K": cfg.k_shots, "C": cfg.channels,
                      "R_per_call": cfg.num_rois * cfg.batch, "mask_P": cfg.mask_size, "episodes_per_step": E,
                      "streams": streams, "us_per_episode_call": round(ms * 1e3 / E, 1),
                      "RoIs_per_s": round(rois / ms * 1e3), "launches_per_episode": runner.launches_per_episode[0]}), flush=True)
    del runner, eps, rpn, head
    torch.cuda.empty_cache()
