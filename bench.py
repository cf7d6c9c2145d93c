#!/usr/bin/env python
"""bench.py -- FGN guided RoIAlign + support-guided fusion throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # reference CPU path (oracle port)

Metric (BASELINE.json): RoIs/s of the guided RoIAlign + fusion hot path (episodes/s reported beside
it).  Workload at every N: cfg3 of BASELINE.json -- COCO2VOC 1-way 1-shot, R50-FPN pyramid, 256
channels, query 800x1344, 1000 proposals per image, random support masks -- the configuration the
metric is quoted on (it fits one GPU).  One *step* = one pass of the whole hot path (AG-RPN
attention on P2-P6, support vectors, guided multi-level RoIAlign + relation fusion + heads, mask
branch RoIAlign with AG-FCN attention) over a block of E distinct episodes per GPU.  Episodes are
sharded over ranks (weak scaling: E per rank fixed), the only collective is the NCCL all_gather of
the per-episode results.

`value`  : inputs resident in HBM (channels_last) when the timed region starts.
`e2e`    : same path through the module API from pinned HOST buffers in the reference's NCHW
           layout; H2D of every input, device repack, D2H of cls/bbox results inside the timed region.
`roofline`: the multi-level RoIAlign kernel, timed alone with CUDA events on its launch stream; `roofline_tensor`: the
           relation contraction the same way; `roofline_step`: the whole step's algorithmic bytes (fused accounting)
           over ms_per_step.  `traffic` / tensor-pipe figures are read from the committed profiles/*.json they came from.
`parity`  : episode 0 through the product path against the CPU oracle on a strided RoI subset, BEFORE any timing.
`sustained`: the same step looped for seconds, with its clock record.  `strong_scaling`, `gather_overhead_w1`: SURVEY 8e.
`cpu_baseline`: the oracle (torchvision CPU roi_align + torch CPU fusion, materialised concat form)
           on this box's host cores, rank 0, N=1 only, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = "cfg3_coco2voc_n1k1_fpn"


# ---- clocks ------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU during the timed region (NVML, 5 ms period)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    _NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
              0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
              0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def _once(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for b, n in self._NAMES.items():
            if bits & b and n != "gpu_idle":
                self.reasons.add(n)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._once()
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- algorithmic bytes ----------------------------------------------------------------------------------
def roi_align_algorithmic_bytes(cfg, rois: torch.Tensor, P: int = 7) -> int:
    """sum_l min(H_l W_l, sum_{roi in l} (ceil(w)+1)(ceil(h)+1)) * C*4 + R*20 + out bytes."""
    from oracle import fgn_oracle as O
    from fgn_b200.episodes import level_hw
    lv = O.map_roi_levels_c(rois, len(cfg.strides))
    total = 0
    for l, s in enumerate(cfg.strides):
        h, w = level_hw(cfg.img_h, cfg.img_w, s)
        r = rois[lv == l]
        cells = ((torch.ceil((r[:, 3] - r[:, 1]) / s) + 1) * (torch.ceil((r[:, 4] - r[:, 2]) / s) + 1)).sum().item()
        total += min(h * w * cfg.batch, int(cells)) * cfg.channels * 4
    return int(total + rois.shape[0] * 20 + rois.shape[0] * cfg.channels * P * P * 4)


# ---- the CPU reference path (oracle) --------------------------------------------------------------------
def cpu_reference_episode(ep, w):
    """The reference's path on CPU for one episode: torchvision CPU roi_align (stand-in for mmcv's CPU
    RoIAlign: same algorithm) + torch CPU fusion exactly as fgn_roi_head.py:253-279,419-449 and
    fgn_ag_rpn_head.py:37-46 (materialised concat form)."""
    from oracle import fgn_oracle as O
    cfg = ep["cfg"]
    n_ext = len(cfg.strides)
    for q, s in zip(ep["qry"], ep["spp"]):
        O.agrpn_attention(q, s, cfg.n_ways, cfg.k_shots)
    if cfg.mode == "fpn":
        cat_mean, mp, _, _ = O.count_spp_fpn(ep["spp"][:n_ext], cfg.strides, ep["spp_bboxes"].clone(), ep["spp_masks"],
                                             cfg.n_ways, cfg.k_shots)
    else:
        cat_mean, mp, _, _ = O.count_spp(ep["spp"][0], ep["spp_bboxes"].clone(), ep["spp_masks"], cfg.n_ways, cfg.k_shots, 16)
    res = O.bbox_forward(ep["qry"][:n_ext], cfg.strides, ep["rois"], cat_mean, cfg.n_ways, w,
                         chunk=max(1, 4000 // cfg.n_ways))
    det = ep["det_rois"]
    labels = [ep["det_labels"][det[:, 0] == b] for b in range(cfg.batch)]
    O.mask_attention(ep["qry"][:n_ext], cfg.strides, det, mp, labels, cfg.n_ways, cfg.mask_size)
    return res


def time_cpu_reference(cfg, steps: int, warmup: int, episodes_per_step: int = 1):
    from fgn_b200.episodes import make_episode, make_weights
    torch.set_num_threads(os.cpu_count() or 1)
    w = make_weights(cfg.channels, 0)
    eps = [make_episode(cfg, seed=i) for i in range(min(4, max(1, episodes_per_step)))]
    with torch.no_grad():
        for i in range(warmup):
            cpu_reference_episode(eps[i % len(eps)], w)
        t0 = time.perf_counter()
        n = 0
        for i in range(steps):
            for j in range(episodes_per_step):
                cpu_reference_episode(eps[(i + j) % len(eps)], w)
                n += 1
        dt = time.perf_counter() - t0
    return n, dt


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ---- main -------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="fgn_b200", choices=["fgn_b200", "reference"])
    ap.add_argument("--episodes-per-step", type=int, default=24, help="distinct episodes per GPU per step")
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch eagerly instead of replaying one CUDA graph per episode")
    ap.add_argument("--no-bf16", action="store_true", help="skip the separately reported bf16 variant")
    ap.add_argument("--no-fold", action="store_true", help="skip the separately reported folded-attention variant")
    ap.add_argument("--streams", type=int, default=8, help="side streams episodes are replayed on round-robin")
    ap.add_argument("--e2e-streams", type=int, default=8, help="streams the end-to-end leg overlaps copies and compute on")
    ap.add_argument("--sustained-seconds", type=float, default=3.0, help="length of the sustained leg (0 = skip)")
    ap.add_argument("--images-per-call", type=int, default=12,
                    help="query images (episodes) batched into one call of the path, as the reference batches them "
                         "(main.py:492-499: batch = 12 for 1-way 1-shot, 10 for N3K1, 8 for N3K3); 1 = one call per episode")
    ap.add_argument("--no-large-launch", action="store_true", help="skip the RoIAlign roofline at cfg5's 8192-RoI launch")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle subset check of episode 0 before timing")
    ap.add_argument("--no-gather-overhead", action="store_true", help="skip the W=1 run with the result gather enabled")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    from fgn_b200.episodes import CONFIGS
    cfg = CONFIGS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{cfg.name}: {cfg.n_ways}-way {cfg.k_shots}-shot, {cfg.mode.upper()} strides {list(cfg.strides)}"
                          f" (+P6 for AG-RPN), C={cfg.channels}, query {cfg.img_h}x{cfg.img_w}, R={cfg.num_rois} proposals/img,"
                          f" {cfg.mask_rois} mask RoIs/img, support {cfg.spp_size}px",
              "episodes_per_gpu_per_step": args.episodes_per_step, "sharding": f"episodes x{max(world, 1)}"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        n, dt = time_cpu_reference(cfg, args.steps, min(args.warmup, 1), 1)
        rois_s = n * cfg.num_rois * cfg.batch / dt
        cores = torch.get_num_threads()
        sample = f"{n} episodes of {cfg.name} ({args.steps} steps x 1 episode), {cpu_model()}"
        line = {"impl": "reference", "metric": "guided RoIAlign+fusion RoIs/s", "value": rois_s, "unit": "RoIs/s",
                "episodes_per_s": n / dt, "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
                "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": dict(config, episodes_per_gpu_per_step=1),
                "cpu_baseline": {"value": rois_s, "unit": "RoIs/s", "cores": cores, "kind": "port", "sample": sample},
                "reference_arm_note": ("one CPU process on rank 0's host cores; at N>1 the driver's ratio is N GPUs against this ONE process"
                                       if args.gpus > 1 else "the oracle port on this box's host cores"),
                "e2e": {"value": rois_s, "unit": "RoIs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the fgn_b200 arm has no CPU fallback (use --impl reference)")
    import torch.distributed as dist
    from fgn_b200 import ops
    from fgn_b200.episodes import (EpisodeRunner, ResultGatherer, batch_episodes, build_heads, episode_to_device,
                                   gather_results, make_episode, make_weights, run_guided_path)

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # Pinned host buffers next to the GPU: bind this rank to the CPUs NVML lists for its GPU BEFORE anything is pinned, so
    # that cudaHostAlloc's first touch lands on the GPU's NUMA node and 8 ranks do not all stream through one socket.
    affinity = None
    full_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        affinity = len(os.sched_getaffinity(0))
    except Exception:
        affinity = None
    config["host_affinity_cpus"] = affinity
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    E = args.episodes_per_step
    n_ext = len(cfg.strides)
    rois_per_episode = cfg.num_rois * cfg.batch

    # ---- inputs: E distinct episodes per rank, resident in HBM (channels_last) and mirrored in pinned host memory
    host_eps = [make_episode(cfg, seed=rank * E + i) for i in range(min(E, 4))]
    dev_eps = []
    for i in range(E):
        ep = episode_to_device(host_eps[i % len(host_eps)], device, channels_last=True)
        if i >= len(host_eps):      # distinct data without the CPU RNG cost: perturb on device
            ep["qry"] = [q + 0.01 * i for q in ep["qry"]]
        dev_eps.append(ep)
    # Episodes are batched into calls of B_call query images -- the reference's own way of running them (B query images x
    # N x K supports per step, main.py:492-499: batch = 12 for 1-way 1-shot); per-image work and results are unchanged (checked below,
    # bit for bit, against one eager call per episode), the launches are B_call times larger.
    B_call = max(1, min(args.images_per_call, E))
    if E % B_call or cfg.batch != 1:
        B_call = 1
    n_calls = E // B_call
    call_eps = [batch_episodes(dev_eps[j * B_call:(j + 1) * B_call]) for j in range(n_calls)] if B_call > 1 else dev_eps
    rpn, head = build_heads(cfg, device, seed=0, shared_head=None if cfg.mode == "fpn" else "c4")
    bytes_resident = sum(t.numel() * 4 for ep in dev_eps for t in ep["qry"] + ep["spp"])
    config["l2"] = f"no flush: episode stream of {E} x {bytes_resident / E / 1e6:.0f} MB distinct inputs > 126 MB L2"

    # ---- parity first: episode 0 through the product path, a strided subset of its rows against the CPU oracle
    #      (oracle/parity.py: the checker, outside every timed region).  A bench number cannot outlive a broken kernel.
    parity_line = None
    if rank == 0 and not args.no_parity:
        from oracle import parity
        with torch.no_grad():
            out0 = run_guided_path(rpn, head, dev_eps[0])
            torch.cuda.synchronize()
            rep = parity.episode_parity(host_eps[0], out0, head, make_weights(cfg.channels, 0), roi_subset=48, det_subset=16)
        parity_line = parity.summarize(rep)
        parity_line["tolerance"] = "|a-b| <= 1e-4 + 1e-5|b| vs oracle/fgn_oracle.py, episode 0, strided RoI subset; attention and support vectors in full"
        if not parity_line["ok"]:
            raise SystemExit(f"bench.py: parity check failed before timing: {json.dumps(parity_line)}")
        del out0

    runner = EpisodeRunner(rpn, head, call_eps, use_graphs=not args.no_graphs, n_streams=min(args.streams, n_calls))
    config["images_per_call"] = B_call
    config["launch"] = (f"{n_calls} call(s) of {B_call} query image(s) each (the reference's image batch: main.py:492-499 sets 12 for 1-way 1-shot), " +
                        ("eager" if args.no_graphs else "one CUDA graph per call") +
                        f", calls round-robin on {max(1, min(args.streams, n_calls))} stream(s)")
    W5 = 5 * cfg.n_ways + 1
    gath = ResultGatherer((E, rois_per_episode, W5), device) if world > 1 else None
    res_single = torch.empty((E, rois_per_episode, W5), device=device)
    state = {"k": 0, "buf": res_single}

    def sink(j, o):                                      # call j = episodes [j*B_call, (j+1)*B_call) of the block
        rows = state["buf"][j * B_call:(j + 1) * B_call].view(B_call * rois_per_episode, W5)
        rows[:, : cfg.n_ways + 1].copy_(o["cls_score"])
        rows[:, cfg.n_ways + 1:].copy_(o["bbox_pred"])

    def run_block(n_eps, buf):
        state["buf"] = buf
        runner.begin()
        for j in range(n_eps // B_call):
            runner.run(j, sink)
        runner.end()

    def step_resident():
        if gath is None:
            run_block(E, res_single)
            return res_single
        k = state["k"]
        run_block(E, gath.local(k))          # waits for the gather that last read this staging buffer (two steps ago)
        gath.submit(k)                       # all_gather on the side stream, overlapped with the next step
        state["k"] = k + 1
        return None

    def sync_all():
        if gath is not None:
            gath.drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None, after=None):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            sync_all()
            if sampler:
                sampler.start()
            l0 = ops.launch_count() + runner.launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            if after is not None:
                after()                      # e.g. the compute stream waits for the gathers still in flight
            e1.record()
            sync_all()
            ms = e0.elapsed_time(e1)
            clocks = sampler.stop() if sampler else None
            launches = ops.launch_count() + runner.launches - l0
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks

    # the overlapped graph replay must reproduce the plain eager, single-stream results bit for bit
    with torch.no_grad():
        run_block(E, res_single)
        torch.cuda.synchronize()
        for i in (0, E // 2, E - 1):
            ref = run_guided_path(rpn, head, dev_eps[i])
            torch.cuda.synchronize()
            if not (torch.equal(res_single[i, :, : cfg.n_ways + 1], ref["cls_score"]) and
                    torch.equal(res_single[i, :, cfg.n_ways + 1:], ref["bbox_pred"])):
                raise SystemExit(f"bench.py: episode {i} differs between graph/multi-stream replay and eager execution")
    if gath is not None:                     # ... and the gathered block of this rank is what this rank computed
        with torch.no_grad():
            step_resident()
            got = gath.result(state["k"] - 1)
            torch.cuda.synchronize()
            if not torch.equal(got[rank * E:(rank + 1) * E], res_single):
                raise SystemExit("bench.py: gathered results differ from the local results")

    sampler = ClockSampler(local_rank)
    ms, launches, clocks = timed(step_resident, args.steps, args.warmup, sampler,
                                 after=(gath.drain if gath is not None else None))
    episodes = E * world * args.steps
    rois_total = episodes * rois_per_episode
    value = rois_total / (ms * 1e-3)

    # ---- sustained: the same step looped for >= args.sustained_seconds (the headline region above is a burst of a few
    #      tens of milliseconds at boost clocks), with its own clock record
    sustained = None
    if args.sustained_seconds > 0:
        n_sus = max(args.steps, int(args.sustained_seconds * 1e3 / max(ms / args.steps, 1e-3)) + 1)
        s2 = ClockSampler(local_rank)
        ms_s, _, clocks_s = timed(step_resident, n_sus, 1, s2, after=(gath.drain if gath is not None else None))
        sustained = {"value": E * world * n_sus * rois_per_episode / (ms_s * 1e-3), "unit": "RoIs/s", "steps": n_sus,
                     "seconds": ms_s * 1e-3, "ms_per_step": ms_s / n_sus, "clocks": clocks_s}

    # ---- strong scaling: the SAME total of E episodes per step split over the ranks (E/world each)
    strong = None
    if world > 1 and E % world == 0:
        e_loc = E // world
        g_strong = ResultGatherer((e_loc, rois_per_episode, W5), device)
        st = {"k": 0}
        if e_loc % B_call == 0:
            run_strong = lambda buf: run_block(e_loc, buf)
            b_strong = B_call
        else:   # this rank's share is not a whole number of B_call-image calls: its own calls of gcd(share, B_call) images
            import math as _m
            b_strong = _m.gcd(e_loc, B_call)
            eps_s = [batch_episodes(dev_eps[j * b_strong:(j + 1) * b_strong]) if b_strong > 1 else dev_eps[j]
                     for j in range(e_loc // b_strong)]
            runner_s = EpisodeRunner(rpn, head, eps_s, use_graphs=not args.no_graphs, n_streams=min(args.streams, len(eps_s)))

            def run_strong(buf):
                def sink_s(j, o):
                    rows = buf[j * b_strong:(j + 1) * b_strong].view(b_strong * rois_per_episode, W5)
                    rows[:, : cfg.n_ways + 1].copy_(o["cls_score"])
                    rows[:, cfg.n_ways + 1:].copy_(o["bbox_pred"])
                runner_s.begin()
                for j in range(len(eps_s)):
                    runner_s.run(j, sink_s)
                runner_s.end()

        def step_strong():
            k = st["k"]
            run_strong(g_strong.local(k))
            g_strong.submit(k)
            st["k"] = k + 1

        ms_st, _, _ = timed(step_strong, args.steps, args.warmup, after=g_strong.drain)
        g_strong.drain()
        strong = {"total_episodes_per_step": E, "episodes_per_gpu_per_step": e_loc, "images_per_call": b_strong,
                  "value": E * args.steps * rois_per_episode / (ms_st * 1e-3), "unit": "RoIs/s",
                  "ms_per_step": ms_st / args.steps,
                  "note": "same total work at every N; efficiency = value(N) / (N * value(1) of this key's N=1 run = the headline value at N=1)"}

    # ---- harness overhead at W=1: the same step with the result gather enabled in a one-rank NCCL group
    gather_w1 = None
    if world == 1 and not args.no_gather_overhead:
        try:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", str(29500 + (os.getpid() % 2000)))
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=device)
            g1 = ResultGatherer((E, rois_per_episode, W5), device)
            st1 = {"k": 0}

            def step_g1():
                k = st1["k"]
                run_block(E, g1.local(k))
                g1.submit(k)
                st1["k"] = k + 1

            ms_g1, _, _ = timed(step_g1, args.steps, args.warmup, after=g1.drain)
            g1.drain()
            torch.cuda.synchronize()
            gather_w1 = {"ms_per_step_with_gather": ms_g1 / args.steps, "ms_per_step_without": ms / args.steps,
                         "overhead_pct": 100.0 * (ms_g1 - ms) / ms,
                         "what": "all_gather_into_tensor of the [E,R,5N+1] results in a one-rank NCCL group on a side stream"}
            dist.destroy_process_group()
        except Exception as e:  # pragma: no cover
            gather_w1 = {"unavailable": repr(e)[:200]}

    # ---- e2e: pinned host buffers -> H2D -> path -> D2H of results.  fp32 NCHW (the reference's layout; the headline
    #      e2e) and, beside it, bf16 channels_last host buffers through the bf16 variant (half the PCIe bytes)
    def pin(ep, bf16=False):
        out = {}
        for k, v in ep.items():
            if isinstance(v, list):
                out[k] = [(t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if bf16 and t.dim() == 4 else t).pin_memory()
                          for t in v]
            elif torch.is_tensor(v):
                out[k] = v.pin_memory()
            else:
                out[k] = v
        return out

    def bytes_of(ep):
        return sum(sum(t.numel() * t.element_size() for t in v) if isinstance(v, list) else v.numel() * v.element_size()
                   for k, v in ep.items() if isinstance(v, list) or torch.is_tensor(v))

    # One staging buffer per episode: every input tensor of an episode at a 256-byte aligned offset of ONE pinned allocation, so
    # an episode crosses PCIe as one copy (14 separate copies of 0.3 KB .. 69 MB leave ~5 % of the link idle between them) into
    # a per-stream device slot that is reused (no allocator work inside the timed region); the tensors are views of the slot.
    class PackedEpisode:
        def __init__(self, ep):
            self.meta, off = [], 0
            for k, v in ep.items():
                for j, t in enumerate(v if isinstance(v, list) else [v]):
                    if torch.is_tensor(t):
                        t = t.contiguous(memory_format=torch.channels_last) if (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)
                                                                                  and not t.is_contiguous()) else t.contiguous()
                        cl = t.dim() == 4 and not t.is_contiguous()
                        src = t.permute(0, 2, 3, 1) if cl else t
                        nbytes = src.numel() * src.element_size()
                        self.meta.append((k, isinstance(v, list), j, off, nbytes, src.dtype, tuple(src.shape), cl, src))
                        off = (off + nbytes + 255) & ~255
            self.nbytes = off
            self.host = torch.empty(off, dtype=torch.uint8).pin_memory()
            for (_, _, _, o, nb, dt, shp, _, src) in self.meta:
                self.host[o:o + nb].view(dt).view(shp).copy_(src)
            self.meta = [m[:8] for m in self.meta]
            self.rest = {k: v for k, v in ep.items() if not (torch.is_tensor(v) or (isinstance(v, list) and v and torch.is_tensor(v[0])))}
            self.label_lens = [int(t.numel()) for t in ep["det_labels_list"]] if "det_labels_list" in ep else None

        def to_device(self, slot):
            slot[: self.nbytes].copy_(self.host, non_blocking=True)                      # ONE H2D copy
            out = dict(self.rest)
            for (k, is_list, j, o, nb, dt, shp, cl) in self.meta:
                t = slot[o:o + nb].view(dt).view(shp)
                t = t.permute(0, 3, 1, 2) if cl else t
                if is_list:
                    out.setdefault(k, []).append(t)
                else:
                    out[k] = t
            return out

    out_host = torch.empty((E, rois_per_episode, W5), dtype=torch.float32).pin_memory()
    d2h = out_host.numel() * 4
    e2e_streams = [torch.cuda.Stream() for _ in range(max(1, args.e2e_streams))]   # copies of episode i+1 overlap compute of episode i

    def make_e2e_step(pinned, channels_last):
        packed = [PackedEpisode(ep) for ep in pinned]
        slots = [torch.empty(max(p.nbytes for p in packed), dtype=torch.uint8, device=device) for _ in e2e_streams]

        def step_e2e():
            main = torch.cuda.current_stream()
            fork = torch.cuda.Event()
            fork.record(main)
            for i in range(E):
                st_ = e2e_streams[i % len(e2e_streams)]
                st_.wait_event(fork)
                with torch.cuda.stream(st_):
                    ep = packed[i % len(packed)].to_device(slots[i % len(e2e_streams)])                    # H2D (one copy)
                    o = run_guided_path(rpn, head, ep)                                                  # repack + path
                    out_host[i].copy_(torch.cat([o["cls_score"], o["bbox_pred"]], 1), non_blocking=True)
            for st_ in e2e_streams:
                ev = torch.cuda.Event()
                ev.record(st_)
                main.wait_event(ev)
            main.synchronize()                                                                          # results on host
        return step_e2e

    e2e_steps = max(2, min(args.steps, 5))
    pinned32 = [pin(ep) for ep in host_eps]
    h2d = bytes_of(pinned32[0]) * E
    step32 = make_e2e_step(pinned32, False)
    ms_e2e, _, _ = timed(step32, e2e_steps, 1)
    # what came back over PCIe is what the resident step computes for the same episodes, bit for bit
    if gath is None:
        with torch.no_grad():
            run_block(E, res_single)
            step32()
            torch.cuda.synchronize()
        for i in range(min(len(pinned32), E)):
            if not torch.equal(out_host[i], res_single[i].cpu()):
                raise SystemExit(f"bench.py: end-to-end results of episode {i} differ from the resident step's")
    e2e_value = E * world * e2e_steps * rois_per_episode / (ms_e2e * 1e-3)
    e2e = {"value": e2e_value, "unit": "RoIs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": e2e_steps, "host_layout": "NCHW fp32 pinned",
           "pcie_gbs_per_gpu": (h2d + d2h) * e2e_steps / (ms_e2e * 1e-3) / 1e9,
           "note": "PCIe-bound: the H2D copy of the fp32 pyramids is the step; per-GPU rate above, all ranks share the host's memory / root complexes"}
    del pinned32
    if not args.no_bf16:
        pinned16 = [pin(ep, bf16=True) for ep in host_eps]
        h2d16 = bytes_of(pinned16[0]) * E
        ms16e, _, _ = timed(make_e2e_step(pinned16, True), e2e_steps, 1)
        e2e["bf16_host_buffers"] = {"value": E * world * e2e_steps * rois_per_episode / (ms16e * 1e-3), "unit": "RoIs/s",
                                    "h2d_bytes_per_step": int(h2d16), "host_layout": "channels_last bf16 pinned (bf16 variant, reported separately)",
                                    "pcie_gbs_per_gpu": (h2d16 + d2h) * e2e_steps / (ms16e * 1e-3) / 1e9}
        del pinned16

    # ---- rooflines.  (1) the dominant kernel: multi-level RoIAlign, timed alone on its stream
    scales = [1.0 / s for s in cfg.strides]
    alg_bytes_call = roi_align_algorithmic_bytes(call_eps[0]["cfg"], call_eps[0]["rois"].cpu())   # one launch = one call's RoIs
    alg_bytes = alg_bytes_call / B_call                                                            # per episode

    def roi_only():
        for ep in call_eps:
            ops.roi_align_multilevel(ep["qry"][:n_ext], ep["rois"], scales, 7, 0, True, out_format="nhwc")

    # the E launches (one per resident episode, 16 x 92 MB of distinct maps) are replayed as one CUDA graph so
    # that the number is the device's: the Python wrapper costs about 30 us per call, the same order as the kernel
    with torch.no_grad():
        roi_only()
        torch.cuda.synchronize()
        roi_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(roi_graph):
            roi_only()
    ms_roi, l_roi, _ = timed(roi_graph.replay, max(args.steps, 10), args.warmup)
    per_launch_s = ms_roi * 1e-3 / (max(args.steps, 10) * n_calls)
    # for the record, the same kernel on single-image launches (1000 RoIs: what rounds 1 and early 2 measured)
    single = None
    if B_call > 1:
        def roi_single():
            for ep in dev_eps[:8]:
                ops.roi_align_multilevel(ep["qry"][:n_ext], ep["rois"], scales, 7, 0, True, out_format="nhwc")
        with torch.no_grad():
            roi_single()
            torch.cuda.synchronize()
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                roi_single()
        ms_1, _, _ = timed(g1.replay, 10, 3)
        t1 = ms_1 * 1e-3 / (10 * min(8, len(dev_eps)))
        single = {"rois_per_launch": rois_per_episode, "us_per_launch": t1 * 1e6,
                  "algorithmic_bytes_per_launch": roi_align_algorithmic_bytes(cfg, host_eps[0]["rois"])}
        del g1
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    achieved = alg_bytes_call / per_launch_s / 1e9

    def committed(name):
        """Per-launch figures of a committed ncu --set full capture (tools/ncu_summary.py --json)."""
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            return None

    ncu_roi = (committed("r02_ncu_roi_align_window_b12.json" if B_call == 12 else "r02_ncu_roi_align_window.json" if B_call == 1 else "none")
               if cfg.name == WORKLOAD else None)
    roofline = {"kernel": "roi_align_window_kernel<7,2,3,2> (level assignment + multi-level RoIAlign, NHWC in/out)", "bound": "hbm",
                "rois_per_launch": B_call * rois_per_episode, "images_per_launch": B_call,
                "us_per_1000_rois": per_launch_s * 1e6 / (B_call * rois_per_episode / 1000.0),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_call, "us_per_launch": per_launch_s * 1e6,
                "single_image_launch": (dict(single, frac=single["algorithmic_bytes_per_launch"] / (single["us_per_launch"] * 1e-6) / 1e9 / peak,
                                             note="one image, 1000 RoIs per launch: bound by load balance across the persistent CTAs (DESIGN.md section 7)")
                                        if single else None),
                "traffic": (ncu_roi or {}).get("traffic"),
                "traffic_source": (("profiles/" + ncu_roi.get("file", "r02_ncu_roi_align_window.json") +
                                    " (dram__bytes_read.sum + dram__bytes_write.sum per launch)") if ncu_roi else None)}

    # (1b) the same kernel on a large launch: cfg5's 16 images x 512 proposals in one call (8192 RoIs).  The fraction at the
    #      benchmark workload (1000 RoIs per launch, every cell re-read ~4.5x through L2) is bound by load balance and by what
    #      the L2 delivers; with 28 instead of 5 items per CTA and half the re-read factor the unchanged kernel sits at the HBM
    #      roofline.  Device-generated maps (the values are irrelevant to the timing; parity at this size: tests/).
    roofline_large = None
    if world == 1 and cfg.name == WORKLOAD and not args.no_large_launch:
        from fgn_b200.episodes import CONFIGS as _CFGS, level_hw as _lhw, synth_rois as _srois
        c5 = _CFGS["cfg5_coco2voc_mask_fpn"]
        g5 = torch.Generator().manual_seed(555)
        sets = []
        for i in range(2):                                                       # 2 x 1.46 GB of distinct maps >> L2
            feats = [torch.randn(c5.batch, *_lhw(c5.img_h, c5.img_w, s_), c5.channels, device=device).permute(0, 3, 1, 2)
                     for s_ in c5.strides]
            sets.append((feats, _srois(g5, c5.num_rois * c5.batch, c5.img_h, c5.img_w, c5.batch)))
        alg5 = roi_align_algorithmic_bytes(c5, sets[0][1])
        sets = [(f, r.to(device)) for f, r in sets]
        sc5 = [1.0 / s_ for s_ in c5.strides]

        def roi_large():
            for f, r in sets:
                ops.roi_align_multilevel(f, r, sc5, 7, 0, True, out_format="nhwc")

        with torch.no_grad():
            roi_large()
            torch.cuda.synchronize()
            g_large = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_large):
                roi_large()
        ms_l, _, _ = timed(g_large.replay, 10, 3)
        t5 = ms_l * 1e-3 / (10 * len(sets))
        ncu5 = committed("r02_ncu_roi_align_window_cfg5.json") or {}
        roofline_large = {"kernel": "roi_align_window_kernel<7,2,3,2> (the same kernel, unchanged)",
                          "workload": "cfg5_coco2voc_mask_fpn: 16 images x 512 proposals = 8192 RoIs in one launch",
                          "bound": "hbm", "us_per_launch": t5 * 1e6, "us_per_1000_rois": t5 * 1e6 / (c5.num_rois * c5.batch / 1000.0),
                          "algorithmic_bytes_per_launch": alg5, "achieved": alg5 / t5 / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": alg5 / t5 / 1e9 / peak, "peak_source": peak_src,
                          "traffic": ncu5.get("traffic"),
                          "frac_of_peak_by_measured_traffic": (ncu5["traffic"] / t5 / 1e9 / peak) if ncu5.get("traffic") else None,
                          "traffic_source": "profiles/r02_ncu_roi_align_window_cfg5.json (dram__bytes_read.sum + dram__bytes_write.sum of one launch)" if ncu5 else None}
        del sets, g_large
        torch.cuda.empty_cache()

    # (2) the tensor-bound kernel: the relation head's contraction (split-weight count: (R + B*N) * 49 * C * C * 2 FLOP)
    M_q = B_call * rois_per_episode * 49
    a_q = [torch.randn(M_q, cfg.channels, device=device) for _ in range(4)]
    w_q = head.cls_reg_shared_conv.weight.detach().reshape(cfg.channels, 2 * cfg.channels)[:, : cfg.channels]

    w_q = w_q.contiguous()
    w_q_split = ops.conv_split_weights(w_q[None])                 # as the head holds it after loading its weights

    def gemm_only():
        for a in a_q:
            ops.gemm_nt(a, w_q, None, "fp32", b_split=w_q_split)

    with torch.no_grad():
        gemm_only()
        torch.cuda.synchronize()
        gemm_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gemm_graph):
            gemm_only()
    ms_g, _, _ = timed(gemm_graph.replay, max(args.steps, 10), args.warmup)
    g_s = ms_g * 1e-3 / (max(args.steps, 10) * len(a_q))
    flops_split = 2.0 * M_q * cfg.channels * cfg.channels
    ncu_gemm = committed("r02_ncu_gemm_tcgen05_pair.json")
    tf32_nominal = 1100.0
    roofline_tensor = {"kernel": "conv_tc2_kernel<3,0> (relation conv as split-weight GEMM on CTA pairs: tcgen05 cta_group::2 kind::tf32, 256-row MMAs, 3xTF32)",
                       "bound": "tensor", "achieved": flops_split / g_s / 1e12, "unit": "TFLOP/s",
                       "achieved_tensor_work": 3 * flops_split / g_s / 1e12, "peak": tf32_nominal,
                       "frac": 3 * flops_split / g_s / 1e12 / tf32_nominal,
                       "peak_source": "nominal dense TF32 (MEASURED_PEAKS.json holds bf16 only: %.0f TFLOP/s burst; TF32 runs at half the bf16 rate)" % float(peaks.get("bf16_tflops", 0.0)),
                       "what": "achieved = split-weight FLOP count / time (SURVEY 8d); achieved_tensor_work counts the three TF32 passes of the fp32-parity scheme, frac = that / nominal TF32",
                       "us_per_launch": g_s * 1e6, "M": M_q, "N": cfg.channels, "K": cfg.channels,
                       "tensor_pipe_pct_of_active": (ncu_gemm or {}).get("tensor_pipe_pct_of_active"),
                       "tensor_pipe_source": "profiles/r02_ncu_gemm_tcgen05_pair.json" if ncu_gemm else None}
    del a_q

    # (3) the whole step against HBM: fused accounting -- every input map read once, every output written
    #     once, RoI features and the split-conv output never counted (they are workspace traffic, not algorithmic)
    q_bytes = sum(t.numel() * 4 for t in host_eps[0]["qry"])
    s_bytes = sum(t.numel() * 4 for t in host_eps[0]["spp"]) + host_eps[0]["spp_masks"].numel()
    step_bytes = ((1 + cfg.n_ways) * q_bytes                     # a1: read qry once, write N attended copies
                  + s_bytes                                       # a1 vectors + a2 support branch: supports read once
                  + (alg_bytes - rois_per_episode * cfg.channels * 49 * 4)   # a4/a5: unique footprint cells + rois
                  + 2 * cfg.channels * cfg.channels * 4 + rois_per_episode * W5 * 4      # a6-a8: weights + logits
                  + cfg.mask_rois * cfg.batch * cfg.channels * cfg.mask_size ** 2 * 4)   # a9: attended mask features out
    step_s = ms * 1e-3 / (args.steps * E)
    roofline_step = {"bound": "hbm", "accounting": "fused (RoI features / split-conv output not counted)",
                     "algorithmic_bytes_per_episode": int(step_bytes), "us_per_episode": step_s * 1e6,
                     "achieved": step_bytes / step_s / 1e9, "peak": peak, "unit": "GB/s", "frac": step_bytes / step_s / 1e9 / peak,
                     "peak_source": peak_src}

    # ---- bf16 variant, reported separately (bf16 NHWC maps + bf16 contraction operands; stated tolerance 3e-2
    #      abs on logits, see tests/test_gpu_parity.py::test_bf16_guided_path_variant)
    bf16_line = None
    if not args.no_bf16:
        eps16 = []
        for ep in call_eps:
            e16 = dict(ep)
            e16["qry"] = [q.bfloat16().contiguous(memory_format=torch.channels_last) for q in ep["qry"]]
            e16["spp"] = [q.bfloat16().contiguous(memory_format=torch.channels_last) for q in ep["spp"]]
            eps16.append(e16)
        runner16 = EpisodeRunner(rpn, head, eps16, use_graphs=not args.no_graphs, n_streams=min(args.streams, n_calls))

        def step16():
            runner16.begin()
            for i in range(n_calls):
                runner16.run(i)
            runner16.end()

        ms16, _, _ = timed(step16, args.steps, args.warmup)
        with torch.no_grad():
            a = run_guided_path(rpn, head, eps16[0])["cls_score"]
            b = run_guided_path(rpn, head, call_eps[0])["cls_score"]
        bf16_line = {"value": E * world * args.steps * rois_per_episode / (ms16 * 1e-3), "unit": "RoIs/s",
                     "ms_per_step": ms16 / args.steps, "max_abs_logit_diff_vs_fp32": float((a - b).abs().max()),
                     "stated_tolerance": 3e-2, "dtype": "bf16 maps/operands, f32 accumulate"}
        del runner16, eps16

    # ---- folded-attention variant, reported separately: the AG-RPN channel attention is folded into the RPN conv's
    #      weights (conv(q*v) = conv(q, W*v): class vectors + B*N*L folded weight sets, 2 + 1 launches) instead of
    #      writing the attended pyramid (the reference's order, which `value` keeps); everything else identical
    fold_line = None
    if not args.no_fold:
        runner_f = EpisodeRunner(rpn, head, call_eps, use_graphs=not args.no_graphs, with_attention="fold",
                                 n_streams=min(args.streams, n_calls))

        def step_fold():
            state["buf"] = res_single
            runner_f.begin()
            for i in range(n_calls):
                runner_f.run(i, sink)
            runner_f.end()

        ms_f, _, _ = timed(step_fold, args.steps, args.warmup)
        l_f = sum(runner_f.launches_per_episode) * args.steps if runner_f.launches_per_episode else 0
        fold_line = {"value": E * world * args.steps * rois_per_episode / (ms_f * 1e-3), "unit": "RoIs/s",
                     "ms_per_step": ms_f / args.steps, "gpu_launches": int(l_f),
                     "what": "AG-RPN attention as rpn_conv weight sets (fgn_fold_attention_weights); attended pyramid not materialised"}
        del runner_f

    if rank == 0:
        line = {"metric": "guided RoIAlign+fusion RoIs/s", "value": value, "unit": "RoIs/s",
                "episodes_per_s": episodes / (ms * 1e-3), "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks, "parity": parity_line,
                "e2e": e2e, "gpu_launches": int(launches),
                "launches_per_call": (runner.launches_per_episode[0] if runner.launches_per_episode else None),
                "launches_per_episode": (runner.launches_per_episode[0] / B_call if runner.launches_per_episode else None),
                "roofline": roofline, "roofline_large_launch": roofline_large, "roofline_tensor": roofline_tensor, "roofline_step": roofline_step,
                "sustained": sustained, "strong_scaling": strong, "gather_overhead_w1": gather_w1,
                "collective": (None if world == 1 else "all_gather_into_tensor of [E,R,5N+1] per step on a side stream, double-buffered, overlapped with the next step"),
                "bf16_variant": bf16_line, "folded_attention_variant": fold_line,
                "reference_arm_note": ("--impl reference runs the CPU path on rank 0 only: at N>1 the driver's ratio is N GPUs against ONE CPU process"
                                       if world > 1 else "--impl reference: the oracle port on this box's host cores")}
        if world == 1 and not args.no_cpu_baseline:
            if full_affinity is not None:                     # the CPU leg gets every host core back
                os.sched_setaffinity(0, full_affinity)
            t0 = time.perf_counter()
            n, dt = time_cpu_reference(cfg, 1, 1, 1)
            reps = max(1, min(20, int(args.cpu_seconds / max(dt, 1e-3)) - 1))
            if reps > 1:
                n, dt = time_cpu_reference(cfg, reps, 0, 1)
            v = n * rois_per_episode / dt
            line["cpu_baseline"] = {"value": v, "unit": "RoIs/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{n} episodes of {cfg.name}, {cpu_model()}, {time.perf_counter() - t0:.1f}s"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
