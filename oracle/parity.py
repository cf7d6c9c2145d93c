"""oracle/parity.py -- TEST INFRASTRUCTURE ONLY (the checker; never measured, never shipped).

Full-size parity of one episode of a BASELINE.json config: the CUDA path has run on the whole
episode (all R RoIs, all supports, every pyramid level); this module re-derives the same quantities
with the CPU oracle (oracle/fgn_oracle.py) and compares.  Quantities that are cheap on the CPU are
compared in full (AG-RPN attention on every level, class means, masked GAP vectors); the per-RoI
quantities (logits, box deltas, RoI features, attended mask features) are compared on a strided RoI
subset -- rows are per-RoI independent (fgn_roi_head.py:266-268), so a subset of rows of the full-size
GPU result is checked against the oracle run on exactly those RoIs.

Used by tests/test_gpu_config_parity.py, bench.py (episode 0, before timing) and smoke().
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import fgn_oracle as O

ATOL, RTOL = 1e-4, 1e-5          # BASELINE.json north_star: 1e-4 max-abs / 1e-5 relative for fp32 features


def _cmp(report: Dict[str, dict], name: str, got: torch.Tensor, want: torch.Tensor, atol: float, rtol: float) -> None:
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    if got.shape != want.shape:
        report[name] = dict(ok=False, max_abs_err=float("inf"), shape_got=list(got.shape), shape_want=list(want.shape))
        return
    if got.numel() == 0:
        report[name] = dict(ok=True, max_abs_err=0.0, n=0)
        return
    err = (got - want).abs()
    bad = err > (atol + rtol * want.abs())
    report[name] = dict(ok=not bool(bad.any()), max_abs_err=float(err.max()), n=int(got.numel()), n_bad=int(bad.sum()),
                        atol=atol, rtol=rtol)


def episode_parity(ep_host: Dict[str, object], out: Dict[str, object], head, w: Dict[str, torch.Tensor],
                   roi_subset: int = 48, det_subset: int = 24, shared_head_cpu=None,
                   atol: float = ATOL, rtol: float = RTOL, post_head_atol: Optional[float] = None,
                   check_attention: bool = True) -> Dict[str, dict]:
    """Compare the CUDA results `out` (dict from episodes.run_guided_path, device tensors) and the
    vectors `head` holds after count_spp with the oracle on the host episode `ep_host`.

    ``shared_head_cpu``: CPU copy of the C4 res5 module when the head has one (the adjacent cuDNN
    module; quantities downstream of it are compared with ``post_head_atol`` because a cuDNN and an
    MKL convolution stack differ by more than one rounding).  Returns {quantity: report}.
    """
    cfg = ep_host["cfg"]
    n_ext = len(cfg.strides)
    rep: Dict[str, dict] = {}
    pa = atol if (shared_head_cpu is None or post_head_atol is None) else post_head_atol

    # a1: AG-RPN channel attention on every RPN level, in full (fgn_ag_rpn_head.py:33-46)
    if check_attention and out.get("qry_fmap_mod") is not None:
        for lvl in range(len(cfg.rpn_strides)):
            _, mod = O.agrpn_attention(ep_host["qry"][lvl], ep_host["spp"][lvl], cfg.n_ways, cfg.k_shots)
            _cmp(rep, f"qry_fmap_mod[{lvl}]", out["qry_fmap_mod"][lvl], mod, atol, rtol)

    # a2: support branch, in full (fgn_roi_head.py:419-449)
    if cfg.mode == "fpn":
        cat_mean, mp, _, _ = O.count_spp_fpn(ep_host["spp"][:n_ext], cfg.strides, ep_host["spp_bboxes"].clone(),
                                             ep_host["spp_masks"], cfg.n_ways, cfg.k_shots)
    else:
        cat_mean, mp, _, _ = O.count_spp(ep_host["spp"][0], ep_host["spp_bboxes"].clone(), ep_host["spp_masks"],
                                         cfg.n_ways, cfg.k_shots, 16, shared_head=shared_head_cpu)
    _cmp(rep, "cat_mean", head.spp_fmaps_roi_aligned_cat_mean, cat_mean, pa, rtol)
    _cmp(rep, "masked_gap", head.spp_fvecs_roi_aligned_cat_mean_mp, mp, pa, rtol)

    # a3-a8: logits and deltas of a strided RoI subset (fgn_roi_head.py:328-342)
    rois = ep_host["rois"]
    r = rois.shape[0]
    step = max(1, r // max(1, roi_subset))
    sel = torch.arange(0, r, step)
    feats = ep_host["qry"][:n_ext]
    want = O.bbox_forward(feats, cfg.strides, rois[sel], cat_mean, cfg.n_ways, w, shared_head=shared_head_cpu,
                          chunk=max(1, 2000 // cfg.n_ways))
    _cmp(rep, "cls_score", out["cls_score"][sel.to(out["cls_score"].device)], want["cls_score"], pa, rtol)
    _cmp(rep, "bbox_pred", out["bbox_pred"][sel.to(out["bbox_pred"].device)], want["bbox_pred"], pa, rtol)
    if out.get("bbox_feats") is not None:
        _cmp(rep, "bbox_feats", out["bbox_feats"][sel.to(out["bbox_feats"].device)], want["bbox_feats"], pa, rtol)
    if out.get("levels") is not None:
        lv = O.map_roi_levels_c(rois, n_ext) if n_ext > 1 else torch.zeros(r, dtype=torch.long)
        same = torch.equal(out["levels"].cpu().long(), lv.long())
        rep["levels"] = dict(ok=bool(same), max_abs_err=0.0 if same else 1.0, n=int(r), bit_exact=True)

    # a9: attended mask features of a strided detection subset (fgn_roi_head.py:360-382, :707-714)
    if out.get("mask_feats") is not None:
        det = ep_host["det_rois"]
        d = det.shape[0]
        dstep = max(1, d // max(1, det_subset))
        dsel = torch.arange(0, d, dstep)
        gather = torch.cat([ep_host["det_labels_list"][i] + cfg.n_ways * i for i in range(cfg.batch)])[dsel]
        vecs = mp.view(cfg.batch * cfg.n_ways, -1, 1, 1)[gather]
        mf, _ = O.single_roi_extractor(feats, det[dsel], cfg.strides, cfg.mask_size, 0, True, 56.0)
        if shared_head_cpu is not None:
            mf = shared_head_cpu(mf)
        _cmp(rep, "mask_feats", out["mask_feats"][dsel.to(out["mask_feats"].device)], mf * vecs, pa, rtol)
    return rep


def summarize(rep: Dict[str, dict]) -> Dict[str, object]:
    """One-line summary for a bench line: worst error and whether every quantity met its bar."""
    worst = max(rep.items(), key=lambda kv: kv[1]["max_abs_err"]) if rep else ("", dict(max_abs_err=0.0))
    return dict(ok=all(v["ok"] for v in rep.values()), parity_max_abs_err=worst[1]["max_abs_err"], worst=worst[0],
                checked={k: round(v["max_abs_err"], 9) for k, v in rep.items()})
