"""Mint tests/golden/mask_paste_kat.npz -- known answers for the test-time mask pasting (FCNMaskHead.get_seg_masks /
_do_paste_mask [3P, mmdet 2.18], called from fgn_roi_head.py:668-671) produced by the torch calls _do_paste_mask
itself makes: sigmoid + F.grid_sample(bilinear, zeros padding, align_corners=False) on CPU, skip_empty=False.

    python tests/golden/make_golden_paste.py

The fixture holds the inputs, the pasted float values (so a test can tell a real mismatch from a pixel whose value
lies within rounding of the threshold), the thresholded masks and their column-major run lengths.  mmdet and
pycocotools are not installable here; the run lengths are counted with plain numpy on torch's masks.
"""
import os

import numpy as np
import torch
import torch.nn.functional as F

H, W, M, THR = 61, 83, 28, 0.5
BOXES = np.array([[10.3, 5.2, 50.7, 40.1], [0, 0, 83, 61], [-5.5, -3.2, 20.1, 70.3], [30, 30, 30, 45.5],
                  [70.2, 50.1, 82.9, 60.9], [40.5, 10.5, 12.5, 33.0], [100, 100, 120, 130], [-40, -40, -10, -5],
                  [0.5, 0.5, 82.5, 60.5]], np.float32)


def torch_paste(logits, boxes, h, w):
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


def run_lengths(mask):
    flat = mask.T.reshape(-1).astype(np.uint8)
    pos = np.concatenate([[0], np.flatnonzero(flat[1:] != flat[:-1]) + 1, [flat.size]])
    c = np.diff(pos).tolist()
    return ([0] + c) if flat[0] else c


if __name__ == "__main__":
    rng = np.random.default_rng(20261018)
    logits = rng.normal(0, 3, (len(BOXES), 1, M, M)).astype(np.float32)
    values = torch_paste(logits, BOXES, H, W)
    masks = values >= np.float32(THR)
    counts = [run_lengths(m) for m in masks]
    flat = np.concatenate([np.asarray(c, np.int64) for c in counts])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mask_paste_kat.npz")
    np.savez_compressed(out, logits=logits, boxes=BOXES, img_hw=np.array([H, W]), thr=np.float32(THR), values=values,
                        masks=masks, counts_flat=flat, counts_len=np.array([len(c) for c in counts]),
                        torch_version=np.array(torch.__version__))
    print(out, masks.sum(axis=(1, 2)).tolist(), [len(c) for c in counts])
