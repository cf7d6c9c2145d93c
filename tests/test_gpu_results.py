"""The wire / on-disk side of the path: the per-image result dict of fgn.py:262-303 with its input
keys and qry_isegmaps_rle, encoded on the device, against the oracle's restatement of encode_mask_results."""
import numpy as np
import pytest
import torch

from oracle import fgn_oracle as O

pytestmark = pytest.mark.gpu


def _masks(seed, d, h, w):
    g = np.random.default_rng(seed)
    m = np.zeros((d, h, w), dtype=bool)
    yy, xx = np.mgrid[:h, :w]
    for i in range(d):
        cy, cx, ry, rx = g.uniform(0, h), g.uniform(0, w), g.uniform(1, max(2, h / 2)), g.uniform(1, max(2, w / 2))
        m[i] = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1
        m[i] ^= g.random((h, w)) < 0.02                      # speckle: many short runs
    return m


@pytest.mark.parametrize("d,h,w", [(5, 37, 53), (3, 480, 480), (2, 800, 1344), (4, 1, 70), (4, 70, 1), (2, 32, 64)])
def test_mask_rle_encode_round_trip_against_encode_mask_results(d, h, w):
    from fgn_b200 import ops
    m = _masks(100 + d + h, d, h, w)
    m[0] = False                                             # all zeros: one run
    if d > 1:
        m[1] = True                                          # all ones: empty first run
    if d > 2:
        m[2, :, -1] = True
        m[2, -1, :] = True                                   # runs that end on the last row / column
    rles, counts = ops.mask_rle_encode(torch.from_numpy(m).cuda(), return_counts=True)
    want = O.encode_mask_results(m)
    for i in range(d):
        assert counts[i] == O.rle_counts(m[i])
        assert rles[i] == want[i]
        assert np.array_equal(O.rle_decode(O.rle_from_string(rles[i]["counts"]), h, w), m[i])
    # a cap that is too small grows instead of truncating
    small, _ = ops.mask_rle_encode(torch.from_numpy(m).cuda(), cap=4, return_counts=True)
    assert small == rles
    assert ops.mask_rle_encode(torch.zeros(0, h, w, dtype=torch.bool).cuda()) == []


def test_format_results_carries_inputs_and_encodes_the_query_masks():
    """fgn.py:283-300: every input key of the image, tensors as numpy, qry_isegmaps -> qry_isegmaps_rle."""
    import fgn_b200
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    det = [torch.rand(4, 5, generator=g).to(dev), torch.rand(0, 5).to(dev)]
    lab = [torch.tensor([0, 1, 0, 2]).to(dev), torch.zeros(0, dtype=torch.long).to(dev)]
    qm = [_masks(7, 3, 40, 56), _masks(8, 1, 40, 56)]
    inputs = dict(qry_img_id=torch.tensor([11, 12]), qry_bboxes=[torch.rand(3, 4), torch.rand(1, 4)],
                  qry_cat_ids=[torch.tensor([1, 2, 3]), torch.tensor([4])],
                  qry_isegmaps=[torch.from_numpy(qm[0]).to(dev), torch.from_numpy(qm[1])],       # device and host masks
                  qry_child_idx=[0, 1], cats_ids_to_sample_real=torch.tensor([[1, 2, 3], [4, 5, 6]]),
                  spp_insts_ids=torch.arange(6).view(2, 3))
    out = fgn_b200.FGN.format_results((det, lab), inputs)
    assert len(out) == 2
    for i, one in enumerate(out):
        assert "qry_isegmaps" not in one and set(inputs) - {"qry_isegmaps"} <= set(one)
        assert one["qry_isegmaps_rle"] == O.encode_mask_results(qm[i])
        assert isinstance(one["qry_bboxes"], np.ndarray) and isinstance(one["spp_insts_ids"], np.ndarray)
        assert int(one["qry_img_id"]) == 11 + i and one["qry_child_idx"] == i
        db = det[i].cpu().numpy()
        assert np.array_equal(one["dt_bboxes"], db[:, [1, 0, 3, 2]]) and np.array_equal(one["dt_scores"], db[:, 4])
