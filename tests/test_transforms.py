"""SURVEY 8(f) row 3 (post-processing side): the coordinate helpers of the evaluation loop against golden vectors from
the unmodified reference lib/utils/transforms.py (cv2.getAffineTransform restated without OpenCV), and the batched
device version of get_final_preds against the same vectors."""
import os

import numpy as np
import pytest
import torch

from litepose_b200.lib.utils import transforms as T
from oracle.make_golden_transforms import CASES, keypoints

GOLD = os.path.join(os.path.dirname(__file__), "golden", "transforms.npz")


def test_host_helpers_match_reference_golden():
    g = np.load(GOLD)
    for i, (h, w, size, cur, mn) in enumerate(CASES):
        pre = "c%02d_" % i
        (wr, hr), c, s = T.get_multi_scale_size(np.zeros((h, w, 3), np.uint8), size, cur, mn)
        assert (wr, hr) == tuple(g[pre + "size"]), i
        assert np.array_equal(c, g[pre + "center"]) and np.array_equal(s, g[pre + "scale"]), i
        assert np.array_equal(T.get_affine_transform(c, s, 0, (wr, hr)), g[pre + "fwd"]), i
        assert np.array_equal(T.get_affine_transform(c, s, 0, [wr, hr], inv=1), g[pre + "inv"]), i
        persons = keypoints(100 + i, 1 + i % 4, 14, 2, wr, hr)
        final = np.stack(T.get_final_preds([persons], c, s, [wr, hr]))
        assert final.dtype == np.float32 and np.array_equal(final, g[pre + "final"]), i


@pytest.mark.gpu
def test_final_preds_device_matches_reference_golden():
    g = np.load(GOLD)
    # images that share a heat-map size form one batch, as in the pipeline
    groups = {}
    for i in range(len(CASES)):
        groups.setdefault(tuple(g["c%02d_size" % i]), []).append(i)
    for (wr, hr), idx in groups.items():
        pcap = 6
        ans = np.full((len(idx), pcap, 14, 5), -3.0, np.float32)
        num = np.zeros(len(idx), np.int32)
        for k, i in enumerate(idx):
            persons = keypoints(100 + i, 1 + i % 4, 14, 2, wr, hr)
            num[k] = len(persons)
            ans[k, :len(persons)] = np.stack(persons)
        d_ans, d_num = torch.from_numpy(ans).cuda(), torch.from_numpy(num).cuda()
        T.final_preds_device(d_ans, d_num, [g["c%02d_center" % i] for i in idx], [g["c%02d_scale" % i] for i in idx],
                             [int(wr), int(hr)])
        out = d_ans.cpu().numpy()
        for k, i in enumerate(idx):
            assert np.array_equal(out[k, :num[k]], g["c%02d_final" % i]), i
            assert (out[k, num[k]:] == -3.0).all()               # rows of absent persons are untouched
