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


@pytest.mark.gpu
def test_preprocessing_device_matches_reference_golden():
    """resize_align_multi_scale (cv2.warpAffine restated in fixed point) and ToTensor + Normalize on the device against
    images / tensors produced by the unmodified reference + torchvision: bit for bit."""
    from oracle.make_golden_transforms import MEAN, PRE_CASES, STD, pre_image
    g = np.load(GOLD)
    for i, (seed, h, w, size) in enumerate(PRE_CASES):
        img = pre_image(seed, h, w)
        pre = "p%02d_" % i
        resized, center, scale = T.resize_align_multi_scale(img, size, 1.0, 1.0)
        assert resized.dtype == np.uint8 and np.array_equal(resized, g[pre + "resized"]), i
        assert np.array_equal(center, g[pre + "center"]) and np.array_equal(scale, g[pre + "scale"]), i
        # a batch of two identical images, normalised on the device
        batch = torch.from_numpy(np.stack([img, img])).cuda()
        t32, _, _ = T.resize_align_normalize_device(batch, size, 1.0, 1.0, MEAN, STD)
        assert t32.dtype == torch.float32 and np.array_equal(t32[1].cpu().numpy(), g[pre + "tensor"]), i
        t16, _, _ = T.resize_align_normalize_device(batch, size, 1.0, 1.0, MEAN, STD, half=True)
        assert torch.equal(t16.cpu(), torch.from_numpy(g[pre + "tensor"]).half()[None].expand(2, -1, -1, -1))


@pytest.mark.gpu
def test_preprocessing_device_matches_cv2_at_full_size():
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(7)
    for h, w, size in ((480, 640, 512), (640, 427, 512), (1080, 1920, 640)):
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        dims, center, scale = T.get_multi_scale_size(img, size, 1.0, 1.0)
        expect = cv2.warpAffine(img, T.get_affine_transform(center, scale, 0, dims), dims)
        got, _, _ = T.resize_align_multi_scale(img, size, 1.0, 1.0)
        assert np.array_equal(got, expect), (h, w, size)
