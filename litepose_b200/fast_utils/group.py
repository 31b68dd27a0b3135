"""Mirror of the reference's "fast inference" parser (nano_demo/fast_utils/group.py:10-47): peaks + KM assignment,
no adjust / refine.  `parse(det, tmap, scale)` keeps the reference's single-image contract; `parse_batch` runs the
whole batch with two kernel launches and one device->host copy."""
import torch

from . import plugins
from ..lib.core.group import Params as _CoreParams


class Params(_CoreParams):
    """The demo's Params (nano_demo/fast_utils/group.py:10-31) = the evaluation parser's fields
    (litepose_b200.lib.core.group.Params) plus the peak window."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.window_size = cfg.TEST.NMS_KERNEL


class HeatmapParser(object):
    def __init__(self, cfg):
        self.params = Params(cfg)
        self.tag_per_joint = cfg.MODEL.TAG_PER_JOINT
        self._jo = {}

    def _joint_order(self, device, c):
        key = (str(device), c)
        if key not in self._jo:
            # the reference hands all 17 entries to assign(), which reads the first C (assign.cpp:81-82); entries that
            # do not name one of the C planes would index out of bounds there, so they are dropped here
            jo = [j for j in self.params.joint_order if j < c][:c]
            self._jo[key] = torch.tensor(jo, dtype=torch.int32, device=device)
        return self._jo[key]

    def parse_batch(self, det, tmap):
        """det [N,C,H,W], tmap [N,C,H,W,T] or [N,C,H,W] (device tensors) -> (num [N] int32, ans [N,M,C,4]) on the device."""
        p = self.params
        if tmap.dim() == 5:
            tmap = tmap[:, :, :, :, 0]
        det = det.contiguous().float()
        tmap = tmap.contiguous().float()
        count, val, tag, ind = plugins.find_peaks(det, tmap, p.detection_threshold, p.window_size, p.max_num_people)
        num, ans = plugins.assign(count, val, tag, ind, self._joint_order(det.device, det.shape[1]), p.tag_threshold,
                                  p.max_num_people)
        return num, ans

    def parse(self, det, tmap, scale):
        """nano_demo/fast_utils/group.py:38-47: image 0 only; returns ans[:num] with x, y multiplied by scale."""
        num, ans = self.parse_batch(det[:1], tmap[:1])
        ans = ans[0, :int(num[0].item())].clone()
        ans[:, :, :2] *= scale
        return ans
