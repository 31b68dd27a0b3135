"""Drop-in replacement of the reference's ``lib/core/group.py``.

``HeatmapParser(cfg).parse(det, tag, adjust, refine)`` has the reference's signature
and return types (reference lib/core/group.py:123-129,269-291) but executes on the
GPU through the C ABI (litepose_b200.parser.DeviceParser); ``parse_batch`` runs a
whole batch as N independent reference calls.  CUDA tensors only.
"""
import numpy as np
import torch

from litepose_b200.parser import DeviceParser


class Params(object):
    """Same fields as the reference's Params (group.py:100-120)."""

    def __init__(self, cfg):
        self.num_joints = cfg.DATASET.NUM_JOINTS
        self.max_num_people = cfg.DATASET.MAX_NUM_PEOPLE
        self.detection_threshold = cfg.TEST.DETECTION_THRESHOLD
        self.tag_threshold = cfg.TEST.TAG_THRESHOLD
        self.use_detection_val = cfg.TEST.USE_DETECTION_VAL
        self.ignore_too_much = cfg.TEST.IGNORE_TOO_MUCH
        with_center = cfg.DATASET.WITH_CENTER
        if with_center and cfg.TEST.IGNORE_CENTER:
            self.num_joints -= 1
        body = [1, 2, 3, 4, 5, 6, 7, 12, 13, 8, 9, 10, 11, 14, 15, 16, 17]
        order = ([18] + body) if (with_center and not cfg.TEST.IGNORE_CENTER) else body
        self.joint_order = [i - 1 for i in order]


class HeatmapParser(object):
    def __init__(self, cfg):
        self.params = Params(cfg)
        self.tag_per_joint = cfg.MODEL.TAG_PER_JOINT
        p = self.params
        self.device_parser = DeviceParser(p.num_joints, p.max_num_people, p.detection_threshold, p.tag_threshold,
                                          p.use_detection_val, p.ignore_too_much, p.joint_order,
                                          cfg.TEST.NMS_KERNEL, cfg.TEST.NMS_PADDING)

    def _tags(self, det, tag):
        """[N,Jt,H,W(,T)] -> float32 [N,J,H,W,T].  MODEL.TAG_PER_JOINT=False: the network emits ONE tag map shared by all
        joints; the reference expands it over the joints for the gather (group.py:150-152) and means to tile it for
        refine (group.py:283-286 - where its own code stops on an unassigned name as soon as a person is found, so
        that branch of the reference returns nothing to compare with: the tiled maps are what it was written to use)."""
        if tag.dim() == 4:
            tag = tag.unsqueeze(4)
        if not self.tag_per_joint:
            if tag.shape[1] != 1:
                raise ValueError("TAG_PER_JOINT=False expects one tag map per image (got %d)" % tag.shape[1])
            tag = tag.expand(-1, det.shape[1], -1, -1, -1)
        return tag.float().contiguous()

    # -- reference sub-methods (public names kept; no external callers in the reference)
    def top_k(self, det, tag):
        """group.py:141-176 -> dict of numpy arrays (canonical tie order, see include/litepose_b200.h)."""
        val_k, ind_k, tag_k = self.device_parser.top_k_device(det.float(), self._tags(det, tag))
        w = det.shape[3]
        ind = ind_k.cpu().numpy().astype(np.int64)
        return {"tag_k": tag_k.cpu().numpy(), "loc_k": np.stack((ind % w, ind // w), axis=3),
                "val_k": val_k.cpu().numpy()}

    def parse_batch(self, det, tag, adjust=True, refine=True):
        """N images -> list of N (ans, scores) pairs, each exactly what the reference's
        ``parse(det[i:i+1], tag[i:i+1])`` returns."""
        res = self.device_parser.run(det.float(), self._tags(det, tag), adjust, refine)
        return [([a], s) for a, s in DeviceParser.to_reference(*res)]

    def parse(self, det, tag, adjust=True, refine=True):
        """group.py:269-291: scores and refine use image 0 only, like the reference."""
        out = self.parse_batch(det, tag, adjust, refine)
        ans0, scores0 = out[0]
        if refine:
            return ans0, scores0
        # without refine the reference returns every image's match result
        return [o[0][0] for o in out], scores0
