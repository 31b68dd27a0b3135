"""Deterministic synthetic weights, frames and planted crowds.

There is no dataset and no checkpoint in this environment, and a random-init
LitePose detects nobody (heat-maps ~ +-0.03 < DETECTION_THRESHOLD), so the
grouping stage is exercised with planted det/tag maps.  The same tensors are fed
to the oracle, the reference arm and the CUDA path.
"""
import numpy as np
import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def randomize_bn_(model, seed=1):
    """Non-trivial BatchNorm statistics so the BN fold is exercised:
    running_mean~N(0,0.1^2), running_var~U(0.75,1.25), weight~U(0.75,1.25),
    bias~N(0,0.1^2)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                c = m.num_features
                m.running_mean.copy_(torch.randn(c, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(c, generator=g) * 0.5 + 0.75)
                m.weight.copy_(torch.rand(c, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(c, generator=g) * 0.1)
    return model


def scale_heads_(model, factor=0.05):
    """Scale the bias-free output 1x1 convs so that a random-weight network's heat-maps stay
    below DETECTION_THRESHOLD (|heat| ~ 0.02): the planted persons are then the only detections."""
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.startswith(("final_refined", "final_raw")) and name.endswith("conv.3.weight"):
                p.mul_(factor)
    return model


def make_frames(n, size, seed=1234, rank=0, width=None):
    """``torch.rand(N,3,S,S)`` then ImageNet normalisation (valid.py:181-184)."""
    g = torch.Generator().manual_seed(seed + rank)
    w = size if width is None else width
    x = torch.rand(n, 3, size, w, generator=g)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    return (x - mean) / std


def plant_crowd(num_joints, h, w, t, num_people=5, seed=0, presence=0.9,
                sigma=2.0, spread=15, margin=20):
    """Planted det [J,H,W] / tag [J,H,W,T] maps for one image (float32 numpy).

    background det~U(0,0.02), tag~N(0,0.05^2); person p: centre uniform in
    [margin, H-margin), each joint present w.p. ``presence`` at centre+U{-spread..spread},
    Gaussian (sigma) amplitude U(0.5,1) max-composited into det, 9x9 tag patch
    = 2*p + N(0,0.05^2) on all T maps.
    """
    rng = np.random.RandomState(seed)
    det = rng.uniform(0.0, 0.02, size=(num_joints, h, w)).astype(np.float32)
    tag = (rng.randn(num_joints, h, w, t) * 0.05).astype(np.float32)
    m = min(margin, h // 4, w // 4)
    sp = min(spread, max(1, m - 5))
    r = int(3 * sigma)
    yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
    gauss = np.exp(-(xx ** 2 + yy ** 2) / (2.0 * sigma * sigma)).astype(np.float32)
    for p in range(num_people):
        cy = rng.randint(m, h - m)
        cx = rng.randint(m, w - m)
        for j in range(num_joints):
            present = rng.rand() < presence
            dy = rng.randint(-sp, sp + 1)
            dx = rng.randint(-sp, sp + 1)
            amp = np.float32(rng.uniform(0.5, 1.0))
            if not present:
                continue
            y = int(np.clip(cy + dy, r, h - r - 1))
            x = int(np.clip(cx + dx, r, w - r - 1))
            patch = det[j, y - r:y + r + 1, x - r:x + r + 1]
            np.maximum(patch, amp * gauss, out=patch)
            t0 = (2.0 * p + rng.randn(9, 9, t) * 0.05).astype(np.float32)
            tag[j, y - 4:y + 5, x - 4:x + 5, :] = t0[: min(9, h - y + 4), : min(9, w - x + 4), :]
    return det, tag


def plant_crowd_batch(n, num_joints, h, w, t, num_people=5, seed=0, **kw):
    dets, tags = [], []
    for i in range(n):
        d, g = plant_crowd(num_joints, h, w, t, num_people=num_people, seed=seed + i, **kw)
        dets.append(d)
        tags.append(g)
    return np.stack(dets), np.stack(tags)
