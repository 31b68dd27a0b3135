"""Drop-in for the coordinate half of the reference's ``lib/utils/transforms.py`` used by valid.py:199-233
(post-processing side): ``get_multi_scale_size`` (:155-180), ``get_affine_transform``
(:60-98), ``affine_transform`` (:101-104), ``transform_preds`` (:50-57) and ``get_final_preds`` (:195-202).

No OpenCV: ``cv2.getAffineTransform`` (a 6x6 solve in double) is restated as OpenCV does it - interleaved x/y
equations, LU with partial pivoting, back substitution - and is bit-identical to cv2 4.13 on the reference's call
patterns (tests/golden/transforms.npz, generated from the unmodified reference).

``get_final_preds`` accepts what ``HeatmapParser.parse`` returns (host arrays) like the reference, or - for whole
batches - the packed DEVICE result of the pipeline, in which case the inverse affine of every keypoint runs in one
kernel (lp_transform_preds_f32) before the device->host copy.

``resize_align_multi_scale`` (:183-192) warps on the GPU: cv2.warpAffine's 8-bit fixed-point arithmetic is restated
exactly in lp_warp_affine_normalize_u8 (bit-identical images); ``resize_align_normalize_device`` additionally applies
torchvision's ToTensor + Normalize (valid.py:172-186,212) in the same kernel and leaves the NCHW tensor on the device.
"""
import numpy as np


def get_dir(src_point, rot_rad):
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    return [src_point[0] * cs - src_point[1] * sn, src_point[0] * sn + src_point[1] * cs]


def get_3rd_point(a, b):
    direct = a - b
    return b + np.array([-direct[1], direct[0]], dtype=np.float32)


def _affine_from_points(src, dst):
    """cv2.getAffineTransform(src, dst) for float32 [3,2] point sets -> float64 [2,3]."""
    a = np.zeros((6, 6), np.float64)
    b = np.zeros(6, np.float64)
    for i in range(3):
        a[2 * i, 0] = a[2 * i + 1, 3] = src[i, 0]
        a[2 * i, 1] = a[2 * i + 1, 4] = src[i, 1]
        a[2 * i, 2] = a[2 * i + 1, 5] = 1.0
        b[2 * i], b[2 * i + 1] = dst[i, 0], dst[i, 1]
    m = 6
    for i in range(m):
        k = i
        for j in range(i + 1, m):
            if abs(a[j, i]) > abs(a[k, i]):
                k = j
        if k != i:
            a[[i, k], i:] = a[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        if a[i, i] == 0.0:
            return np.zeros((2, 3), np.float64)           # singular: OpenCV returns zeros
        d = -1.0 / a[i, i]
        for j in range(i + 1, m):
            alpha = a[j, i] * d
            for c in range(i + 1, m):
                a[j, c] += alpha * a[i, c]
            b[j] += alpha * b[i]
    for i in range(m - 1, -1, -1):
        s = b[i]
        for c in range(i + 1, m):
            s -= a[i, c] * b[c]
        b[i] = s / a[i, i]
    return b.reshape(2, 3)


def _triangle(center, direction, offset):
    """The three float32 control points the reference builds: centre (+offset), centre + direction (+offset) - summed
    in the reference's order - and the point that makes a right angle of the two (get_3rd_point)."""
    pts = np.zeros((3, 2), dtype=np.float32)
    pts[0, :] = center + offset
    pts[1, :] = center + direction + offset
    pts[2:, :] = get_3rd_point(pts[0, :], pts[1, :])
    return pts


def get_affine_transform(center, scale, rot, output_size, shift=np.array([0, 0], dtype=np.float32), inv=0):
    """transforms.py:60-98 (same dtype promotions: float64 arithmetic, float32 control points)."""
    if not isinstance(scale, (np.ndarray, list)):
        scale = np.array([scale, scale])
    box = scale * 200.0
    out_w, out_h = output_size[0], output_size[1]
    src = _triangle(center, get_dir([0, box[0] * -0.5], np.pi * rot / 180), box * shift)
    # destination: centre of the output, direction straight up by half its width
    dst = np.zeros((3, 2), dtype=np.float32)
    dst[0, :] = [out_w * 0.5, out_h * 0.5]
    dst[1, :] = np.array([out_w * 0.5, out_h * 0.5]) + np.array([0, out_w * -0.5], np.float32)
    dst[2:, :] = get_3rd_point(dst[0, :], dst[1, :])
    return _affine_from_points(dst, src) if inv else _affine_from_points(src, dst)


def affine_transform(pt, t):
    """transforms.py:101-104."""
    return np.dot(t, np.array([pt[0], pt[1], 1.]).T)[:2]


def transform_preds(coords, center, scale, output_size):
    """transforms.py:50-57: rows of coords are (x, y, ...); x, y are mapped, the rest is kept."""
    trans = get_affine_transform(center, scale, 0, output_size, inv=1)
    out = coords.copy()
    for row in range(coords.shape[0]):
        out[row, 0:2] = affine_transform(coords[row, 0:2], trans)
    return out


def _up64(v):
    return int((v + 63) // 64 * 64)


def get_multi_scale_size(image, input_size, current_scale, min_scale):
    """transforms.py:155-180: the shorter image side becomes input_size (rounded up to 64), the longer one keeps the
    aspect ratio rounded up to 64; scale is the covered extent in units of 200 px."""
    h, w, _ = image.shape
    center = np.array([int(w / 2.0 + 0.5), int(h / 2.0 + 0.5)])
    short = _up64(min_scale * input_size)
    if w < h:
        w_r = int(short * current_scale / min_scale)
        h_r = int(_up64(short / w * h) * current_scale / min_scale)
        scale = np.array([w / 200.0, h_r / w_r * w / 200.0])
    else:
        h_r = int(short * current_scale / min_scale)
        w_r = int(_up64(short / h * w) * current_scale / min_scale)
        scale = np.array([w_r / h_r * h / 200.0, h / 200.0])
    return (w_r, h_r), center, scale


def get_final_preds(grouped_joints, center, scale, heatmap_size):
    """transforms.py:195-202; grouped_joints = what HeatmapParser.parse returned ([persons]); image 0."""
    return [transform_preds(person, center, scale, heatmap_size) for person in grouped_joints[0]]


def invert_affine(m):
    """The dst->src matrix cv2.warpAffine derives from the src->dst one (OpenCV imgwarp.cpp, same operation order)."""
    m = [float(v) for v in np.asarray(m, np.float64).ravel()]
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0], m[1], m[3], m[4] = a11, m[1] * -d, m[3] * -d, a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return np.array(m, np.float64)


def _warp(images, sizes_hw, input_size, current_scale, min_scale, mode, mean=None, std=None):
    import torch

    from litepose_b200 import _lib
    if not torch.cuda.is_available():
        raise RuntimeError("litepose_b200 pre-processing runs on a CUDA device (no CPU fallback)")
    if torch.is_tensor(images):
        img = images
    else:
        img = torch.from_numpy(np.ascontiguousarray(images))
    if img.dim() == 3:
        img = img.unsqueeze(0)
    if img.dtype != torch.uint8 or img.shape[3] != 3:
        raise TypeError("images must be uint8 [N,H,W,3] (or [H,W,3])")
    n, h, w, _ = img.shape
    size, center, scale = get_multi_scale_size(np.empty((h, w, 3), np.uint8), input_size, current_scale, min_scale)
    minv = invert_affine(get_affine_transform(center, scale, 0, size))
    dev = img.device if img.is_cuda else torch.device("cuda", torch.cuda.current_device())
    d_img = img.to(dev, non_blocking=True).contiguous()
    d_m = torch.from_numpy(np.tile(minv, (n, 1))).to(dev)
    shape = (n, size[1], size[0], 3) if mode == 0 else (n, 3, size[1], size[0])
    out = torch.empty(shape, dtype=(torch.uint8, torch.float32, torch.float16)[mode], device=dev)
    mean_a = np.asarray(mean if mean is not None else (0, 0, 0), np.float32)
    std_a = np.asarray(std if std is not None else (1, 1, 1), np.float32)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.lp_warp_affine_normalize_u8(d_img.data_ptr(), n, h, w, d_m.data_ptr(), size[0], size[1],
                                                   mean_a.ctypes.data, std_a.ctypes.data, out.data_ptr(), mode,
                                                   torch.cuda.current_stream(dev).cuda_stream),
                   "lp_warp_affine_normalize_u8")
    return out, center, scale


def resize_align_multi_scale(image, input_size, current_scale, min_scale):
    """transforms.py:183-192, same return types: (uint8 image [h_resized, w_resized, 3] on the host, center, scale).
    The warp itself runs on the GPU."""
    out, center, scale = _warp(image, None, input_size, current_scale, min_scale, 0)
    return out[0].cpu().numpy(), center, scale


def resize_align_normalize_device(images, input_size, current_scale, min_scale, mean, std, half=False):
    """valid.py:210-213 for a batch of equally sized images without leaving the device: warp, ToTensor, Normalize ->
    NCHW float32 (or float16 = what tofp16 would make of it) CUDA tensor, plus center and scale (shared: same size)."""
    return _warp(images, None, input_size, current_scale, min_scale, 2 if half else 1, mean, std)


def final_preds_device(ans, num, centers, scales, heatmap_size):
    """Whole batch on the device, in place: ans [N,P,J,3+T] float32 CUDA (DeviceParser / pipeline layout), num [N]
    int32 CUDA; centers [N,2], scales [N,2] host arrays.  Every keypoint of every found person gets the inverse
    affine of its image (the same float64 arithmetic as affine_transform, rounded to float32 on store)."""
    import torch

    from litepose_b200 import _lib
    n = ans.shape[0]
    trans = np.stack([get_affine_transform(np.asarray(centers[i]), np.asarray(scales[i]), 0, heatmap_size, inv=1)
                      for i in range(n)]).reshape(n, 6)
    t = torch.from_numpy(np.ascontiguousarray(trans)).to(ans.device)
    lib = _lib.load()
    with torch.cuda.device(ans.device):
        _lib.check(lib.lp_transform_preds_f32(ans.data_ptr(), num.data_ptr(), t.data_ptr(), n, ans.shape[1],
                                              ans.shape[2], ans.shape[3],
                                              torch.cuda.current_stream(ans.device).cuda_stream),
                   "lp_transform_preds_f32")
    return ans


def __getattr__(name):
    """Names of the reference's utils/transforms.py that this module does not restate (flip_back, fliplr_joints, crop,
    resize: training / dataset helpers, no caller on the inference path) resolve from the reference module found
    further down sys.path, when there is one."""
    import importlib.util
    import os
    import sys
    here = os.path.abspath(__file__)
    for base in sys.path:
        cand = os.path.join(base, "utils", "transforms.py")
        if base and os.path.exists(cand) and os.path.abspath(cand) != here:
            spec = importlib.util.spec_from_file_location("_reference_utils_transforms", cand)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            if hasattr(mod, name):
                return getattr(mod, name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
