"""ORACLE - TEST INFRASTRUCTURE ONLY.  Generates tests/golden/transforms.npz from the UNMODIFIED reference
lib/utils/transforms.py (needs /root/reference and cv2: build container only):  python -m oracle.make_golden_transforms

Cases: image shapes x input sizes x scale factors -> get_multi_scale_size, get_affine_transform (forward, inverse),
and get_final_preds on seeded keypoints."""
import importlib.util
import os

import numpy as np

CASES = [  # (h, w, input_size, current_scale, min_scale)
    (480, 640, 512, 1.0, 1.0), (640, 480, 512, 1.0, 1.0), (427, 640, 512, 1.0, 1.0), (333, 500, 448, 1.0, 1.0),
    (1080, 1920, 640, 1.0, 1.0), (500, 375, 256, 1.0, 1.0), (480, 640, 512, 2.0, 0.5), (480, 640, 512, 0.5, 0.5),
    (612, 612, 512, 1.0, 1.0), (97, 1234, 512, 1.0, 1.0),
]


def keypoints(seed, persons, joints, t, hm_w, hm_h):
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(persons):
        a = np.zeros((joints, 3 + t), np.float32)
        a[:, 0] = rs.rand(joints).astype(np.float32) * hm_w
        a[:, 1] = rs.rand(joints).astype(np.float32) * hm_h
        a[:, 2:] = rs.rand(joints, 1 + t).astype(np.float32)
        out.append(a)
    return out


PRE_CASES = [  # (seed, h, w, input_size): small images, full outputs stored
    (1, 97, 123, 128), (2, 150, 101, 128), (3, 64, 64, 64), (4, 45, 200, 64), (5, 240, 320, 256),
]
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]          # valid.py:181-184


def pre_image(seed, h, w):
    rs = np.random.RandomState(seed)
    img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
    img[h // 4:h // 2, w // 4:w // 2] = rs.randint(0, 256, 3).astype(np.uint8)      # a flat patch
    return img


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_transforms", "/root/reference/lib/utils/transforms.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    out = {}
    for i, (h, w, size, cur, mn) in enumerate(CASES):
        img = np.zeros((h, w, 3), np.uint8)
        (wr, hr), center, scale = ref.get_multi_scale_size(img, size, cur, mn)
        fwd = ref.get_affine_transform(center, scale, 0, (wr, hr))
        inv = ref.get_affine_transform(center, scale, 0, [wr, hr], inv=1)
        persons = keypoints(100 + i, 1 + i % 4, 14, 2, wr, hr)
        final = ref.get_final_preds([persons], center, scale, [wr, hr])
        pre = "c%02d_" % i
        out[pre + "size"] = np.array([wr, hr], np.int64)
        out[pre + "center"], out[pre + "scale"] = np.asarray(center), np.asarray(scale)
        out[pre + "fwd"], out[pre + "inv"] = fwd, inv
        out[pre + "final"] = np.stack(final)
    import torchvision.transforms as tvt
    tf = tvt.Compose([tvt.ToTensor(), tvt.Normalize(mean=MEAN, std=STD)])       # valid.py:178-186
    for i, (seed, h, w, size) in enumerate(PRE_CASES):
        img = pre_image(seed, h, w)
        resized, center, scale = ref.resize_align_multi_scale(img, size, 1.0, 1.0)
        pre = "p%02d_" % i
        out[pre + "resized"] = resized
        out[pre + "tensor"] = tf(resized).numpy()
        out[pre + "center"], out[pre + "scale"] = np.asarray(center), np.asarray(scale)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "transforms.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
