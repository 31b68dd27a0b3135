"""ORACLE - TEST INFRASTRUCTURE ONLY.  Generates tests/golden/fast_utils.npz from the reference's own native code
compiled where it lies (oracle/_ref/libfastutils_ref.so, recipe oracle/native/build_native.py).  Run in the build
container (needs /root/reference):  python -m oracle.make_golden_fast_utils"""
import os

import numpy as np

from . import fast_utils_cases as cases
from . import fast_utils_ref as fu
from .native import build_native


def main():
    built = build_native.build()
    assert "ref" in built, "the reference sources are not available here"
    out = {}
    for idx, (seed, kw, thr, win, m, tthr) in enumerate(cases.GOLDEN_CASES):
        det, tm = cases.make_case(seed, **kw)
        count, val, tag, ind = fu.find_peaks(det, tm, thr, win, m, "ref")
        assert count.max(initial=0) <= 10
        jo = cases.joint_order(det.shape[1])
        nums, anss = [], []
        for i in range(det.shape[0]):
            num, ans, _ = fu.assign(count[i], val[i], tag[i], ind[i], jo, tthr, m, "ref")
            assert num <= 10
            nums.append(num)
            anss.append(ans)
        pre = "c%02d_" % idx
        out[pre + "count"], out[pre + "val"], out[pre + "tag"], out[pre + "ind"] = count, val, tag, ind
        out[pre + "num"] = np.asarray(nums, np.int32)
        out[pre + "ans"] = np.stack(anss)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fast_utils.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
