"""CPU checks of the fast_utils oracle: the C restatement against golden vectors produced by the
reference's own native code, and against that code itself where oracle/_ref/libfastutils_ref.so is present."""
import os

import numpy as np
import pytest
import torch

from oracle import fast_utils_cases as cases
from oracle import fast_utils_ref as fu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fast_utils.npz")


def test_port_matches_golden():
    g = np.load(GOLD)
    for idx, (seed, kw, thr, win, m, tthr) in enumerate(cases.GOLDEN_CASES):
        det, tm = cases.make_case(seed, **kw)
        pre = "c%02d_" % idx
        count, val, tag, ind = fu.find_peaks(det, tm, thr, win, m, "port")
        for name, arr in (("count", count), ("val", val), ("tag", tag), ("ind", ind)):
            assert np.array_equal(arr, g[pre + name]), (idx, name)
        jo = cases.joint_order(det.shape[1])
        for i in range(det.shape[0]):
            num, ans, st = fu.assign(count[i], val[i], tag[i], ind[i], jo, tthr, m, "port")
            assert st == 0 and num == g[pre + "num"][i], (idx, i)
            assert np.array_equal(ans, g[pre + "ans"][i]), (idx, i)


@pytest.mark.skipif(not fu.available("ref") and not os.path.isdir("/root/reference"),
                    reason="compiled reference not present")
def test_port_matches_compiled_reference():
    checked = 0
    for seed in range(100, 160):
        det, tm = cases.make_case(seed, people=1 + seed % 6, spread=[2.0, 0.7, 0.3][seed % 3],
                                  tagnoise=[0.05, 0.3][seed % 2], plateau=seed % 5 == 0, clutter=[0, 5, 12][seed % 3])
        a = fu.find_peaks(det, tm, 0.1, 5, 30, "port")
        b = fu.find_peaks(det, tm, 0.1, 5, 30, "ref")
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        count, val, tag, ind = a
        for i in range(det.shape[0]):
            if count[i].max() > 10:
                continue
            n1, a1, s1 = fu.assign(count[i], val[i], tag[i], ind[i], cases.JOINT_ORDER_17, 1.0, 30, "port")
            if n1 > 10:      # the reference's [10] arrays would be overrun
                continue
            n2, a2, _ = fu.assign(count[i], val[i], tag[i], ind[i], cases.JOINT_ORDER_17, 1.0, 30, "ref")
            assert s1 == 0 and n1 == n2 and np.array_equal(a1, a2), (seed, i)
            checked += 1
    assert checked > 80


def test_port_beyond_reference_limit_terminates():
    det, tm = cases.make_case(7, people=20, clutter=30)
    count, val, tag, ind = fu.find_peaks(det, tm, 0.1, 5, 30, "port")
    assert count.max() > 10
    num, ans, st = fu.assign(count[0], val[0], tag[0], ind[0], cases.JOINT_ORDER_17, 1.0, 30, "port")
    assert st == 0 and 10 < num <= 30


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a CUDA device")
def test_plugin_mirror_has_no_cpu_fallback():
    from litepose_b200.fast_utils import plugins
    x = torch.zeros(1, 2, 8, 8)
    with pytest.raises(RuntimeError):
        plugins.find_peaks(x, x, 0.1, 5, 4)
