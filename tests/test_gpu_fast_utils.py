"""GPU parity of the fast_utils plugin (lp_find_peaks_f32 / lp_assign_f32 through the plugin mirror) against golden
vectors of the reference's native code and against the C restatement: bit-exact."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from litepose_b200 import _lib
from litepose_b200.fast_utils import plugins
from litepose_b200.fast_utils.group import HeatmapParser
from oracle import fast_utils_cases as cases
from oracle import fast_utils_ref as fu

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "fast_utils.npz")


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_golden_vectors():
    g = np.load(GOLD)
    for idx, (seed, kw, thr, win, m, tthr) in enumerate(cases.GOLDEN_CASES):
        det, tm = cases.make_case(seed, **kw)
        pre = "c%02d_" % idx
        count, val, tag, ind = plugins.find_peaks(_dev(det), _dev(tm), thr, win, m)
        for name, t in (("count", count), ("val", val), ("tag", tag), ("ind", ind)):
            assert np.array_equal(t.cpu().numpy(), g[pre + name]), (idx, name)
        jo = torch.tensor(cases.joint_order(det.shape[1]), dtype=torch.int32, device="cuda")
        num, ans = plugins.assign(count, val, tag, ind, jo, tthr, m)          # batched
        assert int(plugins.last_status().abs().sum()) == 0
        assert np.array_equal(num.cpu().numpy(), g[pre + "num"]), idx
        assert np.array_equal(ans.cpu().numpy(), g[pre + "ans"]), idx
        # the reference's single-image call shape
        n0, a0 = plugins.assign(count[0], val[0], tag[0], ind[0], jo, tthr, m)
        assert int(n0[0]) == g[pre + "num"][0] and np.array_equal(a0.cpu().numpy(), g[pre + "ans"][0])


@pytest.mark.parametrize("people,clutter,m", [(3, 10, 30), (12, 20, 30), (20, 30, 30), (28, 40, 30), (28, 40, 32), (9, 0, 6)])
def test_against_port_beyond_reference_limit(people, clutter, m):
    det, tm = cases.make_case(1000 + people, n=4, people=people, spread=[2.0, 0.6][people % 2], tagnoise=0.2,
                              plateau=True, clutter=clutter, h=64, w=56)
    exp = fu.find_peaks(det, tm, 0.1, 5, m, "port")
    got = plugins.find_peaks(_dev(det), _dev(tm), 0.1, 5, m)
    for e, t in zip(exp, got):
        assert np.array_equal(t.cpu().numpy(), e)
    jo = torch.tensor(cases.JOINT_ORDER_17, dtype=torch.int32, device="cuda")
    num, ans = plugins.assign(*got, jo, 1.0, m)
    st = plugins.last_status().cpu().numpy()
    for i in range(det.shape[0]):
        n1, a1, s1 = fu.assign(exp[0][i], exp[1][i], exp[2][i], exp[3][i], cases.JOINT_ORDER_17, 1.0, m, "port")
        assert s1 == st[i] == 0
        assert n1 == int(num[i]) and np.array_equal(a1, ans[i].cpu().numpy()), i


def test_out_variants_leave_the_rest_untouched_and_cpu_tensors_round_trip():
    det, tm = cases.make_case(5, people=3)
    n, c, m = det.shape[0], det.shape[1], 12
    count = torch.full((n, c), -7, dtype=torch.int32, device="cuda")
    val = torch.full((n, c, m), -7.0, device="cuda")
    tag = torch.full((n, c, m), -7.0, device="cuda")
    ind = torch.full((n, c, m, 2), -7, dtype=torch.int32, device="cuda")
    plugins.find_peaks_out(count, val, tag, ind, _dev(det), _dev(tm), 0.1, 5, m)
    exp = fu.find_peaks(det, tm, 0.1, 5, m, "port")
    cnt = count.cpu().numpy()
    assert np.array_equal(cnt, exp[0])
    v = val.cpu().numpy()
    for i in range(n):
        for j in range(c):
            k = cnt[i, j]
            assert np.array_equal(v[i, j, :k], exp[1][i, j, :k]) and (v[i, j, k:] == -7.0).all()
    # CPU tensors in (what the reference takes): processed on the device, returned on the CPU
    out = plugins.find_peaks(torch.from_numpy(det), torch.from_numpy(tm), 0.1, 5, m)
    assert all(not t.is_cuda for t in out) and np.array_equal(out[0].numpy(), exp[0])
    assert np.array_equal(out[3].numpy(), exp[3])


def test_error_reporting():
    lib = _lib.load()
    z = torch.zeros(64, dtype=torch.int32, device="cuda")
    f = torch.zeros(64 * 40 * 4, device="cuda")
    rc = lib.lp_assign_f32(z.data_ptr(), f.data_ptr(), f.data_ptr(), z.data_ptr(), z.data_ptr(), 1, 2, 40, 1.0,
                           z.data_ptr(), f.data_ptr(), z.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 5 and b"exceeds" in lib.lp_last_error()          # LP_ERR_CAPACITY
    with pytest.raises(TypeError):
        plugins.find_peaks(torch.zeros(1, 2, 8, 8, dtype=torch.float64, device="cuda"),
                           torch.zeros(1, 2, 8, 8, device="cuda"), 0.1, 5, 4)


def test_group_parser_mirror():
    cfg = NS(DATASET=NS(NUM_JOINTS=17, MAX_NUM_PEOPLE=30, WITH_CENTER=False),
             TEST=NS(DETECTION_THRESHOLD=0.1, TAG_THRESHOLD=1.0, USE_DETECTION_VAL=True, IGNORE_TOO_MUCH=False,
                     NMS_KERNEL=5, IGNORE_CENTER=True),
             MODEL=NS(TAG_PER_JOINT=True))
    det, tm = cases.make_case(21, n=3, people=5, clutter=6)
    parser = HeatmapParser(cfg)
    params = dict(detection_threshold=0.1, window_size=5, max_num_people=30, tag_threshold=1.0,
                  joint_order=cases.JOINT_ORDER_17)
    exp = fu.parse(det, tm[..., None], params, "port")
    num, ans = parser.parse_batch(_dev(det), _dev(tm[..., None]))
    for i, (n1, a1, s1) in enumerate(exp):
        assert s1 == 0 and n1 == int(num[i]) and np.array_equal(a1, ans[i].cpu().numpy())
    one = parser.parse(_dev(det), _dev(tm[..., None]), 4.0)        # reference contract: image 0, coords * scale
    ref0 = exp[0][1][:exp[0][0]].copy()
    ref0[:, :, :2] *= 4.0
    assert np.array_equal(one.cpu().numpy(), ref0)


@pytest.mark.parametrize("seed", list(range(12)))
def test_assign_label_walks_randomised(seed):
    """Round 2: the KM kernel skips rounds of a label walk (same failed search, labels crawling towards a padded column) by
    re-evaluating only the pairs the search looks at - the result must stay bit-identical to the round-by-round C
    restatement.  Cases with missing joints (long walks), close tags (small d), clutter and plateaus."""
    rs = np.random.RandomState(seed)
    people = int(rs.randint(2, 14))
    det, tm = cases.make_case(7000 + seed, n=3, people=people, spread=float(rs.choice([0.3, 0.6, 2.0])),
                              tagnoise=float(rs.choice([0.02, 0.2, 0.6])), plateau=bool(seed & 1), clutter=int(rs.randint(0, 25)),
                              h=64, w=56)
    m = int(rs.choice([16, 30]))
    exp = fu.find_peaks(det, tm, 0.1, 5, m, "port")
    got = plugins.find_peaks(_dev(det), _dev(tm), 0.1, 5, m)
    jo = torch.tensor(cases.JOINT_ORDER_17, dtype=torch.int32, device="cuda")
    tthr = float(rs.choice([0.5, 1.0, 2.0]))
    num, ans = plugins.assign(*got, jo, tthr, m)
    st = plugins.last_status().cpu().numpy()
    for i in range(det.shape[0]):
        n1, a1, s1 = fu.assign(exp[0][i], exp[1][i], exp[2][i], exp[3][i], cases.JOINT_ORDER_17, tthr, m, "port")
        assert s1 == st[i]
        if s1 == 0:
            assert n1 == int(num[i]) and np.array_equal(a1, ans[i].cpu().numpy()), (seed, i)
