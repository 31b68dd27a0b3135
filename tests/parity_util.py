"""Shared comparison helpers of the parity tests (CPU oracle tests and -m gpu tests)."""
import numpy as np


def assert_topk_equal(got, exp, what=""):
    """top-K results (dicts with val_k [N,J,K], loc_k [N,J,K,2], tag_k [N,J,K,T]) must agree on every slot with a
    positive value.  torch.topk leaves the order among EQUAL values unspecified (at 512x512 the
    U(0,0.02) float32 background of the planted maps holds exact ties far below DETECTION_THRESHOLD): values are
    compared on all slots, locations / tags on the slots whose value is unique within its (image, joint) row, and the
    tied slots must hold the same SET of locations."""
    gv, ev = np.asarray(got["val_k"]), np.asarray(exp["val_k"])
    m = ev > 0
    assert np.array_equal(m, gv > 0), what + " positive-slot mask"
    assert np.array_equal(gv[m], ev[m]), what + " val_k"
    n, j, k = ev.shape
    tied = np.zeros_like(m)
    for a in range(n):
        for b in range(j):
            v = ev[a, b]
            _, inv, cnt = np.unique(v, return_inverse=True, return_counts=True)
            tied[a, b] = (cnt[inv] > 1) | (v == v[-1])      # the K-th value may tie with a pixel left outside the list
    u = m & ~tied
    for key in ("loc_k", "tag_k"):
        g, e = np.asarray(got[key]), np.asarray(exp[key])
        assert np.array_equal(g[u], e[u]), "%s %s" % (what, key)
    g, e = np.asarray(got["loc_k"]), np.asarray(exp["loc_k"])
    for a, b in zip(*np.nonzero((m & tied).any(axis=2))):
        t = m[a, b] & tied[a, b] & (ev[a, b] != ev[a, b, -1])   # ties with the K-th value: only the values are determined
        gs = sorted(map(tuple, np.concatenate([gv[a, b][t][:, None], g[a, b][t]], axis=1).tolist()))
        es = sorted(map(tuple, np.concatenate([ev[a, b][t][:, None], e[a, b][t]], axis=1).tolist()))
        assert gs == es, "%s tied slots of image %d joint %d" % (what, a, b)
