"""Oracle Munkres restatement: self-pinned fixtures + optimality vs scipy."""
import os

import numpy as np
import pytest

from oracle.munkres_ref import Munkres


def test_golden_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "munkres_cases.npz"))
    for full, sol in zip(z["mats"], z["sols"]):
        r = int((~np.isnan(full[:, 0])).sum())
        c = int((~np.isnan(full[0, :])).sum())
        m = full[:r, :c]
        pairs = Munkres().compute(m.copy())
        exp = [tuple(p) for p in sol if p[0] >= 0]
        assert pairs == exp


def test_optimal_total_cost_vs_scipy():
    scipy_opt = pytest.importorskip("scipy.optimize")
    rng = np.random.RandomState(0)
    for trial in range(300):
        r, c = rng.randint(1, 10), rng.randint(1, 10)
        if trial % 2:
            m = rng.choice([0.0, 100.0, 200.0], size=(r, c)) - rng.uniform(0.1, 1, size=(r, 1))
        else:
            m = rng.uniform(0, 10, size=(r, c))
        pairs = Munkres().compute(m.copy())
        assert len(pairs) == min(r, c)
        assert len({p[0] for p in pairs}) == len(pairs) and len({p[1] for p in pairs}) == len(pairs)
        ri, ci = scipy_opt.linear_sum_assignment(m)
        assert abs(sum(m[i, j] for i, j in pairs) - m[ri, ci].sum()) < 1e-6 * max(1.0, abs(m).max())


def test_rectangular_and_trivial():
    assert Munkres().compute([[5.0]]) == [(0, 0)]
    assert Munkres().compute([[1.0, 0.0, 2.0]]) == [(0, 1)]
    assert sorted(Munkres().compute([[1.0], [0.0], [2.0]])) == [(1, 0)]
