"""ORACLE (test infrastructure, not product code).

Pure-Python restatement of the Kuhn-Munkres ("Hungarian") assignment solver the
reference reaches through ``from munkres import Munkres`` (reference
lib/core/group.py:13, used at :19-23 ``py_max_match`` and :80).

The ``munkres`` PyPI package is a third-party dependency of the reference: it is
NOT vendored under /root/reference and its version is unpinned
(requirements.txt:12 is the bare word ``munkres``).  It is also absent from this
image (no network).  This file restates the package's published 6-step
algorithm (release line 1.1.x: ``pad_matrix`` -> step1..step6, with
``__find_a_zero(i0, j0)`` scanning cyclically and keeping the LAST uncovered zero
of the first row that has one) as described in SURVEY.md Appendix A.8.

PARITY UNPINNED at this boundary: no reference test, golden vector or fixture
pins the solver's output and the package itself cannot be diffed offline.  The
restatement below IS the oracle; it is property-checked against
``scipy.optimize.linear_sum_assignment`` on total cost (tests/test_munkres.py).
Because optimal assignments on the reference's cost matrices are massively
degenerate (cost = round(dist)*100 - val), the exact step sequence here, including
floating-point round-off in step 6, is part of the behaviour the CUDA matcher
(litepose_b200/csrc/match.cu) must reproduce.
"""


class Munkres(object):
    """Same public surface as ``munkres.Munkres``: ``compute(matrix) -> [(row, col)]``."""

    def compute(self, cost_matrix):
        # pad_matrix: square of side max(rows, max cols), padded with 0.
        # NB the package copies rows with ``row[:]`` which for a numpy row is a
        # view; we never rely on aliasing and work on a private float copy.
        rows = len(cost_matrix)
        cols = 0
        for r in cost_matrix:
            cols = max(cols, len(r))
        n = max(rows, cols)
        C = []
        for r in cost_matrix:
            row = [float(v) for v in r]
            if len(row) < n:
                row += [0.0] * (n - len(row))
            C.append(row)
        while len(C) < n:
            C.append([0.0] * n)
        self.C = C
        self.n = n
        self.row_covered = [False] * n
        self.col_covered = [False] * n
        self.Z0_r = 0
        self.Z0_c = 0
        self.path = [[0, 0] for _ in range(2 * n)]
        self.marked = [[0] * n for _ in range(n)]

        step = 1
        steps = {1: self._step1, 2: self._step2, 3: self._step3,
                 4: self._step4, 5: self._step5, 6: self._step6}
        while step in steps:
            step = steps[step]()

        results = []
        for i in range(rows):
            for j in range(cols):
                if self.marked[i][j] == 1:
                    results.append((i, j))
        return results

    # -- step 1: subtract the row minimum from every row
    def _step1(self):
        for i in range(self.n):
            m = min(self.C[i])
            for j in range(self.n):
                self.C[i][j] -= m
        return 2

    # -- step 2: greedy starring, row-major, one star per row, then clear covers
    def _step2(self):
        n = self.n
        for i in range(n):
            for j in range(n):
                if self.C[i][j] == 0 and not self.col_covered[j] and not self.row_covered[i]:
                    self.marked[i][j] = 1
                    self.col_covered[j] = True
                    self.row_covered[i] = True
                    break
        self._clear_covers()
        return 3

    # -- step 3: cover starred columns; all n covered -> done
    def _step3(self):
        n = self.n
        count = 0
        for i in range(n):
            for j in range(n):
                if self.marked[i][j] == 1 and not self.col_covered[j]:
                    self.col_covered[j] = True
                    count += 1
        return 7 if count >= n else 4

    # -- step 4: prime uncovered zeros
    def _step4(self):
        row = 0
        col = 0
        while True:
            row, col = self._find_a_zero(row, col)
            if row < 0:
                return 6
            self.marked[row][col] = 2
            star_col = self._find_star_in_row(row)
            if star_col >= 0:
                col = star_col
                self.row_covered[row] = True
                self.col_covered[col] = False
            else:
                self.Z0_r = row
                self.Z0_c = col
                return 5

    # -- step 5: augmenting path of alternating primes and stars
    def _step5(self):
        count = 0
        path = self.path
        path[0][0] = self.Z0_r
        path[0][1] = self.Z0_c
        while True:
            row = self._find_star_in_col(path[count][1])
            if row < 0:
                break
            count += 1
            path[count][0] = row
            path[count][1] = path[count - 1][1]
            col = self._find_prime_in_row(path[count][0])
            count += 1
            path[count][0] = path[count - 1][0]
            path[count][1] = col
        for i in range(count + 1):
            r, c = path[i]
            self.marked[r][c] = 0 if self.marked[r][c] == 1 else 1
        self._clear_covers()
        for i in range(self.n):
            for j in range(self.n):
                if self.marked[i][j] == 2:
                    self.marked[i][j] = 0
        return 3

    # -- step 6: add the smallest uncovered value to covered rows, subtract it
    #    from uncovered columns (both applied, in that order, per cell)
    def _step6(self):
        n = self.n
        minval = None
        for i in range(n):
            for j in range(n):
                if not self.row_covered[i] and not self.col_covered[j]:
                    if minval is None or self.C[i][j] < minval:
                        minval = self.C[i][j]
        if minval is None:
            raise RuntimeError("Matrix cannot be solved!")
        events = 0
        for i in range(n):
            for j in range(n):
                if self.row_covered[i]:
                    self.C[i][j] += minval
                    events += 1
                if not self.col_covered[j]:
                    self.C[i][j] -= minval
                    events += 1
                if self.row_covered[i] and not self.col_covered[j]:
                    events -= 2
        if events == 0:
            raise RuntimeError("Matrix cannot be solved!")
        return 4

    def _find_a_zero(self, i0, j0):
        n = self.n
        row = -1
        col = -1
        i = i0
        done = False
        while not done:
            j = j0
            while True:
                if self.C[i][j] == 0 and not self.row_covered[i] and not self.col_covered[j]:
                    row = i
                    col = j
                    done = True
                j = (j + 1) % n
                if j == j0:
                    break
            i = (i + 1) % n
            if i == i0:
                done = True
        return row, col

    def _find_star_in_row(self, row):
        for j in range(self.n):
            if self.marked[row][j] == 1:
                return j
        return -1

    def _find_star_in_col(self, col):
        for i in range(self.n):
            if self.marked[i][col] == 1:
                return i
        return -1

    def _find_prime_in_row(self, row):
        for j in range(self.n):
            if self.marked[row][j] == 2:
                return j
        return -1

    def _clear_covers(self):
        for i in range(self.n):
            self.row_covered[i] = False
            self.col_covered[i] = False
