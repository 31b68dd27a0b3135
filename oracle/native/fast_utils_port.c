/* ORACLE - TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's cpu legs; never by
 * the product path).
 *
 * CPU restatement in C of the reference's native "fast inference" grouping:
 *   find peaks : nano_demo/fast_utils/parse/find_peaks.cpp:9-57   (per plane, first-M peaks in scan order)
 *                nano_demo/fast_utils/parse/find_peaks.cpp:59-97  (CHW / NCHW loops)
 *   assign     : nano_demo/fast_utils/parse/assign.cpp:11-13 (dist), :15-31 (match), :33-43 (update),
 *                :45-66 (KM), :68-122 (assign_out)
 *
 * Pinned: tests/test_fast_utils_oracle.py compares this file bit for bit with the reference's own two .cpp files
 * compiled where they lie (oracle/_ref/libfastutils_ref.so, recipe oracle/native/build_native.py) and with golden
 * vectors produced by that build (tests/golden/fast_utils_*.npz).
 *
 * Deliberate differences from the reference, none of which changes a result the reference defines:
 *   - the reference's fixed [10] stack arrays (assign.cpp:46-48,79-80) become [FU_MAXP]=32; with more than 10
 *     candidates or persons the reference writes out of bounds (undefined), this port is defined up to 32;
 *   - `abs(t) < 1e-2` (assign.cpp:22): with the reference's includes (<cmath>, <cstdio>) and g++ 13 the call binds to
 *     `int abs(int)` (probe: oracle/native/probe_abs.cpp), i.e. the test is trunc(t) == 0; restated as such;
 *   - the `while(true)` of KM (assign.cpp:58-62) has no termination guarantee (update() takes the minimum over ALL
 *     slack entries and slack is never refreshed inside the loop: reaching a padded -1e4 column typically takes
 *     thousands of rounds of the smallest slack); this port stops after FU_MAX_ROUNDS rounds per row and reports
 *     status 1, where the reference would keep spinning.
 */
#include <math.h>
#include <stdint.h>

#define FU_MAXP 32
#define FU_MAX_ROUNDS (1 << 20)

#define FU_MIN(a, b) ((a) < (b) ? (a) : (b))
#define FU_MAX(a, b) ((a) > (b) ? (a) : (b))

/* find_peaks.cpp:9-57 */
static void port_find_peaks_hw(int* count, float* val, float* tag, int* ind, const float* input, const float* tmap,
                               int H, int W, int M, float threshold, int window_size) {
    const int win = window_size / 2;
    int cnt = 0;
    for (int i = 0; i < H && cnt < M; i++)
        for (int j = 0; j < W && cnt < M; j++) {
            const float hval = input[i * W + j];
            if (hval < threshold) continue;
            const int ii_min = FU_MAX(i - win, 0), jj_min = FU_MAX(j - win, 0);
            const int ii_max = FU_MIN(i + win + 1, H), jj_max = FU_MIN(j + win + 1, W);
            int is_peak = 1;
            for (int ii = ii_min; ii < ii_max; ii++)
                for (int jj = jj_min; jj < jj_max; jj++)
                    if (input[ii * W + jj] > hval) is_peak = 0;
            if (is_peak) {
                ind[cnt * 2] = j;
                ind[cnt * 2 + 1] = i;
                val[cnt] = hval;
                tag[cnt] = tmap[i * W + j];
                cnt++;
            }
        }
    *count = cnt;
}

/* find_peaks.cpp:59-97 */
void port_find_peaks_nchw(int* count, float* val, float* tag, int* ind, const float* input, const float* tmap, int N,
                          int C, int H, int W, int M, float threshold, int window_size) {
    for (int p = 0; p < N * C; p++)
        port_find_peaks_hw(count + p, val + (long)p * M, tag + (long)p * M, ind + (long)p * M * 2,
                           input + (long)p * H * W, tmap + (long)p * H * W, H, W, M, threshold, window_size);
}

/* assign.cpp:11-13 */
static float port_dist(float x, float y) { return sqrtf((x - y) * (x - y)); }

typedef struct {
    int n;
    int mat[FU_MAXP];
    float G[FU_MAXP][FU_MAXP];
    float Lx[FU_MAXP], Ly[FU_MAXP], slack[FU_MAXP];
    unsigned char S[FU_MAXP], T[FU_MAXP];
} port_km;

/* assign.cpp:15-31 (recursive, as in the reference) */
static int port_match(port_km* k, int u) {
    k->S[u] = 1;
    for (int i = 0; i < k->n; i++) {
        if (k->T[i]) continue;
        const float t = k->Lx[u] + k->Ly[i] - k->G[u][i];
        if ((int)t == 0) {          /* `abs(t) < 1e-2` with int abs(int): see header */
            k->T[i] = 1;
            if (k->mat[i] == -1 || port_match(k, k->mat[i])) {
                k->mat[i] = u;
                return 1;
            }
        } else
            k->slack[i] = FU_MIN(k->slack[i], t);
    }
    return 0;
}

/* assign.cpp:33-43 */
static void port_update(port_km* k) {
    float d = 1e8f;
    for (int i = 0; i < k->n; i++) d = FU_MIN(d, k->slack[i]);
    for (int i = 0; i < k->n; i++) {
        if (k->S[i]) k->Lx[i] -= d;
        if (k->T[i]) k->Ly[i] += d;
    }
}

/* assign.cpp:45-66; returns 0, or 1 when the round cap was hit (the reference would not return) */
static int port_KM(int* ch, port_km* k) {
    const int n = k->n;
    for (int i = 0; i < n; i++) {
        k->Lx[i] = -1e6f;
        k->Ly[i] = 0;
        k->mat[i] = -1;
        for (int j = 0; j < n; j++) k->Lx[i] = FU_MAX(k->Lx[i], k->G[i][j]);
    }
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) k->slack[j] = 1e6f;
        int rounds = 0;
        for (;;) {
            for (int j = 0; j < n; j++) k->S[j] = k->T[j] = 0;
            if (port_match(k, i)) break;
            if (++rounds >= FU_MAX_ROUNDS) return 1;
            port_update(k);
        }
    }
    for (int i = 0; i < n; i++) ch[k->mat[i]] = i;
    return 0;
}

/* assign.cpp:68-122; one image.  Returns the KM status (0 ok). */
int port_assign(int* num_person, float* ans, const int* cnt, const float* val, const float* tag, const int* ind,
                const int* joint_order, int C, int M, float threshold) {
    int num = 0, nj[FU_MAXP], ch[FU_MAXP];
    float diff[FU_MAXP][FU_MAXP], sum[FU_MAXP];
    port_km km;
    for (int id = 0; id < C; id++) {
        const int i = joint_order[id];
        if (cnt[i] == 0) continue;
        if (num == 0) {
            num = cnt[i];
            for (int j = 0; j < num; j++) {
                const int p = i * M + j, q = (j * C + i) << 2;
                ans[q] = (float)ind[p << 1];
                ans[q | 1] = (float)ind[(p << 1) | 1];
                ans[q | 2] = val[p];
                ans[q | 3] = tag[p];
                nj[j] = 1;
                sum[j] = tag[p];
            }
            continue;
        }
        const int num_add = FU_MAX(num, cnt[i]);
        for (int j = 0; j < num_add; j++)
            for (int k = 0; k < num_add; k++) {
                const int pre = i * M + k;
                if (j >= num || k >= cnt[i]) {
                    diff[j][k] = 1e4f;
                    km.G[j][k] = -1e4f;
                } else {
                    /* `1.0 * sum[j] / nj[j]` is a double quotient narrowed to the float parameter of dist() */
                    const float mean = (float)(1.0 * sum[j] / nj[j]);
                    const float d = port_dist(mean, tag[pre]);
                    diff[j][k] = d;
                    km.G[j][k] = -(d * 100 - val[pre]);
                }
            }
        km.n = num_add;
        if (port_KM(ch, &km)) { *num_person = num; return 1; }
        const int old_num = num;
        for (int j = 0; j < num_add; j++) {
            if (ch[j] >= cnt[i]) continue;
            if (j < old_num && ch[j] < cnt[i] && diff[j][ch[j]] < threshold) {
                const int p = i * M + ch[j], q = (j * C + i) << 2;
                ans[q] = (float)ind[p << 1];
                ans[q | 1] = (float)ind[(p << 1) | 1];
                ans[q | 2] = val[p];
                ans[q | 3] = tag[p];
                nj[j]++;
                sum[j] += tag[p];
            } else {
                if (num == M) continue;
                const int p = i * M + ch[j], q = (num * C + i) << 2;
                ans[q] = (float)ind[p << 1];
                ans[q | 1] = (float)ind[(p << 1) | 1];
                ans[q | 2] = val[p];
                ans[q | 3] = tag[p];
                nj[num] = 1;
                sum[num] = tag[p];
                num++;
            }
        }
    }
    *num_person = num;
    return 0;
}
