// Fused test-time glue between model and parser ("next" row 1 of SURVEY.md §8f).
// Reference: get_multi_stage_outputs + aggregate_results, lib/core/inference.py:75-208, for the evaluation
// configuration of record (SCALE_FACTOR [1], LOSS.WITH_HEATMAPS_LOSS (1,1), TEST.WITH_HEATMAPS (1,1),
// WITH_AE (1,0), TAG_PER_JOINT):
//   plain pass : u0 = bilinear(o0 -> size of o1); ha = (u0[:J] + o1) / 2 ; t0 = u0[J:]
//   flip pass  : uf = flip_x(bilinear(f0)), g1 = flip_x(f1); hf = (uf[:J][fidx] + g1[fidx]) / 2 ; t1 = uf[J:][fidx]
//   project    : each of ha, hf, t0, t1 -> bilinear to (Hd, Wd)   (PROJECT2IMAGE; identity when Hd,Wd == 2h,2w)
//   aggregate  : det = (ha + hf) / 2 (or ha when no flip) ; tag = stack(t0, t1) on the last dim
// The eager reference materialises ~10 full-resolution intermediates; this kernel reads the four model outputs once
// (through L1/L2) and writes det/tag once: algorithmic bytes 4*N*J*(Hd*Wd*(1+T)) written + 4*N*(3J*h*w*5)*passes read.
// Bilinear follows ATen's align_corners=False rule: src = (dst+0.5)*in/out - 0.5 clamped at 0, i1 = min(i0+1, in-1).
#include "common.cuh"

namespace lp {

constexpr int GL_TO = 64;          // output tile side (halved by the host until the mid-level footprint fits GL_TM)
constexpr int GL_TM = 40;          // max mid-level tile side kept in shared memory
constexpr int GL_THREADS = 256;

struct Lerp {
    int i0, i1;
    float l0, l1;
};

__device__ __forceinline__ Lerp lerp_coord(int dst, float scale, int in_size) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    Lerp r;
    r.i0 = min((int)src, in_size - 1);
    r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
    r.l1 = src - (float)r.i0;
    r.l0 = 1.f - r.l1;
    return r;
}

__device__ __forceinline__ float bilerp(const float* __restrict__ p, int W, const Lerp& y, const Lerp& x) {
    return y.l0 * (x.l0 * __ldg(p + y.i0 * W + x.i0) + x.l1 * __ldg(p + y.i0 * W + x.i1)) +
           y.l1 * (x.l0 * __ldg(p + y.i1 * W + x.i0) + x.l1 * __ldg(p + y.i1 * W + x.i1));
}

__global__ void __launch_bounds__(GL_THREADS)
glue_kernel(const float* __restrict__ o0, const float* __restrict__ o1, const float* __restrict__ f0,
            const float* __restrict__ f1, const int32_t* __restrict__ flip_index, int J, int Jm, int tag_shared, int h, int w,
            int flip, int Hd, int Wd, int tiles_x, int to, int accumulate, float divide_by, float* __restrict__ det,
            float* __restrict__ tag) {
    __shared__ float s_ha[GL_TM][GL_TM + 1], s_hf[GL_TM][GL_TM + 1], s_t0[GL_TM][GL_TM + 1], s_t1[GL_TM][GL_TM + 1];
    const int n = blockIdx.z, j = blockIdx.y;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int H2 = 2 * h, W2 = 2 * w;
    const int T = flip ? 2 : 1;
    const bool project = !(Hd == H2 && Wd == W2);
    const float sy = (float)H2 / (float)Hd, sx = (float)W2 / (float)Wd;
    const int oy0 = ty * to, ox0 = tx * to;
    const int oy1 = min(oy0 + to, Hd) - 1, ox1 = min(ox0 + to, Wd) - 1;
    // mid-level footprint of this output tile
    int my0, my1, mx0, mx1;
    if (project) {
        my0 = lerp_coord(oy0, sy, H2).i0; my1 = lerp_coord(oy1, sy, H2).i1;
        mx0 = lerp_coord(ox0, sx, W2).i0; mx1 = lerp_coord(ox1, sx, W2).i1;
    } else {
        my0 = oy0; my1 = oy1; mx0 = ox0; mx1 = ox1;
    }
    const int mh = my1 - my0 + 1, mw = mx1 - mx0 + 1;   // <= GL_TM guaranteed by the host (choice of `to`) unless !project
    const float us_y = (float)h / (float)H2, us_x = (float)w / (float)W2;   // 0.5

    // channel layout of the model outputs (pose_mobilenet.py:86-100): o0 = [Jm heat | Jm tags, or ONE tag map when
    // MODEL.TAG_PER_JOINT is off], o1 = [Jm heat]; Jm > J when the centre joint is ignored (inference.py:147-150).  The
    // shared tag map is written once (by the j == 0 blocks) and is not permuted in the flip pass (inference.py:141-144).
    const int C0 = tag_shared ? Jm + 1 : 2 * Jm;
    if (tag_shared && j != 0) tag = nullptr;
    const size_t hw = (size_t)h * w, HW2 = (size_t)H2 * W2;
    const float* a_heat = o0 + ((size_t)n * C0 + j) * hw;
    const float* a_tag = o0 + ((size_t)n * C0 + Jm + (tag_shared ? 0 : j)) * hw;
    const float* a_o1 = o1 + ((size_t)n * Jm + j) * HW2;
    const float *b_heat = nullptr, *b_tag = nullptr, *b_o1 = nullptr;
    if (flip) {
        const int fj = flip_index[j];
        b_heat = f0 + ((size_t)n * C0 + fj) * hw;
        b_tag = f0 + ((size_t)n * C0 + Jm + (tag_shared ? 0 : fj)) * hw;
        b_o1 = f1 + ((size_t)n * Jm + fj) * HW2;
    }

    if (project) {
        for (int i = threadIdx.x; i < mh * mw; i += GL_THREADS) {
            const int ly = i / mw, lx = i - ly * mw;
            const int y = my0 + ly, x = mx0 + lx;
            const Lerp cy = lerp_coord(y, us_y, h), cx = lerp_coord(x, us_x, w);
            s_ha[ly][lx] = (bilerp(a_heat, w, cy, cx) + __ldg(a_o1 + (size_t)y * W2 + x)) / 2.f;
            s_t0[ly][lx] = bilerp(a_tag, w, cy, cx);
            if (flip) {
                const int xf = W2 - 1 - x;
                const Lerp fx = lerp_coord(xf, us_x, w);
                s_hf[ly][lx] = (bilerp(b_heat, w, cy, fx) + __ldg(b_o1 + (size_t)y * W2 + xf)) / 2.f;
                s_t1[ly][lx] = bilerp(b_tag, w, cy, fx);
            }
        }
        __syncthreads();
    }

    float* dplane = det + ((size_t)n * J + j) * Hd * Wd;
    float* tplane = tag ? tag + ((size_t)n * (tag_shared ? 1 : J) + j) * Hd * Wd * T : nullptr;
    const int th = oy1 - oy0 + 1, tw = ox1 - ox0 + 1;

    // multi-scale aggregation (inference.py:176-208): det accumulates over the scales, the division by the number of
    // scales (valid.py:223) rides on the last call; scales other than 1 write no tags (tag == nullptr)
    const bool plain = !accumulate && divide_by == 1.f && tag != nullptr;
    if (project && Hd == 2 * H2 && Wd == 2 * W2 && flip && plain) {
        // exact x2 projection (the evaluation config): out[2m] = .25 s[m-1] + .75 s[m], out[2m+1] = .75 s[m] + .25 s[m+1]
        // (edge-clamped).  One thread produces a 2x2 output quad from a 3x3 mid-level neighbourhood: 9 LDS per map
        // instead of 16, constant weights, 16-byte tag stores.
        const int qw = tw >> 1, qh = th >> 1;       // Hd, Wd and the tile origin are even
        for (int i = threadIdx.x; i < qh * qw; i += GL_THREADS) {
            const int qy = i / qw, qx = i - qy * qw;
            const int m = (oy0 >> 1) + qy, k = (ox0 >> 1) + qx;              // mid-level pixel of this quad
            const int r0 = max(m - 1, 0) - my0, r1 = m - my0, r2 = min(m + 1, H2 - 1) - my0;
            const int c0 = max(k - 1, 0) - mx0, c1 = k - mx0, c2 = min(k + 1, W2 - 1) - mx0;
            float o[4][4];   // [map][quad pixel: (0,0) (0,1) (1,0) (1,1)]
#define GL_Q(S, idx)                                                                                     \
    {                                                                                                    \
        const float a0 = S[r0][c0], a1 = S[r0][c1], a2 = S[r0][c2];                                      \
        const float b0 = S[r1][c0], b1 = S[r1][c1], b2 = S[r1][c2];                                      \
        const float d0 = S[r2][c0], d1 = S[r2][c1], d2 = S[r2][c2];                                      \
        const float al = 0.25f * a0 + 0.75f * a1, ar = 0.75f * a1 + 0.25f * a2;                          \
        const float bl = 0.25f * b0 + 0.75f * b1, br = 0.75f * b1 + 0.25f * b2;                          \
        const float dl = 0.25f * d0 + 0.75f * d1, dr = 0.75f * d1 + 0.25f * d2;                          \
        o[idx][0] = 0.25f * al + 0.75f * bl;                                                             \
        o[idx][1] = 0.25f * ar + 0.75f * br;                                                             \
        o[idx][2] = 0.75f * bl + 0.25f * dl;                                                             \
        o[idx][3] = 0.75f * br + 0.25f * dr;                                                             \
    }
            GL_Q(s_ha, 0) GL_Q(s_hf, 1) GL_Q(s_t0, 2) GL_Q(s_t1, 3)
#undef GL_Q
            const int Y = oy0 + 2 * qy, X = ox0 + 2 * qx;
            const size_t o0i = (size_t)Y * Wd + X, o1i = o0i + Wd;
            *reinterpret_cast<float2*>(dplane + o0i) = make_float2((o[0][0] + o[1][0]) / 2.f, (o[0][1] + o[1][1]) / 2.f);
            *reinterpret_cast<float2*>(dplane + o1i) = make_float2((o[0][2] + o[1][2]) / 2.f, (o[0][3] + o[1][3]) / 2.f);
            *reinterpret_cast<float4*>(tplane + o0i * 2) = make_float4(o[2][0], o[3][0], o[2][1], o[3][1]);
            *reinterpret_cast<float4*>(tplane + o1i * 2) = make_float4(o[2][2], o[3][2], o[2][3], o[3][3]);
        }
        return;
    }

    for (int i = threadIdx.x; i < th * tw; i += GL_THREADS) {
        const int ly = i / tw, lx = i - ly * tw;
        const int Y = oy0 + ly, X = ox0 + lx;
        float ha, hf = 0.f, t0, t1 = 0.f;
        if (project) {
            const Lerp cy = lerp_coord(Y, sy, H2), cx = lerp_coord(X, sx, W2);
            const int y0 = cy.i0 - my0, y1 = cy.i1 - my0, x0 = cx.i0 - mx0, x1 = cx.i1 - mx0;
#define GL_P(S) (cy.l0 * (cx.l0 * S[y0][x0] + cx.l1 * S[y0][x1]) + cy.l1 * (cx.l0 * S[y1][x0] + cx.l1 * S[y1][x1]))
            ha = GL_P(s_ha);
            t0 = GL_P(s_t0);
            if (flip) {
                hf = GL_P(s_hf);
                t1 = GL_P(s_t1);
            }
#undef GL_P
        } else {
            const Lerp cy = lerp_coord(Y, us_y, h), cx = lerp_coord(X, us_x, w);
            ha = (bilerp(a_heat, w, cy, cx) + __ldg(a_o1 + (size_t)Y * W2 + X)) / 2.f;
            t0 = bilerp(a_tag, w, cy, cx);
            if (flip) {
                const int xf = W2 - 1 - X;
                const Lerp fx = lerp_coord(xf, us_x, w);
                hf = (bilerp(b_heat, w, cy, fx) + __ldg(b_o1 + (size_t)Y * W2 + xf)) / 2.f;
                t1 = bilerp(b_tag, w, cy, fx);
            }
        }
        const size_t o = (size_t)Y * Wd + X;
        float v = flip ? (ha + hf) / 2.f : ha;
        if (accumulate) v = dplane[o] + v;
        if (divide_by != 1.f) v = v / divide_by;
        dplane[o] = v;
        if (tag != nullptr) {
            if (flip) *reinterpret_cast<float2*>(tplane + o * 2) = make_float2(t0, t1);
            else tplane[o] = t0;
        }
    }
}

// ---- exact 4x path (the evaluation config: flip test + PROJECT2IMAGE to 4h x 4w) -------------------------------
// One thread produces a 4x4 output block of one (n, j) plane for all four maps entirely in registers: 3x3
// neighbourhoods of the quarter-resolution maps + 4x4 neighbourhoods of the half-resolution maps, both x2 bilinear
// stages with their constant weights (edge clamping == clamping the source index), flip = mirrored block with
// reversed columns.  No shared memory, no barriers; 16-byte stores.
__device__ __forceinline__ void up4(const float a, const float b, const float c, float (&o)[4]) {
    // mid rows/cols (2i-1, 2i, 2i+1, 2i+2) of a x2 up-sample from neighbours (i-1, i, i+1)
    o[0] = 0.75f * a + 0.25f * b;
    o[1] = 0.25f * a + 0.75f * b;
    o[2] = 0.75f * b + 0.25f * c;
    o[3] = 0.25f * b + 0.75f * c;
}
__device__ __forceinline__ void proj4(const float (&m)[4], float (&o)[4]) {
    // outputs (4i .. 4i+3) from mid samples (2i-1 .. 2i+2)
    o[0] = 0.25f * m[0] + 0.75f * m[1];
    o[1] = 0.75f * m[1] + 0.25f * m[2];
    o[2] = 0.25f * m[1] + 0.75f * m[2];
    o[3] = 0.75f * m[2] + 0.25f * m[3];
}

// mid[4][4] = U(q)[rows 2a-1..2a+2][cols 2b-1..2b+2]   (q: quarter-resolution plane, clamped 3x3 neighbourhood)
__device__ __forceinline__ void mid_from_quarter(const float* __restrict__ q, int w, int ra, int rb, int rc, int ca, int cb,
                                                 int cc, float (&mid)[4][4]) {
    float r[3][4];
    const int rows[3] = {ra, rb, rc};
#pragma unroll
    for (int i = 0; i < 3; ++i) up4(__ldg(q + rows[i] * w + ca), __ldg(q + rows[i] * w + cb), __ldg(q + rows[i] * w + cc), r[i]);
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        float col[4];
        up4(r[0][x], r[1][x], r[2][x], col);
#pragma unroll
        for (int y = 0; y < 4; ++y) mid[y][x] = col[y];
    }
}

__global__ void __launch_bounds__(256)
glue_x4_kernel(const float* __restrict__ o0, const float* __restrict__ o1, const float* __restrict__ f0,
               const float* __restrict__ f1, const int32_t* __restrict__ flip_index, int J, int h, int w,
               float* __restrict__ det, float* __restrict__ tag) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;     // quarter-resolution column
    const int a = blockIdx.y;                                // quarter-resolution row
    const int n = blockIdx.z / J, j = blockIdx.z - n * J;
    if (b >= w) return;
    const int H2 = 2 * h, W2 = 2 * w, Hd = 4 * h, Wd = 4 * w;
    const size_t hw = (size_t)h * w, HW2 = (size_t)H2 * W2;
    const int fj = flip_index[j];
    const int ra = max(a - 1, 0), rc = min(a + 1, h - 1);
    const int bf = w - 1 - b;                                // mirrored block of the flipped pass
    float ha[4][4], hf[4][4], t0[4][4], t1[4][4], tmp[4][4];

    // plain pass
    mid_from_quarter(o0 + ((size_t)n * 2 * J + j) * hw, w, ra, a, rc, max(b - 1, 0), b, min(b + 1, w - 1), ha);
    mid_from_quarter(o0 + ((size_t)n * 2 * J + J + j) * hw, w, ra, a, rc, max(b - 1, 0), b, min(b + 1, w - 1), t0);
    {
        const float* p1 = o1 + ((size_t)n * J + j) * HW2;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int yy = min(max(2 * a - 1 + y, 0), H2 - 1);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int xx = min(max(2 * b - 1 + x, 0), W2 - 1);
                ha[y][x] = (ha[y][x] + __ldg(p1 + (size_t)yy * W2 + xx)) / 2.f;
            }
        }
    }
    // flipped pass: same block pattern at the mirrored column block, columns reversed, channels permuted
    mid_from_quarter(f0 + ((size_t)n * 2 * J + fj) * hw, w, ra, a, rc, max(bf - 1, 0), bf, min(bf + 1, w - 1), tmp);
    {
        const float* p1 = f1 + ((size_t)n * J + fj) * HW2;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int yy = min(max(2 * a - 1 + y, 0), H2 - 1);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int xx = min(max(2 * bf - 1 + x, 0), W2 - 1);
                hf[y][3 - x] = (tmp[y][x] + __ldg(p1 + (size_t)yy * W2 + xx)) / 2.f;
            }
        }
    }
    mid_from_quarter(f0 + ((size_t)n * 2 * J + J + fj) * hw, w, ra, a, rc, max(bf - 1, 0), bf, min(bf + 1, w - 1), tmp);
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int x = 0; x < 4; ++x) t1[y][3 - x] = tmp[y][x];

    // projection x2 (rows then columns) and aggregation
    float od[4][4], ot0[4][4], ot1[4][4];
#define GL_PROJ(SRC, DST)                                                 \
    {                                                                     \
        float v[4][4];                                                    \
        _Pragma("unroll") for (int x = 0; x < 4; ++x) {                   \
            const float m[4] = {SRC[0][x], SRC[1][x], SRC[2][x], SRC[3][x]}; \
            float o[4];                                                   \
            proj4(m, o);                                                  \
            _Pragma("unroll") for (int y = 0; y < 4; ++y) v[y][x] = o[y]; \
        }                                                                 \
        _Pragma("unroll") for (int y = 0; y < 4; ++y) proj4(v[y], DST[y]); \
    }
    float pa[4][4], pf[4][4];
    GL_PROJ(ha, pa) GL_PROJ(hf, pf) GL_PROJ(t0, ot0) GL_PROJ(t1, ot1)
#undef GL_PROJ
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int x = 0; x < 4; ++x) od[y][x] = (pa[y][x] + pf[y][x]) / 2.f;

    float* dplane = det + ((size_t)n * J + j) * Hd * Wd;
    float* tplane = tag + ((size_t)n * J + j) * Hd * Wd * 2;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
        const size_t o = (size_t)(4 * a + y) * Wd + 4 * b;
        *reinterpret_cast<float4*>(dplane + o) = make_float4(od[y][0], od[y][1], od[y][2], od[y][3]);
        *reinterpret_cast<float4*>(tplane + o * 2) = make_float4(ot0[y][0], ot1[y][0], ot0[y][1], ot1[y][1]);
        *reinterpret_cast<float4*>(tplane + o * 2 + 4) = make_float4(ot0[y][2], ot1[y][2], ot0[y][3], ot1[y][3]);
    }
}

}  // namespace lp

using namespace lp;

// picks the output tile side so that the mid-level footprint of a tile fits the shared tile; 0 = ratio not supported
static int glue_tile_side(int h, int w, int Hd, int Wd) {
    for (int to = GL_TO; to >= 8; to >>= 1)
        if ((double)(2 * h) / Hd * to + 3 <= GL_TM && (double)(2 * w) / Wd * to + 3 <= GL_TM) return to;
    return 0;
}

static int glue_launch(const char* who, const float* o0, const float* o1, const float* f0, const float* f1,
                       const int32_t* flip_index, int N, int J, int Jm, int tag_shared, int h, int w, int flip, int Hd,
                       int Wd, int accumulate, float divide_by, float* det, float* tag, bool allow_x4, lp_stream_t stream) {
    LP_CHECK_ARG(o0 && o1 && det, "%s: null pointer", who);
    LP_CHECK_ARG(!flip || (f0 && f1 && flip_index), "%s: flip pass needs f0, f1, flip_index", who);
    LP_CHECK_ARG(N > 0 && N <= 65535 && J > 0 && J <= 65535 && h > 0 && w > 0 && Hd > 0 && Wd > 0,
                 "%s: bad shape N=%d J=%d h=%d w=%d Hd=%d Wd=%d", who, N, J, h, w, Hd, Wd);
    LP_CHECK_ARG(Jm >= J, "%s: model_joints=%d must be >= J=%d", who, Jm, J);
    LP_CHECK_ARG(divide_by > 0.f, "%s: divide_by must be > 0", who);
    allow_x4 = allow_x4 && Jm == J && !tag_shared;
    const bool project = !(Hd == 2 * h && Wd == 2 * w);
    int to = GL_TO;
    if (project) {
        // the shared mid-level tile must cover the footprint of one output tile
        to = glue_tile_side(h, w, Hd, Wd);
        LP_CHECK_ARG(to > 0, "%s: projection %dx%d -> %dx%d shrinks by more than 4.6x (not supported)", who, 2 * h, 2 * w,
                     Hd, Wd);
    }
    if (flip && ((tag && (reinterpret_cast<uintptr_t>(tag) & 15)) || (reinterpret_cast<uintptr_t>(det) & 7))) {
        set_error("%s: tag must be 16-byte and det 8-byte aligned", who);
        return LP_ERR_ALIGN;
    }
    if (allow_x4 && flip && Hd == 4 * h && Wd == 4 * w && (long long)N * J <= 65535 && h <= 65535) {
        dim3 grid((w + 127) / 128, h, N * J);
        glue_x4_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(o0, o1, f0, f1, flip_index, J, h, w, det, tag);
        LP_LAUNCH_CHECK("glue_x4_kernel");
        return LP_OK;
    }
    const int tiles_x = (Wd + to - 1) / to, tiles_y = (Hd + to - 1) / to;
    dim3 grid(tiles_x * tiles_y, J, N);
    glue_kernel<<<grid, GL_THREADS, 0, (cudaStream_t)stream>>>(o0, o1, f0, f1, flip_index, J, Jm, tag_shared ? 1 : 0, h, w, flip,
                                                             Hd, Wd, tiles_x, to, accumulate, divide_by, det, tag);
    LP_LAUNCH_CHECK("glue_kernel");
    return LP_OK;
}

extern "C" int lp_glue_f32(const float* o0, const float* o1, const float* f0, const float* f1, const int32_t* flip_index,
                           int N, int J, int h, int w, int flip, int Hd, int Wd, float* det, float* tag,
                           lp_stream_t stream) {
    LP_CHECK_ARG(tag, "lp_glue_f32: null pointer");
    return glue_launch("lp_glue_f32", o0, o1, f0, f1, flip_index, N, J, J, 0, h, w, flip, Hd, Wd, 0, 1.f, det, tag, true,
                       stream);
}

extern "C" int lp_glue_scale_f32(const float* o0, const float* o1, const float* f0, const float* f1,
                                 const int32_t* flip_index, int N, int J, int model_joints, int tag_shared, int h, int w,
                                 int flip, int Hd, int Wd, int accumulate, float divide_by, float* det, float* tag,
                                 lp_stream_t stream) {
    return glue_launch("lp_glue_scale_f32", o0, o1, f0, f1, flip_index, N, J, model_joints, tag_shared, h, w, flip, Hd, Wd,
                       accumulate ? 1 : 0, divide_by, det, tag, !accumulate && divide_by == 1.f && tag != nullptr, stream);
}
