// Stem: conv3x3 stride 2 (3 -> 32) + folded-BN bias + ReLU6; reads the reference's NCHW input
// (fp32, or fp16 under network_to_half) and writes the NHWC fp16 activation layout used by all
// later kernels.  Reference: convbnrelu(3, 32, ker=3, stride=2), lib/models/pose_mobilenet.py:37,
// lib/models/layers/layers.py:18-24.
//
// CTA = 64 x 8 output pixels; the 3-channel haloed input patch is staged in shared memory as
// fp32 (batched global loads).  Each thread computes two x-adjacent output pixels x 32 channels;
// weights are read as warp-uniform float4 broadcasts from shared memory (1 LDS.128 per 8 FFMA).
// Output: 64 B contiguous per pixel (4 x 16-byte stores).
#include "common.cuh"

namespace lp {

constexpr int ST_TW = 64, ST_TH = 8;                          // output tile; each thread owns 2 x-adjacent pixels
constexpr int ST_IH = ST_TH * 2 + 1;                          // 17 input rows (stride 2, pad 1)
constexpr int ST_IW = ST_TW * 2 + 1;                          // 129 input columns actually used
// the patch is fetched as NB column chunks of 128 bytes (TMA boxes with a wider inner extent faulted on B200)
template <typename TIn> struct StBox {
    static constexpr int W = 128 / sizeof(TIn);                  // columns per chunk: 64 (fp16) / 32 (fp32)
    static constexpr int NB = (ST_IW + W - 1) / W;               // 3 / 5 chunks
};

// The 3-plane haloed input patch arrives with ONE TMA tensor copy (OOB zero fill = conv padding; the box of the flip
// pass is taken from the mirrored column range and read backwards), so the load is a single asynchronous transaction
// instead of a chain of dependent global loads.
template <typename TIn>
__global__ void __launch_bounds__(256)
stem_kernel(const __grid_constant__ CUtensorMap map_x, const __half* __restrict__ w, const float* __restrict__ bias,
            __half* __restrict__ y, int H, int W, int flip_x) {
    constexpr int BW = StBox<TIn>::W, NB = StBox<TIn>::NB;
    __shared__ __align__(128) TIn s_in[NB][3][ST_IH][BW];
    __shared__ __align__(16) float s_w[27][32];   // [tap][co]
    __shared__ __align__(16) float s_b[32];
    __shared__ __align__(8) uint64_t bar;
    const int Ho = H / 2, Wo = W / 2;
    const int n = blockIdx.z;
    const int ox0 = blockIdx.x * ST_TW, oy0 = blockIdx.y * ST_TH;
    const int ix0 = ox0 * 2 - 1, iy0 = oy0 * 2 - 1;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        mbar_expect_tx(&bar, NB * 3 * ST_IH * BW * (int)sizeof(TIn));
        // tile-local column c lives at patch column c (plain) or ST_IW-1-c (flip: box starts at the mirror of ix0+128)
        const int xs = flip_x ? (W - ST_IW - ix0) : ix0;
        for (int b = 0; b < NB; ++b)
            tma_load_4d(&s_in[b][0][0][0], &map_x, &bar, xs + b * BW, iy0, 0, n);
    }
    for (int i = threadIdx.x; i < 27 * 32; i += 256) {
        const int co = i & 31, t = i >> 5;
        s_w[t][co] = __half2float(w[co * 27 + t]);
    }
    if (threadIdx.x < 32) s_b[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
    __syncthreads();
    mbar_wait(&bar, 0);

    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;    // pixels (2*lx, 2*lx+1) of row ly
    float acc0[32], acc1[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc0[i] = acc1[i] = s_b[i];
    const int cbase = flip_x ? (ST_IW - 1 - 4 * lx) : 4 * lx;
    const int cstep = flip_x ? -1 : 1;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            // the two pixels need input columns 4*lx .. 4*lx+4
            float v[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int col = cbase + q * cstep;
                v[q] = (float)s_in[col / BW][c][2 * ly + ky][col % BW];
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4* wr = reinterpret_cast<const float4*>(s_w[c * 9 + ky * 3 + kx]);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 wv = wr[g];
                    acc0[4 * g + 0] = fmaf(v[kx], wv.x, acc0[4 * g + 0]);
                    acc0[4 * g + 1] = fmaf(v[kx], wv.y, acc0[4 * g + 1]);
                    acc0[4 * g + 2] = fmaf(v[kx], wv.z, acc0[4 * g + 2]);
                    acc0[4 * g + 3] = fmaf(v[kx], wv.w, acc0[4 * g + 3]);
                    acc1[4 * g + 0] = fmaf(v[kx + 2], wv.x, acc1[4 * g + 0]);
                    acc1[4 * g + 1] = fmaf(v[kx + 2], wv.y, acc1[4 * g + 1]);
                    acc1[4 * g + 2] = fmaf(v[kx + 2], wv.z, acc1[4 * g + 2]);
                    acc1[4 * g + 3] = fmaf(v[kx + 2], wv.w, acc1[4 * g + 3]);
                }
            }
        }
    const int oy = oy0 + ly;
    if (oy < Ho) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int ox = ox0 + 2 * lx + p;
            if (ox >= Wo) continue;
            const float* acc = p ? acc1 : acc0;
            uint4* op = reinterpret_cast<uint4*>(y + (((size_t)n * Ho + oy) * Wo + ox) * 32);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint4 o;
                __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    h[i] = __floats2half2_rn(fminf(fmaxf(acc[8 * g + 2 * i], 0.f), 6.f),
                                             fminf(fmaxf(acc[8 * g + 2 * i + 1], 0.f), 6.f));
                op[g] = o;
            }
        }
    }
}

template <typename TIn>
static int launch_stem(const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int flip_x,
                       cudaStream_t stream) {
    CUtensorMap map;
    uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, 3u, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)W * sizeof(TIn), (uint64_t)H * W * sizeof(TIn), (uint64_t)3 * H * W * sizeof(TIn)};
    uint32_t box[4] = {(uint32_t)StBox<TIn>::W, (uint32_t)ST_IH, 3u, 1u};
    int rc = make_tmap(&map, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                       sizeof(TIn) == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    if (rc) return rc;
    dim3 grid((W / 2 + ST_TW - 1) / ST_TW, (H / 2 + ST_TH - 1) / ST_TH, N);
    stem_kernel<TIn><<<grid, 256, 0, stream>>>(map, reinterpret_cast<const __half*>(w), bias, reinterpret_cast<__half*>(y),
                                               H, W, flip_x);
    LP_LAUNCH_CHECK("stem_kernel");
    return LP_OK;
}

}  // namespace lp

using namespace lp;

extern "C" int lp_stem_conv3x3_s2(const void* x, int x_is_fp32, int flip_x, const void* w, const float* bias, void* y,
                                  int N, int H, int W, lp_stream_t stream) {
    LP_CHECK_ARG(x && w && y, "lp_stem_conv3x3_s2: null pointer");
    LP_CHECK_ARG(N > 0 && N <= 65535 && H > 0 && W > 0 && H % 2 == 0 && W % 8 == 0,
                 "lp_stem_conv3x3_s2: bad shape N=%d H=%d W=%d (H even, W %% 8 == 0)", N, H, W);
    if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(x)) & 15) {
        set_error("lp_stem_conv3x3_s2: x and y must be 16-byte aligned");
        return LP_ERR_ALIGN;
    }
    if (x_is_fp32) return launch_stem<float>(x, w, bias, y, N, H, W, flip_x, (cudaStream_t)stream);
    return launch_stem<__half>(x, w, bias, y, N, H, W, flip_x, (cudaStream_t)stream);
}
