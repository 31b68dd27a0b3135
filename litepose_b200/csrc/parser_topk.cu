// G1+G2: heat-map NMS + top-K peak pick on the device (reference lib/core/group.py:131-176).
//   nms:   det * (maxpool_kxk(det) == det), -inf padding                  (group.py:131-135)
//   top_k: K largest of the H*W NMS'd values per (n, j), tags gathered     (group.py:141-176)
// Canonical order (torch.topk's tie order is unspecified): value desc, flat index asc over
// survivors with value > 0; unused slots are (0.0f, index 0)  -- same rule as oracle/group_ref.py.
//
// HBM-bound: det is read once (algorithmic bytes 4*N*J*H*W).  Kernel 1: one CTA per strip of
// rows of one plane; separable window max in shared memory, then K rounds of a block-wide
// arg-max over 64-bit keys (value bits << 32 | ~index) using warp-shuffle reductions.
// Kernel 2: one warp per plane merges the per-strip sorted lists and gathers the tags.
#include "common.cuh"

namespace lp {

constexpr int TK_THREADS = 256;
constexpr int TK_MAXK = 64;

__device__ __forceinline__ unsigned long long shfl_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    return v;
}

static inline int strip_rows(int W) {
    // keep both shared planes (rows + halo) under ~96 KB
    int sr = 16;
    while (sr > 2 && (size_t)(sr + 8) * W * 8 > 96 * 1024) sr >>= 1;
    return sr;
}

constexpr int TK_LIST = 2048;   // compacted survivor list per strip (overflow -> dense scan fallback)

// partial: [N*J][strips][K] keys
__global__ void __launch_bounds__(TK_THREADS)
nms_topk_strip_kernel(const float* __restrict__ det, int H, int W, int R /*window radius*/, int SR, int K,
                      unsigned long long* __restrict__ partial) {
    extern __shared__ __align__(16) float sm[];
    const int plane = blockIdx.y;
    const int strip = blockIdx.x, nstrips = gridDim.x;
    const int y0 = strip * SR;
    const int rows = min(SR, H - y0);
    const int hrows = rows + 2 * R;                // rows incl. halo (out-of-image rows hold -inf)
    float* s_val = sm;                             // [SR+2R][W] raw values
    float* s_hmax = sm + (size_t)(SR + 2 * R) * W; // [SR+2R][W] horizontal window max
    __shared__ unsigned long long s_list[TK_LIST];
    __shared__ unsigned long long s_red[TK_THREADS / 32];
    __shared__ unsigned long long s_win;
    __shared__ int s_count;
    const float* p = det + (size_t)plane * H * W;
    const float NEG_INF = __int_as_float(0xff800000);
    if (threadIdx.x == 0) s_count = 0;

    // rows are walked with x = tid + m*256 so no integer division is needed
    for (int r = 0; r < hrows; ++r) {
        const int gy = y0 - R + r;
        const bool in = gy >= 0 && gy < H;
        for (int x = threadIdx.x; x < W; x += TK_THREADS) s_val[r * W + x] = in ? __ldg(p + (size_t)gy * W + x) : NEG_INF;
    }
    __syncthreads();
    for (int r = 0; r < hrows; ++r)
        for (int x = threadIdx.x; x < W; x += TK_THREADS) {
            float m = NEG_INF;
            const int xa = max(x - R, 0), xb = min(x + R, W - 1);
            for (int xx = xa; xx <= xb; ++xx) m = fmaxf(m, s_val[r * W + xx]);
            s_hmax[r * W + x] = m;
        }
    __syncthreads();
    // NMS survivors (v == window max, v > 0): compact their keys; also keep the NMS'd value in place for the
    // dense fallback (each thread touches only its own pixel of s_val and reads s_hmax)
    for (int r = 0; r < rows; ++r)
        for (int x = threadIdx.x; x < W; x += TK_THREADS) {
            float m = NEG_INF;
            for (int d = 0; d <= 2 * R; ++d) m = fmaxf(m, s_hmax[(r + d) * W + x]);
            const float v = s_val[(r + R) * W + x];
            const bool keep = (v == m && v > 0.f);
            s_val[(r + R) * W + x] = keep ? v : 0.f;
            if (keep) {
                const int slot = atomicAdd(&s_count, 1);
                if (slot < TK_LIST) {
                    const unsigned idx = (unsigned)((y0 + r) * W + x);
                    s_list[slot] = ((unsigned long long)__float_as_uint(v) << 32) | (0xffffffffu - idx);
                }
            }
        }
    __syncthreads();
    const int count = s_count;
    const bool dense = count > TK_LIST;
    // compact path: every thread keeps its share of the list in registers
    constexpr int PER = TK_LIST / TK_THREADS;
    unsigned long long mine[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int q = threadIdx.x + i * TK_THREADS;
        mine[i] = (!dense && q < count) ? s_list[q] : 0ull;
    }

    unsigned long long prev = ~0ull;
    unsigned long long* out = partial + ((size_t)plane * nstrips + strip) * K;
    for (int k = 0; k < K; ++k) {
        unsigned long long best = 0ull;
        if (!dense) {
#pragma unroll
            for (int i = 0; i < PER; ++i)
                if (mine[i] < prev && mine[i] > best) best = mine[i];
        } else {
            for (int r = 0; r < rows; ++r)
                for (int x = threadIdx.x; x < W; x += TK_THREADS) {
                    const float v = s_val[(r + R) * W + x];
                    if (v > 0.f) {
                        const unsigned idx = (unsigned)((y0 + r) * W + x);
                        const unsigned long long key =
                            ((unsigned long long)__float_as_uint(v) << 32) | (0xffffffffu - idx);
                        if (key < prev && key > best) best = key;
                    }
                }
        }
        best = shfl_max_u64(best);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned long long v = threadIdx.x < TK_THREADS / 32 ? s_red[threadIdx.x] : 0ull;
            v = shfl_max_u64(v);
            if (threadIdx.x == 0) { s_win = v; out[k] = v; }
        }
        __syncthreads();
        prev = s_win;
        if (prev == 0ull) {   // exhausted: remaining slots are empty
            for (int kk = k + 1 + threadIdx.x; kk < K; kk += TK_THREADS) out[kk] = 0ull;
            break;
        }
    }
}

// one warp per plane
__global__ void __launch_bounds__(32)
topk_merge_kernel(const unsigned long long* __restrict__ partial, const float* __restrict__ tag, int HW, int T,
                  int nstrips, int K, float* __restrict__ val_k, int32_t* __restrict__ ind_k, float* __restrict__ tag_k) {
    const int plane = blockIdx.x;
    const int lane = threadIdx.x;
    const unsigned long long* pl = partial + (size_t)plane * nstrips * K;
    // each lane owns strips lane, lane+32, ...; head[] = cursor into each sorted strip list
    constexpr int MAXS = 8;   // up to 256 strips
    int head[MAXS];
#pragma unroll
    for (int i = 0; i < MAXS; ++i) head[i] = 0;
    for (int k = 0; k < K; ++k) {
        unsigned long long best = 0ull;
        int bi = -1;
#pragma unroll
        for (int i = 0; i < MAXS; ++i) {
            const int s = lane + 32 * i;
            if (s < nstrips && head[i] < K) {
                const unsigned long long v = pl[(size_t)s * K + head[i]];
                if (v > best) { best = v; bi = i; }
            }
        }
        const unsigned long long win = shfl_max_u64(best);
        if (win != 0ull && best == win) {   // keys are unique (index part), exactly one lane matches
#pragma unroll
            for (int i = 0; i < MAXS; ++i)
                if (i == bi) head[i]++;
        }
        if (lane == 0) {
            float v = 0.f;
            int idx = 0;
            if (win != 0ull) {
                v = __uint_as_float((unsigned)(win >> 32));
                idx = (int)(0xffffffffu - (unsigned)(win & 0xffffffffu));
            }
            val_k[(size_t)plane * K + k] = v;
            ind_k[(size_t)plane * K + k] = idx;
            for (int t = 0; t < T; ++t)
                tag_k[((size_t)plane * K + k) * T + t] = __ldg(tag + ((size_t)plane * HW + idx) * T + t);
        }
    }
}

}  // namespace lp

using namespace lp;

extern "C" size_t lp_nms_topk_workspace_bytes(int N, int J, int H, int W, int K) {
    if (N <= 0 || J <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
    const int sr = strip_rows(W);
    const int nstrips = (H + sr - 1) / sr;
    return (size_t)N * J * nstrips * K * sizeof(unsigned long long);
}

extern "C" int lp_nms_topk_f32(const float* det, const float* tag, int N, int J, int H, int W, int T, int nms_kernel,
                               int K, float* val_k, int32_t* ind_k, float* tag_k, void* workspace,
                               size_t workspace_bytes, lp_stream_t stream) {
    LP_CHECK_ARG(det && tag && val_k && ind_k && tag_k && workspace, "lp_nms_topk_f32: null pointer");
    LP_CHECK_ARG(N > 0 && J > 0 && H > 0 && W > 0 && T > 0 && (long long)H * W < (1ll << 31),
                 "lp_nms_topk_f32: bad shape N=%d J=%d H=%d W=%d T=%d", N, J, H, W, T);
    LP_CHECK_ARG(K > 0 && K <= TK_MAXK, "lp_nms_topk_f32: K=%d out of range (1..%d)", K, TK_MAXK);
    LP_CHECK_ARG(nms_kernel >= 1 && nms_kernel <= 9 && (nms_kernel & 1), "lp_nms_topk_f32: NMS kernel %d must be odd, <= 9",
                 nms_kernel);
    LP_CHECK_ARG((long long)N * J <= 65535, "lp_nms_topk_f32: N*J=%lld exceeds the grid limit 65535", (long long)N * J);
    const int R = nms_kernel / 2;
    const int sr = strip_rows(W);
    const int nstrips = (H + sr - 1) / sr;
    LP_CHECK_ARG(nstrips <= 256, "lp_nms_topk_f32: too many strips (%d)", nstrips);
    const size_t need = lp_nms_topk_workspace_bytes(N, J, H, W, K);
    if (workspace_bytes < need) {
        set_error("lp_nms_topk_f32: workspace %zu < required %zu bytes", workspace_bytes, need);
        return LP_ERR_CAPACITY;
    }
    const size_t smem = (size_t)2 * (sr + 2 * R) * W * sizeof(float);
    LP_CHECK_ARG(smem <= 200 * 1024, "lp_nms_topk_f32: W=%d too wide for the strip buffers", W);
    cudaError_t e = cudaFuncSetAttribute((const void*)nms_topk_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(nms_topk)");
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid(nstrips, N * J);
    nms_topk_strip_kernel<<<grid, TK_THREADS, smem, s>>>(det, H, W, R, sr, K,
                                                          reinterpret_cast<unsigned long long*>(workspace));
    LP_LAUNCH_CHECK("nms_topk_strip_kernel");
    topk_merge_kernel<<<N * J, 32, 0, s>>>(reinterpret_cast<const unsigned long long*>(workspace), tag, H * W, T, nstrips, K,
                                           val_k, ind_k, tag_k);
    LP_LAUNCH_CHECK("topk_merge_kernel");
    return LP_OK;
}
