// Fused depthwise 7x7 (stride 1) + ReLU6 + pointwise projection (+ residual) of an inverted-residual block
// (reference lib/models/layers/layers.py:100-118: depth_conv -> point_conv -> optional identity add).
//
// Unfused, the 6x-expanded tensor is written by the depthwise kernel and read back by the projection GEMM; here it
// never leaves the SM.  Per CTA (persistent over 16x16-pixel output tiles):
//   warp 16 (1 thread) TMA producer: haloed 22x22x32-channel input slabs (hardware zero fill = conv padding) into a
//                      4-deep ring, projection-weight K blocks into a 2-deep ring
//   warps 0-15         depthwise on the CUDA cores, two groups of 8 warps working on the two 32-channel halves of a
//                      64-channel K block: one 4x4 micro-block x channel pair per thread per slab; blocks (HEAD = 0):
//                      packed fp16 HFMA2 in chains of two kernel rows folded into a running fp16 total, heads
//                      (HEAD = 1): mixed-precision FHFMA with fp32 accumulation; mirrored conflict-free LDS as in
//                      dwconv.cu; results written as fp16 straight into the 128B-swizzled K-major A-operand tiles
//   warp 17 (1 thread) tcgen05.mma: D[256 px x Co] += A[256 x 64 ch] * Wp^T per 64-channel K block, fp32 in TMEM
//   warps 0-15         epilogue: tcgen05.ld, + folded-BN bias (+ residual row), fp16, 16-byte stores of whole rows
//                      (heads: deferred by one tile over two accumulator buffers, NCHW fp32/fp16 planes)
// HBM traffic per block: read N*H*W*Ce*2 (+ N*H*W*Co*2 residual), write N*H*W*Co*2  -- the depthwise output
// (N*H*W*Ce*2 written + read again) is gone; the kernel is bound by the FMA pipe (2*49 flop per expanded element).
#include "common.cuh"

namespace lp {

constexpr int FP_T = 16;                         // output tile side
constexpr int FP_CB = 32;                        // channels per slab
template <int K> struct FpCfg {
    static constexpr int I = FP_T + K - 1;                       // haloed input side (22 for k = 7, 20 for k = 5)
    static constexpr int IN_BYTES = I * I * FP_CB * 2;           // 30976 / 25600
    static constexpr int W_BYTES = K * K * FP_CB * 2;            // depthwise weights of one slab: 3136 / 1600 B
    static constexpr int IR = 4 + K - 1;                         // input rows / columns of a 4x4 micro-block
};
constexpr int FP_IN_STRIDE = 35840;                           // ring pitch: input tile + weight slab (multiple of 1024)
constexpr int FP_W_OFF = 31744;                               // weights inside a ring stage (128-byte aligned)
constexpr int FP_NIN = 4;                                      // slab ring: even slabs use stages 0/2, odd 1/3
constexpr int FP_A_TILE = 128 * 64 * 2;                       // one M-tile x one 64-channel K block, 16 KiB
constexpr int FP_NB = 2;
constexpr int FP_B_BYTES = 160 * 64 * 2;                      // Co <= 160
constexpr int FP_DW_WARPS = 16;                                // two groups of 8: even / odd 32-channel slabs
constexpr int FP_THREADS = (FP_DW_WARPS + 2) * 32;
static_assert(FP_W_OFF >= FpCfg<7>::IN_BYTES && FP_W_OFF + FpCfg<7>::W_BYTES <= FP_IN_STRIDE, "ring stage layout");
constexpr int FP_MAX_CE = 1024;
constexpr size_t FP_SMEM = (size_t)FP_NIN * FP_IN_STRIDE + 2 * FP_A_TILE + FP_NB * FP_B_BYTES + 1024 + FP_MAX_CE * 4 + 1024;

struct FpBars {
    uint64_t in_full[FP_NIN], in_empty[FP_NIN];
    uint64_t a_full, a_empty;
    uint64_t b_full[FP_NB], b_empty[FP_NB];
    uint64_t tmem_full[2], tmem_empty[2];   // heads: two accumulator buffers (deferred epilogue); blocks use [0]
    uint32_t tmem_base, pad;
};

struct FpParams {
    int N, H, W, Ce, Co, n_tile;      // n_tile = round_up(Co, 16)
    int tiles_x, tiles_y, num_tiles;
    int nslabs, nkb;
    int ns0;                          // slabs taken from the first source (HEAD mode: the rest come from the second)
    int out_fp32;                     // HEAD mode: NCHW output dtype
    const __half* w_dw;               // [k*k][Ce]
    const float* b_dw;                // [Ce] (HEAD mode: padded per slab, nslabs*32)
    const float* b_pj;                // packed, n_tile
    const __half* residual;           // [N,H,W,Co] or null
    void* out;                        // [N,H,W,Co] fp16 (HEAD: [N,Co,H,W] fp32/fp16)
};

__device__ __forceinline__ float fp_fhfma(unsigned short a, unsigned short b, float c) {
    float d;
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
    return d;
}
__device__ __forceinline__ unsigned short fp_lo(__half2 v) { return __half_as_ushort(__low2half(v)); }
__device__ __forceinline__ unsigned short fp_hi(__half2 v) { return __half_as_ushort(__high2half(v)); }

// HEAD = 0: block projection (ReLU6 after the depthwise, NHWC fp16 output with bias and optional residual).
// HEAD = 1: output head (two sources, ReLU after the depthwise, bias-free 1x1, NCHW fp32/fp16 output).
template <int K, int HEAD>
__global__ void __launch_bounds__(FP_THREADS, 1)
dw_project_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_x1,
                  const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_dw,
                  const __grid_constant__ FpParams p) {
    using Cfg = FpCfg<K>;
    constexpr int FP_I = Cfg::I;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sIn = smem;                                         // FP_NIN x [22][22][32] fp16
    uint8_t* sA = smem + FP_NIN * FP_IN_STRIDE;                  // [mtile 2] x 16 KiB, 128B-swizzled
    uint8_t* sB = sA + 2 * FP_A_TILE;                            // FP_NB x [n_tile][64] fp16, 128B-swizzled
    float* sBias = reinterpret_cast<float*>(sB + FP_NB * FP_B_BYTES);    // projection bias (<= 160)
    float* sBdw = sBias + 192;                                            // depthwise bias (<= FP_MAX_CE)
    FpBars* bars = reinterpret_cast<FpBars*>(sBdw + FP_MAX_CE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_x);
        if (HEAD) tma_prefetch_desc(&map_x1);
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_dw);
        for (int i = 0; i < FP_NIN; ++i) { mbar_init(&bars->in_full[i], 1); mbar_init(&bars->in_empty[i], 8); }
        mbar_init(&bars->a_full, FP_DW_WARPS);
        mbar_init(&bars->a_empty, 1);
        for (int i = 0; i < FP_NB; ++i) { mbar_init(&bars->b_full[i], 1); mbar_init(&bars->b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], FP_DW_WARPS); }
        fence_barrier_init();
    }
    if (warp == FP_DW_WARPS + 1) {
        tc_alloc(&bars->tmem_base, 512);
        tc_relinquish();
    }
    for (int i = threadIdx.x; i < p.n_tile; i += FP_THREADS) sBias[i] = p.b_pj ? p.b_pj[i] : 0.f;
    for (int i = threadIdx.x; i < p.nslabs * FP_CB; i += FP_THREADS)
        sBdw[i] = (p.b_dw && (HEAD || i < p.Ce)) ? p.b_dw[i] : 0.f;
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();                   // the expanded input of this block is complete from here on

    if (warp == FP_DW_WARPS) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t iu[2] = {0, 0}, bs = 0, bph = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, n = t / (p.tiles_x * p.tiles_y);
                for (int s = 0; s < p.nslabs; ++s) {
                    if ((s & 1) == 0) {   // projection weights of K block s/2
                        mbar_wait_backoff(&bars->b_empty[bs], bph ^ 1);
                        mbar_expect_tx(&bars->b_full[bs], p.n_tile * 128);
                        tma_load_2d(sB + bs * FP_B_BYTES, &map_w, &bars->b_full[bs], 0, (s >> 1) * p.n_tile);
                        if (++bs == FP_NB) { bs = 0; bph ^= 1; }
                    }
                    const int g = s & 1;                              // even / odd slab group, stages g and g+2
                    const uint32_t is = 2 * (iu[g] & 1) + g;
                    mbar_wait_backoff(&bars->in_empty[is], ((iu[g] >> 1) & 1) ^ 1);
                    mbar_expect_tx(&bars->in_full[is], Cfg::IN_BYTES + Cfg::W_BYTES);
                    const bool second = HEAD && s >= p.ns0;
                    tma_load_4d(sIn + is * FP_IN_STRIDE, second ? &map_x1 : &map_x, &bars->in_full[is],
                                (second ? s - p.ns0 : s) * FP_CB, tx * FP_T - K / 2, ty * FP_T - K / 2, n);
                    tma_load_2d(sIn + is * FP_IN_STRIDE + FP_W_OFF, &map_dw, &bars->in_full[is], s * FP_CB, 0);
                    ++iu[g];
                }
            }
        }
    } else if (warp == FP_DW_WARPS + 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, p.n_tile);
            uint32_t bs = 0, bph = 0, kbc = 0;
            int it = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
                // heads alternate between two accumulator buffers so that the depthwise warps can run the next tile while
                // this tile's MMAs retire (the epilogue of tile i follows the depthwise of tile i+1)
                const int ab = HEAD ? (it & 1) : 0;
                mbar_wait_backoff(&bars->tmem_empty[ab], (HEAD ? ((it >> 1) & 1) : (it & 1)) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < p.nkb; ++kb, ++kbc) {
                    mbar_wait_backoff(&bars->a_full, kbc & 1);
                    mbar_wait_backoff(&bars->b_full[bs], bph);
                    tc_fence_after();
                    const int k16 = 2 * min(2, p.nslabs - 2 * kb);      // 32 channels per slab present in this K block
                    const uint32_t b_base = smem_u32(sB + bs * FP_B_BYTES);
                    for (int mt = 0; mt < 2; ++mt) {
                        const uint32_t a_base = smem_u32(sA + mt * FP_A_TILE);
                        for (int k = 0; k < k16; ++k)
                            tc_mma_f16(tmem_base + ab * 2 * p.n_tile + mt * p.n_tile, umma_desc_sw128(a_base + k * 32),
                                       umma_desc_sw128(b_base + k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(&bars->a_empty);
                    tc_commit(&bars->b_empty[bs]);
                    if (++bs == FP_NB) { bs = 0; bph ^= 1; }
                }
                tc_commit(&bars->tmem_full[ab]);
            }
        }
    } else {
        // ------------------------------------------------------------------ depthwise warps + epilogue
        const int cp = threadIdx.x & 15;
        const int sub = (threadIdx.x >> 4) & 1;
        const bool mir = sub != 0;
        const int grp = warp >> 3;                         // 0: even slabs (channels 0-31 of a K block), 1: odd slabs
        const int gw = warp & 7;
        const int blk = (gw << 1) | sub;                   // 16 micro-blocks: 4 x 4 of 4x4 pixels
        const int by = blk >> 2, bx = blk & 3;
        const int oy = by * 4, ox = bx * 4;
        const int cstep = mir ? -(FP_CB / 2) : (FP_CB / 2);
        uint32_t iu = 0, kbc = 0;                          // slabs consumed by this group, K blocks finished
        // ---- epilogue of tile te (iteration ite): accumulator row = pixel (mt = warp>>2, row = (warp&3)*32 + lane)
        auto epilogue = [&](int te, int ite) {
            const int tx = te % p.tiles_x, ty = (te / p.tiles_x) % p.tiles_y, n = te / (p.tiles_x * p.tiles_y);
            const int ab = HEAD ? (ite & 1) : 0;
            mbar_wait(&bars->tmem_full[ab], HEAD ? ((ite >> 1) & 1) : (ite & 1));
            tc_fence_after();
        {
            const int mt = (warp >> 2) & 1, row = (warp & 3) * 32 + lane;
            const int chalf = warp >> 3;               // the two groups take alternate 16-column chunks
            const int py = mt * 8 + (row >> 4), px = row & 15;
            const int gy = ty * FP_T + py, gx = tx * FP_T + px;
            const bool valid = gy < p.H && gx < p.W;
            const size_t off = (((size_t)n * p.H + gy) * p.W + gx) * p.Co;
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + ab * 2 * p.n_tile + mt * p.n_tile;
            uint32_t r[16];
            for (int c0 = chalf * 16; c0 < p.n_tile; c0 += 32) {
                tc_ld16(taddr + c0, r);
                tc_wait_ld();
                if (HEAD) {
                    if (valid) {
                        const size_t plane = (size_t)p.H * p.W;
                        const size_t o = ((size_t)n * p.Co + c0) * plane + (size_t)gy * p.W + gx;
                        if (p.out_fp32) {
                            float* op = reinterpret_cast<float*>(p.out) + o;
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < p.Co) op[(size_t)i * plane] = __uint_as_float(r[i]);
                        } else {
                            __half* op = reinterpret_cast<__half*>(p.out) + o;
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < p.Co) op[(size_t)i * plane] = __float2half_rn(__uint_as_float(r[i]));
                        }
                    }
                    continue;
                }
                if (valid && c0 < p.Co) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + sBias[c0 + i];
                    const bool two = (c0 + 8) < p.Co;
                    if (p.residual) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off + c0);
                        const uint4 ra = __ldg(rp);
                        const __half2* h = reinterpret_cast<const __half2*>(&ra);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 f = __half22float2(h[i]);
                            v[2 * i] += f.x;
                            v[2 * i + 1] += f.y;
                        }
                        if (two) {
                            const uint4 rb = __ldg(rp + 1);
                            const __half2* g = reinterpret_cast<const __half2*>(&rb);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 f = __half22float2(g[i]);
                                v[8 + 2 * i] += f.x;
                                v[8 + 2 * i + 1] += f.y;
                            }
                        }
                    }
                    uint4 o0, o1;
                    __half2* ph0 = reinterpret_cast<__half2*>(&o0);
                    __half2* ph1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ph0[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                        ph1[i] = __floats2half2_rn(v[8 + 2 * i], v[8 + 2 * i + 1]);
                    }
                    uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + off + c0);
                    op[0] = o0;
                    if (two) op[1] = o1;
                }
            }
        }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[ab]);
        };
        int it = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            for (int kb = 0; kb < p.nkb; ++kb, ++kbc) {
                // (Splitting an odd last slab by rows between the two groups - each on its own M-tile, 2x4 micro-blocks -
                // was measured in round 2: 12 % slower.  The FMA pipe, not the idle group, bounds the kernel, and the
                // smaller micro-blocks need more LDS per MAC.)
                const int s = 2 * kb + grp;
                const bool have = s < p.nslabs;            // the last K block may hold a single slab
                // Accumulators.  Blocks (HEAD = 0): packed fp16 (HFMA2, two channels per instruction).  Measured on B200
                // (tools/microbench/fma_rates.cu, profiles/r2_fma_rates.jsonl): HFMA2 issues every 2 cycles per SM
                // sub-partition but carries 2 MACs per lane = 126 MAC/clk/SM, the FHFMA chain below reaches 108 and needs
                // an issue slot per MAC-lane, leaving none for the LDS/STS of this loop; with HFMA2 half of the issue
                // slots stay free.  The 49 taps are accumulated as four fp16 chains (two kernel rows each) folded into a
                // running fp16 total: the network stays at ~0.3 of the parity tolerance (the error budget is dominated by
                // the fp16 activation storage; emulation in DESIGN.md section 5).  Heads (HEAD = 1) feed the network
                // outputs directly and keep fp32 accumulation (FHFMA).
                float2 acc[4][4];
                __half2 acch[4][4], part[4][4];
                if (have) {
                    const int ch = s * FP_CB + 2 * cp;
                    const float2 b2 = *reinterpret_cast<const float2*>(sBdw + ch);
                    const __half2 bh = __float22half2_rn(b2);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            acc[i][j] = b2;
                            acch[i][j] = bh;
                        }

                    const uint32_t is = 2 * (iu & 1) + grp;          // this group's stages: grp, grp+2
                    mbar_wait(&bars->in_full[is], (iu >> 1) & 1);
                    // weights of this slab (TMA-staged next to the input tile, [49][32] fp16; channels beyond Ce are
                    // zero-filled by the tensor map); mirrored lanes read kx reversed
                    __half2 wreg[K * K];
                    {
                        const __half2* ws = reinterpret_cast<const __half2*>(sIn + is * FP_IN_STRIDE + FP_W_OFF) + cp +
                                            (mir ? (K - 1) * (FP_CB / 2) : 0);
                        const int wstep = mir ? -(FP_CB / 2) : (FP_CB / 2);
#pragma unroll
                        for (int ky = 0; ky < K; ++ky)
#pragma unroll
                            for (int kx = 0; kx < K; ++kx) wreg[ky * K + kx] = ws[ky * K * (FP_CB / 2) + kx * wstep];
                    }
                    const __half2* tile_in = reinterpret_cast<const __half2*>(sIn + is * FP_IN_STRIDE);
                    const __half2* base = tile_in + (oy * FP_I + ox + (mir ? Cfg::IR - 1 : 0)) * (FP_CB / 2) + cp;
#pragma unroll
                    for (int r = 0; r < Cfg::IR; ++r) {
                        __half2 in[Cfg::IR];
#pragma unroll
                        for (int c = 0; c < Cfg::IR; ++c) in[c] = base[r * FP_I * (FP_CB / 2) + c * cstep];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int ky = r - i;
                            if (ky >= 0 && ky < K) {
#pragma unroll
                                for (int kx = 0; kx < K; ++kx) {
                                    const __half2 wv = wreg[ky * K + kx];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        if (HEAD) {
                                            acc[i][j].x = fp_fhfma(fp_lo(in[j + kx]), fp_lo(wv), acc[i][j].x);
                                            acc[i][j].y = fp_fhfma(fp_hi(in[j + kx]), fp_hi(wv), acc[i][j].y);
                                        } else if ((ky & 1) == 0 && kx == 0) {
                                            part[i][j] = __hmul2(in[j + kx], wv);       // a new group of two kernel rows
                                        } else {
                                            part[i][j] = __hfma2(in[j + kx], wv, part[i][j]);
                                        }
                                    }
                                }
                                // fold the finished group (kernel rows {0,1},{2,3},{4,5},{6}) into the running total:
                                // chains of <= 14 roundings at partial magnitude instead of 49 at full magnitude
                                if (!HEAD && ((ky & 1) || ky == K - 1)) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) acch[i][j] = __hadd2(acch[i][j], part[i][j]);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->in_empty[is]);
                    ++iu;
                }
                // the single A buffer is free once the MMAs of the previous K block have retired (long ago in practice)
                mbar_wait(&bars->a_empty, (kbc & 1) ^ 1);
                if (have) {
                    // ReLU6, fp16, into the swizzled A tile: row = pixel, 16-byte chunk j = channels 8j..8j+7 of the K block
                    uint8_t* a_mt = sA + (by >> 1) * FP_A_TILE;
                    const int jch = (grp << 2) | (cp >> 2);
                    const __half2 zero2 = __floats2half2_rn(0.f, 0.f), six2 = __floats2half2_rn(6.f, 6.f);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int r = ((oy + i) & 7) * 16 + ox + (mir ? 3 - j : j);     // row inside the M-tile
                            __half2 v;
                            if (HEAD)       // ReLU (heads)
                                v = __floats2half2_rn(fmaxf(acc[i][j].x, 0.f), fmaxf(acc[i][j].y, 0.f));
                            else            // ReLU6 (blocks), packed
                                v = __hmin2(__hmax2(acch[i][j], zero2), six2);
                            *reinterpret_cast<__half2*>(a_mt + r * 128 + ((jch ^ (r & 7)) << 4) + ((cp & 3) << 2)) = v;
                        }
                    fence_proxy_async();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->a_full);
            }
            if (!HEAD) epilogue(t, it);
            else if (it > 0) epilogue(t - (int)gridDim.x, it - 1);     // deferred: the MMAs of the previous tile retired long ago
        }
        if (HEAD && it > 0) epilogue(blockIdx.x + (it - 1) * (int)gridDim.x, it - 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == FP_DW_WARPS + 1) {
        tc_fence_after();
        tc_dealloc(tmem_base, 512);
    }
}

}  // namespace lp

using namespace lp;

// Fused depthwise-7x7(stride 1, +bias +ReLU6) -> pointwise projection (+bias, +residual).
// x [N,H,W,Ce] fp16 NHWC; w_dw tap-major [49][Ce]; w_proj_packed / b_proj_packed from lp_pw1x1_pack(K=Ce, N=Co);
// out [N,H,W,Co].  Ce % 8 == 0, Co % 8 == 0, Co <= 160.
extern "C" int lp_dw7_project_f16(const void* x, const void* w_dw, const float* b_dw, const void* w_proj_packed,
                                  const float* b_proj_packed, const void* residual, void* out, int N, int H, int W,
                                  int Ce, int Co, lp_stream_t stream) {
    LP_CHECK_ARG(x && w_dw && w_proj_packed && out, "lp_dw7_project_f16: null pointer");
    LP_CHECK_ARG(N > 0 && H > 0 && W > 0 && Ce >= 8 && Ce % 8 == 0 && Ce <= FP_MAX_CE - FP_CB && Co >= 8 && Co % 8 == 0 &&
                     Co <= 160,
                 "lp_dw7_project_f16: bad shape N=%d H=%d W=%d Ce=%d Co=%d (Ce <= 992, Co <= 160)", N, H, W, Ce, Co);
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(w_proj_packed) |
         reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(w_dw)) & 15) {
        set_error("lp_dw7_project_f16: pointers must be 16-byte aligned");
        return LP_ERR_ALIGN;
    }
    FpParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.H = H; p.W = W; p.Ce = Ce; p.Co = Co;
    p.n_tile = (Co + 15) / 16 * 16;
    p.tiles_x = (W + FP_T - 1) / FP_T;
    p.tiles_y = (H + FP_T - 1) / FP_T;
    p.num_tiles = p.tiles_x * p.tiles_y * N;
    p.nslabs = (Ce + FP_CB - 1) / FP_CB;
    p.nkb = (Ce + 63) / 64;
    p.w_dw = reinterpret_cast<const __half*>(w_dw);
    p.b_dw = b_dw;
    p.b_pj = b_proj_packed;
    p.residual = reinterpret_cast<const __half*>(residual);
    p.out = out;
    p.ns0 = p.nslabs;
    CUtensorMap mx, mw;
    {
        uint64_t dims[4] = {(uint64_t)Ce, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t strides[3] = {(uint64_t)Ce * 2, (uint64_t)W * Ce * 2, (uint64_t)H * W * Ce * 2};
        uint32_t box[4] = {(uint32_t)FP_CB, (uint32_t)FpCfg<7>::I, (uint32_t)FpCfg<7>::I, 1u};
        int rc = make_tmap(&mx, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
        // packed projection weights: [kb][n_tile][64] (lp_pw1x1_pack with a single N chunk since Co <= 160)
        uint64_t d2[2] = {64u, (uint64_t)p.nkb * p.n_tile};
        uint64_t s2[1] = {128u};
        uint32_t b2[2] = {64u, (uint32_t)p.n_tile};
        rc = make_tmap(&mw, w_proj_packed, 2, d2, s2, b2, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    CUtensorMap md;
    {
        uint64_t d2[2] = {(uint64_t)Ce, 49u};
        uint64_t s2[1] = {(uint64_t)Ce * 2};
        uint32_t b2[2] = {(uint32_t)FP_CB, 49u};
        int rc = make_tmap(&md, w_dw, 2, d2, s2, b2, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
    }
    cudaError_t e = cudaFuncSetAttribute((const void*)dw_project_kernel<7, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)FP_SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dw7_project)");
    const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    cudaError_t le = launch_pdl(dw_project_kernel<7, 0>, dim3(grid), dim3(FP_THREADS), FP_SMEM, (cudaStream_t)stream, mx, mx,
                                mw, md, p);
    if (le != cudaSuccess) return cuda_fail(le, "launch dw_project_kernel<7,0>");
    LP_LAUNCH_CHECK("dw_project_kernel<7,0>");
    return LP_OK;
}

// ------------------------------------------------------------------ fused output head (M4)
static int head_slabs(int C) { return (C + FP_CB - 1) / FP_CB; }

extern "C" size_t lp_head_fused_dw_elems(int C1, int C2) { return (size_t)25 * (head_slabs(C1) + head_slabs(C2)) * FP_CB; }
extern "C" size_t lp_head_fused_pw_elems(int C1, int C2, int Co) {
    const int ns = head_slabs(C1) + head_slabs(C2);
    return (size_t)((ns + 1) / 2) * ((Co + 15) / 16 * 16) * 64;
}
// dw1/dw2: tap-major [25][C] BN-folded depthwise weights, bdw1/bdw2 their biases; w1 [Co][C1], w2 [Co][C2].
// Outputs (host): dw_cat [25][ns*32] fp16, bdw_cat [ns*32] fp32, pw_packed [kb][n_tile][64] fp16 in slab order.
extern "C" int lp_head_fused_pack(const uint16_t* dw1, const float* bdw1, const uint16_t* dw2, const float* bdw2,
                                  const uint16_t* w1, const uint16_t* w2, int C1, int C2, int Co, uint16_t* dw_cat,
                                  float* bdw_cat, uint16_t* pw_packed) {
    LP_CHECK_ARG(dw1 && dw2 && w1 && w2 && dw_cat && bdw_cat && pw_packed && C1 > 0 && C2 > 0 && Co > 0,
                 "lp_head_fused_pack: bad args");
    const int s1 = head_slabs(C1), s2 = head_slabs(C2), ns = s1 + s2, CP = ns * FP_CB;
    const int nt = (Co + 15) / 16 * 16, nkb = (ns + 1) / 2;
    for (int t = 0; t < 25; ++t)
        for (int c = 0; c < CP; ++c) {
            const bool second = c >= s1 * FP_CB;
            const int cc = second ? c - s1 * FP_CB : c;
            const int C = second ? C2 : C1;
            dw_cat[(size_t)t * CP + c] = cc < C ? (second ? dw2 : dw1)[(size_t)t * C + cc] : (uint16_t)0;
        }
    for (int c = 0; c < CP; ++c) {
        const bool second = c >= s1 * FP_CB;
        const int cc = second ? c - s1 * FP_CB : c;
        const float* b = second ? bdw2 : bdw1;
        bdw_cat[c] = (b && cc < (second ? C2 : C1)) ? b[cc] : 0.f;
    }
    for (int kb = 0; kb < nkb; ++kb)
        for (int r = 0; r < nt; ++r)
            for (int kk = 0; kk < 64; ++kk) {
                const int c = kb * 64 + kk;                    // padded concatenated channel
                uint16_t v = 0;
                if (r < Co && c < CP) {
                    const bool second = c >= s1 * FP_CB;
                    const int cc = second ? c - s1 * FP_CB : c;
                    if (cc < (second ? C2 : C1)) v = (second ? w2 : w1)[(size_t)r * (second ? C2 : C1) + cc];
                }
                pw_packed[((size_t)kb * nt + r) * 64 + kk] = v;
            }
    return LP_OK;
}

// out[N,Co,H,W] = W1 * relu(dw5(a1) + b1) + W2 * relu(dw5(a2) + b2): both SepConv2d heads of one level in ONE kernel
// (reference lib/models/pose_mobilenet.py:151-154, lib/models/layers/layers.py:120-133).
extern "C" int lp_head_fused_f16(const void* a1, const void* a2, const void* dw_cat, const float* bdw_cat,
                                 const void* pw_packed, void* out_nchw, int out_fp32, int N, int H, int W, int C1, int C2,
                                 int Co, lp_stream_t stream) {
    LP_CHECK_ARG(a1 && a2 && dw_cat && bdw_cat && pw_packed && out_nchw, "lp_head_fused_f16: null pointer");
    LP_CHECK_ARG(N > 0 && H > 0 && W > 0 && C1 % 8 == 0 && C2 % 8 == 0 && C1 > 0 && C2 > 0 && Co > 0 && Co <= 128,
                 "lp_head_fused_f16: bad shape N=%d H=%d W=%d C1=%d C2=%d Co=%d", N, H, W, C1, C2, Co);
    if ((reinterpret_cast<uintptr_t>(a1) | reinterpret_cast<uintptr_t>(a2) | reinterpret_cast<uintptr_t>(dw_cat) |
         reinterpret_cast<uintptr_t>(pw_packed)) & 15) {
        set_error("lp_head_fused_f16: pointers must be 16-byte aligned");
        return LP_ERR_ALIGN;
    }
    FpParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.H = H; p.W = W; p.Co = Co;
    p.ns0 = head_slabs(C1);
    p.nslabs = p.ns0 + head_slabs(C2);
    p.Ce = p.nslabs * FP_CB;
    LP_CHECK_ARG(p.Ce <= FP_MAX_CE, "lp_head_fused_f16: too many channels");
    p.nkb = (p.nslabs + 1) / 2;
    p.n_tile = (Co + 15) / 16 * 16;
    p.tiles_x = (W + FP_T - 1) / FP_T;
    p.tiles_y = (H + FP_T - 1) / FP_T;
    p.num_tiles = p.tiles_x * p.tiles_y * N;
    p.w_dw = reinterpret_cast<const __half*>(dw_cat);
    p.b_dw = bdw_cat;
    p.out = out_nchw;
    p.out_fp32 = out_fp32;
    CUtensorMap m0, m1, mw, md;
    for (int i = 0; i < 2; ++i) {
        const int C = i ? C2 : C1;
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
        uint32_t box[4] = {(uint32_t)FP_CB, (uint32_t)FpCfg<5>::I, (uint32_t)FpCfg<5>::I, 1u};
        int rc = make_tmap(i ? &m1 : &m0, i ? a2 : a1, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
    }
    {
        uint64_t d2[2] = {64u, (uint64_t)p.nkb * p.n_tile};
        uint64_t s2[1] = {128u};
        uint32_t b2[2] = {64u, (uint32_t)p.n_tile};
        int rc = make_tmap(&mw, pw_packed, 2, d2, s2, b2, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        uint64_t d3[2] = {(uint64_t)p.Ce, 25u};
        uint64_t s3[1] = {(uint64_t)p.Ce * 2};
        uint32_t b3[2] = {(uint32_t)FP_CB, 25u};
        rc = make_tmap(&md, dw_cat, 2, d3, s3, b3, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
    }
    cudaError_t e = cudaFuncSetAttribute((const void*)dw_project_kernel<5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)FP_SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(head_fused)");
    const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    cudaError_t le = launch_pdl(dw_project_kernel<5, 1>, dim3(grid), dim3(FP_THREADS), FP_SMEM, (cudaStream_t)stream, m0, m1,
                                mw, md, p);
    if (le != cudaSuccess) return cuda_fail(le, "launch dw_project_kernel<5,1>");
    LP_LAUNCH_CHECK("dw_project_kernel<5,1>");
    return LP_OK;
}
