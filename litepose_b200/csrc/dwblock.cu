// One stride-1 inverted-residual block in ONE kernel (reference lib/models/layers/layers.py:90-118:
//   inv (1x1 expand + BN + ReLU6) -> depth_conv (7x7 depthwise + BN + ReLU6) -> point_conv (1x1 + BN) -> (+ identity)).
//
// The 6x-expanded tensor never exists in HBM: per 16x16-pixel output tile the CTA
//   * TMA-loads the NARROW haloed input tile X [22x22 px][Cin <= 64] (hardware zero fill = image border) as a
//     128B-swizzled K-major operand (484 rows of 128 B, channels beyond Cin zero-filled by the tensor map),
//   * expands it 32 channels at a time on the tensor cores: E[px, 32] = X[px, Cin] * We[32, Cin]^T as 4 x
//     (M=128, N=32) tcgen05.mma into one of two 128-column TMEM slots (halo recompute factor 1.89 - the tensor pipe is
//     otherwise idle),
//   * the 8 depthwise warps of the slab's group drain the slot (tcgen05.ld -> +bias -> ReLU6 -> zero outside the image
//     -> fp16) into the [22][22][32ch] slab layout of the depthwise loop (the "epilogue" of the expansion),
//   * run the packed-fp16 7x7 depthwise on the slab (same HFMA2 loop as dwpw.cu), write ReLU6'd fp16 results into
//     the swizzled K-major A tile, and one thread issues the projection tcgen05.mma (fp32 accumulators in TMEM),
//   * all 16 depthwise warps finish the tile: tcgen05.ld, + bias, + identity row (read back from global/L2), fp16, 16-byte
//     stores.
// Warp roles: 0-15 depthwise (two groups of 8 = even / odd 32-channel slabs), 16 TMA producer, 17 MMA issuer.
// HBM traffic of a block: read N*H*W*Cin*2 (x1.9 halo, L2 hits) + identity row, write N*H*W*Co*2.
#include <stdlib.h>

#include "common.cuh"
#include "dw_inner.cuh"

namespace lp {

constexpr int BK_T = 16;                          // output tile side
constexpr int BK_I = BK_T + 6;                    // haloed side (22)
constexpr int BK_PIX = BK_I * BK_I;               // 484 haloed pixels = rows of X
constexpr int BK_CB = 32;                         // channels per slab
constexpr int BK_X_BYTES = 512 * 128;             // 4 M-tiles of 128 rows x 128 B (TMA fills the first 484 rows)
constexpr int BK_X_TX = BK_PIX * 128;             // bytes the TMA box delivers
constexpr int BK_WE_SLAB = BK_CB * 128;           // expansion weights of one slab: 32 rows x 128 B
constexpr int BK_CHUNK = 7840;                    // chunk-major slab: [4 chunks of 8 ch][484 px][16 B], chunk pitch = 7744 + 96
                                                  // (pitch = 32 mod 128: the 4 chunks of a pixel and the pixel of the mirrored
                                                  // half-warp fall into 8 different 16-byte bank groups -> conflict-free LDS;
                                                  // consecutive pixels of a chunk are contiguous -> conflict-free 16-byte STS)
constexpr int BK_SLAB = 31744;                    // 4 * 7840 = 31360 B, padded
constexpr int BK_DW_BYTES = 49 * BK_CB * 2;       // depthwise weights of one slab (tap-major [49][32]) = 3136 B
constexpr int BK_DW_SLAB = 3200;                 // their pitch in shared memory (TMA destinations are 128-byte aligned)
constexpr int BK_A_TILE = 128 * 64 * 2;
constexpr int BK_DW_WARPS = 16;
constexpr int BK_THREADS = (BK_DW_WARPS + 2) * 32;
constexpr int BK_MAX_CE = 416;
constexpr int BK_TMEM_E = 256;                    // first column of the expansion slots (2 x 128 columns); columns 0..255:
                                                  // two projection accumulator buffers (2 M-tiles x n_tile <= 64 each)

struct BkBars {
    uint64_t w_full;
    // streaming mode (the three weight sets do not fit beside X and the slabs): two-slot rings
    uint64_t we_full[2], we_empty[2];          // expansion weights of a slab: producer -> MMA issuer
    uint64_t dw_full[2][2], dw_empty[2][2];    // depthwise weights of a slab, per group: producer -> the group's warps
    uint64_t wp_full[2], wp_empty[2];          // projection weights of a K block: producer -> MMA issuer
    uint64_t x_full, x_empty;
    uint64_t e_full[2], e_empty[2];
    uint64_t a_full, a_empty;
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base, pad;
};

struct BkParams {
    int N, H, W, Cin, Ce, Co, n_tile;
    int tiles_x, tiles_y, num_tiles;
    int nslabs, nkb, k16;             // k16 = K=16 MMA steps that carry input channels (ceil(Cin/16))
    int off_we, off_slab, off_dww, off_a, off_wp, off_bias;   // shared-memory layout (bytes)
    int wp_stage;                     // streaming mode: bytes of one projection-weight ring slot
    int skew_ns;                      // group 1 starts every tile this much later (phase offset of the two groups)
    const float* b_exp;               // [Ce]
    const float* b_dw;                // [Ce]
    const float* b_pj;                // packed, n_tile
    const __half* residual;           // = x when the block has an identity connection, else null
    __half* out;                      // [N,H,W,Co]
};

// STREAM = 0: all weights of the block resident in shared memory; STREAM = 1: expansion / depthwise / projection weights
// travel through two-slot rings (blocks with Ce up to 288 at Cin = 48: stage 2 of LitePose-S)
template <int STREAM>
__global__ void __launch_bounds__(BK_THREADS, 1)
block_s1_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_we,
                const __grid_constant__ CUtensorMap map_dw, const __grid_constant__ CUtensorMap map_wp,
                const __grid_constant__ BkParams p) {
    constexpr int K = 7;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sX = smem;
    uint8_t* sWe = smem + p.off_we;
    uint8_t* sSlab = smem + p.off_slab;
    uint8_t* sDww = smem + p.off_dww;
    uint8_t* sA = smem + p.off_a;
    uint8_t* sWp = smem + p.off_wp;
    float* sBexp = reinterpret_cast<float*>(smem + p.off_bias);
    float* sBdw = sBexp + BK_MAX_CE;
    float* sBpj = sBdw + BK_MAX_CE;
    BkBars* bars = reinterpret_cast<BkBars*>(sBpj + 64);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_we);
        tma_prefetch_desc(&map_dw);
        tma_prefetch_desc(&map_wp);
        mbar_init(&bars->w_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->we_full[i], 1);
            mbar_init(&bars->we_empty[i], 1);
            mbar_init(&bars->wp_full[i], 1);
            mbar_init(&bars->wp_empty[i], 1);
            for (int g = 0; g < 2; ++g) { mbar_init(&bars->dw_full[g][i], 1); mbar_init(&bars->dw_empty[g][i], BK_DW_WARPS / 2); }
        }
        mbar_init(&bars->x_full, 1);
        mbar_init(&bars->x_empty, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->e_full[i], 1); mbar_init(&bars->e_empty[i], BK_DW_WARPS / 2); }
        mbar_init(&bars->a_full, BK_DW_WARPS);
        mbar_init(&bars->a_empty, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], BK_DW_WARPS); }
        fence_barrier_init();
    }
    if (warp == BK_DW_WARPS + 1) {
        tc_alloc(&bars->tmem_base, 512);
        tc_relinquish();
    }
    for (int i = threadIdx.x; i < p.nslabs * BK_CB; i += BK_THREADS) {
        sBexp[i] = (i < p.Ce && p.b_exp) ? p.b_exp[i] : 0.f;
        sBdw[i] = (i < p.Ce && p.b_dw) ? p.b_dw[i] : 0.f;
    }
    for (int i = threadIdx.x; i < p.n_tile; i += BK_THREADS) sBpj[i] = p.b_pj ? p.b_pj[i] : 0.f;
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == BK_DW_WARPS) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            if (!STREAM) {
                // weights are static: they may be fetched before the previous kernel of the stream has finished
                mbar_expect_tx(&bars->w_full, (uint32_t)(p.nslabs * (BK_WE_SLAB + BK_DW_BYTES) + p.nkb * p.n_tile * 128));
                for (int s = 0; s < p.nslabs; ++s) {
                    tma_load_2d(sWe + s * BK_WE_SLAB, &map_we, &bars->w_full, 0, s * BK_CB);
                    tma_load_2d(sDww + s * BK_DW_SLAB, &map_dw, &bars->w_full, s * BK_CB, 0);
                }
                for (int kb = 0; kb < p.nkb; ++kb)
                    tma_load_2d(sWp + kb * p.n_tile * 128, &map_wp, &bars->w_full, 0, kb * p.n_tile);
            }
            pdl_wait();               // the block input is complete from here on
            uint32_t we_n = 0, wp_n = 0, dw_n[2] = {0, 0};
            int it = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
                const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, n = t / (p.tiles_x * p.tiles_y);
                mbar_wait_backoff(&bars->x_empty, (it & 1) ^ 1);
                mbar_expect_tx(&bars->x_full, BK_X_TX);
                tma_load_4d(sX, &map_x, &bars->x_full, 0, tx * BK_T - 3, ty * BK_T - 3, n);
                if (STREAM) {
                    // the tile's weights in the order their consumers want them: per K block the expansion weights of its
                    // slabs (MMA issuer, slab order), their depthwise weights (the group that runs the slab), then the
                    // projection weights of the K block; two-slot rings keep the producer about one K block ahead
                    const int sw = (p.nslabs & 1) ? (it & 1) : 0;
                    for (int kb = 0; kb < p.nkb; ++kb) {
                        for (int s = 2 * kb; s < 2 * kb + 2 && s < p.nslabs; ++s) {
                            const uint32_t q = we_n & 1;
                            mbar_wait_backoff(&bars->we_empty[q], ((we_n >> 1) & 1) ^ 1);
                            mbar_expect_tx(&bars->we_full[q], BK_WE_SLAB);
                            tma_load_2d(sWe + q * BK_WE_SLAB, &map_we, &bars->we_full[q], 0, s * BK_CB);
                            ++we_n;
                        }
                        for (int s = 2 * kb; s < 2 * kb + 2 && s < p.nslabs; ++s) {
                            const int g = (s & 1) ^ sw;
                            const uint32_t q = dw_n[g] & 1;
                            mbar_wait_backoff(&bars->dw_empty[g][q], ((dw_n[g] >> 1) & 1) ^ 1);
                            mbar_expect_tx(&bars->dw_full[g][q], BK_DW_BYTES);
                            tma_load_2d(sDww + (g * 2 + q) * BK_DW_SLAB, &map_dw, &bars->dw_full[g][q], s * BK_CB, 0);
                            ++dw_n[g];
                        }
                        {
                            const uint32_t q = wp_n & 1;
                            mbar_wait_backoff(&bars->wp_empty[q], ((wp_n >> 1) & 1) ^ 1);
                            mbar_expect_tx(&bars->wp_full[q], (uint32_t)p.n_tile * 128);
                            tma_load_2d(sWp + q * p.wp_stage, &map_wp, &bars->wp_full[q], 0, kb * p.n_tile);
                            ++wp_n;
                        }
                    }
                }
            }
        }
    } else if (warp == BK_DW_WARPS + 1) {
        // ------------------------------------------------------------------ MMA issuer (expansions + projections)
        if (lane == 0) {
            const uint32_t idesc_e = umma_idesc_f16(128, BK_CB);
            const uint32_t idesc_p = umma_idesc_f16(128, p.n_tile);
            const uint32_t x_base = smem_u32(sX);
            uint32_t ecount[2] = {0, 0};          // expansion jobs issued per group
            if (!STREAM) mbar_wait_backoff(&bars->w_full, 0);
            uint32_t we_c = 0, wp_c = 0;
            // expansion of slab s of the tile whose X is resident; group = s & 1
            // Odd slab counts (Ce = 96: 3 slabs) leave one group idle in the last round of a tile; the groups therefore swap
            // roles on odd tiles (group 0 takes the odd slabs), so that over two tiles each group runs the same number of
            // slabs and - the groups being only loosely coupled through the A tile - both stay busy.
            auto expand = [&](int s, bool last_of_tile, int tile_it) {
                const int g = (s & 1) ^ ((p.nslabs & 1) ? (tile_it & 1) : 0);
                mbar_wait_backoff(&bars->e_empty[g], (ecount[g] & 1) ^ 1);
                tc_fence_after();
                uint32_t b_base = smem_u32(sWe + s * BK_WE_SLAB);
                if (STREAM) {
                    mbar_wait_backoff(&bars->we_full[we_c & 1], (we_c >> 1) & 1);
                    tc_fence_after();
                    b_base = smem_u32(sWe + (we_c & 1) * BK_WE_SLAB);
                }
                for (int mt = 0; mt < 4; ++mt)
                    for (int k = 0; k < p.k16; ++k)
                        tc_mma_f16(tmem_base + BK_TMEM_E + g * 128 + mt * BK_CB, umma_desc_sw128(x_base + mt * 16384 + k * 32),
                                   umma_desc_sw128(b_base + k * 32), idesc_e, k > 0 ? 1u : 0u);
                tc_commit(&bars->e_full[g]);
                if (STREAM) {
                    tc_commit(&bars->we_empty[we_c & 1]);
                    ++we_c;
                }
                if (last_of_tile) tc_commit(&bars->x_empty);     // X may be overwritten once these MMAs have retired
                ++ecount[g];
            };
            int it = 0;
            uint32_t kbc = 0;
            const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            if (my_tiles > 0) {
                mbar_wait_backoff(&bars->x_full, 0);
                tc_fence_after();
                expand(0, p.nslabs == 1, 0);
                if (p.nslabs > 1) expand(1, p.nslabs == 2, 0);
            }
            for (; it < my_tiles; ++it) {
                for (int kb = 0; kb < p.nkb; ++kb, ++kbc) {
                    // expansions of the NEXT K block (possibly the first of the next tile) go first: they only wait
                    // for the groups' epilogue-1 of the current slabs, which precedes the depthwise work awaited below
                    if (kb + 1 < p.nkb) {
                        const int s0 = 2 * (kb + 1);
                        expand(s0, s0 == p.nslabs - 1, it);
                        if (s0 + 1 < p.nslabs) expand(s0 + 1, s0 + 1 == p.nslabs - 1, it);
                    } else if (it + 1 < my_tiles) {
                        mbar_wait_backoff(&bars->x_full, (it + 1) & 1);
                        tc_fence_after();
                        expand(0, p.nslabs == 1, it + 1);
                        if (p.nslabs > 1) expand(1, p.nslabs == 2, it + 1);
                    }
                    // projection of K block kb
                    if (kb == 0) {
                        mbar_wait_backoff(&bars->tmem_empty[it & 1], ((it >> 1) & 1) ^ 1);
                        tc_fence_after();
                    }
                    mbar_wait_backoff(&bars->a_full, kbc & 1);
                    tc_fence_after();
                    const int k16 = 2 * min(2, p.nslabs - 2 * kb);
                    uint32_t b_base = smem_u32(sWp + kb * p.n_tile * 128);
                    if (STREAM) {
                        mbar_wait_backoff(&bars->wp_full[wp_c & 1], (wp_c >> 1) & 1);
                        tc_fence_after();
                        b_base = smem_u32(sWp + (wp_c & 1) * p.wp_stage);
                    }
                    for (int mt = 0; mt < 2; ++mt) {
                        const uint32_t a_base = smem_u32(sA + mt * BK_A_TILE);
                        for (int k = 0; k < k16; ++k)
                            tc_mma_f16(tmem_base + (it & 1) * 2 * p.n_tile + mt * p.n_tile, umma_desc_sw128(a_base + k * 32),
                                       umma_desc_sw128(b_base + k * 32), idesc_p, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(&bars->a_empty);
                    if (STREAM) {
                        tc_commit(&bars->wp_empty[wp_c & 1]);
                        ++wp_c;
                    }
                }
                tc_commit(&bars->tmem_full[it & 1]);
            }
        }
    } else {
        // ------------------------------------------------------------------ depthwise warps (+ both epilogues)
        const int cp = threadIdx.x & 15;
        const int sub = (threadIdx.x >> 4) & 1;
        const bool mir = sub != 0;
        const int grp = warp >> 3;                         // 0: even slabs, 1: odd slabs
        const int gw = warp & 7;
        const int q = warp & 3;                            // TMEM lane quarter of this warp
        const int hcol = gw >> 2;                          // which 16 channels of the slab this warp converts
        const int blk = (gw << 1) | sub;
        const int by = blk >> 2, bx = blk & 3;
        const int oy = by * 4, ox = bx * 4;
        uint8_t* slab = sSlab + grp * BK_SLAB;
        uint32_t ec = 0, kbc = 0;
        int it = 0;
        // ---- final epilogue of tile te (iteration ite), DEFERRED: it runs after the first K block of the next tile has
        // been handed to the tensor cores, so the depthwise warps never wait for the last projection MMAs of a tile (two
        // accumulator buffers).  accumulator row = pixel (mt = warp>>2 & 1, row = (warp&3)*32 + lane)
        auto epilogue = [&](int te, int ite) {
            const int tx = te % p.tiles_x, ty = (te / p.tiles_x) % p.tiles_y, n = te / (p.tiles_x * p.tiles_y);
            mbar_wait(&bars->tmem_full[ite & 1], (ite >> 1) & 1);
            tc_fence_after();
        {
            const int mt = (warp >> 2) & 1, row = q * 32 + lane;
            const int chalf = warp >> 3;               // the two groups take alternate 16-column chunks
            const int py = mt * 8 + (row >> 4), px = row & 15;
            const int gy = ty * BK_T + py, gx = tx * BK_T + px;
            const bool valid = gy < p.H && gx < p.W;
            const size_t off = (((size_t)n * p.H + gy) * p.W + gx) * p.Co;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (ite & 1) * 2 * p.n_tile + mt * p.n_tile;
            uint32_t r[16];
            for (int c0 = chalf * 16; c0 < p.n_tile; c0 += 32) {
                tc_ld16(taddr + c0, r);
                tc_wait_ld();
                if (valid && c0 < p.Co) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + sBpj[c0 + i];
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
                    uint4* op = reinterpret_cast<uint4*>(p.out + off + c0);
                    op[0] = o0;
                    if (two) op[1] = o1;
                }
            }
        }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[ite & 1]);
        };
        if (!STREAM) mbar_wait(&bars->w_full, 0);           // depthwise weights resident
        pdl_wait();                                        // the identity rows are read from global memory
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y;
            if (grp && p.skew_ns) __nanosleep(p.skew_ns);
            // which of this thread's 4 expansion rows (haloed pixels) lie inside the image
            uint32_t inside = 0;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                const int pi = mt * 128 + q * 32 + lane;
                const int yy = pi / BK_I, xx = pi - yy * BK_I;
                const int gy = ty * BK_T - 3 + yy, gx = tx * BK_T - 3 + xx;
                if (pi < BK_PIX && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) inside |= 1u << mt;
            }
            for (int kb = 0; kb < p.nkb; ++kb, ++kbc) {
                const int vg = grp ^ ((p.nslabs & 1) ? (it & 1) : 0);      // which slab parity this group runs in this tile
                const int s = 2 * kb + vg;
                const bool have = s < p.nslabs;
                __half2 acch[4][4];
                if (have) {
                    // ---- expansion epilogue: TMEM slot -> [22][22][32ch] slab (bias + ReLU6, zero outside the image)
                    if (grp) asm volatile("bar.sync 2, 256;" ::: "memory");     // previous slab fully consumed
                    else asm volatile("bar.sync 1, 256;" ::: "memory");
                    mbar_wait(&bars->e_full[grp], ec & 1);
                    tc_fence_after();
                    {
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + BK_TMEM_E + grp * 128 + hcol * 16;
                        float be[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) be[i] = sBexp[s * BK_CB + hcol * 16 + i];
                        const __half2 zero2 = __floats2half2_rn(0.f, 0.f), six2 = __floats2half2_rn(6.f, 6.f);
#pragma unroll
                        for (int mp = 0; mp < 2; ++mp) {
                            // two M-tiles per round trip: both tcgen05.ld are in flight before the single wait
                            uint32_t ra[16], rb[16];
                            tc_ld16(taddr + (2 * mp) * BK_CB, ra);
                            tc_ld16(taddr + (2 * mp + 1) * BK_CB, rb);
                            tc_wait_ld();
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int mt = 2 * mp + hh;
                                const uint32_t(&r)[16] = hh ? rb : ra;
                                const int pi = mt * 128 + q * 32 + lane;
                                uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
                                if ((inside >> mt) & 1u) {
                                    // + bias in fp32, round to fp16, ReLU6 on the packed halves (clamping commutes with rounding)
                                    __half2* h0 = reinterpret_cast<__half2*>(&o0);
                                    __half2* h1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        h0[i] = __hmin2(__hmax2(__floats2half2_rn(__uint_as_float(r[2 * i]) + be[2 * i],
                                                                                  __uint_as_float(r[2 * i + 1]) + be[2 * i + 1]), zero2), six2);
                                        h1[i] = __hmin2(__hmax2(__floats2half2_rn(__uint_as_float(r[8 + 2 * i]) + be[8 + 2 * i],
                                                                                  __uint_as_float(r[9 + 2 * i]) + be[9 + 2 * i]), zero2), six2);
                                    }
                                }
                                if (pi < BK_PIX) {
                                    // chunk-major slab: chunks 2*hcol, 2*hcol+1 of pixel pi (consecutive lanes -> consecutive 16 B)
                                    *reinterpret_cast<uint4*>(slab + (2 * hcol) * BK_CHUNK + pi * 16) = o0;
                                    *reinterpret_cast<uint4*>(slab + (2 * hcol + 1) * BK_CHUNK + pi * 16) = o1;
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->e_empty[grp]);
                    ++ec;
                    if (grp) asm volatile("bar.sync 2, 256;" ::: "memory");     // slab complete
                    else asm volatile("bar.sync 1, 256;" ::: "memory");

                    // ---- depthwise 7x7 on the slab: packed fp16 (dw_inner.cuh), chunk-major slab addressing
                    const int ch = s * BK_CB + 2 * cp;
                    const __half2 bh = __float22half2_rn(*reinterpret_cast<const float2*>(sBdw + ch));
                    const __half2* tile_in = reinterpret_cast<const __half2*>(slab + (cp >> 2) * BK_CHUNK) + (cp & 3);
                    if (STREAM) {
                        // this slab's depthwise weights arrive through the group's ring (slot = slabs run by the group so far)
                        const uint32_t q = (ec - 1) & 1;
                        mbar_wait(&bars->dw_full[grp][q], ((ec - 1) >> 1) & 1);
                        dw_slab_hfma2<K, 4, BK_I, BK_CB, 4>(tile_in, reinterpret_cast<const __half2*>(sDww + (grp * 2 + q) * BK_DW_SLAB),
                                                            cp, mir, oy, ox, bh, acch);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->dw_empty[grp][q]);
                    } else {
                        dw_slab_hfma2<K, 4, BK_I, BK_CB, 4>(tile_in, reinterpret_cast<const __half2*>(sDww + s * BK_DW_SLAB), cp, mir,
                                                            oy, ox, bh, acch);
                    }
                }
                // the single A buffer is free once the MMAs of the previous K block have retired
                mbar_wait(&bars->a_empty, (kbc & 1) ^ 1);
                if (have) {
                    dw_store_a<4>(sA, BK_A_TILE, acch, oy, ox, mir, (vg << 2) | (cp >> 2), cp);
                    fence_proxy_async();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->a_full);
                if (kb == 0 && it > 0) epilogue(t - (int)gridDim.x, it - 1);
            }
        }
        if (it > 0) epilogue((int)blockIdx.x + (it - 1) * (int)gridDim.x, it - 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == BK_DW_WARPS + 1) {
        tc_fence_after();
        tc_dealloc(tmem_base, 512);
    }
}

static size_t bk_layout(BkParams& p, int stream) {
    size_t off = BK_X_BYTES;
    const int n_we = stream ? 2 : p.nslabs, n_dw = stream ? 4 : p.nslabs;
    p.wp_stage = (p.n_tile * 128 + 1023) & ~1023;
    p.off_we = (int)off;                       off += (size_t)n_we * BK_WE_SLAB;             // 1024-aligned (4 KiB slabs)
    p.off_a = (int)off;                        off += 2 * BK_A_TILE;                          // 1024-aligned
    p.off_wp = (int)off;                       off += stream ? (size_t)2 * p.wp_stage : (((size_t)p.nkb * p.n_tile * 128 + 1023) & ~(size_t)1023);
    p.off_slab = (int)off;                     off += 2 * BK_SLAB;
    p.off_dww = (int)off;                      off += ((size_t)n_dw * BK_DW_SLAB + 127) & ~(size_t)127;
    p.off_bias = (int)off;                     off += (2 * BK_MAX_CE + 64) * 4 + sizeof(BkBars) + 64;
    return off + 1024;                         // alignment slack of the dynamic shared-memory base
}

}  // namespace lp

using namespace lp;

static int bk_shape_ok(int Cin, int Ce, int Co, BkParams* out, int* stream_out = nullptr) {
    if (Cin < 8 || Cin > 64 || Cin % 8 || Ce < 8 || Ce % 8 || Ce > BK_MAX_CE - BK_CB || Co < 8 || Co % 8 || Co > 64) return 0;
    BkParams p;
    memset(&p, 0, sizeof(p));
    p.Cin = Cin; p.Ce = Ce; p.Co = Co;
    p.n_tile = (Co + 15) / 16 * 16;
    p.nslabs = (Ce + BK_CB - 1) / BK_CB;
    p.nkb = (Ce + 63) / 64;
    p.k16 = (Cin + 15) / 16;
    int stream = 0;
    size_t need = bk_layout(p, 0);               // everything resident when it fits (no ring hand-overs)
    if (need > 232448) {                         // 227 KiB of dynamic shared memory per CTA on sm_100
        stream = 1;
        need = bk_layout(p, 1);
        if (need > 232448) return 0;
    }
    if (out) *out = p;
    if (stream_out) *stream_out = stream;
    return (int)need;
}

// 1 when lp_block_s1_f16 can run this block shape (shared-memory / TMEM budget), else 0
extern "C" int lp_block_s1_supported(int Cin, int Ce, int Co) { return bk_shape_ok(Cin, Ce, Co, nullptr) > 0; }

extern "C" size_t lp_block_s1_wexp_elems(int Cin, int Ce) { return (size_t)((Ce + BK_CB - 1) / BK_CB) * BK_CB * 64; }

// w_exp [Ce][Cin] fp16 (BN folded) -> [nslabs*32][64] K-major rows (zero padded), what map_we loads
extern "C" int lp_block_s1_pack_wexp(const uint16_t* w, int Cin, int Ce, uint16_t* out) {
    LP_CHECK_ARG(w && out && Cin > 0 && Cin <= 64 && Ce > 0, "lp_block_s1_pack_wexp: bad args");
    const int rows = (Ce + BK_CB - 1) / BK_CB * BK_CB;
    for (int r = 0; r < rows; ++r)
        for (int k = 0; k < 64; ++k) out[(size_t)r * 64 + k] = (r < Ce && k < Cin) ? w[(size_t)r * Cin + k] : (uint16_t)0;
    return LP_OK;
}

// x [N,H,W,Cin] fp16 NHWC -> out [N,H,W,Co]: relu6(x We^T + be) -> dw7x7 (+bd, relu6) -> Wp (+bp) (+ x when identity)
extern "C" int lp_block_s1_f16(const void* x, const void* w_exp_packed, const float* b_exp, const void* w_dw,
                               const float* b_dw, const void* w_proj_packed, const float* b_proj_packed, int identity,
                               void* out, int N, int H, int W, int Cin, int Ce, int Co, lp_stream_t stream) {
    LP_CHECK_ARG(x && w_exp_packed && w_dw && w_proj_packed && out, "lp_block_s1_f16: null pointer");
    BkParams p;
    int stream_mode = 0;
    const int need = bk_shape_ok(Cin, Ce, Co, &p, &stream_mode);
    LP_CHECK_ARG(N > 0 && H > 0 && W > 0 && need > 0,
                 "lp_block_s1_f16: unsupported shape N=%d H=%d W=%d Cin=%d Ce=%d Co=%d (see lp_block_s1_supported)", N, H, W,
                 Cin, Ce, Co);
    LP_CHECK_ARG(!identity || Cin == Co, "lp_block_s1_f16: identity needs Cin == Co");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(w_proj_packed) |
         reinterpret_cast<uintptr_t>(w_exp_packed) | reinterpret_cast<uintptr_t>(w_dw)) & 15) {
        set_error("lp_block_s1_f16: pointers must be 16-byte aligned");
        return LP_ERR_ALIGN;
    }
    p.N = N; p.H = H; p.W = W;
    p.tiles_x = (W + BK_T - 1) / BK_T;
    p.tiles_y = (H + BK_T - 1) / BK_T;
    p.num_tiles = p.tiles_x * p.tiles_y * N;
    p.b_exp = b_exp;
    p.b_dw = b_dw;
    p.b_pj = b_proj_packed;
    p.residual = identity ? reinterpret_cast<const __half*>(x) : nullptr;
    {
        static int skew = -1;
        if (skew < 0) {
            const char* e = getenv("LP_BLOCK_SKEW_NS");
            skew = e ? atoi(e) : 0;
        }
        p.skew_ns = skew;
    }
    p.out = reinterpret_cast<__half*>(out);
    CUtensorMap mx, mwe, mdw, mwp;
    {
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
        uint32_t box[4] = {64u, (uint32_t)BK_I, (uint32_t)BK_I, 1u};
        int rc = make_tmap(&mx, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        uint64_t d2[2] = {64u, (uint64_t)p.nslabs * BK_CB};
        uint64_t s2[1] = {128u};
        uint32_t b2[2] = {64u, (uint32_t)BK_CB};
        rc = make_tmap(&mwe, w_exp_packed, 2, d2, s2, b2, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        uint64_t d3[2] = {(uint64_t)Ce, 49u};
        uint64_t s3[1] = {(uint64_t)Ce * 2};
        uint32_t b3[2] = {(uint32_t)BK_CB, 49u};
        rc = make_tmap(&mdw, w_dw, 2, d3, s3, b3, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
        uint64_t d4[2] = {64u, (uint64_t)p.nkb * p.n_tile};
        uint64_t s4[1] = {128u};
        uint32_t b4[2] = {64u, (uint32_t)p.n_tile};
        rc = make_tmap(&mwp, w_proj_packed, 2, d4, s4, b4, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    cudaError_t e, le;
    if (stream_mode) {
        e = cudaFuncSetAttribute((const void*)block_s1_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(block_s1)");
        le = launch_pdl(block_s1_kernel<1>, dim3(grid), dim3(BK_THREADS), (size_t)need, (cudaStream_t)stream, mx, mwe, mdw, mwp, p);
    } else {
        e = cudaFuncSetAttribute((const void*)block_s1_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(block_s1)");
        le = launch_pdl(block_s1_kernel<0>, dim3(grid), dim3(BK_THREADS), (size_t)need, (cudaStream_t)stream, mx, mwe, mdw, mwp, p);
    }
    if (le != cudaSuccess) return cuda_fail(le, "launch block_s1_kernel");
    LP_LAUNCH_CHECK("block_s1_kernel");
    return LP_OK;
}
