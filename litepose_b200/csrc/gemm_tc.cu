// tcgen05 (UMMA) tensor-core contraction kernel family for LitePose on sm_100a.
//
// One persistent, warp-specialised kernel covers the three dense contractions of the
// network (reference lib/models/layers/layers.py:95-108, lib/models/pose_mobilenet.py:102-154):
//   MODE_PW     1x1 conv:   out[M,N]   = act(A[M,K] W^T + b) (+res)          (InvBottleneck inv/point_conv, stem 1x1)
//   MODE_DECONV fusion deconv level: 4 sub-pixel phases, each a K = 4*(Cr+Cw) contraction, both branches,
//               folded-BN bias + ReLU, written interleaved into the 2x up-sampled NHWC output
//   MODE_HEAD   head pair:  out_nchw_f32 = A1 W1^T + A2 W2^T
//
// Structure per CTA (320 threads, 1 CTA/SM, grid = min(#tiles, #SMs)):
//   warp 0 lane 0 : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring; MODE_PW may keep the weights of
//                                  its N chunk resident and stream only activation tiles)
//   warp 1 lane 0 : MMA issuer    (tcgen05.mma.cta_group::1.kind::f16, M=128, fp32 accumulators in TMEM,
//                                  tcgen05.commit releases smem stages / publishes accumulators)
//   warps 2..9    : epilogue      (tcgen05.ld 32x32b -> bias/act/residual -> swizzled smem staging -> TMA store, or
//                                  direct stores in the spatial modes), overlapped with the next tile's MMAs through a
//                                  double-buffered TMEM accumulator (2 x 256 columns); two groups of four warps, one
//                                  warp per TMEM lane quarter in each
// A "step" is one 128-row x 64-channel activation tile (one TMA box; spatially shifted boxes with hardware
// zero fill implement the deconv taps and all image borders) multiplied against 1..4 weight sub-tiles.
#include "common.cuh"

namespace lp {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;    // 16 KiB
constexpr int B_BYTES = 256 * BK * 2;   // 32 KiB
constexpr int STAGES = 3;        // MODE_PW (plus OUT_BYTES of output staging)
constexpr int STAGES_NOSTAGE = 4; // MODE_DECONV / MODE_HEAD (no staging buffer)
constexpr int MAX_STAGES = 9;      // resident-weights mode: the whole ring area minus the weights holds 16 KiB A stages
constexpr int RING_BYTES = STAGES * (A_BYTES + B_BYTES);   // 144 KiB
constexpr int OUT_SUB = 128 * 64 * 2;   // one 128-row x 64-column fp16 output sub-tile (128B-swizzled), 16 KiB
constexpr int OUT_BYTES = 4 * OUT_SUB;  // staging for up to 256 output columns
constexpr int MAX_STEPS = 48;
constexpr int GEMM_THREADS = 320;   // warp0 TMA, warp1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int EPI_THREADS = 256;
constexpr int MAX_BIAS = 1024;

enum { MODE_PW = 0, MODE_DECONV = 1, MODE_HEAD = 2 };

struct Step {
    int16_t kc;       // channel offset of this 64-wide K block inside its source tensor
    int8_t dx, dy;    // spatial shift of the activation box (deconv taps)
    uint8_t map;      // activation source (0/1)
    uint8_t nb;       // weight sub-tiles multiplied against this activation tile (1..4)
    uint8_t k16;      // number of K=16 MMAs that carry data in this block (1..4)
    uint8_t pad;
    uint8_t acc[4];   // accumulator index per sub-tile
    uint16_t bt0;     // first weight sub-tile index (rows bt*n_tile of the packed weight matrix)
    uint16_t pad2;
};

struct GemmParams {
    int num_tiles;     // m_tiles * n_chunks
    int n_chunks;
    int n_tile;        // MMA N (multiple of 16, <= 256)
    int num_steps;
    int total_bt;      // weight sub-tiles per chunk
    int b_resident;    // MODE_PW: all K blocks of the CTA's N chunk stay in shared memory, only A tiles stream
    int nst_a;         // MODE_PW resident mode: number of 16 KiB A stages
    int m_tiles;
    int M, N;          // PW: rows, real out channels.  spatial modes: N = Co
    int act;
    int H, W, TH, TW, tiles_x, tiles_y;  // spatial modes
    const float* bias;       // packed, n_chunks*n_tile (may be null)
    const __half* residual;  // PW only (may be null)
    void* out;
    Step steps[MAX_STEPS];
};

struct __align__(8) GemmBarriers {
    uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint64_t res_full[2];
    uint64_t bres_full;
    uint32_t tmem_base;
    uint32_t pad;
};

constexpr size_t GEMM_SMEM = 1024 /*align slack*/ + (size_t)STAGES * (A_BYTES + B_BYTES) + OUT_BYTES + MAX_BIAS * 4 + 256;

template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapOut,
               const __grid_constant__ CUtensorMap mapRes, const __grid_constant__ GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int NST = (MODE == MODE_PW) ? STAGES : STAGES_NOSTAGE;
    uint8_t* sA = smem;
    uint8_t* sB = smem + NST * A_BYTES;
    uint8_t* sOut = smem + NST * (A_BYTES + B_BYTES);      // MODE_PW: swizzled output / residual staging
    float* sBias = reinterpret_cast<float*>(smem + STAGES * (A_BYTES + B_BYTES) + OUT_BYTES);
    GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(sBias + MAX_BIAS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        if (MODE != MODE_PW) tma_prefetch_desc(&mapA1);
        tma_prefetch_desc(&mapB);
        if (MODE == MODE_PW) {
            tma_prefetch_desc(&mapOut);
            if (p.residual) tma_prefetch_desc(&mapRes);
        }
        mbar_init(&bars->res_full[0], 1);
        mbar_init(&bars->res_full[1], 1);
        mbar_init(&bars->bres_full, 1);
        for (int i = 0; i < MAX_STAGES; ++i) {
            mbar_init(&bars->full[i], 1);
            mbar_init(&bars->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->tmem_full[i], 1);
            // MODE_PW with <= 128 output columns: each accumulator buffer is drained by ONE group of four epilogue warps
            mbar_init(&bars->tmem_empty[i], (MODE == MODE_PW && p.n_tile <= 128) ? EPI_THREADS / 64 : EPI_THREADS / 32);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tc_alloc(&bars->tmem_base, 512);
        tc_relinquish();
    }
    {
        const int nb = p.n_chunks * p.n_tile;
        for (int i = threadIdx.x; i < nb && i < MAX_BIAS; i += GEMM_THREADS) sBias[i] = p.bias ? p.bias[i] : 0.f;
    }
    pdl_launch_dependents();      // the next kernel may start its own prologue
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();                   // activations written by the previous kernel are complete and visible from here on

    // Work distribution.  Streaming mode: item t -> (m-tile t / n_chunks, chunk t %% n_chunks), strided over the grid.
    // Resident-weights mode (MODE_PW): the CTA keeps ONE N chunk for its whole life (its weights are loaded once),
    // the CTAs sharing a chunk split the m-tiles round-robin.
    const bool resident = (MODE == MODE_PW) && p.b_resident;
    const int cpc = resident ? (int)gridDim.x / p.n_chunks : 1;           // CTAs per chunk
    const int my_chunk = resident ? (int)blockIdx.x % p.n_chunks : 0;
    const int t_begin = resident ? (int)blockIdx.x / p.n_chunks : (int)blockIdx.x;
    const int t_end = resident ? p.m_tiles : p.num_tiles;
    const int t_step = resident ? cpc : (int)gridDim.x;
    // resident mode: weights at the start of the ring area, A stages behind them
    const int bres_bytes = resident ? ((p.num_steps * p.n_tile * (BK * 2) + 1023) & ~1023) : 0;
    uint8_t* sAres = smem + bres_bytes;
    const int nst = resident ? p.nst_a : NST;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            if (resident) {
                mbar_expect_tx(&bars->bres_full, (uint32_t)p.num_steps * p.n_tile * (BK * 2));
                for (int s = 0; s < p.num_steps; ++s)
                    tma_load_2d(smem + s * p.n_tile * (BK * 2), &mapB, &bars->bres_full, 0,
                                (my_chunk * p.total_bt + s) * p.n_tile);
            }
            for (int t = t_begin; t < t_end; t += t_step) {
                const int chunk = resident ? my_chunk : t % p.n_chunks;
                const int mt = resident ? t : t / p.n_chunks;
                if (resident) {
                    for (int s = 0; s < p.num_steps; ++s) {
                        mbar_wait(&bars->empty[stage], phase ^ 1);
                        mbar_expect_tx(&bars->full[stage], A_BYTES);
                        tma_load_2d(sAres + stage * A_BYTES, &mapA0, &bars->full[stage], p.steps[s].kc, mt * BM);
                        if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1; }
                    }
                    continue;
                }
                int tx = 0, ty = 0, n = 0;
                if (MODE != MODE_PW) {
                    tx = mt % p.tiles_x;
                    ty = (mt / p.tiles_x) % p.tiles_y;
                    n = mt / (p.tiles_x * p.tiles_y);
                }
                for (int s = 0; s < p.num_steps; ++s) {
                    const Step& st = p.steps[s];
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    const uint32_t bytes = A_BYTES + (uint32_t)st.nb * p.n_tile * (BK * 2);
                    mbar_expect_tx(&bars->full[stage], bytes);
                    uint8_t* a_dst = sA + stage * A_BYTES;
                    if (MODE == MODE_PW) {
                        tma_load_2d(a_dst, &mapA0, &bars->full[stage], st.kc, mt * BM);
                    } else {
                        tma_load_4d(a_dst, st.map ? &mapA1 : &mapA0, &bars->full[stage], st.kc,
                                    tx * p.TW + st.dx, ty * p.TH + st.dy, n);
                    }
                    uint8_t* b_dst = sB + stage * B_BYTES;
                    for (int j = 0; j < st.nb; ++j) {
                        tma_load_2d(b_dst + j * p.n_tile * (BK * 2), &mapB, &bars->full[stage], 0,
                                    (chunk * p.total_bt + st.bt0 + j) * p.n_tile);
                    }
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(BM, p.n_tile);
            uint32_t stage = 0, phase = 0;
            int it = 0;
            if (resident) mbar_wait(&bars->bres_full, 0);
            for (int t = t_begin; t < t_end; t += t_step, ++it) {
                const int buf = it & 1;
                mbar_wait(&bars->tmem_empty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                uint32_t used = 0;
                for (int s = 0; s < p.num_steps; ++s) {
                    const Step& st = p.steps[s];
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(resident ? sAres + stage * A_BYTES : sA + stage * A_BYTES);
                    const uint32_t b_base = smem_u32(resident ? smem + s * p.n_tile * (BK * 2) : sB + stage * B_BYTES);
                    for (int j = 0; j < st.nb; ++j) {
                        const uint32_t acc = st.acc[j];
                        const uint32_t d = tmem_base + buf * 256 + acc * p.n_tile;
                        const uint32_t bj = b_base + j * p.n_tile * (BK * 2);
                        for (int k = 0; k < st.k16; ++k) {
                            tc_mma_f16(d, umma_desc_sw128(a_base + k * 32), umma_desc_sw128(bj + k * 32), idesc,
                                       ((used >> acc) & 1u) | (k > 0 ? 1u : 0u));
                        }
                        used |= 1u << acc;
                    }
                    tc_commit(&bars->empty[stage]);
                    if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1; }
                }
                tc_commit(&bars->tmem_full[buf]);
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4)
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;          // the two warps of a quarter take alternate 16-column groups
        const int row = q * 32 + lane;
        // MODE_PW, tiles of <= 128 columns: the two groups of four warps work on ALTERNATE tiles (group g drains TMEM
        // buffer g through staging buffer g with its own named barrier and its own TMA store stream), so two tiles'
        // epilogues are in flight and the latency chain tcgen05.ld -> convert -> st.shared -> TMA store of one tile
        // hides behind the other's.  Wider tiles keep both groups on the same tile (alternate 16-column groups).
        const bool split = (MODE == MODE_PW) && p.n_tile <= 128;
        const int it0 = split ? half : 0, it_inc = split ? 2 : 1;
        int it = it0;
        for (int t = t_begin + it0 * t_step; t < t_end; t += it_inc * t_step, it += it_inc) {
            const int buf = it & 1;
            const int chunk = resident ? my_chunk : t % p.n_chunks;
            const int mt = resident ? t : t / p.n_chunks;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256;
            uint32_t r[16];
            if (MODE != MODE_PW) {
                mbar_wait(&bars->tmem_full[buf], (it >> 1) & 1);
                tc_fence_after();
            }
            if (MODE == MODE_PW) {
                // Output tile goes through a 128B-swizzled shared-memory staging buffer (conflict-free 16-byte
                // st.shared) and leaves with TMA tensor stores (full 128-byte lines, rows/columns beyond M/N
                // clipped by the tensor map); the residual tile arrives the same way.
                const bool issuer = split ? (threadIdx.x == 64 + 128 * half) : (threadIdx.x == 64);
                const int ncols = min(p.n_tile, p.N - chunk * p.n_tile);       // valid columns of this chunk
                const int nsub = (ncols + 63) >> 6;
                // staging: 4 sub-tile slots.  split: group g owns slots 2g, 2g+1.  Otherwise the whole tile uses slots 0..
                uint8_t* sStage = sOut + (split ? half * 2 * OUT_SUB : 0);
                uint64_t* res_bar = &bars->res_full[split ? half : 0];
                const uint32_t res_par = split ? ((it >> 1) & 1) : (it & 1);
                if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // this issuer's previous store
                if (split) {
                    if (half) asm volatile("bar.sync 2, 128;" ::: "memory");
                    else asm volatile("bar.sync 1, 128;" ::: "memory");
                } else {
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
                if (p.residual && issuer) {   // residual tile is fetched while this tile's MMAs are still running
                    mbar_expect_tx(res_bar, nsub * OUT_SUB);
                    for (int g = 0; g < nsub; ++g)
                        tma_load_2d(sStage + g * OUT_SUB, &mapRes, res_bar, chunk * p.n_tile + g * 64, mt * BM);
                }
                mbar_wait(&bars->tmem_full[buf], (it >> 1) & 1);
                tc_fence_after();
                if (p.residual) mbar_wait(res_bar, res_par);
                uint8_t* srow = sStage + row * 128;
                for (int c0 = split ? 0 : half * 16; c0 < p.n_tile; c0 += split ? 16 : 32) {
                    tc_ld16(taddr + c0, r);
                    tc_wait_ld();
                    if (c0 < ncols) {
                        const int n0 = chunk * p.n_tile + c0;
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 bv = *reinterpret_cast<const float4*>(sBias + n0 + 4 * i);
                            v[4 * i + 0] = __uint_as_float(r[4 * i + 0]) + bv.x;
                            v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bv.y;
                            v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bv.z;
                            v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bv.w;
                        }
                        if (p.act != LP_ACT_NONE) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = act_apply(v[i], p.act);
                        }
                        uint8_t* sub = srow + (c0 >> 6) * OUT_SUB;
                        const int j0 = (c0 & 63) >> 3;                        // 16-byte chunk index inside the row
                        uint4* d0 = reinterpret_cast<uint4*>(sub + (((j0) ^ (row & 7)) << 4));
                        uint4* d1 = reinterpret_cast<uint4*>(sub + (((j0 + 1) ^ (row & 7)) << 4));
                        if (p.residual) {
                            const uint4 ra = *d0, rb = *d1;
                            const __half2* h = reinterpret_cast<const __half2*>(&ra);
                            const __half2* g = reinterpret_cast<const __half2*>(&rb);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 f = __half22float2(h[i]), e = __half22float2(g[i]);
                                v[2 * i] += f.x;
                                v[2 * i + 1] += f.y;
                                v[8 + 2 * i] += e.x;
                                v[8 + 2 * i + 1] += e.y;
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
                        *d0 = o0;
                        *d1 = o1;
                    }
                }
                // accumulator drained: hand the TMEM buffer back before the store is issued
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);
                fence_proxy_async();
                if (split) {
                    if (half) asm volatile("bar.sync 2, 128;" ::: "memory");
                    else asm volatile("bar.sync 1, 128;" ::: "memory");
                } else {
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
                if (issuer) {
                    for (int g = 0; g < nsub; ++g) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                         reinterpret_cast<uint64_t>(&mapOut)),
                                     "r"(smem_u32(sStage + g * OUT_SUB)), "r"(chunk * p.n_tile + g * 64), "r"(mt * BM)
                                     : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                continue;
            } else {
                const int tx = mt % p.tiles_x;
                const int ty = (mt / p.tiles_x) % p.tiles_y;
                const int n = mt / (p.tiles_x * p.tiles_y);
                const int ly = row / p.TW, lx = row % p.TW;
                const int y = ty * p.TH + ly, x = tx * p.TW + lx;
                const bool valid = (y < p.H) && (x < p.W);
                if (MODE == MODE_DECONV) {
                    __half* out = reinterpret_cast<__half*>(p.out);
                    const int Co = p.N;
                    for (int ph = 0; ph < 4; ++ph) {
                        const int a = ph >> 1, b = ph & 1;
                        __half* op = out + ((((long long)n * 2 * p.H + 2 * y + a) * (2 * p.W)) + 2 * x + b) * Co;
                        for (int c0 = half * 16; c0 < p.n_tile; c0 += 32) {
                            tc_ld16(taddr + ph * p.n_tile + c0, r);
                            tc_wait_ld();
                            if (valid && c0 < Co) {
                                uint4 o0, o1;
                                __half2* ph0 = reinterpret_cast<__half2*>(&o0);
                                __half2* ph1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    ph0[i] = __floats2half2_rn(
                                        fmaxf(__uint_as_float(r[2 * i]) + sBias[c0 + 2 * i], 0.f),
                                        fmaxf(__uint_as_float(r[2 * i + 1]) + sBias[c0 + 2 * i + 1], 0.f));
                                    ph1[i] = __floats2half2_rn(
                                        fmaxf(__uint_as_float(r[8 + 2 * i]) + sBias[c0 + 8 + 2 * i], 0.f),
                                        fmaxf(__uint_as_float(r[8 + 2 * i + 1]) + sBias[c0 + 8 + 2 * i + 1], 0.f));
                                }
                                uint4* o = reinterpret_cast<uint4*>(op + c0);
                                o[0] = o0;
                                if (c0 + 8 < Co) o[1] = o1;
                            }
                        }
                    }
                } else {  // MODE_HEAD: NCHW, fp32 (act == 1) or fp16 (act == 0)
                    const int Co = p.N;
                    const long long plane = (long long)p.H * p.W;
                    const long long off = (long long)n * Co * plane + (long long)y * p.W + x;
                    float* op32 = reinterpret_cast<float*>(p.out) + off;
                    __half* op16 = reinterpret_cast<__half*>(p.out) + off;
                    for (int c0 = half * 16; c0 < p.n_tile; c0 += 32) {
                        tc_ld16(taddr + c0, r);
                        tc_wait_ld();
                        if (valid) {
                            if (p.act) {
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if (c0 + i < Co) op32[(long long)(c0 + i) * plane] = __uint_as_float(r[i]);
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if (c0 + i < Co) op16[(long long)(c0 + i) * plane] = __float2half_rn(__uint_as_float(r[i]));
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);
        }
        if (MODE == MODE_PW && (threadIdx.x == 64 || threadIdx.x == 192)) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tc_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------ host side
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// N <= 160 (every projection, incl. the fused depthwise+projection kernel's single-chunk layout): one chunk of
// round_up(N,16) MMA columns.  Wider layers (the 6x expansions) are cut into 128-column chunks: a multiple of the
// 64-column TMA store sub-tile (stores of neighbouring chunks never overlap), and tiles of <= 128 columns let the two
// epilogue warp groups drain alternate tiles.  (192/256-column chunks measured 0.9 % slower on the whole step.)
static void pw_tiling(int N, int* n_chunks, int* n_tile) {
    const int np = round_up(N, 16);
    if (np <= 160) {
        *n_chunks = 1;
        *n_tile = np;
        return;
    }
    *n_chunks = (np + 127) / 128;
    *n_tile = 128;
}

static int set_smem_attr_once(const void* fn) {
    // cudaFuncSetAttribute is per-device state; cheap enough to set on every call
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_tc)");
    return LP_OK;
}

static int make_b_map(CUtensorMap* m, const void* w, int total_rows, int n_tile) {
    uint64_t dims[2] = {(uint64_t)BK, (uint64_t)total_rows};
    uint64_t strides[1] = {(uint64_t)BK * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)n_tile};
    return make_tmap(m, w, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

static int make_act_map4(CUtensorMap* m, const void* x, int N, int H, int W, int C, int TW, int TH) {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)TW, (uint32_t)TH, 1u};
    return make_tmap(m, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

static void pick_spatial_tile(int H, int W, int* TH, int* TW) {
    int best = -1, bw = 32;
    const int cands[3] = {32, 16, 8};
    for (int i = 0; i < 3; ++i) {
        const int tw = cands[i], th = BM / tw;
        const int cover = round_up(W, tw) * round_up(H, th);
        if (best < 0 || cover < best) { best = cover; bw = tw; }
    }
    *TW = bw;
    *TH = BM / bw;
}

template <int MODE>
static int launch_gemm(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& mo,
                       const CUtensorMap& mr, const GemmParams& p, cudaStream_t stream) {
    int rc = set_smem_attr_once((const void*)gemm_tc_kernel<MODE>);
    if (rc) return rc;
    int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    if (MODE == MODE_PW && p.b_resident) {
        // a multiple of n_chunks CTAs, at most one per SM and no more CTAs per chunk than m-tiles
        int cpc = num_sms() / p.n_chunks;
        if (cpc > p.m_tiles) cpc = p.m_tiles;
        grid = cpc * p.n_chunks;
    }
    if (grid < 1) return LP_OK;
    cudaError_t le = launch_pdl(gemm_tc_kernel<MODE>, dim3(grid), dim3(GEMM_THREADS), GEMM_SMEM, stream, a0, a1, b, mo, mr, p);
    if (le != cudaSuccess) return cuda_fail(le, "launch gemm_tc_kernel");
    LP_LAUNCH_CHECK("gemm_tc_kernel");
    return LP_OK;
}

}  // namespace lp

using namespace lp;

// ------------------------------------------------------------------ pointwise 1x1
extern "C" size_t lp_pw1x1_packed_elems(int K, int N) {
    int nc, nt;
    pw_tiling(N, &nc, &nt);
    return (size_t)nc * ((K + BK - 1) / BK) * nt * BK;
}
extern "C" size_t lp_pw1x1_packed_bias_elems(int N) {
    int nc, nt;
    pw_tiling(N, &nc, &nt);
    return (size_t)nc * nt;
}
extern "C" int lp_pw1x1_pack(const uint16_t* w, const float* bias, int K, int N, uint16_t* wp, float* bp) {
    LP_CHECK_ARG(w && wp && bp && K > 0 && N > 0, "lp_pw1x1_pack: null pointer or bad shape K=%d N=%d", K, N);
    int nc, nt;
    pw_tiling(N, &nc, &nt);
    const int kb = (K + BK - 1) / BK;
    for (int c = 0; c < nc; ++c)
        for (int s = 0; s < kb; ++s)
            for (int r = 0; r < nt; ++r) {
                const int n = c * nt + r;
                uint16_t* dst = wp + (((size_t)c * kb + s) * nt + r) * BK;
                for (int kk = 0; kk < BK; ++kk) {
                    const int k = s * BK + kk;
                    dst[kk] = (n < N && k < K) ? w[(size_t)n * K + k] : (uint16_t)0;
                }
            }
    for (int i = 0; i < nc * nt; ++i) bp[i] = (bias && i < N) ? bias[i] : 0.f;
    return LP_OK;
}

extern "C" int lp_pw1x1_f16(const void* a, const void* w_packed, const float* bias_packed, const void* residual,
                            void* out, int M, int K, int N, int act, lp_stream_t stream) {
    LP_CHECK_ARG(a && w_packed && out, "lp_pw1x1_f16: null pointer");
    LP_CHECK_ARG(M > 0 && K >= 8 && N >= 8 && K % 8 == 0 && N % 8 == 0,
                 "lp_pw1x1_f16: need M>0, K%%8==0, N%%8==0 (M=%d K=%d N=%d)", M, K, N);
    LP_CHECK_ARG(act >= LP_ACT_NONE && act <= LP_ACT_RELU6, "lp_pw1x1_f16: bad act %d", act);
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(w_packed) |
         reinterpret_cast<uintptr_t>(residual)) & 15) {
        set_error("lp_pw1x1_f16: pointers must be 16-byte aligned");
        return LP_ERR_ALIGN;
    }
    GemmParams p;
    memset(&p, 0, sizeof(p));
    pw_tiling(N, &p.n_chunks, &p.n_tile);
    const int kb = (K + BK - 1) / BK;
    LP_CHECK_ARG(kb <= MAX_STEPS && p.n_chunks * p.n_tile <= MAX_BIAS, "lp_pw1x1_f16: K=%d or N=%d too large", K, N);
    const int m_tiles = (M + BM - 1) / BM;
    p.num_tiles = m_tiles * p.n_chunks;
    p.m_tiles = m_tiles;
    p.num_steps = kb;
    {
        // resident-weights mode when one N chunk's weights leave room for >= 5 A stages in the ring area
        const int bres = ((kb * p.n_tile * (BK * 2)) + 1023) & ~1023;
        const int nst = (RING_BYTES - bres) / A_BYTES;
        if (nst >= 5 && p.n_chunks <= num_sms()) {
            p.b_resident = 1;
            p.nst_a = nst < MAX_STAGES ? nst : MAX_STAGES;
        }
    }
    p.total_bt = kb;
    p.M = M;
    p.N = N;
    p.act = act;
    p.bias = bias_packed;
    p.residual = reinterpret_cast<const __half*>(residual);
    p.out = out;
    for (int s = 0; s < kb; ++s) {
        Step& st = p.steps[s];
        st.kc = (int16_t)(s * BK);
        st.nb = 1;
        const int kv = (K - s * BK) < BK ? (K - s * BK) : BK;
        st.k16 = (uint8_t)((kv + 15) / 16);
        st.acc[0] = 0;
        st.bt0 = (uint16_t)s;
    }
    CUtensorMap ma, mb;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
        uint64_t strides[1] = {(uint64_t)K * 2};
        uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
        int rc = make_tmap(&ma, a, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        rc = make_b_map(&mb, w_packed, p.n_chunks * kb * p.n_tile, p.n_tile);
        if (rc) return rc;
    }
    CUtensorMap mo, mr;
    {
        uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
        uint64_t strides[1] = {(uint64_t)N * 2};
        uint32_t box[2] = {64u, (uint32_t)BM};
        int rc = make_tmap(&mo, out, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        mr = mo;
        if (residual) {
            rc = make_tmap(&mr, residual, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
            if (rc) return rc;
        }
    }
    return launch_gemm<MODE_PW>(ma, ma, mb, mo, mr, p, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ fusion deconv
// Two kernels share one packed-weight buffer: [tiled-kernel weights | row-kernel weights (when the channels qualify)].
// Wide maps run the row-streaming kernel (deconv_rows.cu), the others the tiled MODE_DECONV kernel above.
namespace lp {
bool deconv_rows_channels_ok(int Cr, int Cw, int Co);
bool deconv_rows_shape_ok(int W);
size_t deconv_rows_packed_elems(int Cr, int Cw, int Co);
void deconv_rows_pack(const uint16_t* wr, const uint16_t* ww, int Cr, int Cw, int Co, uint16_t* wp);
int launch_deconv_rows(const void* refined, const void* raw, const void* w_rows, const float* bias_packed, void* out, int N,
                       int H, int W, int Cr, int Cw, int Co, cudaStream_t stream);
}  // namespace lp

static int deconv_ntile(int Co) { return round_up(Co, 16); }

// enumerate steps; optionally emit packed weights
static int deconv_program(int Cr, int Cw, int Co, Step* steps, int* total_bt, const uint16_t* wr, const uint16_t* ww,
                          uint16_t* wp) {
    const int nt = deconv_ntile(Co);
    int s = 0, bt = 0;
    for (int br = 0; br < 2; ++br) {
        const int C = br ? Cw : Cr;
        const uint16_t* w = br ? ww : wr;
        for (int kb = 0; kb * BK < C; ++kb) {
            const int kv = (C - kb * BK) < BK ? (C - kb * BK) : BK;
            for (int di = -1; di <= 1; ++di)
                for (int dj = -1; dj <= 1; ++dj) {
                    if (s >= MAX_STEPS) return -1;
                    Step st;
                    memset(&st, 0, sizeof(st));
                    st.kc = (int16_t)(kb * BK);
                    st.dx = (int8_t)dj;
                    st.dy = (int8_t)di;
                    st.map = (uint8_t)br;
                    st.k16 = (uint8_t)((kv + 15) / 16);
                    st.bt0 = (uint16_t)bt;
                    int nb = 0;
                    for (int a = 0; a < 2; ++a) {
                        int ki;
                        if (a == 0) { if (di == 0) ki = 1; else if (di == -1) ki = 3; else continue; }
                        else        { if (di == 0) ki = 2; else if (di == 1) ki = 0; else continue; }
                        for (int b = 0; b < 2; ++b) {
                            int kj;
                            if (b == 0) { if (dj == 0) kj = 1; else if (dj == -1) kj = 3; else continue; }
                            else        { if (dj == 0) kj = 2; else if (dj == 1) kj = 0; else continue; }
                            st.acc[nb] = (uint8_t)(a * 2 + b);
                            if (wp) {
                                uint16_t* dst = wp + (size_t)(bt + nb) * nt * BK;
                                for (int co = 0; co < nt; ++co)
                                    for (int kk = 0; kk < BK; ++kk) {
                                        const int ci = kb * BK + kk;
                                        dst[(size_t)co * BK + kk] =
                                            (co < Co && ci < C) ? w[(((size_t)ci * Co + co) * 4 + ki) * 4 + kj] : (uint16_t)0;
                                    }
                            }
                            ++nb;
                        }
                    }
                    st.nb = (uint8_t)nb;
                    bt += nb;
                    if (steps) steps[s] = st;
                    ++s;
                }
        }
    }
    *total_bt = bt;
    return s;
}

extern "C" size_t lp_deconv_packed_elems(int Cr, int Cw, int Co) {
    int bt = 0;
    if (deconv_program(Cr, Cw, Co, nullptr, &bt, nullptr, nullptr, nullptr) < 0) return 0;
    return (size_t)bt * deconv_ntile(Co) * BK + deconv_rows_packed_elems(Cr, Cw, Co);
}
extern "C" size_t lp_deconv_packed_bias_elems(int Co) { return (size_t)deconv_ntile(Co); }
extern "C" int lp_deconv_pack(const uint16_t* wr, const uint16_t* ww, const float* bias, int Cr, int Cw, int Co,
                              uint16_t* wp, float* bp) {
    LP_CHECK_ARG(wr && ww && wp && bp, "lp_deconv_pack: null pointer");
    LP_CHECK_ARG(Cr % 8 == 0 && Cw % 8 == 0 && Co % 8 == 0 && Co <= 64, "lp_deconv_pack: bad channels %d %d %d", Cr, Cw, Co);
    int bt = 0;
    LP_CHECK_ARG(deconv_program(Cr, Cw, Co, nullptr, &bt, wr, ww, wp) > 0, "lp_deconv_pack: too many K blocks");
    if (deconv_rows_channels_ok(Cr, Cw, Co)) deconv_rows_pack(wr, ww, Cr, Cw, Co, wp + (size_t)bt * deconv_ntile(Co) * BK);
    for (int i = 0; i < deconv_ntile(Co); ++i) bp[i] = (bias && i < Co) ? bias[i] : 0.f;
    return LP_OK;
}

extern "C" int lp_fusion_deconv_f16(const void* refined, const void* raw, const void* w_packed, const float* bias_packed,
                                    void* out, int N, int H, int W, int Cr, int Cw, int Co, lp_stream_t stream) {
    LP_CHECK_ARG(refined && raw && w_packed && out, "lp_fusion_deconv_f16: null pointer");
    LP_CHECK_ARG(N > 0 && H > 0 && W > 0 && Cr % 8 == 0 && Cw % 8 == 0 && Co % 8 == 0 && Co >= 8 && Co <= 64,
                 "lp_fusion_deconv_f16: bad shape N=%d H=%d W=%d Cr=%d Cw=%d Co=%d (Co<=64)", N, H, W, Cr, Cw, Co);
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.n_chunks = 1;
    p.n_tile = deconv_ntile(Co);
    p.num_steps = deconv_program(Cr, Cw, Co, p.steps, &p.total_bt, nullptr, nullptr, nullptr);
    LP_CHECK_ARG(p.num_steps > 0, "lp_fusion_deconv_f16: too many K blocks");
    if ((reinterpret_cast<uintptr_t>(refined) | reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(out) |
         reinterpret_cast<uintptr_t>(w_packed)) & 15) {
        set_error("lp_fusion_deconv_f16: pointers must be 16-byte aligned");
        return LP_ERR_ALIGN;
    }
    if (deconv_rows_channels_ok(Cr, Cw, Co) && deconv_rows_shape_ok(W)) {
        const uint16_t* w_rows = reinterpret_cast<const uint16_t*>(w_packed) + (size_t)p.total_bt * p.n_tile * BK;
        return launch_deconv_rows(refined, raw, w_rows, bias_packed, out, N, H, W, Cr, Cw, Co, (cudaStream_t)stream);
    }
    pick_spatial_tile(H, W, &p.TH, &p.TW);
    p.tiles_x = (W + p.TW - 1) / p.TW;
    p.tiles_y = (H + p.TH - 1) / p.TH;
    p.num_tiles = p.tiles_x * p.tiles_y * N;
    p.H = H;
    p.W = W;
    p.N = Co;
    p.act = LP_ACT_RELU;
    p.bias = bias_packed;
    p.out = out;
    CUtensorMap m0, m1, mb;
    int rc = make_act_map4(&m0, refined, N, H, W, Cr, p.TW, p.TH);
    if (rc) return rc;
    rc = make_act_map4(&m1, raw, N, H, W, Cw, p.TW, p.TH);
    if (rc) return rc;
    rc = make_b_map(&mb, w_packed, p.total_bt * p.n_tile, p.n_tile);
    if (rc) return rc;
    return launch_gemm<MODE_DECONV>(m0, m1, mb, mb, mb, p, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ heads
extern "C" size_t lp_head_packed_elems(int C1, int C2, int Co) {
    return (size_t)((C1 + BK - 1) / BK + (C2 + BK - 1) / BK) * round_up(Co, 16) * BK;
}
extern "C" int lp_head_pack(const uint16_t* w1, const uint16_t* w2, int C1, int C2, int Co, uint16_t* wp) {
    LP_CHECK_ARG(w1 && w2 && wp && C1 > 0 && C2 > 0 && Co > 0, "lp_head_pack: bad args");
    const int nt = round_up(Co, 16);
    int bt = 0;
    for (int br = 0; br < 2; ++br) {
        const int C = br ? C2 : C1;
        const uint16_t* w = br ? w2 : w1;
        for (int kb = 0; kb * BK < C; ++kb, ++bt)
            for (int co = 0; co < nt; ++co)
                for (int kk = 0; kk < BK; ++kk) {
                    const int ci = kb * BK + kk;
                    wp[((size_t)bt * nt + co) * BK + kk] = (co < Co && ci < C) ? w[(size_t)co * C + ci] : (uint16_t)0;
                }
    }
    return LP_OK;
}

extern "C" int lp_head_pw_dual_f16(const void* a1, const void* a2, const void* w_packed, void* out_nchw, int out_fp32,
                                   int N, int H, int W, int C1, int C2, int Co, lp_stream_t stream) {
    LP_CHECK_ARG(a1 && a2 && w_packed && out_nchw, "lp_head_pw_dual_f16: null pointer");
    LP_CHECK_ARG(N > 0 && H > 0 && W > 0 && C1 % 8 == 0 && C2 % 8 == 0 && Co > 0 && Co <= 256,
                 "lp_head_pw_dual_f16: bad shape N=%d H=%d W=%d C1=%d C2=%d Co=%d", N, H, W, C1, C2, Co);
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.n_chunks = 1;
    p.n_tile = round_up(Co, 16);
    int s = 0;
    for (int br = 0; br < 2; ++br) {
        const int C = br ? C2 : C1;
        for (int kb = 0; kb * BK < C; ++kb, ++s) {
            LP_CHECK_ARG(s < MAX_STEPS, "lp_head_pw_dual_f16: too many K blocks");
            Step& st = p.steps[s];
            st.kc = (int16_t)(kb * BK);
            st.map = (uint8_t)br;
            st.nb = 1;
            const int kv = (C - kb * BK) < BK ? (C - kb * BK) : BK;
            st.k16 = (uint8_t)((kv + 15) / 16);
            st.bt0 = (uint16_t)s;
        }
    }
    p.num_steps = s;
    p.total_bt = s;
    p.TW = 32;
    p.TH = 4;
    if (W < 32) pick_spatial_tile(H, W, &p.TH, &p.TW);
    p.tiles_x = (W + p.TW - 1) / p.TW;
    p.tiles_y = (H + p.TH - 1) / p.TH;
    p.num_tiles = p.tiles_x * p.tiles_y * N;
    p.H = H;
    p.W = W;
    p.N = Co;
    p.act = out_fp32 ? 1 : 0;   // MODE_HEAD reuses `act` as the output-dtype flag
    p.out = out_nchw;
    CUtensorMap m0, m1, mb;
    int rc = make_act_map4(&m0, a1, N, H, W, C1, p.TW, p.TH);
    if (rc) return rc;
    rc = make_act_map4(&m1, a2, N, H, W, C2, p.TW, p.TH);
    if (rc) return rc;
    rc = make_b_map(&mb, w_packed, p.total_bt * p.n_tile, p.n_tile);
    if (rc) return rc;
    return launch_gemm<MODE_HEAD>(m0, m1, mb, mb, mb, p, (cudaStream_t)stream);
}
