// Issue-rate microbenchmark of the CUDA-core multiply-add flavours a depthwise convolution can use on sm_100a:
//   FFMA (fp32 x fp32 + fp32), FHFMA (fma.rn.f32.f16: fp16 x fp16 + fp32), HFMA2 (packed fp16), FFMA2 (fma.rn.f32x2).
// One CTA of 512 threads per SM, every thread runs ITERS iterations of an unrolled body of NB independent accumulator
// chains; cycles are read with clock64() inside the kernel (max over CTAs), so the result is instructions / clk / SM
// independent of the clock the box happens to run at.  Output: one JSON line per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rates fma_rates.cu ; ./fma_rates
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int ITERS = 2048;

__device__ __forceinline__ float fhfma(unsigned short a, unsigned short b, float c) {
    float d;
    asm volatile("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
    return d;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ unsigned hfma2(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// conv-like operand pattern: 4 "weights" w[k] x 4 "inputs" x[j] -> 16 accumulators (each weight reused 4 times in a row)
template <int MODE>
__global__ void __launch_bounds__(512, 1) rate_kernel(const float* __restrict__ in, float* out, long long* cycles) {
    const int t = threadIdx.x;
    float xf[8], wf[4];
    for (int i = 0; i < 8; ++i) xf[i] = in[(t + i * 37) & 1023];
    for (int i = 0; i < 4; ++i) wf[i] = in[(t * 3 + i * 11) & 1023];
    float acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    unsigned xh[8], wh[4], acch[16];
    unsigned long long x2[8], w2[4], acc2[16];
    for (int i = 0; i < 8; ++i) {
        __half2 h = __floats2half2_rn(xf[i], xf[(i + 1) & 7]);
        xh[i] = *reinterpret_cast<unsigned*>(&h);
        x2[i] = ((unsigned long long)__float_as_uint(xf[i]) << 32) | __float_as_uint(xf[(i + 3) & 7]);
    }
    for (int i = 0; i < 4; ++i) {
        __half2 h = __floats2half2_rn(wf[i], wf[(i + 1) & 3]);
        wh[i] = *reinterpret_cast<unsigned*>(&h);
        w2[i] = ((unsigned long long)__float_as_uint(wf[i]) << 32) | __float_as_uint(wf[(i + 1) & 3]);
    }
    for (int i = 0; i < 16; ++i) { acch[i] = 0; acc2[i] = 0; }
    __syncthreads();
    const long long c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (MODE == 0) {            // FFMA, 3 register operands
                    acc[k * 4 + j] = ffma(xf[j + k], wf[k], acc[k * 4 + j]);
                    acc[k * 4 + j] = ffma(xf[j + k + 1], wf[(k + 1) & 3], acc[k * 4 + j]);
                } else if (MODE == 1) {     // FHFMA lo + hi (the r1 depthwise inner loop)
                    acc[k * 4 + j] = fhfma((unsigned short)(xh[j + k] & 0xffff), (unsigned short)(wh[k] & 0xffff), acc[k * 4 + j]);
                    acc[(k * 4 + j + 8) & 15] = fhfma((unsigned short)(xh[j + k] >> 16), (unsigned short)(wh[k] >> 16), acc[(k * 4 + j + 8) & 15]);
                } else if (MODE == 2) {     // HFMA2
                    acch[k * 4 + j] = hfma2(xh[j + k], wh[k], acch[k * 4 + j]);
                    acch[k * 4 + j] = hfma2(xh[j + k + 1], wh[(k + 1) & 3], acch[k * 4 + j]);
                } else if (MODE == 3) {     // FFMA2 (packed fp32 pair)
                    acc2[k * 4 + j] = ffma2(x2[j + k], w2[k], acc2[k * 4 + j]);
                    acc2[k * 4 + j] = ffma2(x2[j + k + 1], w2[(k + 1) & 3], acc2[k * 4 + j]);
                } else if (MODE == 4) {     // FFMA and HFMA2 interleaved 1:1 (do they share the pipe?)
                    acc[k * 4 + j] = ffma(xf[j + k], wf[k], acc[k * 4 + j]);
                    acch[k * 4 + j] = hfma2(xh[j + k], wh[k], acch[k * 4 + j]);
                } else if (MODE == 5) {     // FFMA2 and HFMA2 interleaved 1:1
                    acc2[k * 4 + j] = ffma2(x2[j + k], w2[k], acc2[k * 4 + j]);
                    acch[k * 4 + j] = hfma2(xh[j + k], wh[k], acch[k * 4 + j]);
                }
            }
    }
    const long long c1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 16; ++i) {
        s += acc[i] + __uint_as_float(acch[i]) + __uint_as_float((unsigned)acc2[i]) + __uint_as_float((unsigned)(acc2[i] >> 32));
    }
    out[blockIdx.x * blockDim.x + t] = s;
    if (t == 0) cycles[blockIdx.x] = c1 - c0;
}

template <int MODE>
static void run(const char* name, int macs_per_instr_a, const float* in, float* out, long long* cyc, int sms) {
    rate_kernel<MODE><<<sms, 512>>>(in, out, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate_kernel<MODE><<<sms, 512>>>(in, out, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[256], mx = 0;
    cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    const double instr = 32.0 * ITERS * 512 / 32;            // warp instructions per SM
    printf("{\"variant\": \"%s\", \"warp_instr_per_clk_per_sm\": %.3f, \"lane_ops_per_clk_per_sm\": %.1f, "
           "\"macs_per_clk_per_sm\": %.1f, \"cycles\": %lld, \"ms\": %.4f, \"err\": \"%s\"}\n",
           name, instr / mx, instr * 32 / mx, instr * 32 * macs_per_instr_a / 2.0 / mx, mx, ms,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    float *in, *out;
    long long* cyc;
    cudaMalloc(&in, 1024 * 4);
    cudaMalloc(&out, sms * 512 * 4);
    cudaMalloc(&cyc, 256 * 8);
    float h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = 0.001f * (i % 97);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    // macs_per_instr_a: sum of MACs of the two instructions of a pair (x2 later divided by 2)
    run<0>("FFMA (3 regs)", 2, in, out, cyc, sms);
    run<1>("FHFMA lo/hi (fma.rn.f32.f16)", 2, in, out, cyc, sms);
    run<2>("HFMA2 (fma.rn.f16x2)", 4, in, out, cyc, sms);
    run<3>("FFMA2 (fma.rn.f32x2)", 4, in, out, cyc, sms);
    run<4>("FFMA + HFMA2 interleaved", 3, in, out, cyc, sms);
    run<5>("FFMA2 + HFMA2 interleaved", 4, in, out, cyc, sms);
    return 0;
}
