// FFMA vs FFMA2 (fma.rn.f32x2, sm_100) issue / execution rate on B200, in the operand pattern the conv kernels use:
//   acc[c] += x * w[c]   with x a per-thread scalar and w[c] warp-uniform (kernel-parameter constant bank -> uniform registers).
// Variants: scalar FFMA, FFMA2 (scalar-broadcast x, UR weight pair), FFMA2 + interleaved ALU work (does the freed issue slot pay?),
// FFMA2 with per-thread (register) weight pairs.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

struct P { float w[512]; };

__device__ __forceinline__ void fma2(float& a0, float& a1, float x0, float x1, float w0, float w1) {
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%0,%1}; fma.rn.f32x2 rc, ra, rb, rc; mov.b64 {%0,%1}, rc; }"
        : "+f"(a0), "+f"(a1) : "f"(x0), "f"(x1), "f"(w0), "f"(w1));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(const float4* __restrict__ in, float* __restrict__ out, int iters, const __grid_constant__ P prm) {
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
    float4 v = in[threadIdx.x];
    float xs[4] = {v.x, v.y, v.z, v.w};
    int junk = threadIdx.x;
    float wr[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) wr[c] = in[threadIdx.x + 256].x * (c + 1);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {                        // 8 x 16 = 128 MACs per iteration and thread
            const float x = xs[r & 3];
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
                const float w0 = prm.w[(r * 16 + c) & 511], w1 = prm.w[(r * 16 + c + 1) & 511];
                if (MODE == 0) { acc[c] = fmaf(x, w0, acc[c]); acc[c + 1] = fmaf(x, w1, acc[c + 1]); }
                if (MODE == 1) fma2(acc[c], acc[c + 1], x, x, w0, w1);
                if (MODE == 2) { fma2(acc[c], acc[c + 1], x, x, w0, w1); junk = (junk ^ (junk << 1)) + c; }
                if (MODE == 3) fma2(acc[c], acc[c + 1], x, x, wr[c], wr[c + 1]);
                if (MODE == 4) { acc[c] = fmaf(x, w0, acc[c]); acc[c + 1] = fmaf(x, w1, acc[c + 1]); junk = (junk ^ (junk << 1)) + c; }
                if (MODE == 5) { fma2(acc[c], acc[c + 1], x, x, w0, w1); junk = (junk ^ (junk << 1)) + c; junk = (junk ^ (junk >> 3)) + r; }
            }
        }
    }
    float s = (float)junk;
#pragma unroll
    for (int c = 0; c < 16; ++c) s += acc[c];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, const float4* in, float* out, const P& prm, int sms, double clk_ghz) {
    const int iters = 4096, blocks = sms * 8;
    k<MODE><<<blocks, 256>>>(in, out, 16, prm);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<blocks, 256>>>(in, out, iters, prm);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double macs = (double)blocks * 256 * iters * 128;
    printf("%-34s %8.3f ms  %7.2f TFLOP/s  %6.1f MAC/clk/SM (at %.2f GHz)\n", name, ms, 2 * macs / ms * 1e-9, macs / (ms * 1e-3) / sms / (clk_ghz * 1e9), clk_ghz);
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    float4* in; float* out; cudaMalloc(&in, 8192 * 16); cudaMemset(in, 0, 8192 * 16); cudaMalloc(&out, (size_t)pr.multiProcessorCount * 8 * 256 * 4);
    P prm; for (int i = 0; i < 512; ++i) prm.w[i] = 1.0f / (i + 1);
    printf("%s, %d SMs, max clock %.3f GHz\n", pr.name, pr.multiProcessorCount, ghz);
    run<0>("FFMA  (x scalar, w uniform)", in, out, prm, pr.multiProcessorCount, ghz);
    run<1>("FFMA2 (x broadcast, w UR pair)", in, out, prm, pr.multiProcessorCount, ghz);
    run<3>("FFMA2 (x broadcast, w reg pair)", in, out, prm, pr.multiProcessorCount, ghz);
    run<4>("FFMA  + 2 ALU per 2 MAC", in, out, prm, pr.multiProcessorCount, ghz);
    run<2>("FFMA2 + 2 ALU per 2 MAC", in, out, prm, pr.multiProcessorCount, ghz);
    run<5>("FFMA2 + 4 ALU per 2 MAC", in, out, prm, pr.multiProcessorCount, ghz);
    return 0;
}
