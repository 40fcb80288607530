// Micro-benchmark: how should warp-uniform conv weights be fed to the FP32 FMA pipe on sm_100a?
//
// The GTCRN-Micro layers are tiny channel contractions (K = 8/16/80/144, N = 16) applied to a
// huge number of independent positions (batch x time x freq).  One thread owns P positions and
// all 16 output channels, so activations live in registers and the only per-FFMA operand that
// must come from somewhere else is the (warp-uniform) weight.  Candidates measured here:
//   reg    : pure register FFMA chain                        -> measured FP32 peak (roofline denominator)
//   const  : weights as constant-bank operands  c[3][imm]    -> no load instruction at all
//   smem   : weights via broadcast LDS.128 from shared memory
//   ldg    : weights via uniform LDG.128 (L1 resident)
// Each "layer" is a 16->16 1x1 conv (256 weights) followed by a leaky-relu; NL distinct layers are
// streamed per iteration (footprint NL KB) to expose constant-cache / I-cache thrash.
//
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ffma_feed ffma_feed.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int MAXL = 60;                 // 60 KB of weights max (constant bank is 64 KB)
__constant__ float cW[MAXL * 256];

__device__ __forceinline__ float lrelu(float v) { return fmaxf(v, 0.25f * v); }

// ---------------------------------------------------------------- pure register peak
__global__ void __launch_bounds__(256) k_reg(float* out, int iters) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
    float m = 0.999f, c = 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------- constant-bank operands
// UNROLL_L = true : every layer is its own straight-line code (immediate constant offsets)
// UNROLL_L = false: one layer body, layer index is a runtime loop (uniform-register indexed constants)
template <int P, int NL, bool UNROLL_L>
__global__ void __launch_bounds__(256) k_const(float* out, int iters) {
    float x[P][16];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[p][i] = (threadIdx.x + p) * 1e-3f + i * 0.01f;
    for (int it = 0; it < iters; ++it) {
        if (UNROLL_L) {
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                float y[P][16];
#pragma unroll
                for (int co = 0; co < 16; ++co)
#pragma unroll
                    for (int p = 0; p < P; ++p) y[p][co] = 0.f;
#pragma unroll
                for (int ci = 0; ci < 16; ++ci)
#pragma unroll
                    for (int co = 0; co < 16; ++co)
#pragma unroll
                        for (int p = 0; p < P; ++p) y[p][co] = fmaf(cW[l * 256 + ci * 16 + co], x[p][ci], y[p][co]);
#pragma unroll
                for (int co = 0; co < 16; ++co)
#pragma unroll
                    for (int p = 0; p < P; ++p) x[p][co] = lrelu(y[p][co]);
            }
        } else {
#pragma unroll 1
            for (int l = 0; l < NL; ++l) {
                const float* w = cW + l * 256;
                float y[P][16];
#pragma unroll
                for (int co = 0; co < 16; ++co)
#pragma unroll
                    for (int p = 0; p < P; ++p) y[p][co] = 0.f;
#pragma unroll
                for (int ci = 0; ci < 16; ++ci)
#pragma unroll
                    for (int co = 0; co < 16; ++co)
#pragma unroll
                        for (int p = 0; p < P; ++p) y[p][co] = fmaf(w[ci * 16 + co], x[p][ci], y[p][co]);
#pragma unroll
                for (int co = 0; co < 16; ++co)
#pragma unroll
                    for (int p = 0; p < P; ++p) x[p][co] = lrelu(y[p][co]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int i = 0; i < 16; ++i) s += x[p][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------- broadcast LDS.128 / uniform LDG.128
template <int P, int NL, bool GLOBAL>
__global__ void __launch_bounds__(256) k_ld(const float* __restrict__ gW, float* out, int iters) {
    extern __shared__ float4 sW4[];
    if (!GLOBAL) {
        const float4* g4 = reinterpret_cast<const float4*>(gW);
        for (int i = threadIdx.x; i < NL * 64; i += blockDim.x) sW4[i] = g4[i];
        __syncthreads();
    }
    const float4* W4 = GLOBAL ? reinterpret_cast<const float4*>(gW) : sW4;
    float x[P][16];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[p][i] = (threadIdx.x + p) * 1e-3f + i * 0.01f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
            const float4* w = W4 + l * 64;
            float y[P][16];
#pragma unroll
            for (int co = 0; co < 16; ++co)
#pragma unroll
                for (int p = 0; p < P; ++p) y[p][co] = 0.f;
#pragma unroll
            for (int ci = 0; ci < 16; ++ci) {
                float4 w0 = GLOBAL ? __ldg(w + ci * 4 + 0) : w[ci * 4 + 0];
                float4 w1 = GLOBAL ? __ldg(w + ci * 4 + 1) : w[ci * 4 + 1];
                float4 w2 = GLOBAL ? __ldg(w + ci * 4 + 2) : w[ci * 4 + 2];
                float4 w3 = GLOBAL ? __ldg(w + ci * 4 + 3) : w[ci * 4 + 3];
                const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                                      w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
                for (int co = 0; co < 16; ++co)
#pragma unroll
                    for (int p = 0; p < P; ++p) y[p][co] = fmaf(wv[co], x[p][ci], y[p][co]);
            }
#pragma unroll
            for (int co = 0; co < 16; ++co)
#pragma unroll
                for (int p = 0; p < P; ++p) x[p][co] = lrelu(y[p][co]);
        }
    }
    float s = 0;
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int i = 0; i < 16; ++i) s += x[p][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

static float* d_out; static float* d_W;
static int g_blocks;

template <int P, int NL, bool U>
static void run_const(const char* tag, int threads, int smem_limit_kb = 0) {
    int iters = 12000 / (NL * P) + 1;
    // smem_limit_kb > 0: request that much dynamic shared memory per block to cap the number of resident blocks per SM
    if (smem_limit_kb) CK(cudaFuncSetAttribute(k_const<P, NL, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit_kb * 1024));
    float ms = time_ms([&] { k_const<P, NL, U><<<g_blocks, threads, smem_limit_kb * 1024>>>(d_out, iters); });
    double flop = 2.0 * 256 * P * NL * (double)iters * g_blocks * threads;
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_const<P, NL, U>, threads, smem_limit_kb * 1024);
    printf("const  %-8s P=%d NL=%2d (%2d KB) thr=%3d warps/SM=%2d : %7.3f ms  %6.2f TFLOP/s\n", tag, P, NL, NL, threads, occ * threads / 32, ms, flop / ms * 1e-9);
}
template <int P, int NL, bool G>
static void run_ld(int threads) {
    int iters = 12000 / (NL * P) + 1;
    size_t smem = G ? 0 : NL * 1024;
    CK(cudaFuncSetAttribute(k_ld<P, NL, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float ms = time_ms([&] { k_ld<P, NL, G><<<g_blocks, threads, smem>>>(d_W, d_out, iters); });
    double flop = 2.0 * 256 * P * NL * (double)iters * g_blocks * threads;
    printf("%-6s          P=%d NL=%2d (%2d KB) thr=%3d : %7.3f ms  %6.2f TFLOP/s\n", G ? "ldg" : "smem", P, NL, NL, threads, ms, flop / ms * 1e-9);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s  SMs %d  clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    const int sms = prop.multiProcessorCount;
    g_blocks = sms * 8;
    CK(cudaMalloc(&d_out, sizeof(float) * g_blocks * 256));
    float* hW = (float*)malloc(sizeof(float) * MAXL * 256);
    srand(1);
    for (int i = 0; i < MAXL * 256; ++i) hW[i] = ((rand() % 2001) - 1000) * 1e-4f;   // |w| <= 0.1 keeps values bounded
    CK(cudaMemcpyToSymbol(cW, hW, sizeof(float) * MAXL * 256));
    CK(cudaMalloc(&d_W, sizeof(float) * MAXL * 256));
    CK(cudaMemcpy(d_W, hW, sizeof(float) * MAXL * 256, cudaMemcpyHostToDevice));

    for (int threads : {128, 256}) {
        int iters = 4000;
        float ms = time_ms([&] { k_reg<<<g_blocks, threads>>>(d_out, iters); });
        double flop = 2.0 * 256 * (double)iters * g_blocks * threads;
        printf("reg    peak                        thr=%3d : %7.3f ms  %6.2f TFLOP/s\n", threads, ms, flop / ms * 1e-9);
    }
    // constant operands: footprint sweep, straight-line vs looped layer body
    run_const<1, 1, true>("unroll", 256);
    run_const<2, 1, true>("unroll", 256);
    run_const<2, 2, true>("unroll", 256);
    run_const<2, 8, true>("unroll", 256);
    run_const<2, 16, true>("unroll", 256);
    run_const<1, 8, false>("loop", 256);
    run_const<2, 8, false>("loop", 256);
    run_const<2, 8, false>("loop", 128);
    run_const<4, 8, false>("loop", 128);
    run_const<2, 32, false>("loop", 256);
    run_const<2, 60, false>("loop", 256);
    run_const<4, 60, false>("loop", 128);
    // occupancy sensitivity of the constant-bank feed (LDCU + FFMA.UR): 64 / 32 / 16 / 8 warps per SM
    run_const<2, 8, false>("loop", 256, 0);
    run_const<2, 8, false>("loop", 256, 50);
    run_const<2, 8, false>("loop", 256, 100);
    run_const<2, 8, false>("loop", 256, 200);
    run_const<4, 8, false>("loop", 256, 100);
    run_const<4, 8, false>("loop", 256, 200);
    run_const<1, 8, false>("loop", 256, 0);
    run_const<1, 8, false>("loop", 256, 50);
    // broadcast LDS.128
    run_ld<1, 8, false>(256);
    run_ld<2, 8, false>(256);
    run_ld<4, 8, false>(256);
    run_ld<4, 8, false>(128);
    run_ld<2, 32, false>(256);
    run_ld<4, 32, false>(128);
    // uniform LDG.128 through L1
    run_ld<2, 8, true>(256);
    run_ld<4, 8, true>(128);
    run_ld<4, 32, true>(128);
    printf("done\n");
    return 0;
}
